"""Point-to-voxel Gaussian splat (reference: svox_t/p2v.py:33-53, forward only; the backward is a next-rank
component)."""
from torch import autograd

from . import csrc as _C


class _VoxelizationFunction(autograd.Function):
    @staticmethod
    def forward(ctx, points, point_features, volume_corner, volume_size, n_voxels, kernel_radius, conv_radius):
        return _C.p2v(points, point_features, volume_corner, volume_size, n_voxels, kernel_radius, conv_radius)

    @staticmethod
    def backward(ctx, grad_output):
        raise RuntimeError("p2v backward is not implemented in svox_t_b200 (SURVEY 8f rank 2)")


def voxelize(points, point_features, volume_corner, volume_size, n_voxels, kernel_radius, conv_radius):
    return _VoxelizationFunction.apply(points, point_features, volume_corner, volume_size, n_voxels, kernel_radius,
                                       conv_radius)
