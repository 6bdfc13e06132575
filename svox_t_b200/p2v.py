"""Point-to-voxel Gaussian splat with its backward (reference: svox_t/p2v.py:33-53)."""
from torch import autograd

from . import csrc as _C


class _VoxelizationFunction(autograd.Function):
    @staticmethod
    def forward(ctx, points, point_features, volume_corner, volume_size, n_voxels, kernel_radius, conv_radius):
        ctx.save_for_backward(points, point_features, volume_corner, volume_size)
        ctx.n_voxels, ctx.kernel_radius, ctx.conv_radius = n_voxels, kernel_radius, conv_radius
        return _C.p2v(points, point_features, volume_corner, volume_size, n_voxels, kernel_radius, conv_radius)

    @staticmethod
    def backward(ctx, grad_output):
        points, feats, corner, size = ctx.saved_tensors
        g_p, g_f = _C.p2v_backward(grad_output.contiguous(), points, feats, corner, size, ctx.n_voxels,
                                   ctx.kernel_radius, ctx.conv_radius)
        return g_p, g_f, None, None, None, None, None


def voxelize(points, point_features, volume_corner, volume_size, n_voxels, kernel_radius, conv_radius):
    return _VoxelizationFunction.apply(points, point_features, volume_corner, volume_size, n_voxels, kernel_radius,
                                       conv_radius)
