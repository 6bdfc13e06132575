// svoxb_points.cu -- per-point kernels of the animated-frame rebuild: linear blend skinning of the voxel centres
// and the Gaussian point-to-voxel splat.
//
// Replaces (reference paths relative to /root/reference/svox_t/csrc):
//   warp_vertices_kernel / warp_vertices      svox_kernel.cu:123-154, 354-378
//   p2v_kernel / p2v                          p2v_kernel.cu:103-151, 240-261
#include "svoxb_common.cuh"

namespace svoxb {

// Four threads per point: thread r builds row r of the blended 4x4 (one float4 store, consecutive threads write
// consecutive 16-byte segments of mats_out) and, for r < 3, the r-th warped coordinate. The reference uses one
// thread per point with 12*B atomicAdds on its own output row (svox_kernel.cu:139-146).
__global__ void __launch_bounds__(256)
warp_vertices_kernel(const float* __restrict__ T, const float* __restrict__ coords, const float* __restrict__ w,
                     const int32_t* __restrict__ jidx, int64_t P, int B, float* __restrict__ coords_out,
                     float4* __restrict__ mats_out) {
    for (int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gid < P * 4;
         gid += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = gid >> 2;
        const int r = (int)(gid & 3);
        float4 m = make_float4(0.f, 0.f, 0.f, r == 3 ? 1.f : 0.f);
        if (r < 3) {
            for (int b = 0; b < B; ++b) {
                const float wb = __ldg(w + p * B + b);
                if (wb > 0.f) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(T) + (int64_t)__ldg(jidx + p * B + b) * 4 + r);
                    m.x += wb * t.x; m.y += wb * t.y; m.z += wb * t.z; m.w += wb * t.w;
                }
            }
            const float x = __ldg(coords + 3 * p), y = __ldg(coords + 3 * p + 1), z = __ldg(coords + 3 * p + 2);
            coords_out[3 * p + r] = x * m.x + y * m.y + z * m.z + m.w;
        }
        mats_out[gid] = m;
    }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Eight threads per point, one per x-slice of the footprint (the footprint of the reference's configurations is
// at most 6 cells per axis; wider ones loop). The Gaussian is separable, exp(-r^2/2k^2) = ex*ey*ez, so a thread
// evaluates ~13 exponentials instead of one per cell; the cutoff r <= conv_radius is the reference's own test.
// The reference walks the whole footprint serially in one thread per point with an expf per cell.
__global__ void __launch_bounds__(256)
p2v_kernel(const float* __restrict__ points, const float* __restrict__ feat, int64_t P, int F,
           const float* __restrict__ corner, const float* __restrict__ size, int n, float kr, float cr,
           float* __restrict__ voxels) {
    const float c0 = __ldg(corner), c1 = __ldg(corner + 1), c2 = __ldg(corner + 2);
    const float vs0 = __ldg(size) / (float)(n - 1), vs1 = __ldg(size + 1) / (float)(n - 1),
                vs2 = __ldg(size + 2) / (float)(n - 1);
    const float inv2k = 1.0f / (2 * kr * kr);
    const float cr2 = cr * cr * 1.0001f;          // slack: the early-outs must never drop a cell the exact test keeps
    for (int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gid < P * 8;
         gid += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = gid >> 3;
        const int xi = (int)(gid & 7);
        const float x = __ldg(points + 3 * p), y = __ldg(points + 3 * p + 1), z = __ldg(points + 3 * p + 2);
        const float sg = __ldg(feat + p * F + (F - 1));
        const int lx = clampi((int)floorf((x - cr - c0) / vs0), 0, n - 1), hx = clampi((int)ceilf((x + cr - c0) / vs0), 0, n - 1);
        const int ly = clampi((int)floorf((y - cr - c1) / vs1), 0, n - 1), hy = clampi((int)ceilf((y + cr - c1) / vs1), 0, n - 1);
        const int lz = clampi((int)floorf((z - cr - c2) / vs2), 0, n - 1), hz = clampi((int)ceilf((z + cr - c2) / vs2), 0, n - 1);
        for (int vx = lx + xi; vx <= hx; vx += 8) {
            const float dx = x - ((float)vx * vs0 + c0);
            const float dx2 = dx * dx;
            if (dx2 > cr2) continue;
            const float ex = expf(-dx2 * inv2k) * sg;
            for (int vy = ly; vy <= hy; ++vy) {
                const float dy = y - ((float)vy * vs1 + c1);
                const float dxy2 = dx2 + dy * dy;
                if (dxy2 > cr2) continue;
                const float exy = ex * expf(-(dy * dy) * inv2k);
                float* rowp = voxels + ((int64_t)vx * n + vy) * n;
                for (int vz = lz; vz <= hz; ++vz) {
                    const float dz = z - ((float)vz * vs2 + c2);
                    // the cutoff exactly as the reference states it (p2v_kernel.cu:137-139): r = sqrt(...) <= conv_radius
                    const float r = sqrtf(dx * dx + dy * dy + dz * dz);
                    if (r <= cr) atomicAdd(rowp + vz, exy * expf(-(dz * dz) * inv2k));
                }
            }
        }
    }
}

}  // namespace svoxb

using namespace svoxb;

extern "C" int svoxb_warp_vertices(const float* T, const float* coords, const float* w, const int32_t* joint_index,
                                   int64_t P, int32_t B, float* coords_out, float* mats_out, void* stream) {
    SVOXB_REQUIRE(P >= 0 && B >= 0, "bad sizes");
    if (P == 0) return 0;
    SVOXB_REQUIRE(T && coords && w && joint_index && coords_out && mats_out, "NULL tensor");
    SVOXB_REQUIRE(((uintptr_t)T & 15) == 0 && ((uintptr_t)mats_out & 15) == 0, "matrices must be 16-byte aligned");
    const int grid = (int)min((P * 4 + 255) / 256, (int64_t)sm_count() * 16);
    warp_vertices_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(T, coords, w, joint_index, P, B, coords_out,
                                                                reinterpret_cast<float4*>(mats_out));
    count_launch();
    return check_cuda(cudaGetLastError(), "warp_vertices_kernel launch");
}

extern "C" int svoxb_p2v(const float* points, const float* point_features, int64_t P, int32_t F, const float* corner,
                         const float* size, int32_t n_voxels, float kernel_radius, float conv_radius, float* voxels,
                         void* stream) {
    SVOXB_REQUIRE(n_voxels >= 2 && F >= 1 && P >= 0, "bad sizes");
    SVOXB_REQUIRE(voxels && corner && size && (P == 0 || (points && point_features)), "NULL tensor");
    cudaStream_t st = (cudaStream_t)stream;
    SVOXB_CUDA(cudaMemsetAsync(voxels, 0, sizeof(float) * (size_t)n_voxels * n_voxels * n_voxels, st));
    if (P == 0) return 0;
    const int grid = (int)min((P * 8 + 255) / 256, (int64_t)sm_count() * 32);
    p2v_kernel<<<grid, 256, 0, st>>>(points, point_features, P, F, corner, size, n_voxels, kernel_radius,
                                     conv_radius, voxels);
    count_launch();
    return check_cuda(cudaGetLastError(), "p2v_kernel launch");
}
