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

// ---- backward of the two point kernels -------------------------------------------------------------
// warp_vertices_kernel_backward (svox_kernel.cu:156-211). One thread per point; the J x 12 bone-matrix gradient,
// which in the reference receives 24*B global atomics per point on a few dozen addresses, is first reduced in a
// per-CTA shared-memory table and flushed with one global atomic per entry per CTA.
__global__ void __launch_bounds__(256)
warp_vertices_bwd_kernel(const float* __restrict__ T, const float* __restrict__ coords, const float* __restrict__ w,
                         const int32_t* __restrict__ jidx, const float* __restrict__ g_coords /* [P,3] */,
                         const float* __restrict__ g_mats /* [P,4,4] */, int64_t P, int B, int J, int use_smem,
                         float* __restrict__ grad_T /* [J,4,4] */, float* __restrict__ grad_coords /* [P,3] */,
                         float* __restrict__ grad_w /* [P,B] */) {
    extern __shared__ float sT[];                    // [J][12] partial bone gradients of this CTA
    if (use_smem) {
        for (int i = threadIdx.x; i < J * 12; i += blockDim.x) sT[i] = 0.0f;
        __syncthreads();
    }
    for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x; p0 < P; p0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = p0 + threadIdx.x;
        if (p < P) {
            const float x[3] = {__ldg(coords + 3 * p), __ldg(coords + 3 * p + 1), __ldg(coords + 3 * p + 2)};
            const float gc[3] = {__ldg(g_coords + 3 * p), __ldg(g_coords + 3 * p + 1), __ldg(g_coords + 3 * p + 2)};
            float m[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            float gm[3][4];          // total upstream gradient of the blended matrix rows 0..2
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(g_mats) + p * 4 + r);
                gm[r][0] = g.x + gc[r] * x[0]; gm[r][1] = g.y + gc[r] * x[1];
                gm[r][2] = g.z + gc[r] * x[2]; gm[r][3] = g.w + gc[r];
            }
            for (int b = 0; b < B; ++b) {
                const float wb = __ldg(w + p * B + b);
                float gw = 0.0f;
                if (wb > 0.f) {
                    const int j = __ldg(jidx + p * B + b);
                    const float4* Tj = reinterpret_cast<const float4*>(T) + (int64_t)j * 4;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float4 t = __ldg(Tj + r);
                        m[r][0] += wb * t.x; m[r][1] += wb * t.y; m[r][2] += wb * t.z; m[r][3] += wb * t.w;
                        gw += t.x * gm[r][0] + t.y * gm[r][1] + t.z * gm[r][2] + t.w * gm[r][3];
                        float* dst = use_smem ? sT + j * 12 + r * 4 : grad_T + (int64_t)j * 16 + r * 4;
                        atomicAdd(dst + 0, wb * gm[r][0]); atomicAdd(dst + 1, wb * gm[r][1]);
                        atomicAdd(dst + 2, wb * gm[r][2]); atomicAdd(dst + 3, wb * gm[r][3]);
                    }
                }
                grad_w[p * B + b] = gw;
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) grad_coords[3 * p + i] = gc[0] * m[0][i] + gc[1] * m[1][i] + gc[2] * m[2][i];
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < J * 12; i += blockDim.x) {
            const float v = sT[i];
            if (v != 0.0f) atomicAdd(grad_T + (int64_t)(i / 12) * 16 + (i % 12), v);
        }
    }
}

// p2v_kernel_backward (p2v_kernel.cu:153-234). Eight threads per point (x-slices of the footprint), partial sums
// combined with three shuffles, no atomics. The feature gradient goes to channel 0, as in the reference (App. B12).
__global__ void __launch_bounds__(256)
p2v_bwd_kernel(const float* __restrict__ g_vox, const float* __restrict__ points, const float* __restrict__ feat,
               int64_t P, int F, const float* __restrict__ corner, const float* __restrict__ size, int n, float kr,
               float cr, float* __restrict__ grad_points, float* __restrict__ grad_feat) {
    const float c0 = __ldg(corner), c1 = __ldg(corner + 1), c2 = __ldg(corner + 2);
    const float vs0 = __ldg(size) / (float)(n - 1), vs1 = __ldg(size + 1) / (float)(n - 1),
                vs2 = __ldg(size + 2) / (float)(n - 1);
    const float inv2k = 1.0f / (2 * kr * kr), invk2 = 1.0f / (kr * kr);
    const float cr2 = cr * cr * 1.0001f;
    const int64_t total = ((P * 8 + 31) / 32) * 32;                  // whole warps: the shuffles need every lane
    for (int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
         gid += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = gid >> 3;
        const int xi = (int)(gid & 7);
        float gx = 0.f, gy = 0.f, gz = 0.f, gf = 0.f;
        if (p < P) {
            const float x = __ldg(points + 3 * p), y = __ldg(points + 3 * p + 1), z = __ldg(points + 3 * p + 2);
            const float sg = __ldg(feat + p * F + (F - 1));
            const int lx = clampi((int)floorf((x - cr - c0) / vs0), 0, n - 1), hx = clampi((int)ceilf((x + cr - c0) / vs0), 0, n - 1);
            const int ly = clampi((int)floorf((y - cr - c1) / vs1), 0, n - 1), hy = clampi((int)ceilf((y + cr - c1) / vs1), 0, n - 1);
            const int lz = clampi((int)floorf((z - cr - c2) / vs2), 0, n - 1), hz = clampi((int)ceilf((z + cr - c2) / vs2), 0, n - 1);
            for (int vx = lx + xi; vx <= hx; vx += 8) {
                const float dx = x - ((float)vx * vs0 + c0);
                if (dx * dx > cr2) continue;
                const float ex = expf(-(dx * dx) * inv2k);
                for (int vy = ly; vy <= hy; ++vy) {
                    const float dy = y - ((float)vy * vs1 + c1);
                    if (dx * dx + dy * dy > cr2) continue;
                    const float exy = ex * expf(-(dy * dy) * inv2k);
                    const float* rowp = g_vox + ((int64_t)vx * n + vy) * n;
                    for (int vz = lz; vz <= hz; ++vz) {
                        const float dz = z - ((float)vz * vs2 + c2);
                        const float r = sqrtf(dx * dx + dy * dy + dz * dz);
                        if (r <= cr) {
                            const float wgt = exy * expf(-(dz * dz) * inv2k);
                            const float og = __ldg(rowp + vz);
                            gf += og * wgt;                                  // p2v_kernel.cu:203
                            const float k = -(og * sg) * wgt * invk2;        // p2v_kernel.cu:206-226
                            gx += k * dx; gy += k * dy; gz += k * dz;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int s = 4; s > 0; s >>= 1) {
            gx += __shfl_xor_sync(FULL, gx, s); gy += __shfl_xor_sync(FULL, gy, s);
            gz += __shfl_xor_sync(FULL, gz, s); gf += __shfl_xor_sync(FULL, gf, s);
        }
        if (xi == 0 && p < P) {
            grad_points[3 * p] = gx; grad_points[3 * p + 1] = gy; grad_points[3 * p + 2] = gz;
            grad_feat[p * F] = gf;
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

extern "C" int svoxb_warp_vertices_bwd(const float* T, const float* coords, const float* w, const int32_t* joint_index,
                                       const float* grad_coords_out, const float* grad_mats_out, int64_t P, int32_t B,
                                       int32_t J, float* grad_T, float* grad_coords, float* grad_w, void* stream) {
    SVOXB_REQUIRE(P >= 0 && B >= 0 && J >= 1, "bad sizes");
    SVOXB_REQUIRE(grad_T != nullptr, "grad_T is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    SVOXB_CUDA(cudaMemsetAsync(grad_T, 0, sizeof(float) * 16 * (size_t)J, st));
    if (P == 0) return 0;
    SVOXB_REQUIRE(T && coords && w && joint_index && grad_coords_out && grad_mats_out && grad_coords && grad_w, "NULL tensor");
    SVOXB_REQUIRE(((uintptr_t)T & 15) == 0 && ((uintptr_t)grad_mats_out & 15) == 0, "matrices must be 16-byte aligned");
    const size_t smem = sizeof(float) * 12 * (size_t)J;
    const int use_smem = smem <= 48 * 1024;
    const int grid = (int)min((P + 255) / 256, (int64_t)sm_count() * 4);
    warp_vertices_bwd_kernel<<<grid, 256, use_smem ? smem : 0, st>>>(T, coords, w, joint_index, grad_coords_out,
                                                                    grad_mats_out, P, B, J, use_smem, grad_T,
                                                                    grad_coords, grad_w);
    count_launch();
    return check_cuda(cudaGetLastError(), "warp_vertices_bwd_kernel launch");
}

extern "C" int svoxb_p2v_bwd(const float* grad_voxels, const float* points, const float* point_features, int64_t P,
                             int32_t F, const float* corner, const float* size, int32_t n_voxels, float kernel_radius,
                             float conv_radius, float* grad_points, float* grad_features, void* stream) {
    SVOXB_REQUIRE(n_voxels >= 2 && F >= 1 && P >= 0, "bad sizes");
    if (P == 0) return 0;
    SVOXB_REQUIRE(grad_voxels && points && point_features && corner && size && grad_points && grad_features, "NULL tensor");
    cudaStream_t st = (cudaStream_t)stream;
    SVOXB_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)P * F, st));
    const int grid = (int)min((P * 8 + 255) / 256, (int64_t)sm_count() * 32);
    p2v_bwd_kernel<<<grid, 256, 0, st>>>(grad_voxels, points, point_features, P, F, corner, size, n_voxels,
                                         kernel_radius, conv_radius, grad_points, grad_features);
    count_launch();
    return check_cuda(cudaGetLastError(), "p2v_bwd_kernel launch");
}
