// svoxb_vertical.cu -- the remaining point-wise operators of the reference's `svox_t.csrc` module:
//   query_vertical_backward (svox_kernel.cu:83-95, 380-403), assign_vertical (svox_kernel.cu:97-108, 326-339),
//   calc_corners (svox_kernel.cu:213-237, 436-457) and grid_weight_render (rt_kernel.cu:1240-1344, 1454-1478).
// The first two fault in the reference (get_tree_leaf_ptr writes through a null data_id, svox_kernel.cu:61-62) and
// calc_corners dispatches on an int32 tensor and raises (svox_kernel.cu:446-448); they are built here to the semantics
// their source states. Layout follows the rest of the library: lane = point for the descent, whole warps for row
// traffic (coalesced 4*K-byte rows instead of one thread looping over a row).
#include "svoxb_march.cuh"

namespace svoxb {

int make_tree_args(const svoxb_tree* t, TreeArgs& a);   // svoxb_tree.cu

// Leaf row of one world-space point, or -1 (empty leaf / lane without a point).
__device__ __forceinline__ int point_row(const TreeArgs& tr, const float* __restrict__ pts, int64_t q) {
    const float px = fmaf(__ldg(tr.scaling), __ldg(pts + 3 * q), __ldg(tr.offset));
    const float py = fmaf(__ldg(tr.scaling + 1), __ldg(pts + 3 * q + 1), __ldg(tr.offset + 1));
    const float pz = fmaf(__ldg(tr.scaling + 2), __ldg(pts + 3 * q + 2), __ldg(tr.offset + 2));
    float rx, ry, rz, cube;
    const int64_t slot = descend_ref(tr.child, tr.N, px, py, pz, rx, ry, rz, cube);
    const int di = __ldg(tr.data + slot);
    return (di >= 0 && (int64_t)di < tr.M) ? di : -1;                 // svox_kernel.cu:61
}

// grad_data[row(p_q), :K] += grad_out[q, :K]: one reduction instruction per 32 channels of a row.
__global__ void __launch_bounds__(256)
query_bwd_kernel(TreeArgs tr, const float* __restrict__ pts, int64_t Q, const float* __restrict__ grad_out, int K,
                 float* __restrict__ grad_data) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t base = warp_id * 32; base < Q; base += warps_total * 32) {
        const int64_t q = base + lane;
        const int idx = q < Q ? point_row(tr, pts, q) : -1;
        unsigned vm = __ballot_sync(FULL, idx >= 0);
        while (vm) {
            const int r = __ffs(vm) - 1;
            vm &= vm - 1;
            const int idx_r = __shfl_sync(FULL, idx, r);
            SVOXB_DBG(idx_r >= 0 && idx_r < tr.M);
            const float* src = grad_out + (base + r) * K;
            float* dst = grad_data + (int64_t)idx_r * K;
            for (int c = lane; c < K; c += 32) atomicAdd(dst + c, __ldg(src + c));
        }
    }
}

// assign_vertical, deterministic: phase 0 elects the largest point index per row, phase 1 lets the winner copy.
__global__ void __launch_bounds__(256)
assign_kernel(TreeArgs tr, float* __restrict__ features_mut, const float* __restrict__ pts, int64_t Q,
              const float* __restrict__ values, int K, int* __restrict__ winner, int phase) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t base = warp_id * 32; base < Q; base += warps_total * 32) {
        const int64_t q = base + lane;
        int idx = q < Q ? point_row(tr, pts, q) : -1;
        if (phase == 0) {
            if (idx >= 0) atomicMax(winner + idx, (int)q);
            continue;
        }
        if (idx >= 0 && winner[idx] != (int)q) idx = -1;
        unsigned vm = __ballot_sync(FULL, idx >= 0);
        while (vm) {
            const int r = __ffs(vm) - 1;
            vm &= vm - 1;
            const int idx_r = __shfl_sync(FULL, idx, r);
            SVOXB_DBG(idx_r >= 0 && idx_r < tr.M && base + r < Q);
            const float* src = values + (base + r) * K;
            float* dst = features_mut + (int64_t)idx_r * tr.D;
            for (int c = lane; c < K; c += 32) dst[c] = __ldg(src + c);
        }
    }
}

// Lower corner of cell [node, i, j, k] in tree coordinates: walk parent_depth to the root, (corner + ijk) / N per level.
__global__ void __launch_bounds__(256)
calc_corners_kernel(const int32_t* __restrict__ parent_depth, int N, int64_t n_nodes,
                    const int64_t* __restrict__ indexer, int64_t Q, float* __restrict__ out) {
    const float fN = (float)N;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
        int node = (int)indexer[4 * q];
        int c1 = (int)indexer[4 * q + 1], c2 = (int)indexer[4 * q + 2], c3 = (int)indexer[4 * q + 3];
        float x = 0.0f, y = 0.0f, z = 0.0f;
        while (true) {
            x = (x + (float)c1) / fN; y = (y + (float)c2) / fN; z = (z + (float)c3) / fN;
            if (node <= 0 || node >= n_nodes) break;
            int packed = __ldg(parent_depth + 2 * (int64_t)node);
            c3 = packed % N; packed /= N;
            c2 = packed % N; packed /= N;
            c1 = packed % N; packed /= N;
            node = packed;
        }
        out[3 * q] = x; out[3 * q + 1] = y; out[3 * q + 2] = z;
    }
}

// grid_weight_render: one lane per pixel, warps cover 8x4 pixel tiles (coherent cells). Weights are >= 0 whenever they
// can win against the zero-initialised grid, so the reference's CAS-loop float max is one integer atomicMax.
__global__ void __launch_bounds__(256)
grid_weight_kernel(const float* __restrict__ grid, int reso, RaySource src, const float* __restrict__ offset,
                   const float* __restrict__ scaling, float step, float sigma_thresh, float* __restrict__ grid_weight,
                   float* __restrict__ grid_hit) {
    const int lane = threadIdx.x & 31;
    const int tiles_x = (src.width + 7) >> 3, tiles_y = (src.height + 3) >> 2;
    const int64_t n_tiles = (int64_t)tiles_x * tiles_y;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float fr = (float)reso;
    for (int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < n_tiles; tile += warps_total) {
        const int px = (int)(tile % tiles_x) * 8 + (lane & 7), py = (int)(tile / tiles_x) * 4 + (lane >> 3);
        if (px >= src.width || py >= src.height) continue;
        float ox, oy, oz, dx, dy, dz;
        camera_ray(src, px, py, ox, oy, oz, dx, dy, dz);
        if (src.ndc_w >= 0) world2ndc(src, ox, oy, oz, dx, dy, dz);
        Ray r;
        ray_setup(offset, scaling, ox, oy, oz, dx, dy, dz, r);      // a miss leaves t == tmax: no samples
        float T = 1.0f;
        while (r.t < r.tmax) {
            float x = clamp01(fmaf(r.t, r.dx, r.ox)) * fr, y = clamp01(fmaf(r.t, r.dy, r.oy)) * fr,
                  z = clamp01(fmaf(r.t, r.dz, r.oz)) * fr;
            const float fu = floorf(x), fv = floorf(y), fw = floorf(z);
            x -= fu; y -= fv; z -= fw;
            const int64_t cell = ((int64_t)(int)fu * reso + (int)fv) * reso + (int)fw;
            SVOXB_DBG(cell >= 0 && cell < (int64_t)reso * reso * reso);
            float smin, smax;
            dda_unit(x, y, z, r.ix, r.iy, r.iz, smin, smax);
            const float delta_t = (smax - smin) / fr + step;
            const float sigma = __ldg(grid + cell);
            if (sigma > sigma_thresh) {
                const float att = expf(-delta_t * r.ds * sigma);
                const float w = T * (1.0f - att);
                T *= att;
                atomicMax(reinterpret_cast<int*>(grid_weight) + cell, __float_as_int(w));
                atomicAdd(grid_hit + cell, 1.0f);
            }
            r.t += delta_t;
        }
    }
}

}  // namespace svoxb

using namespace svoxb;

extern "C" int svoxb_query_bwd(const svoxb_tree* tree, const float* pts, int64_t Q, const float* grad_out, int32_t K,
                               float* grad_data, void* stream) {
    TreeArgs tr;
    int rc = make_tree_args(tree, tr); if (rc) return rc;
    SVOXB_REQUIRE(Q >= 0 && K >= 1, "query_bwd: bad sizes");
    if (Q == 0 || tree->M == 0) return 0;
    SVOXB_REQUIRE(pts && grad_out && grad_data, "query_bwd: NULL tensor");
    const int grid = (int)min((Q + 255) / 256, (int64_t)sm_count() * 8);
    query_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tr, pts, Q, grad_out, K, grad_data);
    count_launch();
    return check_cuda(cudaGetLastError(), "query_bwd_kernel launch");
}

extern "C" int svoxb_assign(const svoxb_tree* tree, float* features_mut, const float* pts, int64_t Q,
                            const float* values, int32_t K, void* stream) {
    TreeArgs tr;
    int rc = make_tree_args(tree, tr); if (rc) return rc;
    SVOXB_REQUIRE(Q >= 0 && Q < (1ll << 31), "assign: point count out of range");
    SVOXB_REQUIRE(K >= 1 && K <= tree->D, "assign: values have %d channels, rows have %d", K, tree->D);
    if (Q == 0 || tree->M == 0) return 0;
    SVOXB_REQUIRE(features_mut && pts && values, "assign: NULL tensor");
    cudaStream_t st = (cudaStream_t)stream;
    int* winner = nullptr;
    SVOXB_CUDA(cudaMallocAsync((void**)&winner, sizeof(int) * (size_t)tree->M, st));
    rc = check_cuda(cudaMemsetAsync(winner, 0xff, sizeof(int) * (size_t)tree->M, st), "assign: memset");
    if (rc == 0) {
        const int grid = (int)min((Q + 255) / 256, (int64_t)sm_count() * 8);
        assign_kernel<<<grid, 256, 0, st>>>(tr, features_mut, pts, Q, values, K, winner, 0);
        assign_kernel<<<grid, 256, 0, st>>>(tr, features_mut, pts, Q, values, K, winner, 1);
        count_launch(2);
        rc = check_cuda(cudaGetLastError(), "assign_kernel launch");
    }
    cudaFreeAsync(winner, st);
    return rc;
}

extern "C" int svoxb_calc_corners(const int32_t* parent_depth, int32_t N, int64_t n_nodes, const int64_t* indexer,
                                  int64_t Q, float* out, void* stream) {
    SVOXB_REQUIRE(N >= 2 && N <= 16 && n_nodes >= 1 && Q >= 0, "calc_corners: bad sizes");
    if (Q == 0) return 0;
    SVOXB_REQUIRE(parent_depth && indexer && out, "calc_corners: NULL tensor");
    const int grid = (int)min((Q + 255) / 256, (int64_t)sm_count() * 8);
    calc_corners_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(parent_depth, N, n_nodes, indexer, Q, out);
    count_launch();
    return check_cuda(cudaGetLastError(), "calc_corners_kernel launch");
}

extern "C" int svoxb_grid_weight_render(const float* grid, int32_t reso, const svoxb_camera* cam,
                                        const svoxb_render_options* opt, const float* offset, const float* scaling,
                                        float* grid_weight, float* grid_hit, void* stream) {
    SVOXB_REQUIRE(grid && cam && opt && offset && scaling && grid_weight && grid_hit, "grid_weight_render: NULL argument");
    SVOXB_REQUIRE(reso >= 1 && reso <= 2048, "grid_weight_render: reso=%d out of range", reso);
    SVOXB_REQUIRE(cam->c2w && cam->width >= 1 && cam->height >= 1, "grid_weight_render: bad camera");
    RaySource src;
    memset(&src, 0, sizeof(src));
    src.c2w = cam->c2w; src.fx = cam->fx; src.fy = cam->fy; src.width = cam->width; src.height = cam->height;
    src.ndc_w = opt->ndc_width; src.ndc_h = opt->ndc_height; src.ndc_focal = opt->ndc_focal;
    const int64_t n_tiles = (int64_t)((cam->width + 7) / 8) * ((cam->height + 3) / 4);
    const int grid_dim = (int)min((n_tiles + 7) / 8, (int64_t)sm_count() * 8);
    grid_weight_kernel<<<grid_dim, 256, 0, (cudaStream_t)stream>>>(grid, reso, src, offset, scaling, opt->step_size,
                                                                  opt->sigma_thresh, grid_weight, grid_hit);
    count_launch();
    return check_cuda(cudaGetLastError(), "grid_weight_kernel launch");
}
