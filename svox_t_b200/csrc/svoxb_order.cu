// svoxb_order.cu -- longest-first ray order for the BACKWARD over explicit ray batches (no reference counterpart: the
// reference runs one thread per ray, rt_kernel.cu:654-671, and has no queue to order).
//
// The march kernels are persistent: warps pull 32-ray entries from a global queue and refill a lane as soon as its ray
// ends. Marching the rays sorted by length -- longest first -- helps the backward twice (measured on B200, C3 tree):
//   * short batches (one GPU's share of a batch split over 8 GPUs: 131 072 rays on 113 664 backward lanes): the rays
//     that do not fit the first wave start whenever a lane frees up and the kernel ends with whichever of them happens to
//     be long -- 1.10 ms; longest first leaves the short rays for the second wave (longest-processing-time scheduling):
//     0.79 ms;
//   * ANY batch size: the 32 rays of a warp are alike, so they are inside the object -- hits, gradient reductions -- and
//     outside of it -- nothing to do -- at the same time; with random neighbours nearly every iteration of a warp has a
//     few hits and pays for the whole gradient stage. 2^20 rays: 6.15 -> 4.94 ms, 10 % fewer instructions per ray.
// Per-ray results do not depend on which lane serves a ray.
//
// The cost of a ray is its EXACT number of march iterations, which the forward over the same batch writes as a
// by-product (RaySource::steps_out: one counter register per lane, stored through a __noinline__ call so that the
// 72-register loop does not spill; svoxb_render_rays_fwd_cost -> svoxb_render_rays_bwd_cost). Two small kernels:
// histogram, then a counting-sort scatter (descending cost; the order within a cost bin is whatever the atomics
// produce). At 64 k rays and below the two extra kernels cost more than the order gains (0.52 -> 0.56 ms).
// The FORWARD is not ordered: its costs are not known before it runs (an estimate is a kernel of its own and measured
// slower than it gained).
#include <stdlib.h>
#include "svoxb_march.cuh"

namespace svoxb {

constexpr int ORD_BINS = 1024;
constexpr int ORD_THREADS = 256;
constexpr int ORD_ITEMS = 4;          // rays per thread and tile of the scatter

// Histogram of the per-ray costs (the forward's exact iteration counts).
__global__ void __launch_bounds__(ORD_THREADS)
ray_hist_kernel(const int* __restrict__ cost, int Q, unsigned* __restrict__ hist) {
    __shared__ unsigned h[ORD_BINS];
    for (int i = threadIdx.x; i < ORD_BINS; i += ORD_THREADS) h[i] = 0;
    __syncthreads();
    for (int r = blockIdx.x * ORD_THREADS + threadIdx.x; r < Q; r += gridDim.x * ORD_THREADS)
        atomicAdd(&h[max(0, min(ORD_BINS - 1, __ldg(cost + r)))], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < ORD_BINS; i += ORD_THREADS)
        if (h[i]) atomicAdd(hist + i, h[i]);
}

// order[first position of the ray's bin in descending-cost order + running count of the bin] = ray
__global__ void __launch_bounds__(ORD_THREADS)
ray_order_scatter_kernel(const int* __restrict__ cost, int Q, const unsigned* __restrict__ hist,
                         unsigned* __restrict__ cursor, int* __restrict__ order) {
    __shared__ unsigned base[ORD_BINS];
    __shared__ unsigned part[ORD_THREADS];
    // exclusive scan of the histogram read from the top bin down: 4 bins per thread, then a scan of the partials
    constexpr int PER = ORD_BINS / ORD_THREADS;
    unsigned loc[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        loc[j] = sum;
        sum += __ldg(hist + (ORD_BINS - 1 - (threadIdx.x * PER + j)));
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int s = 1; s < ORD_THREADS; s <<= 1) {
        const unsigned v = threadIdx.x >= s ? part[threadIdx.x - s] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    const unsigned before = part[threadIdx.x] - sum;
#pragma unroll
    for (int j = 0; j < PER; ++j) base[ORD_BINS - 1 - (threadIdx.x * PER + j)] = before + loc[j];
    __syncthreads();
    if (__ldg(cost) < 0) {      // the forward could not count (svoxb_render_rays_fwd_cost): keep the caller's order
        for (int r = blockIdx.x * ORD_THREADS + threadIdx.x; r < Q; r += gridDim.x * ORD_THREADS) order[r] = r;
        return;
    }
    // Tiles of ORD_THREADS * ORD_ITEMS rays: ranks inside the tile from shared-memory counters, then ONE global
    // reservation per occupied bin of the tile. (A global atomic per ray lands on the ~200 distinct step counts of a
    // batch: 0.21 ms for 2^20 rays, as much as the table pass of the whole step.)
    __shared__ unsigned lcount[ORD_BINS];
    constexpr int TILE = ORD_THREADS * ORD_ITEMS;
    for (int t0 = blockIdx.x * TILE; t0 < Q; t0 += gridDim.x * TILE) {
        for (int b = threadIdx.x; b < ORD_BINS; b += ORD_THREADS) lcount[b] = 0;
        __syncthreads();
        int key[ORD_ITEMS];
        unsigned rank[ORD_ITEMS];
#pragma unroll
        for (int j = 0; j < ORD_ITEMS; ++j) {
            const int r = t0 + j * ORD_THREADS + threadIdx.x;
            key[j] = -1;
            if (r < Q) {
                key[j] = max(0, min(ORD_BINS - 1, __ldg(cost + r)));
                rank[j] = atomicAdd(&lcount[key[j]], 1u);
            }
        }
        __syncthreads();
        for (int b = threadIdx.x; b < ORD_BINS; b += ORD_THREADS) {
            const unsigned n = lcount[b];
            if (n) lcount[b] = base[b] + atomicAdd(cursor + b, n);      // first position of the tile's rays of bin b
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ORD_ITEMS; ++j)
            if (key[j] >= 0) order[lcount[key[j]] + rank[j]] = t0 + j * ORD_THREADS + threadIdx.x;
        __syncthreads();
    }
}

int scratch_alloc(void** p, size_t bytes, cudaStream_t st);   // svoxb_tree.cu: stream-ordered pool

// Batches from 0.75 rays per resident backward lane upwards are marched longest first; shorter ones end before the two
// extra kernels pay off. SVOXB_ORDER_MAX_RAYS sets an upper bound (0 disables the ordering).
int64_t ray_order_max_rays() {
    static const long long forced = getenv("SVOXB_ORDER_MAX_RAYS") ? atoll(getenv("SVOXB_ORDER_MAX_RAYS")) : -1;
    return forced >= 0 ? forced : ((long long)1 << 31) - 1;
}
int64_t ray_order_min_rays() { return (long long)sm_count() * 24 * 32 * 3 / 4; }
bool want_ray_order(const TreeArgs& tr, int64_t Q) {
    return tr.use_accel && Q >= ray_order_min_rays() && Q <= ray_order_max_rays();
}

// Builds the permutation in stream-ordered scratch memory; the caller releases *order with cudaFreeAsync on `st`
// after the march that reads it has been launched. `cost` (device, [Q] int32): the per-ray costs to order by.
int build_ray_order(const int* cost, int64_t Q, int** order, cudaStream_t st) {
    *order = nullptr;
    const size_t order_bytes = (sizeof(int) * (size_t)Q + 15) / 16 * 16;
    char* mem = nullptr;
    int rc = scratch_alloc((void**)&mem, order_bytes + 2 * sizeof(unsigned) * ORD_BINS, st);
    if (rc) return rc;
    int* ord = reinterpret_cast<int*>(mem);
    unsigned* hist = reinterpret_cast<unsigned*>(mem + order_bytes);
    cudaError_t e = cudaMemsetAsync(hist, 0, 2 * sizeof(unsigned) * ORD_BINS, st);
    if (e == cudaSuccess) {
        const int grid = (int)min((Q + ORD_THREADS - 1) / ORD_THREADS, (int64_t)sm_count() * 8);
        ray_hist_kernel<<<grid, ORD_THREADS, 0, st>>>(cost, (int)Q, hist);
        const int sgrid = (int)min((Q + ORD_THREADS * ORD_ITEMS - 1) / (ORD_THREADS * ORD_ITEMS), (int64_t)sm_count() * 8);
        ray_order_scatter_kernel<<<sgrid, ORD_THREADS, 0, st>>>(cost, (int)Q, hist, hist + ORD_BINS, ord);
        count_launch(2);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        cudaFreeAsync(mem, st);
        return check_cuda(e, "ray order kernels");
    }
    *order = ord;
    return 0;
}

}  // namespace svoxb
