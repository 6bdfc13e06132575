// svoxb_order.cu -- longest-first ray order for the BACKWARD over short explicit ray batches (no reference counterpart:
// the reference runs one thread per ray, rt_kernel.cu:654-671, and has no queue to order).
//
// The march kernels are persistent: warps pull 32-ray entries from a global queue and refill a lane as soon as its ray
// ends. With millions of rays the lanes stay busy; with about one ray per resident lane (a strong-scaling shard of the
// 2^20-ray training batch: 128 k rays on 148 SMs x 24 warps x 32 lanes) every warp marches until the LONGEST of its 32
// rays ends while the other lanes idle. Handing the rays out longest first makes the rays of a warp alike (lanes end
// together) and leaves the short rays for last (they fill the tail): classic longest-processing-time scheduling.
// Per-ray results do not depend on which lane serves a ray.
//
// The cost of a ray is its EXACT number of march iterations, which the forward over the same batch writes as a
// by-product (RaySource::steps_out; svoxb_render_rays_fwd_cost -> svoxb_render_rays_bwd_cost). Two small kernels:
// histogram, then a counting-sort scatter (descending cost; the order within a cost bin is whatever the atomics
// produce). Measured on B200, C3 tree (profiles/NOTES_r02.md): backward of 128 k rays 1.10 -> 0.79 ms, of 256 k rays
// 1.88 -> 1.43 ms; at 64 k rays and below the two extra kernels cost more than the order gains (0.52 -> 0.56 ms), so
// the order is only built for batches of 0.75 .. 3 rays per resident lane. The FORWARD is never ordered: an exact
// order gains nothing there (128 k rays: 0.45 ms unordered, 0.49 ms ordered by exact counts -- every ray has its own
// lane and the kernel ends with the longest ray either way), and an estimate costs a kernel of its own (a top-grid
// march per ray, 0.15 ms) on top.
#include <stdlib.h>
#include "svoxb_march.cuh"

namespace svoxb {

constexpr int ORD_BINS = 1024;
constexpr int ORD_THREADS = 256;

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
    for (int r = blockIdx.x * ORD_THREADS + threadIdx.x; r < Q; r += gridDim.x * ORD_THREADS) {
        const int key = max(0, min(ORD_BINS - 1, __ldg(cost + r)));
        order[base[key] + atomicAdd(cursor + key, 1u)] = r;
    }
}

int scratch_alloc(void** p, size_t bytes, cudaStream_t st);   // svoxb_tree.cu: stream-ordered pool

// Batches of 0.75 .. 3 rays per resident lane are handed out longest first; longer ones keep every lane busy anyway,
// shorter ones end before the two extra kernels pay off. SVOXB_ORDER_MAX_RAYS overrides the upper bound (0 disables
// the ordering).
int64_t ray_order_max_rays() {
    static const long long forced = getenv("SVOXB_ORDER_MAX_RAYS") ? atoll(getenv("SVOXB_ORDER_MAX_RAYS")) : -1;
    return forced >= 0 ? forced : (long long)sm_count() * 24 * 32 * 3;
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
        ray_order_scatter_kernel<<<grid, ORD_THREADS, 0, st>>>(cost, (int)Q, hist, hist + ORD_BINS, ord);
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
