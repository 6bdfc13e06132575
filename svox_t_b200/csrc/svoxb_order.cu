// svoxb_order.cu -- longest-first ray order for SHORT explicit ray batches (no reference counterpart: the reference
// runs one thread per ray, rt_kernel.cu:654-671, and has no queue to order).
//
// The march kernels are persistent: warps pull 32-ray entries from a global queue and refill a lane as soon as its ray
// ends. With millions of rays the lanes stay busy; with about one ray per resident lane (a strong-scaling shard of the
// 2^20-ray training batch: 128 k rays on 148 SMs x 24 warps x 32 lanes) every warp marches until the LONGEST of its 32
// rays ends while the other lanes idle -- measured 0.73 of the large-batch per-ray rate. Handing the rays out longest
// first makes the rays of a warp alike (lanes end together) and leaves the short rays for last (they fill the tail):
// classic longest-processing-time scheduling. Per-ray results do not depend on which lane serves a ray.
//
// Cost estimate per ray: a march through the accelerator's TOP GRID only (8^bits[0] cells, <= 16 KB, L1-resident):
// a leaf cell costs one sample, a refined cell (pointer to a brick) costs the number of finest-level cells the chord
// crosses, n = chord * 2^lmax * (|dx| + |dy| + |dz|), less the cells the march's step_size skips: n / (1 + step * n /
// chord). Two small kernels: cost + histogram, then a counting-sort scatter (descending cost). The order within a cost
// bin is whatever the atomics produce. On the C3 scene the estimate orders well enough to cut the idle lane-iterations
// of a 128 k-ray batch from 33 % to 15 %; the EXACT iteration counts, which the forward march can write as a by-product
// (RaySource::steps_out), cut them to under 1 % -- that is what the backward over the same batch is ordered by when
// the caller passes the forward's counts on (svoxb_render_rays_fwd_cost / _bwd_cost).
#include <stdlib.h>
#include "svoxb_march.cuh"

namespace svoxb {

constexpr int ORD_BINS = 1024;
constexpr int ORD_THREADS = 256;

__global__ void __launch_bounds__(ORD_THREADS)
ray_cost_kernel(TreeArgs tr, const float* __restrict__ origins, const float* __restrict__ dirs, int Q, float step,
                int* __restrict__ cost_out, unsigned* __restrict__ hist) {
    __shared__ unsigned h[ORD_BINS];
    for (int i = threadIdx.x; i < ORD_BINS; i += ORD_THREADS) h[i] = 0;
    __syncthreads();
    const AccelView& a = tr.acc;
    const uint32_t* __restrict__ top = a.cells[0];
    const int b0 = a.bits[0];
    const float s0 = __int_as_float((127 + b0) << 23), inv0 = __int_as_float((127 - b0) << 23);   // 2^b0, 2^-b0
    const float fine = __int_as_float((127 + a.lmax) << 23);                                        // 2^lmax
    for (int r = blockIdx.x * ORD_THREADS + threadIdx.x; r < Q; r += gridDim.x * ORD_THREADS) {
        Ray ray;
        const float* o = origins + (int64_t)r * 3;
        const float* d = dirs + (int64_t)r * 3;
        ray_setup(tr.offset, tr.scaling, __ldg(o), __ldg(o + 1), __ldg(o + 2), __ldg(d), __ldg(d + 1), __ldg(d + 2), ray);
        const float l1 = fabsf(ray.dx) + fabsf(ray.dy) + fabsf(ray.dz);
        float cost = 0.0f, t = ray.t;
        for (int it = 0; it < 4 * (1 << b0) && t < ray.tmax; ++it) {
            const float px = clamp01(fmaf(t, ray.dx, ray.ox)), py = clamp01(fmaf(t, ray.dy, ray.oy)),
                        pz = clamp01(fmaf(t, ray.dz, ray.oz));
            const float qx = px * s0, qy = py * s0, qz = pz * s0;
            const float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
            const uint32_t cell = __ldg(top + ((((int)fx << b0) | (int)fy) << b0 | (int)fz));
            float smin, smax;
            dda_unit(qx - fx, qy - fy, qz - fz, ray.ix, ray.iy, ray.iz, smin, smax);
            const float chord = (smax - smin) * inv0;
            const float cells = chord * fine * l1;            // finest-level cells crossed; the march skips step per sample
            cost += (cell & ACC_PTR) ? fmaxf(1.0f, cells / (1.0f + step * fine * l1)) : 1.0f;
            t += chord + step;
        }
        cost_out[r] = (int)cost;
        atomicAdd(&h[min(ORD_BINS - 1, (int)cost)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ORD_BINS; i += ORD_THREADS)
        if (h[i]) atomicAdd(hist + i, h[i]);
}

// Histogram of per-ray costs somebody else produced (the forward's exact iteration counts).
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
    for (int r = blockIdx.x * ORD_THREADS + threadIdx.x; r < Q; r += gridDim.x * ORD_THREADS) {
        const int key = max(0, min(ORD_BINS - 1, __ldg(cost + r)));
        order[base[key] + atomicAdd(cursor + key, 1u)] = r;
    }
}

int scratch_alloc(void** p, size_t bytes, cudaStream_t st);   // svoxb_tree.cu: stream-ordered pool

// Batches this short (rays per resident lane) are handed out longest first; longer ones keep every lane busy anyway
// and would only pay for the two extra kernels. SVOXB_ORDER_MAX_RAYS overrides the bound (0 disables the ordering).
int64_t ray_order_max_rays();
bool want_ray_order(const TreeArgs& tr, int64_t Q) { return tr.use_accel && Q >= 2048 && Q <= ray_order_max_rays(); }

int64_t ray_order_max_rays() {
    static const long long forced = getenv("SVOXB_ORDER_MAX_RAYS") ? atoll(getenv("SVOXB_ORDER_MAX_RAYS")) : -1;
    return forced >= 0 ? forced : (long long)sm_count() * 24 * 32 * 3;
}

// Builds the permutation in stream-ordered scratch memory; the caller releases *order with cudaFreeAsync on `st`
// after the march that reads it has been launched. `cost` (device, [Q] int32): with `cost_is_input` the per-ray costs
// to order by (the forward's exact counts); otherwise optional -- the top-grid estimate is written there (scratch if
// NULL) and then ordered by.
int build_ray_order(const TreeArgs& tr, const float* origins, const float* dirs, int64_t Q, float step, int* cost,
                    bool cost_is_input, int** order, cudaStream_t st) {
    *order = nullptr;
    const size_t order_bytes = sizeof(int) * (size_t)Q, cost_bytes = cost ? 0 : (sizeof(int) * (size_t)Q + 15) / 16 * 16;
    char* mem = nullptr;
    int rc = scratch_alloc((void**)&mem, order_bytes + cost_bytes + 2 * sizeof(unsigned) * ORD_BINS, st);
    if (rc) return rc;
    int* ord = reinterpret_cast<int*>(mem);
    if (!cost) cost = reinterpret_cast<int*>(mem + order_bytes);
    unsigned* hist = reinterpret_cast<unsigned*>(mem + order_bytes + cost_bytes);
    cudaError_t e = cudaMemsetAsync(hist, 0, 2 * sizeof(unsigned) * ORD_BINS, st);
    if (e == cudaSuccess) {
        const int grid = (int)min((Q + ORD_THREADS - 1) / ORD_THREADS, (int64_t)sm_count() * 8);
        if (cost_is_input) ray_hist_kernel<<<grid, ORD_THREADS, 0, st>>>(cost, (int)Q, hist);
        else ray_cost_kernel<<<grid, ORD_THREADS, 0, st>>>(tr, origins, dirs, (int)Q, step, cost, hist);
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
