// svoxb_order.cu -- longest-first ray order for the BACKWARD over short explicit ray batches (no reference counterpart:
// the reference runs one thread per ray, rt_kernel.cu:654-671, and has no queue to order).
//
// The march kernels are persistent: warps pull 32-ray entries from a global queue and refill a lane as soon as its ray
// ends. With millions of rays the lanes stay busy; with about one ray per resident lane (a strong-scaling shard of the
// 2^20-ray training batch: 131 072 rays on 148 SMs x 24 warps x 32 lanes = 113 664 backward lanes) the rays that do
// not fit the first wave start whenever a lane frees up, and the kernel ends with whichever of them happens to be
// long: 1.10 ms against 0.76 ms at the full batch's per-ray rate. Handing the rays out longest first leaves the short
// rays for the second wave (classic longest-processing-time scheduling): 0.79 ms. Per-ray results do not depend on
// which lane serves a ray.
//
// Where the order comes from: the FORWARD over the same batch (svoxb_render_rays_fwd_cost) runs 28 warps per SM --
// 132 608 lanes, every ray of such a batch starts at once -- and appends each ray's index to a list at the moment the
// ray ends (one warp-aggregated atomicAdd per group of rays ending together). The list is therefore sorted by march
// length, shortest first, and the backward (svoxb_render_rays_bwd_cost) simply reads it back to front. No sort, no
// extra kernel, and nothing carried through the forward's loop (an iteration counter per lane pushed its 72-register
// instantiation into spilling: 0.45 -> 0.53 ms). The forward itself is never ordered: an exact order gains nothing
// there (every ray has its own lane and the kernel ends with the longest ray either way).
// Batches of more rays than the forward has lanes are marched in K launches of equal consecutive ranges (each keeps its
// own list, every ray of a launch starts at once); the backward interleaves the K lists from their tails, which is
// close to the globally sorted order (256 k rays, K = 2: backward 1.73 -> see profiles/NOTES_r02.md).
// A forward that cannot keep the list (view-dependent formats, fused depth, odd widths) leaves the identity order.
#include <stdlib.h>
#include "svoxb_march.cuh"

namespace svoxb {

// Batches of 0.75 .. 3 rays per resident backward lane are handed out longest first; longer ones keep every lane busy
// anyway, shorter ones end before the order pays off. SVOXB_ORDER_MAX_RAYS overrides the upper bound (0 disables it).
int64_t ray_order_max_rays() {
    static const long long forced = getenv("SVOXB_ORDER_MAX_RAYS") ? atoll(getenv("SVOXB_ORDER_MAX_RAYS")) : -1;
    return forced >= 0 ? forced : (long long)sm_count() * 24 * 32 * 3;
}
int64_t ray_order_min_rays() { return (long long)sm_count() * 24 * 32 * 3 / 4; }
bool want_ray_order(const TreeArgs& tr, int64_t Q) {
    return tr.use_accel && Q >= ray_order_min_rays() && Q <= ray_order_max_rays();
}

// Per range of the balanced K-way split: list[a + i] = len-1-i. Read back to front (and interleaved) this is close to
// the caller's own order -- what the backward sees when the forward that ran could not keep completion lists (a
// list-keeping forward overwrites every entry: each ray ends exactly once).
__global__ void __launch_bounds__(256) reverse_identity_kernel(int* __restrict__ list, int Q, int K) {
    const int base = Q / K, rem = Q % K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Q; i += gridDim.x * blockDim.x) {
        int k = rem ? min(i / (base + 1), rem) : 0;                 // ranges 0 .. rem-1 hold base+1 entries
        if (k == rem) k = rem + (i - rem * (base + 1)) / max(base, 1);
        const int a = k * base + min(k, rem), len = base + (k < rem ? 1 : 0);
        list[i] = len - 1 - (i - a);
    }
}

int fill_reverse_identity(int* list, int64_t Q, int K, cudaStream_t st) {
    const int grid = (int)min((Q + 255) / 256, (int64_t)sm_count() * 8);
    reverse_identity_kernel<<<grid, 256, 0, st>>>(list, (int)Q, K);
    count_launch();
    return check_cuda(cudaGetLastError(), "reverse_identity_kernel launch");
}

}  // namespace svoxb
