// svoxb_render_q.cu -- "quad-lane" march kernels: the fast path for feature widths D % 4 == 0, 4 <= D <= 128.
//
// Same semantics as the scalar-lane kernels in svoxb_render.cu (reference: rt_kernel.cu:221-328 forward,
// 330-496 backward, 781-834 depth); what changes is how a warp touches the feature table:
//   * a feature row is covered by LPR = pow2ceil(D/4) lanes holding one float4 each, so ONE load instruction fetches
//     RPI = 32/LPR whole rows (4 rows = 512 B at D = 32) and one red.global.add.v4.f32 per lane scatters RPI gradient
//     rows (256-bit row blocks, V4 = 2, exist behind SVOXB_WIDE_ROWS: fewer instructions but measured slower);
//   * the lanes' traversal never reads the feature table: every leaf that holds a row becomes a candidate, the rows
//     of the candidates are requested as a batch and each owner lane picks its sample's sigma out of the loaded row
//     with one shuffle (the reference's separate 4-byte sigma gather and second dependent round trip are gone); rows the
//     accelerator's hit marks flag as sigma <= 0 never become candidates;
//   * software pipelining: the brick lookup of sample i+1 is issued before, and consumed after, the compositing of
//     sample i, whose rows were requested one step earlier -- both latencies hide behind arithmetic; for widths that
//     take several batches per iteration (D > 32) the rows of batch b+1 are requested as soon as batch b has left the
//     registers, ahead of the wait for the brick word;
//   * the 32 x D partial outputs of a warp's rays (forward) / the staged grad_out rows (backward) live in shared
//     memory: ONE CTA per SM, 28 warps at 72 registers (forward, D <= 32) or 24 warps at 80 (backward, depth
//     variants); the shared-memory carve-out is set to what the CTA needs so that the rest of the 228 KB stays L1;
//   * the per-hit channel dot product of the backward is reduced with a transposing butterfly over the LPR lanes
//     of a row (LPR-1 shuffles for LPR hits instead of log2(LPR) per hit).
// Lane layout: lane = q * LPR + c; q = which of the RPI rows of a load, c = channel block (channels VEC*c ..
// VEC*c+VEC-1, VEC = 4*V4). Ray r of the warp (owner lane r) is served as row q = r % RPI of group j = r / RPI.
#include <stdlib.h>
#include "svoxb_march.cuh"

namespace svoxb {

__device__ __forceinline__ float4 sigmoid4(const float4 x) {
    return make_float4(fast_sigmoid(x.x), fast_sigmoid(x.y), fast_sigmoid(x.z), fast_sigmoid(x.w));
}
// `act` (warp-uniform): the rows come from the pre-activated table, the sigmoid has been applied already.
__device__ __forceinline__ float4 activated(const float4 x, bool act) { return act ? x : sigmoid4(x); }

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// L2 eviction-priority hints for the backward (measured on B200, C3: 7.21 -> 6.61 ms). The gradient table is
// read-modify-written by the reductions, so an L2 miss on it costs a DRAM read AND a write-back, a miss on the
// feature table one read: the reductions ask L2 to keep their lines (evict_last), the row gathers to drop theirs first
// (evict_first). Each hint alone gains 0.2 ms, both together 0.6 ms; the fraction of lines marked evict_last
// (0.25 / 0.5 / 1.0) makes no difference.
#ifndef SVOXB_BWD_HINTS
#define SVOXB_BWD_HINTS 1
#endif
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void red_add_v4_hint(float* p, float a, float b, float c, float d, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(a), "f"(b), "f"(c),
                 "f"(d), "l"(pol) : "memory");
}
// Camera rays, forward: a warp takes 32 new pixels only when ALL its lanes have finished (the rays of a pixel tile stay in
// step, so the empty-space step below is taken by whole warps); 0 = refill lane by lane like the explicit ray batches.
#ifndef SVOXB_TILE_SYNC
#define SVOXB_TILE_SYNC 1
#endif
#ifndef SVOXB_TILE_SYNC_BWD
#define SVOXB_TILE_SYNC_BWD 1
#endif
#ifndef SVOXB_FWD_STREAM_ROWS
#define SVOXB_FWD_STREAM_ROWS 1
#endif
// Empty-space step of the forward (see the kernel): 0 disables it.
#ifndef SVOXB_EMPTY_STEP
#define SVOXB_EMPTY_STEP 1
#endif
#ifndef SVOXB_ROWS_NO_L1
#define SVOXB_ROWS_NO_L1 0
#endif
#if SVOXB_ROWS_NO_L1
#define SVOXB_L1_HINT ".L1::no_allocate"
#else
#define SVOXB_L1_HINT ""
#endif
__device__ __forceinline__ void red_add_f32_hint(float* p, float a, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(a), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ldg_hint(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc" SVOXB_L1_HINT ".L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
// Row gather of the forward: the rows stream through (a row is not touched twice by the same SM within its L1
// lifetime), so they are kept out of L1, which then holds the brick sectors consecutive samples of a ray re-read.
__device__ __forceinline__ float4 ldg_row(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc" SVOXB_L1_HINT ".v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// NB per-lane partials -> totals over the LPR consecutive lanes of a row group. On return the lane holds the
// total of value index (lane % NB). Transposing steps (each halves the live values), then plain butterfly steps.
template <int NB, int LPR>
__device__ __forceinline__ float quad_reduce(float (&v)[NB], int lane) {
    static_assert(NB == 1 || NB == 2 || NB == 4 || NB == 8, "NB");
    if constexpr (NB == 8) {
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float send = up ? v[j] : v[j + 4];
            const float keep = up ? v[j + 4] : v[j];
            v[j] = keep + __shfl_xor_sync(FULL, send, 4);
        }
    }
    if constexpr (NB >= 4) {
        const bool up = lane & 2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float send = up ? v[j] : v[j + 2];
            const float keep = up ? v[j + 2] : v[j];
            v[j] = keep + __shfl_xor_sync(FULL, send, 2);
        }
    }
    if constexpr (NB >= 2) {
        const bool up = lane & 1;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULL, send, 1);
    }
    float r = v[0];
#pragma unroll
    for (int w = NB; w < LPR; w <<= 1) r += __shfl_xor_sync(FULL, r, w);
    return r;
}

template <int BITS>
__device__ __forceinline__ constexpr unsigned low_mask() { return BITS >= 32 ? 0xffffffffu : ((1u << (BITS & 31)) - 1u); }

template <int LPR, int V4>
struct Quad {
    static constexpr int VEC = 4 * V4;                   // channels per lane
    static constexpr int RPI = 32 / LPR;                 // rows per load instruction
    static constexpr int NB = (LPR * V4 <= 8) ? LPR : 8 / V4;   // row groups in flight per batch (4*V4*NB registers)
    static constexpr int NBATCH = LPR / NB;
    static constexpr int RAYS_PER_BATCH = RPI * NB;
    static constexpr int DP = VEC * LPR;                 // padded row width
    // ONE CTA per SM (a single copy of the top grid in shared memory); its size is what registers (80 per thread)
    // and the per-warp rows in shared memory allow
#ifndef SVOXB_THREADS32
#define SVOXB_THREADS32 768
#endif
#ifndef SVOXB_THREADS64
#define SVOXB_THREADS64 640
#endif
#ifndef SVOXB_THREADS128
#define SVOXB_THREADS128 384
#endif
    static constexpr int THREADS = DP <= 32 ? SVOXB_THREADS32 : (DP <= 64 ? SVOXB_THREADS64 : SVOXB_THREADS128);
    static constexpr int NWARPS = THREADS / 32;
    // The forward fits 72 registers without spills, so it takes 28 warps (7 per scheduler): 2.97 -> 2.81 ms. 26 and 30
    // warps (uneven per scheduler) and 32 (64 registers, spills, smaller L1) are slower; the backward spills at 72.
#ifndef SVOXB_FWD_THREADS32
#define SVOXB_FWD_THREADS32 896
#endif
    static constexpr int FWD_THREADS = DP <= 32 ? SVOXB_FWD_THREADS32 : THREADS;
};

// V4 float4 of one row block. 256-bit form: LDG.E.256 (sm_100+), which carries its L2 eviction priority inline.
template <int V4>
struct RowBlk {
    float4 v[V4];
};

template <int V4, bool EVICT_FIRST>
__device__ __forceinline__ RowBlk<V4> load_row_block(const char* p, uint64_t pol_first) {
    RowBlk<V4> r;
    if constexpr (V4 == 2) {
        if constexpr (EVICT_FIRST)
            asm volatile("ld.global.nc.L2::evict_first.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=f"(r.v[0].x), "=f"(r.v[0].y), "=f"(r.v[0].z), "=f"(r.v[0].w), "=f"(r.v[1].x), "=f"(r.v[1].y),
                           "=f"(r.v[1].z), "=f"(r.v[1].w) : "l"(p));
        else
            asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=f"(r.v[0].x), "=f"(r.v[0].y), "=f"(r.v[0].z), "=f"(r.v[0].w), "=f"(r.v[1].x), "=f"(r.v[1].y),
                           "=f"(r.v[1].z), "=f"(r.v[1].w) : "l"(p));
    } else {
        if constexpr (EVICT_FIRST) r.v[0] = ldg_hint(reinterpret_cast<const float4*>(p), pol_first);
        else r.v[0] = ldg_row(reinterpret_cast<const float4*>(p));
    }
    return r;
}

// ------------------------------------------------------------------------------------------------------------
// Loop structure (NBATCH = 1 for D <= 32, 2 for D <= 64, 4 for D <= 128; x holds the rows of ONE batch):
//   S0   request the rows of batch 0 of the pending candidates (found in the PREVIOUS iteration)
//   S1   probe_begin : next sample position, top-grid lookup in shared memory, brick lookup ISSUED
//   S2.0 composite batch 0 (its rows have been in flight since S0)
//   S2.0'request the rows of batch 1 (the registers of batch 0 are free)
//   S3   probe_end   : brick word consumed -> new candidates, t advanced   (the rows of batch 1 are in flight)
//   S2.b composite batch b, request the rows of batch b+1  (b = 1 .. NBATCH-1)
//   S4   flush finished rays, refill; the rows of batch 0 of the new candidates are requested first thing in the
//        next iteration (S0), ahead of S1
// AL (aligned): D % VEC == 0, rows are read in the caller's [M, D] layout (raw or activated). !AL: any other D -- the
// rows come from the PADDED activated table, which holds the D-1 payload channels only (stride = D-1 rounded up to a
// multiple of 4 floats: 16-byte aligned rows, D = 33 is served like D = 32), sigma comes from the compact sigma[M]
// array next to it (one 4-byte gather per candidate, L2-resident), and the output rows / gradient rows of the
// caller's unaligned [M, D] layout are written channel by channel.
// COUNT: also write the number of march iterations of every ray (RaySource::steps_out) -- the exact per-ray cost the
// backward over the same batch is ordered by (svoxb_order.cu). One counter register per lane; the store is a real call
// on purpose: inlined, those few instructions perturb the register allocation of the whole 72-register loop into
// spilling (16-24 bytes, reloaded every iteration: 0.45 -> 0.53 ms at 131 072 rays); as a call the live registers are
// saved around this cold site only and the kernel keeps 72 registers, no stack, 28 warps.
__device__ __noinline__ void store_steps(int* __restrict__ steps_out, bool mine, int row, int steps) {
    if (mine) steps_out[row] = steps;
}

template <int LPR, int V4, bool ACCEL, bool IMAGE, bool DEPTH, bool AL, bool COUNT = false>
__global__ void __launch_bounds__(((DEPTH || !AL) ? Quad<LPR, V4>::THREADS : Quad<LPR, V4>::FWD_THREADS), 1)
march_fwd_quad_kernel(TreeArgs tr, RaySource src, MarchOpts opt, float* __restrict__ out, float* __restrict__ depth,
                      unsigned long long* counter) {
    using G = Quad<LPR, V4>;
    constexpr int RPI = G::RPI, NB = G::NB, NBATCH = G::NBATCH, RPB = G::RAYS_PER_BATCH, VEC = G::VEC;
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    const int top_words = ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31;
    // this warp's 32 x DP partial outputs: accs[(j * V4 + h) * 32] = float4 h of this lane's block of ray RPI*j + q
    float4* accs = reinterpret_cast<float4*>(smem_u32 + top_words) + (size_t)(threadIdx.x >> 5) * 32 * LPR * V4 + lane;
    static_assert(AL || V4 == 1, "padded rows use 128-bit blocks");
    const int q = lane / LPR, c = lane % LPR;
    const int D = tr.D, DV = AL ? D / VEC : (D - 1 + VEC - 1) / VEC;
    const bool lane_ok = c < DV, is_sig = AL && c == DV - 1;
    const int sig_src = (lane % RPI) * LPR + (DV - 1);   // lane that holds sigma of this owner lane's row
    const bool act = AL ? tr.feat_act != nullptr : true;
    const char* fbase = reinterpret_cast<const char*>(act ? tr.feat_act : tr.features) + 4 * VEC * min(c, DV - 1);
    const unsigned row_bytes = AL ? (unsigned)D * 4u : (unsigned)tr.act_stride * 4u;
    const float* off = tr.offset;
    const float* scl = tr.scaling;
    // Wide rows of random ray batches (D > 32: the tables of such scenes are many times the L2 and a row is not touched
    // twice while it is resident): the row gathers ask L2 to drop their lines first, which keeps the accelerator's bricks
    // resident. C5 scene, 2^20 random rays: 5.58 -> 5.44 ms. Camera rays re-use rows between neighbouring pixels: not hinted.
    constexpr bool FWD_STREAM_ROWS = SVOXB_FWD_STREAM_ROWS && LPR >= 16 && !IMAGE;
    [[maybe_unused]] uint64_t pol_first = 0;
    if constexpr (FWD_STREAM_ROWS) pol_first = policy_evict_first();

    RowBlk<V4> x[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int h = 0; h < V4; ++h) x[j].v[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < LPR * V4; ++j) accs[j * 32] = make_float4(0.f, 0.f, 0.f, 0.f);

    Ray ray;
    float T = 1.0f, p_dt = 0.0f, p_t = 0.0f;
    int row = 0, p_idx = -1;
    bool active = false, got_depth = false, trav_done = true;
    [[maybe_unused]] int steps = 0;
    Queue qu{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (SVOXB_TILE_SYNC && IMAGE ? need == FULL : need != 0u) {
            const unsigned got = refill<IMAGE>(src, off, scl, counter, qu, need, lane, ray, row);
            if ((got >> lane) & 1u) {
                active = true; trav_done = false; T = 1.0f; got_depth = false;
                if constexpr (COUNT) steps = 0;
                if (DEPTH) depth[row] = 0.0f;           // overwritten at the first hit, if any
            }
            need = 0;
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        const unsigned pm = __ballot_sync(FULL, p_idx >= 0);
        bool stopped = false;
        int n_idx = -1;
        float n_dt = 0.0f, n_t = 0.0f;
        if (SVOXB_EMPTY_STEP && IMAGE && pm == 0u) {
            // ---- empty space: no lane of the warp holds a pending candidate (the coherent rays of a pixel tile in front
            // of / behind the object, sparse scenes): nothing to fetch and nothing to composite, the lanes only traverse.
            // Camera rays only -- measured on B200: C5 1080p view 3.96 -> 3.82 ms, C4 1080p render 1.88 -> 1.85 ms; random
            // ray batches never have 32 idle lanes at once and only pay for the test (C3 forward 2.87 -> 2.93 ms).
            if (active && !trav_done) {
                if (!(ray.t < ray.tmax)) trav_done = true;
                else {
                    Probe pb;
                    probe_begin<ACCEL>(tr, top, ray, pb);
                    if (DEPTH) n_t = ray.t;
                    probe_end<ACCEL>(tr, pb, ray, opt.step, n_idx, n_dt);
                    ray.t += n_dt;
                    if constexpr (COUNT) ++steps;
                    if (!(ray.t < ray.tmax)) trav_done = true;
                }
            }
        } else {
        // ---- S0: request the rows of batch 0 of the pending candidates. First thing in the iteration (a load left
        // in flight across the loop back-edge is waited for at the loop header); unconditional on purpose (a guarded
        // load turns x into a phi whose register copies stall on the data); row 0 stands in for "no candidate".
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
            const int idx = max(__shfl_sync(FULL, p_idx, RPI * jj + q), 0);
            SVOXB_DBG((int64_t)idx < max(tr.M, (int64_t)1));
            x[jj] = load_row_block<V4, FWD_STREAM_ROWS>(fbase + (size_t)(unsigned)idx * row_bytes, pol_first);
        }
        float sig_own = 0.0f;                                   // !AL: sigma of this lane's pending candidate
        if constexpr (!AL) sig_own = __ldg(tr.sigma_c + max(p_idx, 0));

        // ---- S1 -------------------------------------------------------------------------------------------------
        bool trav = active && !trav_done;
        Probe pb;
        if (trav) {
            if (!(ray.t < ray.tmax)) { trav_done = true; trav = false; }
            else probe_begin<ACCEL>(tr, top, ray, pb);
        }

        if (trav) probe_mid<ACCEL>(tr, pb);

#pragma unroll
        for (int b = 0; b < NBATCH; ++b) {
            // ---- S2.b: composite ---------------------------------------------------------------------------------
            const unsigned bm = NBATCH == 1 ? pm : (pm >> ((RPB * b) & 31)) & low_mask<RPB>();
            if (bm) {
                float sig = sig_own;
                if constexpr (AL) {
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        const float v = __shfl_sync(FULL, x[jj].v[V4 - 1].w, sig_src);
                        if (lane / RPI == b * NB + jj) sig = v;
                    }
                }
                float w = 0.0f;
                if ((NBATCH == 1 || lane / RPB == b) && p_idx >= 0 && sig > opt.sigma_thresh) {
                    const float att = expf(-p_dt * ray.ds * sig);                     // rt_kernel.cu:280
                    w = T * (1.0f - att);
                    if (DEPTH && !got_depth) { depth[row] = ray.ds * p_t; got_depth = true; }   // rt_kernel.cu:826-830
                    T *= att;
                    if (T <= opt.stop_thresh) stopped = true;                         // rt_kernel.cu:313
                }
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) {
                    const int j = b * NB + jj;
                    const float w_j = __shfl_sync(FULL, w, RPI * j + q);
                    if (w_j != 0.0f) {
#pragma unroll
                        for (int h = 0; h < V4; ++h) {
                            const float4 s = activated(x[jj].v[h], act);
                            float4 a = accs[(j * V4 + h) * 32];
                            a.x = fmaf(w_j, s.x, a.x);
                            a.y = fmaf(w_j, s.y, a.y);
                            a.z = fmaf(w_j, s.z, a.z);
                            a.w = fmaf(w_j, s.w, a.w);
                            accs[(j * V4 + h) * 32] = a;
                        }
                    }
                }
            }
            // rows of batch b+1 of the pending candidates: requested as soon as batch b has left the registers, so that
            // they are in flight while S3 waits for its brick word (batch 0 was requested at the top of the iteration)
            if (b + 1 < NBATCH) {
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) {
                    const int idx = max(__shfl_sync(FULL, p_idx, RPI * ((b + 1) * NB + jj) + q), 0);
                    SVOXB_DBG((int64_t)idx < max(tr.M, (int64_t)1));
                    x[jj] = load_row_block<V4, FWD_STREAM_ROWS>(fbase + (size_t)(unsigned)idx * row_bytes, pol_first);
                }
            }
            // ---- S3 (once, after the first batch) ------------------------------------------------------------------
            if (b == 0 && trav) {
                if (DEPTH) n_t = ray.t;
                probe_end<ACCEL, 3>(tr, pb, ray, opt.step, n_idx, n_dt);
                ray.t += n_dt;
                if constexpr (COUNT) ++steps;
                if (!(ray.t < ray.tmax)) trav_done = true;
            }
        }
        }   // pm != 0
        if (stopped) { n_idx = -1; trav_done = true; }
        p_idx = n_idx; p_dt = n_dt; p_t = n_t;
        const int fin = (active && trav_done && p_idx < 0) ? (stopped ? 2 : 1) : 0;

        const unsigned fm = __ballot_sync(FULL, fin != 0);
        if (fm) {
#pragma unroll
            for (int j = 0; j < LPR; ++j) {
                const unsigned gm = (fm >> ((RPI * j) & 31)) & low_mask<RPI>();
                if (gm) {
                    const int r = RPI * j + q;
                    const float T_r = __shfl_sync(FULL, T, r);
                    const int fin_r = __shfl_sync(FULL, fin, r);
                    const int row_r = __shfl_sync(FULL, row, r);
                    SVOXB_DBG(fin_r == 0 || (row_r >= 0 && (IMAGE ? row_r < src.width * (src.row_end - src.row_begin) : row_r < src.total)));
                    if (fin_r != 0) {
#pragma unroll
                        for (int h = 0; h < V4; ++h) {
                            float4 v = accs[(j * V4 + h) * 32];
                            if (fin_r == 2) {
                                const float scale = (float)(1.0 / (1.0 - (double)T_r));   // rt_kernel.cu:315
                                v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
                            } else {
                                const float add = T_r * opt.bg;                           // rt_kernel.cu:323-325
                                v.x += add; v.y += add; v.z += add; v.w += add;
                            }
                            if constexpr (AL) {
                                if (is_sig && h == V4 - 1) v.w = 1.0f - T_r;              // rt_kernel.cu:317,326
                                // written once, never re-read by this kernel: streaming store
                                if (lane_ok) __stcs(reinterpret_cast<float4*>(out + (int64_t)row_r * D + VEC * c + 4 * h), v);
                            } else {
                                float* orow = out + (int64_t)row_r * D;
                                const float ve[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (4 * c + e < D - 1) __stcs(orow + 4 * c + e, ve[e]);
                            }
                            accs[(j * V4 + h) * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
            }
            if constexpr (COUNT) store_steps(src.steps_out, fin != 0, row, steps);
            if (fin != 0) {
                if constexpr (!AL) __stcs(out + (int64_t)row * D + (D - 1), 1.0f - T);   // opacity, by the owner lane
                active = false;
            }
            need |= fm;
        }
    }
}

template <int LPR, int V4, bool ACCEL, bool IMAGE, bool AL>
__global__ void __launch_bounds__((Quad<LPR, V4>::THREADS), 1)
march_bwd_quad_kernel(TreeArgs tr, RaySource src, MarchOpts opt, const float* __restrict__ grad_out,
                      const float* __restrict__ saved_out, float* __restrict__ grad, float* __restrict__ grad_sigma,
                      unsigned long long* counter) {
    // AL: `grad` is the caller's [M, D] table. !AL: `grad` is a scratch table in the layout of the payload-only
    // activated table (aligned rows of tr.act_stride floats -> red.v4) and `grad_sigma` a compact [M] array; the
    // launcher folds both into the caller's table afterwards.
    using G = Quad<LPR, V4>;
    constexpr int RPI = G::RPI, NB = G::NB, NBATCH = G::NBATCH, RPB = G::RAYS_PER_BATCH, VEC = G::VEC, DP = G::DP;
    extern __shared__ __align__(128) uint32_t smem_u32[];
    const int top_words = ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0;
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // staged grad_out rows [32 rays][DP]; with 256-bit blocks the two float4 of a block swap places in the rows read
    // by lanes 4..7 of every quarter-warp, which keeps the 128-bit shared loads conflict-free
    float* gs = reinterpret_cast<float*>(smem_u32 + top_words) + (size_t)warp * 32 * DP;
    static_assert(AL || V4 == 1, "padded rows use 128-bit blocks");
    const int q = lane / LPR, c = lane % LPR;
    const int D = tr.D, DV = AL ? D / VEC : (D - 1 + VEC - 1) / VEC;
    const bool lane_ok = c < DV, is_sig = AL && c == DV - 1;
    const int sig_src = (lane % RPI) * LPR + (DV - 1);
    const int red_src = (lane % RPI) * LPR + ((lane / RPI) % NB);     // lane holding this owner's reduced dot product
    const int swz = V4 == 2 ? (lane >> 2) & 1 : 0;
    const bool act = AL ? tr.feat_act != nullptr : true;
    const char* fbase = reinterpret_cast<const char*>(act ? tr.feat_act : tr.features) + 4 * VEC * min(c, DV - 1);
    char* gbase = reinterpret_cast<char*>(grad) + 4 * VEC * c;
    const unsigned row_bytes = AL ? (unsigned)D * 4u : (unsigned)tr.act_stride * 4u;   // feature rows
    const unsigned grow_bytes = row_bytes;                                             // gradient rows: same layout
    const float* off = tr.offset;
    const float* scl = tr.scaling;

    RowBlk<V4> x[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int h = 0; h < V4; ++h) x[j].v[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint64_t pol_last = policy_evict_last(), pol_first = policy_evict_first();
    Ray ray;
    float T = 1.0f, accum = 0.0f, T_end = 0.0f, gop = 0.0f, p_dt = 0.0f;
    int row = 0, p_idx = -1;
    bool active = false, trav_done = true;
    Queue qu{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (SVOXB_TILE_SYNC_BWD && IMAGE ? need == FULL : need != 0u) {      // camera rays: whole tiles, as in the forward
            unsigned got = refill<IMAGE>(src, off, scl, counter, qu, need, lane, ray, row);
            if ((got >> lane) & 1u) { active = true; trav_done = false; T = 1.0f; }
            need = 0;
            while (got) {       // per new ray: stage grad_out row, accum = <g, out>, T_end, g_opacity
                const int r = __ffs(got) - 1;
                got &= got - 1;
                const int row_r = __shfl_sync(FULL, row, r);
                const float* g = grad_out + (int64_t)row_r * D;
                const float* so = saved_out + (int64_t)row_r * D;
                float part = 0.0f, g_last = 0.0f, o_last = 0.0f;
                const int DG = AL ? D : D - 1;                           // staged channels: !AL keeps the payload only
                for (int e = lane; e < DP; e += 32) {
                    const float gv = (e < DG) ? __ldcs(g + e) : 0.0f;     // read once: do not let them displace the tables
                    const float ov = (e < DG) ? __ldcs(so + e) : 0.0f;
                    int slot = e >> 2;                                   // float4 slot inside the row
                    if (V4 == 2) slot ^= ((((r % RPI) * LPR + (e / VEC)) >> 2) & 1);   // reader lane = q*LPR + c
                    gs[r * DP + 4 * slot + (e & 3)] = gv;
                    if (e < D - 1) part = fmaf(gv, ov, part);
                    if (e == D - 1) { g_last = gv; o_last = ov; }
                }
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(FULL, part, s);
                if constexpr (AL) {
                    g_last = __shfl_sync(FULL, g_last, (D - 1) & 31);
                    o_last = __shfl_sync(FULL, o_last, (D - 1) & 31);
                } else {
                    g_last = __ldcs(g + (D - 1));
                    o_last = __ldcs(so + (D - 1));
                }
                if (lane == r) { accum = part; T_end = 1.0f - o_last; gop = g_last; }
            }
            __syncwarp();
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        const unsigned pm = __ballot_sync(FULL, p_idx >= 0);
        int n_idx = -1;
        float n_dt = 0.0f;
        if (SVOXB_TILE_SYNC_BWD && IMAGE && pm == 0u) {
            // ---- empty space (camera rays; see the forward): the lanes only traverse
            if (active && !trav_done) {
                if (!(ray.t < ray.tmax)) trav_done = true;
                else {
                    Probe pb;
                    probe_begin<ACCEL>(tr, top, ray, pb);
                    probe_end<ACCEL>(tr, pb, ray, opt.step, n_idx, n_dt);
                    ray.t += n_dt;
                    if (!(ray.t < ray.tmax)) trav_done = true;
                }
            }
        } else {
        // ---- S0: rows of batch 0 of the pending candidates (see the forward kernel) ---------------------------------
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
            const int idx = max(__shfl_sync(FULL, p_idx, RPI * jj + q), 0);
            SVOXB_DBG((int64_t)idx < max(tr.M, (int64_t)1));
            x[jj] = load_row_block<V4, SVOXB_BWD_HINTS != 0>(fbase + (size_t)(unsigned)idx * row_bytes, pol_first);
        }
        float sig_own = 0.0f;
        if constexpr (!AL) sig_own = __ldg(tr.sigma_c + max(p_idx, 0));

        // ---- S1 -------------------------------------------------------------------------------------------------
        bool trav = active && !trav_done;
        Probe pb;
        if (trav) {
            if (!(ray.t < ray.tmax)) { trav_done = true; trav = false; }
            else probe_begin<ACCEL>(tr, top, ray, pb);
        }

        if (trav) probe_mid<ACCEL>(tr, pb);

#pragma unroll
        for (int b = 0; b < NBATCH; ++b) {
            // ---- S2.b: gradient of batch b of the pending candidates ------------------------------------------------
            const unsigned bm = NBATCH == 1 ? pm : (pm >> ((RPB * b) & 31)) & low_mask<RPB>();
            if (bm) {
                float sig = sig_own;
                if constexpr (AL) {
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        const float v = __shfl_sync(FULL, x[jj].v[V4 - 1].w, sig_src);
                        if (lane / RPI == b * NB + jj) sig = v;
                    }
                }
                float w = 0.0f, dd = 0.0f;
                const bool hit = (NBATCH == 1 || lane / RPB == b) && p_idx >= 0 && sig > 0.0f;   // rt_kernel.cu:382,456
                if (hit) {
                    const float att = expf(-p_dt * sig * ray.ds);
                    w = T * (1.0f - att);
                    dd = p_dt * ray.ds;
                    T *= att;
                }
                const unsigned hb = __ballot_sync(FULL, hit);
                if (hb) {
                    float cp[NB];
                    RowBlk<V4> sv[NB];
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        // Rows of rays without a hit (stand-in row 0) are reduced too: their totals are never read.
                        const int r = RPI * (b * NB + jj) + q;
                        float tot = 0.0f;
#pragma unroll
                        for (int h = 0; h < V4; ++h) {
                            const float4 g4 = *reinterpret_cast<const float4*>(gs + r * DP + 4 * ((c * V4 + h) ^ swz));
                            const float4 s = activated(x[jj].v[h], act);
                            const float sx = s.x * g4.x, sy = s.y * g4.y, sz = s.z * g4.z, sw = s.w * g4.w;
                            tot += (sx + sy) + (sz + ((AL && is_sig && h == V4 - 1) ? 0.0f : sw));   // !AL: g is 0 there
                            sv[jj].v[h] = make_float4(fmaf(-sx, s.x, sx), fmaf(-sy, s.y, sy), fmaf(-sz, s.z, sz),
                                                      fmaf(-sw, s.w, sw));               // s (1 - s) g
                        }
                        cp[jj] = lane_ok ? tot : 0.0f;
                    }
                    const float c_tot = quad_reduce<NB, LPR>(cp, lane);
                    const float c_own = __shfl_sync(FULL, c_tot, red_src);
                    float sgrad = 0.0f;
                    if (hit) {
                        accum -= w * c_own;                                              // rt_kernel.cu:479-480
                        sgrad = dd * (c_own * T - accum) + dd * gop * T_end;             // rt_kernel.cu:486-490
                        if constexpr (!AL) red_add_f32_hint(grad_sigma + p_idx, sgrad, pol_last);   // by the owner lane
                    }
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        const int r = RPI * (b * NB + jj) + q;
                        const float w_j = __shfl_sync(FULL, w, r);
                        const float sg_j = __shfl_sync(FULL, sgrad, r);
                        const int idx_j = __shfl_sync(FULL, p_idx, r);
                        if (((hb >> r) & 1u) && lane_ok) {
                            SVOXB_DBG(idx_j >= 0 && (int64_t)idx_j < tr.M);
                            float* grow = reinterpret_cast<float*>(gbase + (size_t)(unsigned)idx_j * grow_bytes);
                            if constexpr (AL) {
#pragma unroll
                                for (int h = 0; h < V4; ++h) {
                                    const float4 t = sv[jj].v[h];
                                    const float last = (is_sig && h == V4 - 1) ? sg_j : w_j * t.w;
#if SVOXB_BWD_HINTS
                                    red_add_v4_hint(grow + 4 * h, w_j * t.x, w_j * t.y, w_j * t.z, last, pol_last);
#else
                                    red_add_v4(grow + 4 * h, w_j * t.x, w_j * t.y, w_j * t.z, last);
#endif
                                }
                            } else {        // payload-only scratch rows: padding lanes carry zeros (g is 0 there)
                                const float4 t = sv[jj].v[0];
                                red_add_v4_hint(grow, w_j * t.x, w_j * t.y, w_j * t.z, w_j * t.w, pol_last);
                            }
                        }
                    }
                }
            }
            if (b + 1 < NBATCH) {     // rows of batch b+1: in flight while S3 waits for its brick word (see the forward)
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) {
                    const int idx = max(__shfl_sync(FULL, p_idx, RPI * ((b + 1) * NB + jj) + q), 0);
                    SVOXB_DBG((int64_t)idx < max(tr.M, (int64_t)1));
                    x[jj] = load_row_block<V4, SVOXB_BWD_HINTS != 0>(fbase + (size_t)(unsigned)idx * row_bytes, pol_first);
                }
            }
            // ---- S3 (once, after the first batch) ------------------------------------------------------------------
            if (b == 0 && trav) {
                probe_end<ACCEL, 3>(tr, pb, ray, opt.step, n_idx, n_dt);
                ray.t += n_dt;
                if (!(ray.t < ray.tmax)) trav_done = true;
            }
        }
        }   // pm != 0
        p_idx = n_idx; p_dt = n_dt;
        const bool fin = active && trav_done && p_idx < 0;

        const unsigned fm = __ballot_sync(FULL, fin);
        if (fm) {
            if (fin) active = false;
            need |= fm;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
// D % 4 == 0: always. Any other D: when the padded activated table is attached (its rows are 16-byte aligned).
bool quad_supported(const TreeArgs& tr) {
    if (tr.D < 2 || tr.D > 128) return false;
    if (tr.D % 4 == 0) return true;
    return tr.feat_act != nullptr && tr.sigma_c != nullptr && tr.act_stride == (tr.D - 1 + 3) / 4 * 4;
}

// Shared memory this one-CTA-per-SM kernel needs (+1 KB the runtime reserves per CTA), in KB; the rest of the SM's
// 228 KB stays L1, which is what bounds the number of row gathers in flight (measured: forcing the carve-out to the
// maximum slows the forward from 3.0 to 5.2 ms).
static int carveout_kb(size_t smem) {
    int kb = (int)((smem + 1024 + 1023) / 1024);
    if (const char* e = getenv("SVOXB_CARVEOUT_KB")) kb = atoi(e);
    return kb;
}

// Short queues: spread the warps over the SMs (smaller CTAs) instead of filling a few SMs with 24 warps each.
static int threads_for(int max_threads, int64_t queue_len, int chunk, const char* cap_env = nullptr) {
    if (cap_env && getenv(cap_env)) max_threads = min(max_threads, max(64, atoi(getenv(cap_env)) / 32 * 32));   // dev knob
    const int64_t warps_needed = (queue_len + chunk - 1) / chunk;
    const int64_t per_sm = (warps_needed + sm_count() - 1) / sm_count();
    return (int)max((int64_t)64, min((int64_t)max_threads, per_sm * 32));
}

// Queue entries of explicit ray batches: 32 rays (one per lane). Measured on B200 against the image tiles' 64 (C3 tree):
// 16 k rays fwd 0.57 -> 0.32 ms, 64 k rays 0.59 -> 0.36 / bwd 0.88 -> 0.56 ms, 512 k rays bwd 3.42 -> 3.08 ms, 2^20 rays
// 2.79 -> 2.77 / 5.97 -> 5.92 ms: twice the warps for short batches, a finer-grained tail for long ones.
static int chunk_for(const RaySource& src, bool image) {
    static const int forced = getenv("SVOXB_CHUNK") ? atoi(getenv("SVOXB_CHUNK")) : 0;
    (void)src;
    if (image) return CHUNK;
    return (forced >= 1 && forced <= 64) ? forced : RAY_CHUNK;
}


static int pow2ceil(int n) {
    int l = 1;
    while (l < n) l <<= 1;
    return l;
}

template <int LPR, int V4, bool ACCEL, bool IMAGE, bool AL>
static int launch_fwd_q(const TreeArgs& tr, const RaySource& src_in, const MarchOpts& m, float* out, float* depth,
                        cudaStream_t st) {
    using G = Quad<LPR, V4>;
    RaySource src = src_in;
    src.chunk = chunk_for(src, IMAGE);
    bool count = false;
    if constexpr (ACCEL && !IMAGE && AL) count = src.steps_out != nullptr && !depth;
    const int threads = threads_for((depth || !AL) ? G::THREADS : G::FWD_THREADS, src.total, src.chunk, "SVOXB_FWD_THREADS_CAP");
    const size_t smem = (ACCEL ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0) + sizeof(float4) * (threads / 32) * 32 * LPR * V4;
    void (*kern)(TreeArgs, RaySource, MarchOpts, float*, float*, unsigned long long*);
    if (depth) kern = march_fwd_quad_kernel<LPR, V4, ACCEL, IMAGE, true, AL>;
    else kern = march_fwd_quad_kernel<LPR, V4, ACCEL, IMAGE, false, AL>;
    if constexpr (ACCEL && !IMAGE && AL) {
        if (count) kern = march_fwd_quad_kernel<LPR, V4, ACCEL, IMAGE, false, AL, true>;
    }
    if (!count) src.steps_out = nullptr;
    int grid = 0;
    int rc = persistent_grid(kern, smem, src.total, grid, threads, carveout_kb(smem), src.chunk);
    if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, threads, smem, st>>>(tr, src, m, out, depth, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "march_fwd_quad_kernel launch");
}

// grad[M, D] += (payload scratch rows, compact sigma gradients): folds the !AL backward's aligned scratch into the
// caller's table. One thread per element of the caller's table.
__global__ void __launch_bounds__(256)
merge_padded_grad_kernel(const float* __restrict__ gpay, const float* __restrict__ gsig, int64_t M, int D, int S,
                         float* __restrict__ grad) {
    const int64_t n = M * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / D;
        const int c = (int)(i - r * D);
        grad[i] += c < D - 1 ? gpay[r * S + c] : gsig[r];
    }
}

int scratch_alloc(void** p, size_t bytes, cudaStream_t st);   // svoxb_tree.cu: stream-ordered pool

template <int LPR, int V4, bool ACCEL, bool IMAGE, bool AL>
static int launch_bwd_q(const TreeArgs& tr, const RaySource& src_in, const MarchOpts& m, const float* go, const float* so,
                        float* grad, cudaStream_t st) {
    using G = Quad<LPR, V4>;
    RaySource src = src_in;
    src.chunk = chunk_for(src, IMAGE);
    const int threads = threads_for(G::THREADS, src.total, src.chunk, "SVOXB_BWD_THREADS_CAP");
    const size_t smem = (ACCEL ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0) + sizeof(float) * (threads / 32) * 32 * G::DP;
    auto kern = march_bwd_quad_kernel<LPR, V4, ACCEL, IMAGE, AL>;
    int grid = 0;
    int rc = persistent_grid(kern, smem, src.total, grid, threads, carveout_kb(smem), src.chunk);
    if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    if constexpr (AL) {
        kern<<<grid, threads, smem, st>>>(tr, src, m, go, so, grad, nullptr, counter);
        count_launch();
        return check_cuda(cudaGetLastError(), "march_bwd_quad_kernel launch");
    } else {
        // The caller's rows (D floats) are not 16-byte aligned: reduce into an aligned scratch copy (stream-ordered
        // pool memory, released in stream order) with red.v4, then fold it into the caller's table in one pass.
        const size_t pay = sizeof(float) * (size_t)tr.M * tr.act_stride, sig = sizeof(float) * (size_t)tr.M;
        float* scratch = nullptr;
        if ((rc = scratch_alloc((void**)&scratch, pay + sig, st))) return rc;
        float* gsig = scratch + (size_t)tr.M * tr.act_stride;
        cudaError_t e = cudaMemsetAsync(scratch, 0, pay + sig, st);
        if (e == cudaSuccess) {
            kern<<<grid, threads, smem, st>>>(tr, src, m, go, so, scratch, gsig, counter);
            const int mg = (int)min(((int64_t)tr.M * tr.D + 255) / 256, (int64_t)sm_count() * 16);
            merge_padded_grad_kernel<<<mg, 256, 0, st>>>(scratch, gsig, tr.M, tr.D, tr.act_stride, grad);
            count_launch(2);
            e = cudaGetLastError();
        }
        cudaFreeAsync(scratch, st);
        return check_cuda(e, "march_bwd_quad_kernel (padded) launch");
    }
}

#define SVOXB_Q_CASES(FN, L, V, ...)                                                                   \
    case ((L) * 2 + (V)-1) * 4 + 0: return FN<L, V, false, false, true>(__VA_ARGS__);                  \
    case ((L) * 2 + (V)-1) * 4 + 1: return FN<L, V, false, true, true>(__VA_ARGS__);                   \
    case ((L) * 2 + (V)-1) * 4 + 2: return FN<L, V, true, false, true>(__VA_ARGS__);                   \
    case ((L) * 2 + (V)-1) * 4 + 3: return FN<L, V, true, true, true>(__VA_ARGS__);
// padded rows (D % 4 != 0): 128-bit blocks, LPR = pow2ceil(ceil(D / 4))
#define SVOXB_Q_PAD_CASES(FN, L, ...)                                                                  \
    case (L) * 4 + 0: return FN<L, 1, false, false, false>(__VA_ARGS__);                               \
    case (L) * 4 + 1: return FN<L, 1, false, true, false>(__VA_ARGS__);                                \
    case (L) * 4 + 2: return FN<L, 1, true, false, false>(__VA_ARGS__);                                \
    case (L) * 4 + 3: return FN<L, 1, true, true, false>(__VA_ARGS__);

// Measured on B200 (C3): 256-bit row blocks execute 9 % (forward) / 18 % (backward) fewer instructions but run SLOWER
// (fwd 2.98 -> 3.18 ms, bwd 6.14 -> 7.43 ms): the march is latency-bound, and eight small independent row groups per
// iteration overlap better than four large ones (short- and long-scoreboard stalls per issue roughly double).
#ifndef SVOXB_WIDE_ROWS
#define SVOXB_WIDE_ROWS 0
#endif

#if SVOXB_WIDE_ROWS
#define SVOXB_Q_WIDE_CASES(FN, ...)                                                                    \
    SVOXB_Q_CASES(FN, 1, 2, __VA_ARGS__) SVOXB_Q_CASES(FN, 2, 2, __VA_ARGS__)                          \
    SVOXB_Q_CASES(FN, 4, 2, __VA_ARGS__) SVOXB_Q_CASES(FN, 8, 2, __VA_ARGS__) SVOXB_Q_CASES(FN, 16, 2, __VA_ARGS__)
#else
#define SVOXB_Q_WIDE_CASES(FN, ...)
#endif

// 256-bit row blocks when enabled and the rows allow it (D % 8 == 0, 32-byte aligned tables), else 128-bit blocks.
#define SVOXB_Q_DISPATCH(FN, wide, ...)                                                                \
    do {                                                                                               \
        if (tr.D % 4 != 0) {                                                                           \
            const int lpr = pow2ceil((tr.D - 1 + 3) / 4);                                              \
            switch (lpr * 4 + ((tr.use_accel ? 2 : 0) | (image ? 1 : 0))) {                            \
                SVOXB_Q_PAD_CASES(FN, 1, __VA_ARGS__) SVOXB_Q_PAD_CASES(FN, 2, __VA_ARGS__)            \
                SVOXB_Q_PAD_CASES(FN, 4, __VA_ARGS__) SVOXB_Q_PAD_CASES(FN, 8, __VA_ARGS__)            \
                SVOXB_Q_PAD_CASES(FN, 16, __VA_ARGS__) SVOXB_Q_PAD_CASES(FN, 32, __VA_ARGS__)          \
                default: break;                                                                        \
            }                                                                                          \
            set_error("quad kernels: unsupported feature width D=%d", tr.D);                          \
            return SVOXB_EINVAL;                                                                       \
        }                                                                                              \
        const int v4 = (wide) ? 2 : 1;                                                                 \
        const int lpr = pow2ceil(tr.D / (4 * v4));                                                     \
        const int sel = (tr.use_accel ? 2 : 0) | (image ? 1 : 0);                                      \
        switch ((lpr * 2 + v4 - 1) * 4 + sel) {                                                        \
            SVOXB_Q_CASES(FN, 1, 1, __VA_ARGS__) SVOXB_Q_CASES(FN, 2, 1, __VA_ARGS__)                  \
            SVOXB_Q_CASES(FN, 4, 1, __VA_ARGS__) SVOXB_Q_CASES(FN, 8, 1, __VA_ARGS__)                  \
            SVOXB_Q_CASES(FN, 16, 1, __VA_ARGS__) SVOXB_Q_CASES(FN, 32, 1, __VA_ARGS__)                \
            SVOXB_Q_WIDE_CASES(FN, __VA_ARGS__)                                                        \
            default: break;                                                                            \
        }                                                                                              \
        set_error("quad kernels: unsupported feature width D=%d", tr.D);                              \
        return SVOXB_EINVAL;                                                                           \
    } while (0)


int launch_fwd_quad(const TreeArgs& tr_in, const RaySource& src, const MarchOpts& m, bool image, float* out,
                    float* depth, cudaStream_t st) {
    TreeArgs tr = tr_in;
    if (m.sigma_thresh < 0.0f) tr.acc_miss_mask = 0;     // the marks encode sigma > 0: too strict for this predicate
    if (tr.D % 4 == 0)
        SVOXB_REQUIRE(((uintptr_t)tr.features & 15) == 0 && ((uintptr_t)out & 15) == 0, "features/out must be 16-byte aligned");
    SVOXB_REQUIRE(((uintptr_t)tr.feat_act & 15) == 0, "the activated table must be 16-byte aligned");
    const bool wide = SVOXB_WIDE_ROWS && tr.D % 8 == 0 && (((uintptr_t)tr.features | (uintptr_t)tr.feat_act) & 31) == 0;
    SVOXB_Q_DISPATCH(launch_fwd_q, wide, tr, src, m, out, depth, st);
}

int launch_bwd_quad(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, bool image, const float* grad_out,
                    const float* saved_out, float* grad, cudaStream_t st) {
    if (tr.D % 4 == 0)
        SVOXB_REQUIRE(((uintptr_t)tr.features & 15) == 0 && ((uintptr_t)grad & 15) == 0,
                      "features/grad_features must be 16-byte aligned");
    SVOXB_REQUIRE(((uintptr_t)tr.feat_act & 15) == 0, "the activated table must be 16-byte aligned");
    const bool wide = SVOXB_WIDE_ROWS && tr.D % 8 == 0 && (((uintptr_t)tr.features | (uintptr_t)tr.feat_act) & 31) == 0;
    SVOXB_Q_DISPATCH(launch_bwd_q, wide, tr, src, m, grad_out, saved_out, grad, st);
}

}  // namespace svoxb
