// svoxb_render_shrgb.cu -- fast path of the view-dependent render for the classic layout: spherical harmonics with
// three output channels (RGB) -- rows of 3*B coefficients + sigma, B in {1, 4, 9, 16, 25} (D = 4, 13, 28, 49, 76).
//
// Replaces the same reference code as svoxb_render_sh.cu (trace_ray / trace_ray_backward non-RGBA branches,
// rt_kernel.cu:283-301, 388-417, 463-473; maybe_precalc_basis 109-185) for FORMAT_SH with C == 3.
//
// Here the per-hit work is LANE-private: a hit costs 3 dot products of length B against the ray's basis, so the lane
// that owns the ray keeps the basis (B registers) and its three partial outputs in registers and reads its hit row
// itself -- the 3B+1 floats of a row are contiguous (one or two 128-byte lines, 128-bit loads when D % 4 == 0), sigma
// comes with them, nothing goes through shared memory and no shuffle is needed. The traversal is the packed
// accelerator walk of the other kernels (rows marked sigma <= 0 are skipped), finished lanes refill from the global
// ray queue. Backward: one re-march with accum = <g, out> from the saved output; the row gradient is w g_t s_t(1-s_t)
// basis_i, which needs no row data, and leaves as red.global.add.v4 (D % 4 == 0) or scalar reductions.
#include "svoxb_march.cuh"

namespace svoxb {

#ifndef SVOXB_TILE_SYNC
#define SVOXB_TILE_SYNC 1
#endif

// L2 eviction priority of the gradient table (read-modify-written: a miss costs a DRAM read and a write-back), as in the
// quad backward (svoxb_render_q.cu).
__device__ __forceinline__ uint64_t sh_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

struct ShArgs {
    int min_comp, max_comp;
    const float* tm;          // [M,4,4] per-row view rotation or nullptr
    int format;               // SVOXB_FORMAT_SH / _SG / _ASG: only the per-ray basis differs (rt_kernel.cu:109-185)
    const float* extra;       // SG [B,>=4]: (lambda, mu) ; ASG [B,>=11]: (a, b, x, y, z)
    int extra_cols;
};

template <int B>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float (&out)[B]) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    out[0] = 0.28209479177387814f;
    if constexpr (B >= 4) {
        out[1] = -0.4886025119029199f * y;
        out[2] = 0.4886025119029199f * z;
        out[3] = -0.4886025119029199f * x;
    }
    if constexpr (B >= 9) {
        out[4] = 1.0925484305920792f * xy;
        out[5] = -1.0925484305920792f * yz;
        out[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
        out[7] = -1.0925484305920792f * xz;
        out[8] = 0.5462742152960396f * (xx - yy);
    }
    if constexpr (B >= 16) {
        out[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
        out[10] = 2.890611442640554f * xy * z;
        out[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
        out[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
        out[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
        out[14] = 1.445305721320277f * z * (xx - yy);
        out[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
    }
    if constexpr (B >= 25) {
        out[16] = 2.5033429417967046f * xy * (xx - yy);
        out[17] = -1.7701307697799304f * yz * (3.0f * xx - yy);
        out[18] = 0.9461746957575601f * xy * (7.0f * zz - 1.0f);
        out[19] = -0.6690465435572892f * yz * (7.0f * zz - 3.0f);
        out[20] = 0.10578554691520431f * (zz * (35.0f * zz - 30.0f) + 3.0f);
        out[21] = -0.6690465435572892f * xz * (7.0f * zz - 3.0f);
        out[22] = 0.47308734787878004f * (xx - yy) * (7.0f * zz - 1.0f);
        out[23] = -1.7701307697799304f * xz * (xx - 3.0f * yy);
        out[24] = 0.6258357354491761f * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy));
    }
}

// The ray's B basis values for any of the three view-dependent formats (the format is uniform over the launch).
template <int B>
__device__ __forceinline__ void fmt_basis(const ShArgs& sa, float x, float y, float z, float (&out)[B]) {
    if (sa.format == SVOXB_FORMAT_SH) {
        sh_basis<B>(x, y, z, out);
    } else if (sa.format == SVOXB_FORMAT_SG) {                             // rt_kernel.cu:116-124
#pragma unroll
        for (int i = 0; i < B; ++i) {
            const float* p = sa.extra + (size_t)i * sa.extra_cols;
            const float dt = x * __ldg(p + 1) + y * __ldg(p + 2) + z * __ldg(p + 3);
            out[i] = expf(__ldg(p) * (dt - 1.0f)) / (float)B;
        }
    } else {                                                               // ASG, rt_kernel.cu:125-140
#pragma unroll
        for (int i = 0; i < B; ++i) {
            const float* p = sa.extra + (size_t)i * sa.extra_cols;
            const float S = x * __ldg(p + 8) + y * __ldg(p + 9) + z * __ldg(p + 10);
            const float dx = x * __ldg(p + 2) + y * __ldg(p + 3) + z * __ldg(p + 4);
            const float dy = x * __ldg(p + 5) + y * __ldg(p + 6) + z * __ldg(p + 7);
            out[i] = S * expf(-__ldg(p) * dx * dx - __ldg(p + 1) * dy * dy) / (float)B;
        }
    }
}

template <int B>
__device__ __forceinline__ void sh_basis_row(const ShArgs& sa, int idx, const ViewDir& v, float (&out)[B]) {
    const float* m = sa.tm + (size_t)(unsigned)idx * 16;                 // rt_kernel.cu:283-291
    fmt_basis<B>(sa, __ldg(m + 0) * v.x + __ldg(m + 1) * v.y + __ldg(m + 2) * v.z,
                __ldg(m + 4) * v.x + __ldg(m + 5) * v.y + __ldg(m + 6) * v.z,
                __ldg(m + 8) * v.x + __ldg(m + 9) * v.y + __ldg(m + 10) * v.z, out);
}

// The 3B + 1 floats of a row in registers (static indexing); 128-bit loads when the rows are 16-byte aligned.
template <int B, bool VEC>
__device__ __forceinline__ void load_row(const float* __restrict__ rowp, float (&v)[3 * B + 1]) {
    constexpr int N = 3 * B + 1;
    if constexpr (VEC) {
        static_assert(N % 4 == 0, "vector rows need D % 4 == 0");
#pragma unroll
        for (int k = 0; k < N / 4; ++k) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(rowp) + k);
            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = __ldg(rowp + k);
    }
}

// tmp_t = sum_{i in window} basis_i * row[t*B + i], in the reference's order (rt_kernel.cu:295-299).
template <int B>
__device__ __forceinline__ void sh_dots(const float (&v)[3 * B + 1], const float (&basis)[B], int lo, int hi, float (&tmp)[3]) {
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        float a = 0.0f;
#pragma unroll
        for (int i = 0; i < B; ++i)
            if (i >= lo && i <= hi) a = fmaf(basis[i], v[t * B + i], a);
        tmp[t] = a;
    }
}

template <int B, bool VEC, bool ACCEL, bool IMAGE>
__global__ void __launch_bounds__(BLOCK)
sh_rgb_fwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, ShArgs sa, float* __restrict__ out, unsigned long long* counter) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31;
    constexpr int D = 3 * B + 1;
    const float* off = tr.offset;
    const float* scl = tr.scaling;

    // Software pipeline (as in svoxb_render_q.cu): the candidate found in iteration i is composited in iteration i+1; its
    // row is requested at the top of that iteration, the brick lookup of the next sample is issued right after it and
    // consumed after the compositing -- the two dependent load chains of a sample run side by side. Rows the accelerator
    // marks "sigma <= 0" never become candidates.
    Ray ray;
    ViewDir vd{0.f, 0.f, 0.f};
    float basis[B];
    float T = 1.0f, a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, p_dt = 0.0f;
    int row = 0, p_idx = -1;
    bool active = false, trav_done = true;
    Queue q{0, 0, false};
    unsigned need = FULL;

    while (true) {
        // Camera rays: a warp takes 32 new pixels only when ALL its lanes have finished -- the rays of a pixel tile then
        // stay in step and whole warps skip the row stage while they cross empty space (as in svoxb_render_q.cu).
        if (SVOXB_TILE_SYNC && IMAGE ? need == FULL : need != 0u) {
            const unsigned got = refill<IMAGE, true>(src, off, scl, counter, q, need, lane, ray, row, &vd);
            if ((got >> lane) & 1u) {
                active = true; trav_done = false; T = 1.0f; a0 = a1 = a2 = 0.0f;
                fmt_basis<B>(sa, vd.x, vd.y, vd.z, basis);
            }
            need = 0;
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        // S0: row of the pending candidate (row 0 stands in for "none": unconditional loads, except that a camera-ray warp
        // without any pending candidate -- empty space -- skips the stage)
        float v[D];
        SVOXB_DBG((int64_t)max(p_idx, 0) < max(tr.M, (int64_t)1));
        const bool rows_wanted = !(SVOXB_TILE_SYNC && IMAGE) || __ballot_sync(FULL, p_idx >= 0) != 0u;
        if (rows_wanted) load_row<B, VEC>(tr.features + (size_t)(unsigned)max(p_idx, 0) * D, v);
        // S1: next sample, brick lookup issued
        bool trav = active && !trav_done;
        Probe pb;
        if (trav) {
            if (!(ray.t < ray.tmax)) { trav_done = true; trav = false; }
            else probe_begin<ACCEL>(tr, top, ray, pb);
        }
        // S2: composite the pending candidate
        bool stopped = false;
        if (rows_wanted && p_idx >= 0 && v[D - 1] > opt.sigma_thresh) {                // rt_kernel.cu:279-320
            float tmp[3];
            const float att = expf(-p_dt * ray.ds * v[D - 1]);
            const float w = T * (1.0f - att);
            if (sa.tm) sh_basis_row<B>(sa, p_idx, vd, basis);
            sh_dots<B>(v, basis, sa.min_comp, sa.max_comp, tmp);
            a0 = fmaf(w, fast_sigmoid(tmp[0]), a0);
            a1 = fmaf(w, fast_sigmoid(tmp[1]), a1);
            a2 = fmaf(w, fast_sigmoid(tmp[2]), a2);
            T *= att;
            if (T <= opt.stop_thresh) stopped = true;
        }
        // S3: brick word consumed -> next candidate
        int n_idx = -1;
        float n_dt = 0.0f;
        if (trav) {
            probe_end<ACCEL>(tr, pb, ray, opt.step, n_idx, n_dt);
            ray.t += n_dt;
            if (!(ray.t < ray.tmax)) trav_done = true;
        }
        if (stopped) { n_idx = -1; trav_done = true; }
        p_idx = n_idx; p_dt = n_dt;
        const int fin = (active && trav_done && p_idx < 0) ? (stopped ? 2 : 1) : 0;   // 1 = left the volume, 2 = stopped early
        if (fin != 0) {                                                              // rt_kernel.cu:313-326
            float4 o;
            if (fin == 2) {
                const float scale = (float)(1.0 / (1.0 - (double)T));
                o = make_float4(a0 * scale, a1 * scale, a2 * scale, 1.0f - T);
            } else {
                const float add = T * opt.bg;
                o = make_float4(a0 + add, a1 + add, a2 + add, 1.0f - T);
            }
            __stcs(reinterpret_cast<float4*>(out) + row, o);
            active = false;
        }
        need |= __ballot_sync(FULL, fin != 0);
    }
}

template <int B, bool VEC, bool ACCEL, bool IMAGE>
__global__ void __launch_bounds__(BLOCK)
sh_rgb_bwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, ShArgs sa, const float* __restrict__ grad_out,
                  const float* __restrict__ saved_out, float* __restrict__ grad, unsigned long long* counter) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31;
    constexpr int D = 3 * B + 1;
    float* stage = reinterpret_cast<float*>(smem_u32 + (ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0)) + (threadIdx.x >> 5) * 32 * D;
    const uint64_t pol_last = sh_policy_evict_last();
    const float* off = tr.offset;
    const float* scl = tr.scaling;

    Ray ray;
    ViewDir vd{0.f, 0.f, 0.f};
    float basis[B];
    float T = 1.0f, accum = 0.0f, T_end = 0.0f, g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, gop = 0.0f, p_dt = 0.0f;
    int row = 0, p_idx = -1;
    bool active = false, trav_done = true;
    Queue q{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (need) {
            const unsigned got = refill<IMAGE, true>(src, off, scl, counter, q, need, lane, ray, row, &vd);
            if ((got >> lane) & 1u) {
                active = true; trav_done = false; T = 1.0f;
                fmt_basis<B>(sa, vd.x, vd.y, vd.z, basis);
                const float4 g = __ldcs(reinterpret_cast<const float4*>(grad_out) + row);
                const float4 so = __ldcs(reinterpret_cast<const float4*>(saved_out) + row);
                g0 = g.x; g1 = g.y; g2 = g.z; gop = g.w;
                accum = g.x * so.x + g.y * so.y + g.z * so.z;       // = the reference's pass-1 total (rt:428-436)
                T_end = 1.0f - so.w;
            }
            need = 0;
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        // S0 / S1 as in the forward kernel
        float v[D];
        SVOXB_DBG((int64_t)max(p_idx, 0) < max(tr.M, (int64_t)1));
        load_row<B, VEC>(tr.features + (size_t)(unsigned)max(p_idx, 0) * D, v);
        bool trav = active && !trav_done;
        Probe pb;
        if (trav) {
            if (!(ray.t < ray.tmax)) { trav_done = true; trav = false; }
            else probe_begin<ACCEL>(tr, top, ray, pb);
        }
        // S2: gradient of the pending candidate
        int hit_idx = -1;
        if (p_idx >= 0 && v[D - 1] > 0.0f) {                                          // rt_kernel.cu:382,456
            float tmp[3];
            const float sigma = v[D - 1];
            const float att = expf(-p_dt * sigma * ray.ds);
            const float w = T * (1.0f - att), dd = p_dt * ray.ds;
            if (sa.tm) sh_basis_row<B>(sa, p_idx, vd, basis);
            sh_dots<B>(v, basis, sa.min_comp, sa.max_comp, tmp);
            const float s0 = fast_sigmoid(tmp[0]), s1 = fast_sigmoid(tmp[1]), s2 = fast_sigmoid(tmp[2]);
            const float c = s0 * g0 + s1 * g1 + s2 * g2;                              // rt_kernel.cu:416
            float gs[3] = {w * s0 * (1.0f - s0) * g0, w * s1 * (1.0f - s1) * g1, w * s2 * (1.0f - s2) * g2};
            T *= att;
            accum -= w * c;                                                          // rt_kernel.cu:479-480
            const float sgrad = dd * (c * T - accum) + dd * gop * T_end;              // rt_kernel.cu:486-490
            // row gradient: d/d coef[t*B+i] = gs_t * basis_i inside the component window; sigma last
#pragma unroll
            for (int k = 0; k < D - 1; ++k) {
                const int i = k % B;
                v[k] = (i >= sa.min_comp && i <= sa.max_comp) ? gs[k / B] * basis[i] : 0.0f;
            }
            v[D - 1] = sgrad;
            hit_idx = p_idx;
            if constexpr (VEC) {
#pragma unroll
                for (int k = 0; k < D / 4; ++k)
                    reinterpret_cast<float4*>(stage + lane * D)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) stage[lane * D + k] = v[k];
            }
        }
        // The 32 lanes' gradient rows leave through shared memory: consecutive lanes reduce consecutive pieces of a row, so
        // one instruction covers 32 / (D/4) whole rows (full sectors) instead of one 16-byte piece of 32 different rows.
        const unsigned hm = __ballot_sync(FULL, hit_idx >= 0);
        if (hm) {
            __syncwarp();
            constexpr int PR = VEC ? D / 4 : D;                       // pieces per row
#pragma unroll
            for (int i = 0; i < PR; ++i) {
                const int e = lane + 32 * i, r = e / PR, k = e - r * PR;
                const int idx_r = __shfl_sync(FULL, hit_idx, r);
                if (idx_r >= 0) {
                    SVOXB_DBG((int64_t)idx_r < tr.M);
                    float* grow = grad + (size_t)(unsigned)idx_r * D;
                    if constexpr (VEC) {
                        const float4 t = reinterpret_cast<const float4*>(stage)[e];
                        asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(grow + 4 * k),
                                     "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w), "l"(pol_last) : "memory");
                    } else {
                        const float t = stage[e];
                        if (t != 0.0f)
                            asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(grow + k), "f"(t), "l"(pol_last)
                                         : "memory");
                    }
                }
            }
            __syncwarp();
        }
        // S3
        int n_idx = -1;
        float n_dt = 0.0f;
        if (trav) {
            probe_end<ACCEL>(tr, pb, ray, opt.step, n_idx, n_dt);
            ray.t += n_dt;
            if (!(ray.t < ray.tmax)) trav_done = true;
        }
        p_idx = n_idx; p_dt = n_dt;
        const bool fin = active && trav_done && p_idx < 0;
        if (fin) active = false;
        need = __ballot_sync(FULL, fin);
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
template <int B, bool VEC, bool ACCEL, bool IMAGE>
static int launch_sh_fwd(const TreeArgs& tr_in, const RaySource& src, const MarchOpts& m, const ShArgs& sa, float* out,
                         cudaStream_t st) {
    TreeArgs tr = tr_in;
    if (m.sigma_thresh < 0.0f) tr.acc_miss_mask = 0;     // the marks encode sigma > 0: too strict for this predicate
    const size_t smem = ACCEL ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0;
    auto kern = sh_rgb_fwd_kernel<B, VEC, ACCEL, IMAGE>;
    int grid = 0;
    int rc = persistent_grid(kern, smem, src.total, grid);
    if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, BLOCK, smem, st>>>(tr, src, m, sa, out, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "sh_rgb_fwd_kernel launch");
}

template <int B, bool VEC, bool ACCEL, bool IMAGE>
static int launch_sh_bwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, const ShArgs& sa, const float* go,
                         const float* so, float* grad, cudaStream_t st) {
    const size_t smem = (ACCEL ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0) + sizeof(float) * WARPS * 32 * (3 * B + 1);
    auto kern = sh_rgb_bwd_kernel<B, VEC, ACCEL, IMAGE>;
    int grid = 0;
    int rc = persistent_grid(kern, smem, src.total, grid);
    if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, BLOCK, smem, st>>>(tr, src, m, sa, go, so, grad, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "sh_rgb_bwd_kernel launch");
}

#define SVOXB_SH_AI(FN, BB, VV, ...)                                                     \
    (tr.use_accel ? (image ? FN<BB, VV, true, true>(__VA_ARGS__) : FN<BB, VV, true, false>(__VA_ARGS__)) \
                  : (image ? FN<BB, VV, false, true>(__VA_ARGS__) : FN<BB, VV, false, false>(__VA_ARGS__)))

// True when (format, basis_dim, D) is the SH-RGB layout these kernels cover.
bool sh_rgb_supported(int format, int B, int D) {
    return (format == SVOXB_FORMAT_SH || format == SVOXB_FORMAT_SG || format == SVOXB_FORMAT_ASG) &&
           (B == 1 || B == 4 || B == 9 || B == 16 || B == 25) && D == 3 * B + 1;
}

int sh_rgb_fwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, int B, int min_comp, int max_comp,
               const float* tm, int format, const float* extra, int extra_cols, bool image, float* out,
               cudaStream_t st) {
    const ShArgs sa{min_comp, max_comp, tm, format, extra, extra_cols};
    const bool al = (((uintptr_t)tr.features | (uintptr_t)out) & 15) == 0;
    SVOXB_REQUIRE(((uintptr_t)out & 15) == 0, "out must be 16-byte aligned");
    switch (B) {
        case 1: return al ? SVOXB_SH_AI(launch_sh_fwd, 1, true, tr, src, m, sa, out, st) : SVOXB_SH_AI(launch_sh_fwd, 1, false, tr, src, m, sa, out, st);
        case 4: return SVOXB_SH_AI(launch_sh_fwd, 4, false, tr, src, m, sa, out, st);
        case 9: return al ? SVOXB_SH_AI(launch_sh_fwd, 9, true, tr, src, m, sa, out, st) : SVOXB_SH_AI(launch_sh_fwd, 9, false, tr, src, m, sa, out, st);
        case 16: return SVOXB_SH_AI(launch_sh_fwd, 16, false, tr, src, m, sa, out, st);
        case 25: return al ? SVOXB_SH_AI(launch_sh_fwd, 25, true, tr, src, m, sa, out, st) : SVOXB_SH_AI(launch_sh_fwd, 25, false, tr, src, m, sa, out, st);
        default: break;
    }
    set_error("sh_rgb_fwd: unsupported basis_dim %d", B);
    return SVOXB_EINVAL;
}

int sh_rgb_bwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, int B, int min_comp, int max_comp,
               const float* tm, int format, const float* extra, int extra_cols, bool image, const float* go,
               const float* so, float* grad, cudaStream_t st) {
    const ShArgs sa{min_comp, max_comp, tm, format, extra, extra_cols};
    SVOXB_REQUIRE((((uintptr_t)go | (uintptr_t)so) & 15) == 0, "grad_out / saved_out must be 16-byte aligned");
    const bool al = (((uintptr_t)tr.features | (uintptr_t)grad) & 15) == 0;
    switch (B) {
        case 1: return al ? SVOXB_SH_AI(launch_sh_bwd, 1, true, tr, src, m, sa, go, so, grad, st) : SVOXB_SH_AI(launch_sh_bwd, 1, false, tr, src, m, sa, go, so, grad, st);
        case 4: return SVOXB_SH_AI(launch_sh_bwd, 4, false, tr, src, m, sa, go, so, grad, st);
        case 9: return al ? SVOXB_SH_AI(launch_sh_bwd, 9, true, tr, src, m, sa, go, so, grad, st) : SVOXB_SH_AI(launch_sh_bwd, 9, false, tr, src, m, sa, go, so, grad, st);
        case 16: return SVOXB_SH_AI(launch_sh_bwd, 16, false, tr, src, m, sa, go, so, grad, st);
        case 25: return al ? SVOXB_SH_AI(launch_sh_bwd, 25, true, tr, src, m, sa, go, so, grad, st) : SVOXB_SH_AI(launch_sh_bwd, 25, false, tr, src, m, sa, go, so, grad, st);
        default: break;
    }
    set_error("sh_rgb_bwd: unsupported basis_dim %d", B);
    return SVOXB_EINVAL;
}

}  // namespace svoxb
