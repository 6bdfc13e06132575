// svoxb_render_sh.cu -- the march for the reference's remaining render variants:
//   * view-dependent leaf formats: every output channel is sigmoid(<basis(view dir), B coefficients of the row>)
//     with a spherical-harmonics (1/4/9/16/25), spherical-Gaussian or anisotropic-SG basis, optionally re-evaluated
//     per hit after rotating the view direction by the hit row's 3x3 matrix;
//   * motion-feature render: every hit blends a small per-joint feature table with the hit row's skinning weights.
//
// Replaces (reference paths relative to /root/reference/svox_t/csrc):
//   maybe_precalc_basis                                         rt_kernel.cu:109-185
//   trace_ray / trace_ray_backward, non-RGBA branches           rt_kernel.cu:283-301, 388-417, 463-473
//   motion_feature_trace_ray (+ the backward it meant to be)    rt_kernel.cu:885-1064, 1525-1572
//
// Same skeleton as the scalar-lane RGBA kernels (svoxb_render.cu): persistent warps, lane = ray for the traversal,
// finished lanes refilled from the global queue. What differs is the per-hit arithmetic:
//   view-dependent : LANE-private arithmetic over a warp-cooperative ROW STAGE -- the rows of the lanes' pending candidates
//                    are copied into shared memory with cp.async (coalesced; in flight during the traversal of the next
//                    sample), then the owner lane evaluates its hit against its ray's basis (a lane-private slot in
//                    shared memory, re-evaluated per hit when per-row rotations are given): C dot products of length
//                    <= B over the row's coefficients, in the reference's order; the backward writes its gradient row
//                    (w g_t s_t(1-s_t) basis_i, sigma gradient last) over the staged row and the warp reduces the rows
//                    out coalesced. (Lane-private global reads of the rows cost D-1 load instructions x 32 L1 wavefronts
//                    per iteration: 2.6x slower. A first, fully warp-cooperative version that served the 32 hits of an
//                    iteration one after another was slower than the reference's thread-per-ray code.)
//                    SH rows with three output channels take the register-only kernels of svoxb_render_shrgb.cu.
//   motion feature : lane k < F accumulates sum_j w_j * JF[joint_j][k] (the joint table is a few KB, L1-resident);
//                    the backward reduces dL/dJF in a per-CTA shared-memory table (J x F addresses receive every
//                    contribution of every ray -- global atomics would serialise on them) and flushes it once.
#include <stdlib.h>
#include "svoxb_march.cuh"

namespace svoxb {

#ifndef SVOXB_TILE_SYNC
#define SVOXB_TILE_SYNC 1
#endif
constexpr int MAXB = 25;      // basis functions per channel (SH degree 4)
constexpr int MAXC = 31;      // view-dependent output channels: C + 1 (opacity) values per ray are served by one warp

struct FmtArgs {
    int format, B, C, min_comp, max_comp, extra_cols;
    int S, SB, SC;            // odd row strides (floats) of the per-warp shared-memory tables: row stage (>= D), basis
                              // (>= B), partial outputs / staged grad_out (>= C + 1)
    const float* extra;       // SG [B,>=4]: (lambda, mu) ; ASG [B,>=11]: (a, b, x, y, z)
    const float* tm;          // [M,4,4] per-row view rotation or nullptr
};

__constant__ float kSH[22] = {
    1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f,
    -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f, -0.4570457994644658f,
    1.445305721320277f, -0.5900435899266435f,
    2.5033429417967046f, -1.7701307697799304f, 0.9461746957575601f, -0.6690465435572892f, 0.10578554691520431f,
    -0.6690465435572892f, 0.47308734787878004f, -1.7701307697799304f, 0.6258357354491761f, 0.0f};

// rt_kernel.cu:109-185. `out` is this ray's slot in shared memory.
__device__ __forceinline__ void eval_basis(const FmtArgs& f, float x, float y, float z, float* out) {
    if (f.format == SVOXB_FORMAT_SH) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        out[0] = 0.28209479177387814f;
        if (f.B >= 4) {
            out[1] = -0.4886025119029199f * y;
            out[2] = 0.4886025119029199f * z;
            out[3] = -0.4886025119029199f * x;
        }
        if (f.B >= 9) {
            out[4] = kSH[0] * xy;
            out[5] = kSH[1] * yz;
            out[6] = kSH[2] * (2.0f * zz - xx - yy);
            out[7] = kSH[3] * xz;
            out[8] = kSH[4] * (xx - yy);
        }
        if (f.B >= 16) {
            out[9] = kSH[5] * y * (3.0f * xx - yy);
            out[10] = kSH[6] * xy * z;
            out[11] = kSH[7] * y * (4.0f * zz - xx - yy);
            out[12] = kSH[8] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
            out[13] = kSH[9] * x * (4.0f * zz - xx - yy);
            out[14] = kSH[10] * z * (xx - yy);
            out[15] = kSH[11] * x * (xx - 3.0f * yy);
        }
        if (f.B >= 25) {
            out[16] = kSH[12] * xy * (xx - yy);
            out[17] = kSH[13] * yz * (3.0f * xx - yy);
            out[18] = kSH[14] * xy * (7.0f * zz - 1.0f);
            out[19] = kSH[15] * yz * (7.0f * zz - 3.0f);
            out[20] = kSH[16] * (zz * (35.0f * zz - 30.0f) + 3.0f);
            out[21] = kSH[17] * xz * (7.0f * zz - 3.0f);
            out[22] = kSH[18] * (xx - yy) * (7.0f * zz - 1.0f);
            out[23] = kSH[19] * xz * (xx - 3.0f * yy);
            out[24] = kSH[20] * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy));
        }
    } else if (f.format == SVOXB_FORMAT_SG) {
        for (int i = 0; i < f.B; ++i) {
            const float* p = f.extra + (size_t)i * f.extra_cols;
            const float dt = x * __ldg(p + 1) + y * __ldg(p + 2) + z * __ldg(p + 3);
            out[i] = expf(__ldg(p) * (dt - 1.0f)) / (float)f.B;
        }
    } else {    // ASG
        for (int i = 0; i < f.B; ++i) {
            const float* p = f.extra + (size_t)i * f.extra_cols;
            const float S = x * __ldg(p + 8) + y * __ldg(p + 9) + z * __ldg(p + 10);
            const float dx = x * __ldg(p + 2) + y * __ldg(p + 3) + z * __ldg(p + 4);
            const float dy = x * __ldg(p + 5) + y * __ldg(p + 6) + z * __ldg(p + 7);
            out[i] = S * expf(-__ldg(p) * dx * dx - __ldg(p + 1) * dy * dy) / (float)f.B;
        }
    }
}

// rt_kernel.cu:283-291: rotate the view direction by the upper 3x3 of the hit row's matrix, then re-evaluate.
__device__ __forceinline__ void eval_basis_rotated(const FmtArgs& f, int idx, const ViewDir& v, float* out) {
    const float* m = f.tm + (size_t)(unsigned)idx * 16;
    const float x = __ldg(m + 0) * v.x + __ldg(m + 1) * v.y + __ldg(m + 2) * v.z;
    const float y = __ldg(m + 4) * v.x + __ldg(m + 5) * v.y + __ldg(m + 6) * v.z;
    const float z = __ldg(m + 8) * v.x + __ldg(m + 9) * v.y + __ldg(m + 10) * v.z;
    eval_basis(f, x, y, z, out);
}

// Shared memory per warp of the view-dependent kernels, sized for the format at hand (FmtArgs::SB / SC / S are odd row
// strides, so lane-per-row accesses are conflict-free):
//   basis[32][SB]  basis of each lane's ray (B values);
//   acc[32][SC]    forward: partial outputs of each lane's ray; backward: the ray's staged grad_out row (C + 1 values);
//   rows[32][S]    row stage: row r = the feature row of lane r's pending candidate, D values.
// Row stage. The candidates of a warp's lanes are 32 different rows; read lane-privately, each of the D-1 coefficient
// loads of an iteration would touch 32 different sectors (D-1 instructions x 32 L1 wavefronts -- what bounded the first
// version of these kernels). Instead every lane copies ITS channel (+32k) of every candidate row with cp.async: coalesced
// global reads that land in shared memory without passing through registers, in flight while the lanes traverse to their
// next sample. The lane-private dot products then read shared memory. The backward overwrites each staged row with its
// gradient row and sends those out the same way, as coalesced reductions.
struct FmtWarp {
    float* basis;
    float* acc;
    float* rows;
};

__device__ __forceinline__ FmtWarp fmt_warp(uint32_t* smem_after_top, const FmtArgs& fa) {
    const int per_warp = 32 * (fa.SB + fa.SC + fa.S);
    float* base = reinterpret_cast<float*>(smem_after_top) + (size_t)(threadIdx.x >> 5) * per_warp;
    return FmtWarp{base, base + 32 * fa.SB, base + 32 * (fa.SB + fa.SC)};
}

__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Request the feature rows (all D channels, sigma included) of the lanes' candidates: row of lane r -> rows[r * S ...].
template <int K>
__device__ __forceinline__ void stage_rows_async(const float* __restrict__ features, int D, int S, unsigned cm, int idx,
                                                 int lane, float* rows) {
    while (cm) {
        const int r = __ffs(cm) - 1;
        cm &= cm - 1;
        const int idx_r = __shfl_sync(FULL, idx, r);
        SVOXB_DBG(idx_r >= 0 && S >= D);
        const float* src = features + (size_t)(unsigned)idx_r * D + lane;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(rows + r * S + lane);
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (lane + 32 * k < D) cp_async4(dst + 128u * k, src + 32 * k);
    }
    cp_async_commit();
}

// The gradient rows the lanes left in the stage (row r = lane r's hit) go out as coalesced reductions, row by row.
template <int K>
__device__ __forceinline__ void reduce_hit_rows(float* __restrict__ grad, int D, int S, unsigned hm, int hidx, int lane,
                                                const float* rows) {
    while (hm) {
        const int r = __ffs(hm) - 1;
        hm &= hm - 1;
        const int idx_r = __shfl_sync(FULL, hidx, r);
        float* grow = grad + (size_t)(unsigned)idx_r * D + lane;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (lane + 32 * k < D) {
                const float v = rows[r * S + lane + 32 * k];
                if (v != 0.0f) atomicAdd(grow + 32 * k, v);
            }
        }
    }
}

// One traversal step of the software pipeline shared by both kernels: advances the lane's ray by one sample and returns
// the leaf row of that sample (or -1: empty leaf, or a row the hit marks flag as sigma <= 0) with its delta_t.
template <bool ACCEL>
__device__ __forceinline__ void fmt_next_sample(const TreeArgs& tr, const uint32_t* top, float step, bool active, Ray& ray,
                                                bool& trav_done, int& n_idx, float& n_dt) {
    n_idx = -1; n_dt = 0.0f;
    if (active && !trav_done) {
        if (!(ray.t < ray.tmax)) trav_done = true;
        else {
            Probe pb;
            probe_begin<ACCEL>(tr, top, ray, pb);
            probe_end<ACCEL>(tr, pb, ray, step, n_idx, n_dt);
            ray.t += n_dt;
            if (!(ray.t < ray.tmax)) trav_done = true;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Forward (rt_kernel.cu:261-326), software-pipelined: the rows of the PENDING candidates are requested (cp.async into the
// stage), the lanes traverse to their next sample while the rows are in flight, then every lane composites its pending
// sample: sigma from the staged row, C dot products of length <= B against its ray's basis (lane-private slot).
template <int K, bool ACCEL, bool IMAGE>
__global__ void __launch_bounds__(BLOCK)
march_fmt_fwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, FmtArgs fa, float* __restrict__ out,
                     unsigned long long* counter) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    const int top_words = ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31;
    const FmtWarp sm = fmt_warp(smem_u32 + top_words, fa);
    const int D = tr.D, B = fa.B, C = fa.C, Co = C + 1, S = fa.S, SB = fa.SB, SC = fa.SC;
    const float* off = tr.offset;
    const float* scl = tr.scaling;
    float* my_basis = sm.basis + lane * SB;
    float* my_acc = sm.acc + lane * SC;
    const float* my_row = sm.rows + lane * S;
    for (int t = 0; t < C; ++t) my_acc[t] = 0.0f;

    Ray ray;
    ViewDir vd{0.f, 0.f, 0.f};
    float T = 1.0f, p_dt = 0.0f;
    int row = 0, p_idx = -1;
    bool active = false, trav_done = true;
    Queue q{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (SVOXB_TILE_SYNC && IMAGE ? need == FULL : need != 0u) {   // camera rays: whole tiles (svoxb_render_q.cu)
            const unsigned got = refill<IMAGE, true>(src, off, scl, counter, q, need, lane, ray, row, &vd);
            if ((got >> lane) & 1u) {
                active = true; trav_done = false; T = 1.0f; p_idx = -1;
                eval_basis(fa, vd.x, vd.y, vd.z, my_basis);
            }
            need = 0;
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        stage_rows_async<K>(tr.features, D, S, __ballot_sync(FULL, p_idx >= 0), p_idx, lane, sm.rows);
        int n_idx; float n_dt;
        fmt_next_sample<ACCEL>(tr, top, opt.step, active, ray, trav_done, n_idx, n_dt);
        cp_async_wait_all();
        __syncwarp();

        int fin = 0;                  // 1 = the ray left the volume, 2 = stopped early (T <= stop_thresh)
        SVOXB_DBG(p_idx < 0 || (int64_t)p_idx < tr.M);
        if (p_idx >= 0) {
            const float sigma = my_row[D - 1];
            if (sigma > opt.sigma_thresh) {                                      // rt_kernel.cu:279
                const float att = expf(-p_dt * ray.ds * sigma);
                const float w = T * (1.0f - att);
                if (fa.tm) eval_basis_rotated(fa, p_idx, vd, my_basis);         // rt_kernel.cu:283-291
                for (int t = 0; t < C; ++t) {                                    // rt_kernel.cu:293-301
                    float tmp = 0.0f;
                    for (int i = fa.min_comp; i <= fa.max_comp; ++i) tmp += my_basis[i] * my_row[t * B + i];
                    my_acc[t] = fmaf(w, fast_sigmoid(tmp), my_acc[t]);
                }
                T *= att;
                if (T <= opt.stop_thresh) fin = 2;
            }
        }
        p_idx = n_idx; p_dt = n_dt;
        if (fin == 0 && active && trav_done && p_idx < 0) fin = 1;
        __syncwarp();

        // ---- finished rays (rt_kernel.cu:313-326) -----------------------------------------------------------------
        unsigned fm = __ballot_sync(FULL, fin != 0);
        if (fm) {
            need |= fm;
            if (fin != 0) { active = false; p_idx = -1; }
            while (fm) {
                const int r = __ffs(fm) - 1;
                fm &= fm - 1;
                const float T_r = __shfl_sync(FULL, T, r);
                const int fin_r = __shfl_sync(FULL, fin, r);
                const int row_r = __shfl_sync(FULL, row, r);
                if (lane < Co) {
                    float v;
                    if (lane == C) v = 1.0f - T_r;
                    else if (fin_r == 2) v = sm.acc[r * SC + lane] * (float)(1.0 / (1.0 - (double)T_r));
                    else v = sm.acc[r * SC + lane] + T_r * opt.bg;
                    out[(int64_t)row_r * Co + lane] = v;
                    if (lane < C) sm.acc[r * SC + lane] = 0.0f;
                }
            }
            __syncwarp();
        }
    }
}

// Backward of the above: ONE re-march with accum = <g, out> from the saved forward output (see svoxb_render.cu), the same
// pipeline. With per-row rotations the basis is re-evaluated at every hit -- the reference's second pass keeps the basis
// of the last hit of its first pass (rt_kernel.cu:446-494 never call maybe_precalc_basis), which is not the gradient.
template <int K, bool ACCEL, bool IMAGE>
__global__ void __launch_bounds__(BLOCK)
march_fmt_bwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, FmtArgs fa, const float* __restrict__ grad_out,
                     const float* __restrict__ saved_out, float* __restrict__ grad, unsigned long long* counter) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    const int top_words = ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31;
    const FmtWarp sm = fmt_warp(smem_u32 + top_words, fa);
    const int D = tr.D, B = fa.B, C = fa.C, Co = C + 1, S = fa.S, SB = fa.SB, SC = fa.SC;
    const float* off = tr.offset;
    const float* scl = tr.scaling;
    float* my_basis = sm.basis + lane * SB;
    const float* my_g = sm.acc + lane * SC;
    float* my_row = sm.rows + lane * S;

    Ray ray;
    ViewDir vd{0.f, 0.f, 0.f};
    float T = 1.0f, accum = 0.0f, T_end = 0.0f, gop = 0.0f, p_dt = 0.0f;
    int row = 0, p_idx = -1;
    bool active = false, trav_done = true;
    Queue q{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (SVOXB_TILE_SYNC && IMAGE ? need == FULL : need != 0u) {   // camera rays: whole tiles (svoxb_render_q.cu)
            unsigned got = refill<IMAGE, true>(src, off, scl, counter, q, need, lane, ray, row, &vd);
            if ((got >> lane) & 1u) {
                active = true; trav_done = false; T = 1.0f; p_idx = -1;
                eval_basis(fa, vd.x, vd.y, vd.z, my_basis);
            }
            need = 0;
            while (got) {       // per new ray: stage grad_out, accum = sum_{t<C} g_t out_t, T_end, g_opacity
                const int r = __ffs(got) - 1;
                got &= got - 1;
                const int row_r = __shfl_sync(FULL, row, r);
                float gv = 0.0f, ov = 0.0f;
                if (lane < Co) {
                    gv = __ldg(grad_out + (int64_t)row_r * Co + lane);
                    ov = __ldg(saved_out + (int64_t)row_r * Co + lane);
                    sm.acc[r * SC + lane] = gv;
                }
                float part = lane < C ? gv * ov : 0.0f;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(FULL, part, s);
                const float g_last = __shfl_sync(FULL, gv, C), o_last = __shfl_sync(FULL, ov, C);
                if (lane == r) { accum = part; T_end = 1.0f - o_last; gop = g_last; }
            }
            __syncwarp();
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        stage_rows_async<K>(tr.features, D, S, __ballot_sync(FULL, p_idx >= 0), p_idx, lane, sm.rows);
        int n_idx; float n_dt;
        fmt_next_sample<ACCEL>(tr, top, opt.step, active, ray, trav_done, n_idx, n_dt);
        cp_async_wait_all();
        __syncwarp();

        // Every lane serves its own pending sample: C dot products, the coefficient gradients w g_t s_t(1-s_t) basis_i
        // (rt_kernel.cu:403-417) overwrite the row's coefficients in the stage, the sigma gradient (rt_kernel.cu:479-490)
        // its last slot; the finished gradient rows leave as coalesced reductions.
        bool hit = false;
        const int hidx = max(p_idx, 0);
        SVOXB_DBG(p_idx < 0 || (int64_t)p_idx < tr.M);
        if (p_idx >= 0) {
            const float sig = my_row[D - 1];
            if (sig > 0.0f) {                                                    // rt_kernel.cu:382,456
                const float att = expf(-p_dt * sig * ray.ds);
                const float w = T * (1.0f - att), dd = p_dt * ray.ds;
                hit = true;
                if (fa.tm) eval_basis_rotated(fa, p_idx, vd, my_basis);
                T *= att;
                float c = 0.0f;
                for (int t = 0; t < C; ++t) {
                    float tmp = 0.0f;
                    for (int i = fa.min_comp; i <= fa.max_comp; ++i) tmp += my_basis[i] * my_row[t * B + i];
                    const float s = fast_sigmoid(tmp), gv = my_g[t];
                    c = fmaf(s, gv, c);
                    const float gs = w * s * (1.0f - s) * gv;
                    for (int i = 0; i < B; ++i)
                        my_row[t * B + i] = (i >= fa.min_comp && i <= fa.max_comp) ? gs * my_basis[i] : 0.0f;
                }
                for (int j = C * B; j < D - 1; ++j) my_row[j] = 0.0f;     // channels past the last whole basis block
                accum -= w * c;
                my_row[D - 1] = dd * (c * T - accum) + dd * gop * T_end;
            }
        }
        const unsigned hm = __ballot_sync(FULL, hit);
        if (hm) {
            __syncwarp();
            reduce_hit_rows<K>(grad, D, S, hm, hidx, lane, sm.rows);
        }
        p_idx = n_idx; p_dt = n_dt;
        const bool fin = active && trav_done && p_idx < 0;
        __syncwarp();

        const unsigned fm = __ballot_sync(FULL, fin);
        if (fm) {
            if (fin) active = false;
            need |= fm;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Motion-feature render. jf[J,F] joint features, sw[M,NB] skinning weights and ji[M,NB] joint indices per leaf row.
struct JointArgs {
    const float* jf;
    const float* sw;
    const int32_t* ji;
    int J, F, NB;
};

// pos_joint_feature[k] of row idx (rt_kernel.cu:953-958); lane k < F.
__device__ __forceinline__ float blend_joint_feature(const JointArgs& ja, int idx, int lane) {
    float pj = 0.0f;
    const float* swr = ja.sw + (size_t)(unsigned)idx * ja.NB;
    const int32_t* jir = ja.ji + (size_t)(unsigned)idx * ja.NB;
    for (int j = 0; j < ja.NB; ++j) {
        const float wj = __ldg(swr + j);
        if (wj > 0.0f) pj += wj * __ldg(ja.jf + (size_t)__ldg(jir + j) * ja.F + lane);
    }
    return pj;
}

template <bool ACCEL>
__global__ void __launch_bounds__(BLOCK)
motion_feature_fwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, JointArgs ja, float* __restrict__ out,
                          unsigned long long* counter) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    const int top_words = ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31;
    float (*acc)[32] = reinterpret_cast<float (*)[32]>(smem_u32 + top_words) + (threadIdx.x >> 5) * 32;
    const int F = ja.F;
    const float* off = tr.offset;
    const float* scl = tr.scaling;
    for (int r = 0; r < 32; ++r) acc[r][lane] = 0.0f;

    Ray ray;
    float T = 1.0f;
    int row = 0;
    bool active = false, missed = false;
    Queue q{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (need) {
            const unsigned got = refill<false>(src, off, scl, counter, q, need, lane, ray, row);
            if ((got >> lane) & 1u) {
                active = true; T = 1.0f;
                float a, b;                                      // a ray that misses the cube returns zeros, not the
                dda_unit(ray.ox, ray.oy, ray.oz, ray.ix, ray.iy, ray.iz, a, b);   // background (rt_kernel.cu:911-916)
                missed = b < 0.0f || a > b;
            }
            need = 0;
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        bool hit = false;
        float w = 0.0f;
        int hidx = 0, fin = 0;
        if (active) {
            if (!(ray.t < ray.tmax)) {
                fin = missed ? 3 : 1;
            } else {
                int64_t idx; float delta_t, sigma;
                sample<ACCEL>(tr, top, ray, opt.step, idx, delta_t, sigma);
                // idx >= 0: an EMPTY leaf counts as sigma = 0 in the reference, which then dereferences a null row when the
                // threshold is negative (rt_kernel.cu:278-304); here it is simply not a hit
                if (idx >= 0 && sigma > opt.sigma_thresh) {
                    const float att = expf(-delta_t * ray.ds * sigma);
                    w = T * (1.0f - att);
                    hit = true; hidx = (int)idx;
                    T *= att;
                    if (T <= opt.stop_thresh) fin = 2;
                }
                ray.t += delta_t;
                if (fin == 0 && !(ray.t < ray.tmax)) fin = 1;
            }
        }

        unsigned hm = __ballot_sync(FULL, hit);
        while (hm) {
            const int r = __ffs(hm) - 1;
            hm &= hm - 1;
            const int idx_r = __shfl_sync(FULL, hidx, r);
            const float w_r = __shfl_sync(FULL, w, r);
            if (lane < F) acc[r][lane] = fmaf(w_r, fast_sigmoid(blend_joint_feature(ja, idx_r, lane)), acc[r][lane]);
        }

        unsigned fm = __ballot_sync(FULL, fin != 0);
        if (fm) {
            need = fm;
            if (fin != 0) active = false;
            while (fm) {
                const int r = __ffs(fm) - 1;
                fm &= fm - 1;
                const float T_r = __shfl_sync(FULL, T, r);
                const int fin_r = __shfl_sync(FULL, fin, r);
                const int row_r = __shfl_sync(FULL, row, r);
                if (lane < F) {
                    float v = acc[r][lane];
                    if (fin_r == 2) v *= (float)(1.0 / (1.0 - (double)T_r));     // rt_kernel.cu:966-970
                    else if (fin_r == 1) v += T_r * opt.bg;                      // rt_kernel.cu:975-977
                    else v = 0.0f;
                    out[(int64_t)row_r * F + lane] = v;
                    acc[r][lane] = 0.0f;
                }
            }
        }
    }
}

// dL/dJF[joint_j][k] += w_j * weight * s_k (1 - s_k) * g[k] at every sample with sigma > 0: what
// rt_kernel.cu:981-1064 set out to compute (Appendix B3 lists what its code does instead).
template <bool ACCEL>
__global__ void __launch_bounds__(BLOCK)
motion_feature_bwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, JointArgs ja, const float* __restrict__ grad_out,
                          float* __restrict__ grad_jf, int use_table, unsigned long long* counter) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    const int top_words = ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31;
    float (*gs)[32] = reinterpret_cast<float (*)[32]>(smem_u32 + top_words) + (threadIdx.x >> 5) * 32;
    float* table = reinterpret_cast<float*>(smem_u32 + top_words) + WARPS * 32 * 32;     // [J][F], this CTA's partial sums
    const int F = ja.F, JF = ja.J * ja.F;
    const float* off = tr.offset;
    const float* scl = tr.scaling;
    if (use_table) {
        for (int i = threadIdx.x; i < JF; i += blockDim.x) table[i] = 0.0f;
        __syncthreads();
    }
    float* dst = use_table ? table : grad_jf;

    Ray ray;
    float T = 1.0f;
    int row = 0;
    bool active = false;
    Queue q{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (need) {
            unsigned got = refill<false>(src, off, scl, counter, q, need, lane, ray, row);
            if ((got >> lane) & 1u) { active = true; T = 1.0f; }
            need = 0;
            while (got) {
                const int r = __ffs(got) - 1;
                got &= got - 1;
                const int row_r = __shfl_sync(FULL, row, r);
                gs[r][lane] = lane < F ? __ldg(grad_out + (int64_t)row_r * F + lane) : 0.0f;
            }
            __syncwarp();
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        bool hit = false, fin = false;
        float w = 0.0f;
        int hidx = 0;
        if (active) {
            if (!(ray.t < ray.tmax)) {
                fin = true;
            } else {
                int64_t idx; float delta_t, sigma;
                sample<ACCEL>(tr, top, ray, opt.step, idx, delta_t, sigma);
                if (sigma > 0.0f) {                                              // rt_kernel.cu:1029
                    const float att = expf(-delta_t * sigma * ray.ds);
                    w = T * (1.0f - att);
                    hit = true; hidx = (int)idx;
                    T *= att;
                }
                ray.t += delta_t;
                if (!(ray.t < ray.tmax)) fin = true;
            }
        }

        unsigned hm = __ballot_sync(FULL, hit);
        while (hm) {
            const int r = __ffs(hm) - 1;
            hm &= hm - 1;
            const int idx_r = __shfl_sync(FULL, hidx, r);
            const float w_r = __shfl_sync(FULL, w, r);
            if (lane < F) {
                const float s = fast_sigmoid(blend_joint_feature(ja, idx_r, lane));
                const float gt = w_r * s * (1.0f - s) * gs[r][lane];
                const float* swr = ja.sw + (size_t)(unsigned)idx_r * ja.NB;
                const int32_t* jir = ja.ji + (size_t)(unsigned)idx_r * ja.NB;
                for (int j = 0; j < ja.NB; ++j) {
                    const float wj = __ldg(swr + j);
                    if (wj > 0.0f) atomicAdd(dst + (size_t)__ldg(jir + j) * F + lane, wj * gt);
                }
            }
        }

        const unsigned fm = __ballot_sync(FULL, fin);
        if (fm) {
            if (fin) active = false;
            need = fm;
        }
    }
    if (use_table) {
        __syncthreads();
        for (int i = threadIdx.x; i < JF; i += blockDim.x) {
            const float v = table[i];
            if (v != 0.0f) atomicAdd(grad_jf + i, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Motion-feature render, staged form (the one the entry points launch whenever the joint table fits shared memory).
// Same arithmetic as the kernels above; what changes is where the per-hit operands live:
//   * jf[J,F] sits in shared memory for the whole kernel (one copy per CTA), rows padded to FP = 4 * LPR floats,
//     LPR = pow2ceil(F / 4) lanes x float4 covering a row;
//   * the march is software-pipelined like the feature render: a candidate found in iteration i is composited in
//     iteration i+1, its sigma / skinning row / joint-index row are requested (lane-private gathers) BEFORE the traversal
//     step of iteration i+1 and consumed after it; rows the accelerator marks "sigma <= 0" never become candidates;
//   * the owner lane of a hit stages (weights clamped at 0 -- the reference skips w <= 0, rt_kernel.cu:955 --, joint row
//     offsets, compositing weight, its lane id) in the warp's shared-memory slot number rank = #hits in lower lanes:
//     the hits of an iteration form a dense list, served without any ballot / find-first-set bookkeeping;
//   * forward, NB == 4: 32 / LPR hits are served at once, LPR lanes x float4 each (4 hits x 8 lanes at F = 32): operands
//     by broadcast LDS.128, four LDS.128 from the joint table, 16 FMA, 4 sigmoids, one float4 read-modify-write of the
//     ray's partial output row;
//   * backward: lane = channel, one hit at a time; every warp reduces dL/dJF into ITS OWN [J,FP] table in shared memory
//     with plain read-modify-writes (lane k owns column k, so no two lanes ever touch one address) -- shared-memory
//     float atomics are a compare-and-swap loop on this architecture (ATOMS.CAST.SPIN) -- and flushes it at the end.
// NB4: NB == 4 with 16-byte aligned rows -> the operands travel as one float4 + one int4 per lane.
struct MfSmem {
    uint32_t* top; float* jf; float* rows; float* st_w; int* st_j; float2* st_wr; float* table;
};

__host__ __device__ __forceinline__ int mf_lpr(int F) { return F <= 4 ? 1 : (F <= 8 ? 2 : (F <= 16 ? 4 : 8)); }

template <int TABLES>       // gradient tables per warp (0 in the forward)
__device__ __forceinline__ MfSmem mf_carve(uint32_t* base, int top_words, int JFP, int NB, int warp) {
    MfSmem m;
    m.top = base;
    m.jf = reinterpret_cast<float*>(base + top_words);
    const int nb_pad = (NB + 3) & ~3;
    float* w0 = m.jf + JFP;                                          // per-warp regions follow, 16-byte aligned
    const int per_warp = 32 * 32 + 2 * 32 * nb_pad + 64 + TABLES * JFP;
    float* mine = w0 + (size_t)warp * per_warp;
    m.rows = mine;                                                   // fwd: partial outputs [32][32]; bwd: grad_out rows
    m.st_w = mine + 32 * 32;
    m.st_j = reinterpret_cast<int*>(m.st_w + 32 * nb_pad);
    m.st_wr = reinterpret_cast<float2*>(m.st_j + 32 * nb_pad);
    m.table = reinterpret_cast<float*>(m.st_wr + 32);
    return m;
}

static size_t mf_smem_bytes(int tables, int top_words, int J, int F, int NB, int warps) {
    const int JFP = J * 4 * mf_lpr(F), nb_pad = (NB + 3) & ~3;
    const size_t per_warp = 32 * 32 + 2 * 32 * nb_pad + 64 + (size_t)tables * JFP;
    return sizeof(float) * ((size_t)top_words + JFP + per_warp * warps);
}

// The owner lane's operands of its pending candidate -> slot `slot` of this warp's hit list.
template <bool NB4>
__device__ __forceinline__ void mf_stage(const MfSmem& sm, const JointArgs& ja, int FP, int slot, int lane, int idx,
                                         float w, const float4& wv, const int4& jv) {
    const int nb_pad = (ja.NB + 3) & ~3, jmax = ja.J - 1;
    SVOXB_DBG(slot >= 0 && slot < 32 && idx >= 0);
    sm.st_wr[slot] = make_float2(w, __int_as_float(lane));
    if (NB4) {
        reinterpret_cast<float4*>(sm.st_w)[slot] = make_float4(fmaxf(wv.x, 0.f), fmaxf(wv.y, 0.f), fmaxf(wv.z, 0.f), fmaxf(wv.w, 0.f));
        reinterpret_cast<int4*>(sm.st_j)[slot] = make_int4(min(max(jv.x, 0), jmax) * FP, min(max(jv.y, 0), jmax) * FP,
                                                           min(max(jv.z, 0), jmax) * FP, min(max(jv.w, 0), jmax) * FP);
    } else {
        const float* swr = ja.sw + (size_t)(unsigned)idx * ja.NB;
        const int32_t* jir = ja.ji + (size_t)(unsigned)idx * ja.NB;
        for (int b = 0; b < ja.NB; ++b) {
            sm.st_w[slot * nb_pad + b] = fmaxf(__ldg(swr + b), 0.0f);
            sm.st_j[slot * nb_pad + b] = min(max(__ldg(jir + b), 0), jmax) * FP;
        }
    }
}

// pos_joint_feature[lane] of the hit in slot `slot` (rt_kernel.cu:953-958); jfl = joint table + lane.
template <bool NB4>
__device__ __forceinline__ float mf_blend(const MfSmem& sm, const float* jfl, int NB, int slot) {
    if (NB4) {
        const float4 w = reinterpret_cast<const float4*>(sm.st_w)[slot];
        const int4 j = reinterpret_cast<const int4*>(sm.st_j)[slot];
        return fmaf(w.w, jfl[j.w], fmaf(w.z, jfl[j.z], fmaf(w.y, jfl[j.y], w.x * jfl[j.x])));
    }
    const int nb_pad = (NB + 3) & ~3;
    float pj = 0.0f;
    for (int b = 0; b < NB; ++b) pj = fmaf(sm.st_w[slot * nb_pad + b], jfl[sm.st_j[slot * nb_pad + b]], pj);
    return pj;
}

__device__ __forceinline__ void mf_load_jf(const MfSmem& sm, const JointArgs& ja, int FP) {
    for (int i = threadIdx.x; i < ja.J * FP; i += blockDim.x) {
        const int j = i / FP, k = i - j * FP;
        sm.jf[i] = k < ja.F ? __ldg(ja.jf + j * ja.F + k) : 0.0f;
    }
}

template <bool ACCEL, bool NB4>
__global__ void __launch_bounds__(BLOCK)
mf_fwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, JointArgs ja, float* __restrict__ out,
              unsigned long long* counter) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    const int top_words = ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0;
    if (ACCEL) load_top(tr, smem_u32);
    const int lane = threadIdx.x & 31, F = ja.F, lpr = mf_lpr(F), FP = 4 * lpr;
    const MfSmem sm = mf_carve<0>(smem_u32, top_words, ja.J * FP, ja.NB, threadIdx.x >> 5);
    mf_load_jf(sm, ja, FP);
    for (int r = 0; r < 32; ++r) sm.rows[r * 32 + lane] = 0.0f;
    __syncthreads();
    const float* off = tr.offset;
    const float* scl = tr.scaling;
    const size_t sig_stride = (size_t)tr.D;
    const float* sig_base = tr.features + (tr.D - 1);
    const float* jfl = sm.jf + lane;
    const int q = lane / lpr, c = lane - q * lpr, rpi = 32 / lpr;
    const float4* jf4 = reinterpret_cast<const float4*>(sm.jf) + c;     // float4 block c of every joint row
    float4* rows4 = reinterpret_cast<float4*>(sm.rows) + c;             // ... and of every partial output row

    Ray ray;
    float T = 1.0f, p_dt = 0.0f;
    int row = 0, p_idx = -1;
    bool active = false, missed = false, trav_done = true;
    Queue qu{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (need) {
            const unsigned got = refill<false>(src, off, scl, counter, qu, need, lane, ray, row);
            if ((got >> lane) & 1u) {
                active = true; trav_done = false; T = 1.0f;
                float a, b;                                      // a ray that misses the cube returns zeros, not the
                dda_unit(ray.ox, ray.oy, ray.oz, ray.ix, ray.iy, ray.iz, a, b);   // background (rt_kernel.cu:911-916)
                missed = b < 0.0f || a > b;
            }
            need = 0;
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        // S0: operands of the pending candidate (row 0 stands in for "none": unconditional loads, see svoxb_render_q.cu)
        const int pi = max(p_idx, 0);
        SVOXB_DBG((int64_t)pi < max(tr.M, (int64_t)1));
        const float sig = __ldg(sig_base + (size_t)(unsigned)pi * sig_stride);
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        int4 jv = make_int4(0, 0, 0, 0);
        if (NB4) {
            wv = __ldg(reinterpret_cast<const float4*>(ja.sw) + pi);
            jv = __ldg(reinterpret_cast<const int4*>(ja.ji) + pi);
        }
        // S1: next sample of the traversal
        int n_idx = -1;
        float n_dt = 0.0f;
        if (active && !trav_done) {
            if (!(ray.t < ray.tmax)) trav_done = true;
            else {
                Probe pb;
                probe_begin<ACCEL>(tr, sm.top, ray, pb);
                probe_end<ACCEL>(tr, pb, ray, opt.step, n_idx, n_dt);
                ray.t += n_dt;
                if (!(ray.t < ray.tmax)) trav_done = true;
            }
        }
        // S2: composite the pending candidates
        const bool hit = p_idx >= 0 && sig > opt.sigma_thresh;
        const unsigned hm = __ballot_sync(FULL, hit);
        bool stopped = false;
        if (hm) {
            if (hit) {
                const float att = expf(-p_dt * ray.ds * sig);
                mf_stage<NB4>(sm, ja, FP, __popc(hm & ((1u << lane) - 1u)), lane, p_idx, T * (1.0f - att), wv, jv);
                T *= att;
                if (T <= opt.stop_thresh) stopped = true;
            }
            __syncwarp();
            const int nh = __popc(hm);
            if (NB4) {
                for (int slot = q; slot < nh; slot += rpi) {
                    const float4 w = reinterpret_cast<const float4*>(sm.st_w)[slot];
                    const int4 j = reinterpret_cast<const int4*>(sm.st_j)[slot];
                    const float2 wr = sm.st_wr[slot];
                    const float4 a0 = jf4[j.x >> 2], a1 = jf4[j.y >> 2], a2 = jf4[j.z >> 2], a3 = jf4[j.w >> 2];
                    float4 pj;
                    pj.x = fmaf(w.w, a3.x, fmaf(w.z, a2.x, fmaf(w.y, a1.x, w.x * a0.x)));
                    pj.y = fmaf(w.w, a3.y, fmaf(w.z, a2.y, fmaf(w.y, a1.y, w.x * a0.y)));
                    pj.z = fmaf(w.w, a3.z, fmaf(w.z, a2.z, fmaf(w.y, a1.z, w.x * a0.z)));
                    pj.w = fmaf(w.w, a3.w, fmaf(w.z, a2.w, fmaf(w.y, a1.w, w.x * a0.w)));
                    float4* ar = rows4 + __float_as_int(wr.y) * 8;
                    float4 acc = *ar;
                    acc.x = fmaf(wr.x, fast_sigmoid(pj.x), acc.x);
                    acc.y = fmaf(wr.x, fast_sigmoid(pj.y), acc.y);
                    acc.z = fmaf(wr.x, fast_sigmoid(pj.z), acc.z);
                    acc.w = fmaf(wr.x, fast_sigmoid(pj.w), acc.w);
                    *ar = acc;
                }
            } else if (lane < F) {
                for (int slot = 0; slot < nh; ++slot) {
                    const float2 wr = sm.st_wr[slot];
                    float* ar = sm.rows + __float_as_int(wr.y) * 32 + lane;
                    *ar = fmaf(wr.x, fast_sigmoid(mf_blend<false>(sm, jfl, ja.NB, slot)), *ar);
                }
            }
            __syncwarp();
        }
        if (stopped) { n_idx = -1; trav_done = true; }
        p_idx = n_idx; p_dt = n_dt;
        const int fin = (active && trav_done && p_idx < 0) ? (missed ? 3 : (stopped ? 2 : 1)) : 0;

        unsigned fm = __ballot_sync(FULL, fin != 0);
        if (fm) {
            need = fm;
            if (fin != 0) active = false;
            while (fm) {
                const int r = __ffs(fm) - 1;
                fm &= fm - 1;
                const float T_r = __shfl_sync(FULL, T, r);
                const int fin_r = __shfl_sync(FULL, fin, r);
                const int row_r = __shfl_sync(FULL, row, r);
                float v = sm.rows[r * 32 + lane];
                sm.rows[r * 32 + lane] = 0.0f;                                   // padding channels included
                if (lane < F) {
                    if (fin_r == 2) v *= (float)(1.0 / (1.0 - (double)T_r));     // rt_kernel.cu:966-970
                    else if (fin_r == 1) v += T_r * opt.bg;                      // rt_kernel.cu:975-977
                    else v = 0.0f;
                    __stcs(out + (int64_t)row_r * F + lane, v);
                }
            }
            __syncwarp();
        }
    }
}

// HP = hits served at once in the backward = private gradient tables per warp: 1 (lane = channel), or -- NB == 4 and
// 16 < F <= 32 -- 2 / 4 with 16 / 8 lanes x float2 / float4 per hit.
template <int HP> struct MfVec;
template <> struct MfVec<2> { using T = float2; };
template <> struct MfVec<4> { using T = float4; };
__device__ __forceinline__ float2 mf_axpy(float a, float2 x, float2 y) { return make_float2(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y)); }
__device__ __forceinline__ float4 mf_axpy(float a, float4 x, float4 y) {
    return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}
__device__ __forceinline__ float2 mf_scale(float a, float2 x) { return make_float2(a * x.x, a * x.y); }
__device__ __forceinline__ float4 mf_scale(float a, float4 x) { return make_float4(a * x.x, a * x.y, a * x.z, a * x.w); }
// weight * s (1 - s) * g per component, s = sigmoid(pj)
__device__ __forceinline__ float mf_gt1(float wgt, float pj, float g) {
    const float s = fast_sigmoid(pj), sg = s * g;
    return wgt * fmaf(-sg, s, sg);
}
__device__ __forceinline__ float2 mf_gt(float wgt, float2 pj, float2 g) { return make_float2(mf_gt1(wgt, pj.x, g.x), mf_gt1(wgt, pj.y, g.y)); }
__device__ __forceinline__ float4 mf_gt(float wgt, float4 pj, float4 g) {
    return make_float4(mf_gt1(wgt, pj.x, g.x), mf_gt1(wgt, pj.y, g.y), mf_gt1(wgt, pj.z, g.z), mf_gt1(wgt, pj.w, g.w));
}

template <bool ACCEL, bool NB4, int HP>
__global__ void __launch_bounds__(BLOCK)
mf_bwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, JointArgs ja, const float* __restrict__ grad_out,
              float* __restrict__ grad_jf, unsigned long long* counter) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    const int top_words = ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0;
    if (ACCEL) load_top(tr, smem_u32);
    const int lane = threadIdx.x & 31, F = ja.F, FP = 4 * mf_lpr(F), JFP = ja.J * FP;
    const MfSmem sm = mf_carve<HP>(smem_u32, top_words, JFP, ja.NB, threadIdx.x >> 5);
    mf_load_jf(sm, ja, FP);
    for (int i = lane; i < HP * JFP; i += 32) sm.table[i] = 0.0f;
    __syncthreads();
    const float* off = tr.offset;
    const float* scl = tr.scaling;
    const size_t sig_stride = (size_t)tr.D;
    const float* sig_base = tr.features + (tr.D - 1);
    const float* jfl = sm.jf + lane;
    float* tbl = sm.table + lane;
    const float* gl = sm.rows + lane;

    Ray ray;
    float T = 1.0f, p_dt = 0.0f;
    int row = 0, p_idx = -1;
    bool active = false, trav_done = true;
    Queue qu{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (need) {
            unsigned got = refill<false>(src, off, scl, counter, qu, need, lane, ray, row);
            if ((got >> lane) & 1u) { active = true; trav_done = false; T = 1.0f; }
            need = 0;
            while (got) {
                const int r = __ffs(got) - 1;
                got &= got - 1;
                const int row_r = __shfl_sync(FULL, row, r);
                sm.rows[r * 32 + lane] = lane < F ? __ldcs(grad_out + (int64_t)row_r * F + lane) : 0.0f;
            }
            __syncwarp();
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        const int pi = max(p_idx, 0);
        SVOXB_DBG((int64_t)pi < max(tr.M, (int64_t)1));
        const float sig = __ldg(sig_base + (size_t)(unsigned)pi * sig_stride);
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        int4 jv = make_int4(0, 0, 0, 0);
        if (NB4) {
            wv = __ldg(reinterpret_cast<const float4*>(ja.sw) + pi);
            jv = __ldg(reinterpret_cast<const int4*>(ja.ji) + pi);
        }
        int n_idx = -1;
        float n_dt = 0.0f;
        if (active && !trav_done) {
            if (!(ray.t < ray.tmax)) trav_done = true;
            else {
                Probe pb;
                probe_begin<ACCEL>(tr, sm.top, ray, pb);
                probe_end<ACCEL>(tr, pb, ray, opt.step, n_idx, n_dt);
                ray.t += n_dt;
                if (!(ray.t < ray.tmax)) trav_done = true;
            }
        }
        const bool hit = p_idx >= 0 && sig > 0.0f;                                // rt_kernel.cu:1029
        const unsigned hm = __ballot_sync(FULL, hit);
        if (hm) {
            if (hit) {
                const float att = expf(-p_dt * sig * ray.ds);
                mf_stage<NB4>(sm, ja, FP, __popc(hm & ((1u << lane) - 1u)), lane, p_idx, T * (1.0f - att), wv, jv);
                T *= att;
            }
            __syncwarp();
            const int nh = __popc(hm);
            if constexpr (HP > 1) {
                using V = typename MfVec<HP>::T;
                constexpr int LPH = 32 / HP;                                      // lanes per hit; FP == 32 == LPH * HP
                const int h = lane / LPH, c = lane % LPH;
                const V* jfv = reinterpret_cast<const V*>(sm.jf) + c;
                const V* gv = reinterpret_cast<const V*>(sm.rows) + c;
                V* tb = reinterpret_cast<V*>(sm.table + h * JFP) + c;             // this hit slot's private table
                for (int slot = h; slot < nh; slot += HP) {
                    const float2 wr = sm.st_wr[slot];
                    const float4 w = reinterpret_cast<const float4*>(sm.st_w)[slot];
                    int4 j = reinterpret_cast<const int4*>(sm.st_j)[slot];
                    j.x /= HP; j.y /= HP; j.z /= HP; j.w /= HP;                   // float offsets -> vector offsets
                    const V pj = mf_axpy(w.w, jfv[j.w], mf_axpy(w.z, jfv[j.z], mf_axpy(w.y, jfv[j.y], mf_scale(w.x, jfv[j.x]))));
                    const V gt = mf_gt(wr.x, pj, gv[__float_as_int(wr.y) * LPH]);
                    const bool distinct = j.x != j.y && j.x != j.z && j.x != j.w && j.y != j.z && j.y != j.w && j.z != j.w;
                    if (distinct) {
                        const V t0 = tb[j.x], t1 = tb[j.y], t2 = tb[j.z], t3 = tb[j.w];
                        tb[j.x] = mf_axpy(w.x, gt, t0);
                        tb[j.y] = mf_axpy(w.y, gt, t1);
                        tb[j.z] = mf_axpy(w.z, gt, t2);
                        tb[j.w] = mf_axpy(w.w, gt, t3);
                    } else {
                        tb[j.x] = mf_axpy(w.x, gt, tb[j.x]);
                        tb[j.y] = mf_axpy(w.y, gt, tb[j.y]);
                        tb[j.z] = mf_axpy(w.z, gt, tb[j.z]);
                        tb[j.w] = mf_axpy(w.w, gt, tb[j.w]);
                    }
                }
            } else
            if (lane < F) {
                for (int slot = 0; slot < nh; ++slot) {
                    const float2 wr = sm.st_wr[slot];
                    const float s = fast_sigmoid(mf_blend<NB4>(sm, jfl, ja.NB, slot));
                    const float sg = s * gl[__float_as_int(wr.y) * 32];
                    const float gt = wr.x * fmaf(-sg, s, sg);                     // weight * s (1 - s) * g
                    if (NB4) {
                        const float4 w = reinterpret_cast<const float4*>(sm.st_w)[slot];
                        const int4 j = reinterpret_cast<const int4*>(sm.st_j)[slot];
                        const bool distinct = j.x != j.y && j.x != j.z && j.x != j.w && j.y != j.z && j.y != j.w && j.z != j.w;
                        if (distinct) {           // four independent read-modify-writes: loads first, then the stores
                            const float t0 = tbl[j.x], t1 = tbl[j.y], t2 = tbl[j.z], t3 = tbl[j.w];
                            tbl[j.x] = fmaf(w.x, gt, t0);
                            tbl[j.y] = fmaf(w.y, gt, t1);
                            tbl[j.z] = fmaf(w.z, gt, t2);
                            tbl[j.w] = fmaf(w.w, gt, t3);
                        } else {                  // a joint repeats within the row: keep the updates in order
                            tbl[j.x] = fmaf(w.x, gt, tbl[j.x]);
                            tbl[j.y] = fmaf(w.y, gt, tbl[j.y]);
                            tbl[j.z] = fmaf(w.z, gt, tbl[j.z]);
                            tbl[j.w] = fmaf(w.w, gt, tbl[j.w]);
                        }
                    } else {
                        const int nb_pad = (ja.NB + 3) & ~3;
                        for (int b = 0; b < ja.NB; ++b) {
                            float* t = tbl + sm.st_j[slot * nb_pad + b];
                            *t = fmaf(sm.st_w[slot * nb_pad + b], gt, *t);
                        }
                    }
                }
            }
            __syncwarp();
        }
        p_idx = n_idx; p_dt = n_dt;
        const bool fin = active && trav_done && p_idx < 0;
        const unsigned fm = __ballot_sync(FULL, fin);
        if (fm) {
            if (fin) active = false;
            need = fm;
        }
    }
    __syncwarp();
    for (int i = lane; i < JFP; i += 32) {
        const int j = i / FP, k = i - j * FP;
        float v = sm.table[i];
#pragma unroll
        for (int t = 1; t < HP; ++t) v += sm.table[t * JFP + i];
        if (k < F && v != 0.0f) atomicAdd(grad_jf + j * F + k, v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Motion-feature render, table form (large ray batches). The blended joint feature of a hit depends on the leaf row
// only: pos_joint_feature[idx][k] = sum_b sw[idx][b] jf[ji[idx][b]][k] (rt_kernel.cu:953-958). With Q * 32 >= M rays a
// row is hit many times (77 at C3), so the blend + sigmoid is evaluated ONCE PER ROW into a table laid out exactly like
// the pre-activated feature table of the RGBA render ([M, F+1] rows with the raw sigma last when (F+1) % 4 == 0, else
// payload rows padded to a multiple of 4 floats + a compact sigma array) and the march itself is the feature render's
// quad kernel on that table. Backward: the quad backward reduces dL/d(pre-sigmoid blend) per row (its sigma gradients
// are computed into a column nobody reads), and one pass over the rows folds them into dL/dJF through the rows'
// skinning weights -- M * NB * F multiply-adds instead of (#hits) * NB * F.
__global__ void __launch_bounds__(256)
mf_table_kernel(JointArgs ja, const float* __restrict__ features, int64_t M, int D, int stride, int full_rows,
                float* __restrict__ table, float* __restrict__ sigma) {
    // one thread per table element: the threads of a row re-read its NB weights / joint indices through L1
    const int64_t n = M * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / stride;
        const int k = (int)(i - r * stride);
        float v = 0.0f;
        if (k < ja.F) {
            const float* swr = ja.sw + r * ja.NB;
            const int32_t* jir = ja.ji + r * ja.NB;
            float pj = 0.0f;
            for (int b = 0; b < ja.NB; ++b) {
                const float w = __ldg(swr + b);
                if (w > 0.0f) pj += w * __ldg(ja.jf + (size_t)min(max(__ldg(jir + b), 0), ja.J - 1) * ja.F + k);
            }
            v = fast_sigmoid(pj);
        } else if (full_rows && k == ja.F) {
            v = __ldg(features + r * D + (D - 1));
        }
        table[i] = v;
        if (!full_rows && k == 0) sigma[r] = __ldg(features + r * D + (D - 1));
    }
}

// out[q, 0:F] = tmp[q, 0:F], zeros for rays that miss the cube (rt_kernel.cu:911-916). Lane l tests ray base + l (the
// set-up divides in double precision: once per ray, not once per lane), then the warp copies the 32 rows.
__global__ void __launch_bounds__(256)
mf_finish_kernel(const float* __restrict__ tmp, const float* __restrict__ origins, const float* __restrict__ dirs,
                 const float* __restrict__ off, const float* __restrict__ scl, int64_t Q, int F, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < Q; base += warps_total * 32) {
        const int64_t q = base + lane;
        bool missed = true;
        if (q < Q) {
            Ray ray;
            ray_setup(off, scl, __ldg(origins + 3 * q), __ldg(origins + 3 * q + 1), __ldg(origins + 3 * q + 2),
                      __ldg(dirs + 3 * q), __ldg(dirs + 3 * q + 1), __ldg(dirs + 3 * q + 2), ray);
            float a, b;
            dda_unit(ray.ox, ray.oy, ray.oz, ray.ix, ray.iy, ray.iz, a, b);
            missed = b < 0.0f || a > b;
        }
        const unsigned mm = __ballot_sync(FULL, missed);
        const int nr = (int)min((int64_t)32, Q - base);
        for (int e = lane; e < nr * F; e += 32) {           // the 32 output rows are one contiguous span
            const int rr = e / F, k = e - rr * F;
            __stcs(out + base * F + e, ((mm >> rr) & 1u) ? 0.0f : __ldcs(tmp + (base + rr) * (F + 1) + k));
        }
    }
}

// gpad[q, 0:F] = grad_out[q, 0:F], gpad[q, F] = 0 (no gradient enters through the opacity column).
__global__ void __launch_bounds__(256)
mf_pad_grad_kernel(const float* __restrict__ g, int64_t Q, int F, float* __restrict__ gpad) {
    const int64_t n = Q * (F + 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / (F + 1);
        const int k = (int)(i - q * (F + 1));
        gpad[i] = k < F ? __ldcs(g + q * F + k) : 0.0f;
    }
}

// grad_jf[ji[r][b]][k] += sw[r][b] * gpre[r][k]: lane = channel, one [J,F] table per warp in shared memory (plain
// read-modify-writes), flushed with one atomic per entry and warp. use_table == 0: global atomics per contribution.
__global__ void __launch_bounds__(256)
mf_fold_kernel(JointArgs ja, const float* __restrict__ gpre, int64_t M, int D2, int use_table, float* __restrict__ grad_jf) {
    extern __shared__ __align__(16) float fold_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, JF = ja.J * ja.F, NB = ja.NB;
    float* tbl = use_table ? fold_smem + (size_t)warp * JF : grad_jf;
    if (use_table) for (int i = lane; i < JF; i += 32) tbl[i] = 0.0f;
    __syncwarp();
    constexpr int R = 4;                                     // rows in flight per warp: their loads are independent
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * R; r0 < M; r0 += warps_total * R) {
        for (int k0 = 0; k0 < ja.F; k0 += 32) {
            const int k = k0 + lane;
            float g[R];
#pragma unroll
            for (int i = 0; i < R; ++i) g[i] = (k < ja.F && r0 + i < M) ? __ldcs(gpre + (r0 + i) * D2 + k) : 0.0f;
            // the R rows' weights / joint indices are one contiguous span of R * NB entries: lane l fetches entry l
            const bool lanes_hold = R * NB <= 32;
            float w_l = 0.0f;
            int j_l = 0;
            if (lanes_hold && lane < R * NB && r0 * NB + lane < M * NB) {
                w_l = __ldg(ja.sw + r0 * NB + lane);
                j_l = min(max(__ldg(ja.ji + r0 * NB + lane), 0), ja.J - 1);
            }
#pragma unroll
            for (int i = 0; i < R; ++i) {
                if (r0 + i >= M) break;
                for (int b = 0; b < NB; ++b) {
                    float w;
                    int j;
                    if (lanes_hold) {
                        w = __shfl_sync(FULL, w_l, i * NB + b);
                        j = __shfl_sync(FULL, j_l, i * NB + b);
                    } else {
                        w = __ldg(ja.sw + (r0 + i) * NB + b);
                        j = min(max(__ldg(ja.ji + (r0 + i) * NB + b), 0), ja.J - 1);
                    }
                    if (!(w > 0.0f) || k >= ja.F || g[i] == 0.0f) continue;
                    float* t = tbl + (size_t)j * ja.F + k;
                    if (use_table) *t = fmaf(w, g[i], *t);
                    else atomicAdd(t, w * g[i]);
                }
            }
        }
    }
    if (use_table) {
        __syncwarp();
        for (int i = lane; i < JF; i += 32)
            if (tbl[i] != 0.0f) atomicAdd(grad_jf + i, tbl[i]);
    }
}

int scratch_alloc(void** p, size_t bytes, cudaStream_t st);   // svoxb_tree.cu: stream-ordered pool

// Worth it when every row is visited many times (the same rule the renderer uses for the activated table).
static bool mf_table_route(const TreeArgs& tr, const JointArgs& ja, int64_t Q) {
    static const int force = getenv("SVOXB_MF_TABLE") ? atoi(getenv("SVOXB_MF_TABLE")) : -1;
    if (force >= 0) return force != 0 && ja.F + 1 <= 128;
    return ja.F + 1 <= 128 && tr.M > 0 && Q * 32 >= tr.M;
}

struct MfTable {
    TreeArgs tr2;        // the tree as the quad kernels see it: D = F + 1, rows = blend table
    float* mem = nullptr;
};

static int mf_make_table(const TreeArgs& tr, const JointArgs& ja, cudaStream_t st, MfTable& t) {
    const int D2 = ja.F + 1;
    const bool full = D2 % 4 == 0;
    const int stride = full ? D2 : (ja.F + 3) / 4 * 4;
    const size_t n_tab = (size_t)tr.M * stride, n_sig = full ? 0 : (size_t)tr.M;
    int rc = scratch_alloc((void**)&t.mem, sizeof(float) * (n_tab + n_sig), st);
    if (rc) return rc;
    const int grid = (int)min(((int64_t)n_tab + 255) / 256, (int64_t)sm_count() * 32);
    mf_table_kernel<<<grid, 256, 0, st>>>(ja, tr.features, tr.M, tr.D, stride, full ? 1 : 0, t.mem, t.mem + n_tab);
    count_launch();
    t.tr2 = tr;
    t.tr2.D = D2;
    t.tr2.features = t.mem;
    t.tr2.feat_act = t.mem;
    t.tr2.act_stride = stride;
    t.tr2.sigma_c = full ? nullptr : t.mem + n_tab;
    return check_cuda(cudaGetLastError(), "mf_table_kernel launch");
}

static int mf_table_fwd(const TreeArgs& tr, const JointArgs& ja, const RaySource& src, const MarchOpts& m, float* out,
                        cudaStream_t st) {
    MfTable t;
    int rc = mf_make_table(tr, ja, st, t);
    if (rc) return rc;
    float* tmp = nullptr;
    rc = scratch_alloc((void**)&tmp, sizeof(float) * (size_t)src.total * (ja.F + 1), st);
    if (rc == 0) rc = launch_fwd_quad(t.tr2, src, m, false, tmp, nullptr, st);
    if (rc == 0) {
        const int grid = (int)min((src.total + 255) / 256, (int64_t)sm_count() * 16);
        mf_finish_kernel<<<grid, 256, 0, st>>>(tmp, src.origins, src.dirs, tr.offset, tr.scaling, src.total, ja.F, out);
        count_launch();
        rc = check_cuda(cudaGetLastError(), "mf_finish_kernel launch");
    }
    if (tmp) cudaFreeAsync(tmp, st);
    cudaFreeAsync(t.mem, st);
    return rc;
}

static int mf_table_bwd(const TreeArgs& tr, const JointArgs& ja, const RaySource& src, const MarchOpts& m,
                        const float* grad_out, float* grad_jf, cudaStream_t st) {
    MfTable t;
    int rc = mf_make_table(tr, ja, st, t);
    if (rc) return rc;
    const int D2 = ja.F + 1;
    const size_t n_pad = (size_t)src.total * D2, n_pre = (size_t)tr.M * D2;
    float* buf = nullptr;
    rc = scratch_alloc((void**)&buf, sizeof(float) * (n_pad + n_pre), st);
    if (rc == 0) {
        float* gpad = buf;
        float* gpre = buf + n_pad;
        rc = check_cuda(cudaMemsetAsync(gpre, 0, sizeof(float) * n_pre, st), "memset");
        if (rc == 0) {
            const int grid = (int)min(((int64_t)n_pad + 255) / 256, (int64_t)sm_count() * 16);
            mf_pad_grad_kernel<<<grid, 256, 0, st>>>(grad_out, src.total, ja.F, gpad);
            count_launch();
            // saved_out only feeds the sigma gradients, which land in a column nobody reads: any finite rows will do
            rc = launch_bwd_quad(t.tr2, src, m, false, gpad, gpad, gpre, st);
        }
        if (rc == 0) {
            const size_t tab = sizeof(float) * (size_t)ja.J * ja.F * 8;
            const int use_table = tab <= 96 * 1024;
            auto kern = mf_fold_kernel;
            if (use_table && tab > 47 * 1024)
                rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab), "smem opt-in");
            if (rc == 0) {
                const int grid = (int)min((tr.M * 8 + 255) / 256, (int64_t)sm_count() * 8);
                kern<<<grid, 256, use_table ? tab : 0, st>>>(ja, gpre, tr.M, D2, use_table, grad_jf);
                count_launch();
                rc = check_cuda(cudaGetLastError(), "mf_fold_kernel launch");
            }
        }
    }
    if (buf) cudaFreeAsync(buf, st);
    cudaFreeAsync(t.mem, st);
    return rc;
}

// ---- host side ---------------------------------------------------------------------------------------------------
int make_tree_args(const svoxb_tree* t, TreeArgs& a, void* use_stream);   // svoxb_tree.cu

static int make_fmt(const svoxb_tree* tree, const svoxb_render_options* opt, FmtArgs& f) {
    const int D = tree->D, B = opt->basis_dim;
    SVOXB_REQUIRE(opt->format == SVOXB_FORMAT_SH || opt->format == SVOXB_FORMAT_SG || opt->format == SVOXB_FORMAT_ASG,
                  "unknown data format %d", opt->format);
    SVOXB_REQUIRE(B >= 1 && B <= MAXB, "basis_dim=%d out of range [1,%d]", B, MAXB);
    if (opt->format == SVOXB_FORMAT_SH)
        SVOXB_REQUIRE(B == 1 || B == 4 || B == 9 || B == 16 || B == 25, "SH basis_dim=%d must be 1, 4, 9, 16 or 25", B);
    SVOXB_REQUIRE(D <= 128, "feature width D=%d not supported (max 128)", D);
    f.format = opt->format; f.B = B; f.C = (D - 1) / B;                    // rt_kernel.cu:1352-1358
    SVOXB_REQUIRE(f.C >= 1 && f.C <= MAXC, "%d output channels out of range [1,%d]", f.C, MAXC);
    f.min_comp = opt->min_comp; f.max_comp = opt->max_comp;
    SVOXB_REQUIRE(f.min_comp >= 0 && f.max_comp < B && f.min_comp <= f.max_comp,
                  "min_comp=%d / max_comp=%d out of range for basis_dim=%d", f.min_comp, f.max_comp, B);
    f.extra = tree->extra_data; f.extra_cols = tree->extra_cols;
    if (opt->format != SVOXB_FORMAT_SH) {
        const int need = opt->format == SVOXB_FORMAT_SG ? 4 : 11;
        SVOXB_REQUIRE(f.extra != nullptr && tree->extra_rows >= B && tree->extra_cols >= need,
                      "SG/ASG formats need extra_data [>=%d, >=%d]", B, need);
    }
    f.tm = tree->transformation_matrices;
    f.S = D | 1; f.SB = B | 1; f.SC = (f.C + 1) | 1;
    return 0;
}

static size_t fmt_smem_bytes(const TreeArgs& tr, const FmtArgs& f, bool accel) {
    return (accel ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0) + sizeof(float) * 32 * (f.S + f.SB + f.SC) * WARPS;
}

template <int K, bool ACCEL, bool IMAGE>
static int launch_fmt_fwd(const TreeArgs& tr_in, const RaySource& src, const MarchOpts& m, const FmtArgs& f, float* out,
                          cudaStream_t st) {
    TreeArgs tr = tr_in;
    if (m.sigma_thresh < 0.0f) tr.acc_miss_mask = 0;     // the marks encode sigma > 0: too strict for this predicate
    const size_t smem = fmt_smem_bytes(tr, f, ACCEL);
    auto kern = march_fmt_fwd_kernel<K, ACCEL, IMAGE>;
    int grid = 0;
    int rc = persistent_grid(kern, smem, src.total, grid);
    if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, BLOCK, smem, st>>>(tr, src, m, f, out, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "march_fmt_fwd_kernel launch");
}

template <int K, bool ACCEL, bool IMAGE>
static int launch_fmt_bwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, const FmtArgs& f,
                          const float* go, const float* so, float* grad, cudaStream_t st) {
    const size_t smem = fmt_smem_bytes(tr, f, ACCEL);
    auto kern = march_fmt_bwd_kernel<K, ACCEL, IMAGE>;
    int grid = 0;
    int rc = persistent_grid(kern, smem, src.total, grid);
    if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, BLOCK, smem, st>>>(tr, src, m, f, go, so, grad, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "march_fmt_bwd_kernel launch");
}

#define SVOXB_FMT_DISPATCH(FN, ...)                                                                       \
    do {                                                                                                  \
        const int K = (tr.D + 31) / 32;                                                                   \
        const int sel = K * 4 + (tr.use_accel ? 2 : 0) + (image ? 1 : 0);                                 \
        switch (sel) {                                                                                    \
            case 4: return FN<1, false, false>(__VA_ARGS__);  case 5: return FN<1, false, true>(__VA_ARGS__);   \
            case 6: return FN<1, true, false>(__VA_ARGS__);   case 7: return FN<1, true, true>(__VA_ARGS__);    \
            case 8: return FN<2, false, false>(__VA_ARGS__);  case 9: return FN<2, false, true>(__VA_ARGS__);   \
            case 10: return FN<2, true, false>(__VA_ARGS__);  case 11: return FN<2, true, true>(__VA_ARGS__);   \
            case 12: return FN<3, false, false>(__VA_ARGS__); case 13: return FN<3, false, true>(__VA_ARGS__);  \
            case 14: return FN<3, true, false>(__VA_ARGS__);  case 15: return FN<3, true, true>(__VA_ARGS__);   \
            case 16: return FN<4, false, false>(__VA_ARGS__); case 17: return FN<4, false, true>(__VA_ARGS__);  \
            case 18: return FN<4, true, false>(__VA_ARGS__);  case 19: return FN<4, true, true>(__VA_ARGS__);   \
            default: break;                                                                               \
        }                                                                                                 \
        set_error("feature width D=%d not supported (max 128)", tr.D);                                   \
        return SVOXB_EINVAL;                                                                              \
    } while (0)

// svoxb_render_shrgb.cu: lane-private fast path for SH / SG / ASG rows with three output channels
bool sh_rgb_supported(int format, int B, int D);
int sh_rgb_fwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, int B, int min_comp, int max_comp,
               const float* tm, int format, const float* extra, int extra_cols, bool image, float* out,
               cudaStream_t st);
int sh_rgb_bwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, int B, int min_comp, int max_comp,
               const float* tm, int format, const float* extra, int extra_cols, bool image, const float* go,
               const float* so, float* grad, cudaStream_t st);

// Entry points used by svoxb_render.cu when opt->format != RGBA.
int fmt_render_fwd(const svoxb_tree* tree, const TreeArgs& tr, const RaySource& src, const MarchOpts& m,
                   const svoxb_render_options* opt, bool image, float* out, cudaStream_t st) {
    FmtArgs f;
    int rc = make_fmt(tree, opt, f); if (rc) return rc;
    if (sh_rgb_supported(f.format, f.B, tr.D))
        return sh_rgb_fwd(tr, src, m, f.B, f.min_comp, f.max_comp, f.tm, f.format, f.extra, f.extra_cols, image, out, st);
    SVOXB_FMT_DISPATCH(launch_fmt_fwd, tr, src, m, f, out, st);
}

int fmt_render_bwd(const svoxb_tree* tree, const TreeArgs& tr, const RaySource& src, const MarchOpts& m,
                   const svoxb_render_options* opt, bool image, const float* go, const float* so, float* grad,
                   cudaStream_t st) {
    FmtArgs f;
    int rc = make_fmt(tree, opt, f); if (rc) return rc;
    if (sh_rgb_supported(f.format, f.B, tr.D))
        return sh_rgb_bwd(tr, src, m, f.B, f.min_comp, f.max_comp, f.tm, f.format, f.extra, f.extra_cols, image, go, so, grad, st);
    SVOXB_FMT_DISPATCH(launch_fmt_bwd, tr, src, m, f, go, so, grad, st);
}

static int make_joint_args(const svoxb_tree* tree, const float* jf, const float* sw, const int32_t* ji, int J, int F,
                           int NB, JointArgs& ja) {
    SVOXB_REQUIRE(jf && sw && ji, "joint_features / skinning_weights / joint_index are NULL");
    SVOXB_REQUIRE(J >= 1 && F >= 1 && F <= 127 && NB >= 1, "motion feature render: J=%d, F=%d (max 127), B=%d", J, F, NB);
    (void)tree;
    ja.jf = jf; ja.sw = sw; ja.ji = ji; ja.J = J; ja.F = F; ja.NB = NB;
    return 0;
}

// The staged kernels need the joint table (and, backward, one gradient table per warp) in shared memory next to the
// top grid and the per-warp rows; two CTAs per SM should still fit.
static bool mf_staged_ok(const TreeArgs& tr, const JointArgs& ja, int tables) {
    const int top_words = tr.use_accel ? (1 << (3 * tr.acc.bits[0])) : 0;
    static const int limit_kb = getenv("SVOXB_MF_SMEM_KB") ? atoi(getenv("SVOXB_MF_SMEM_KB")) : 110;   // two CTAs per SM
    return ja.NB <= 16 && mf_smem_bytes(tables, top_words, ja.J, ja.F, ja.NB, WARPS) <= (size_t)limit_kb * 1024;
}
static bool mf_nb4(const JointArgs& ja) {
    return ja.NB == 4 && (((uintptr_t)ja.sw | (uintptr_t)ja.ji) & 15) == 0;
}

}  // namespace svoxb

using namespace svoxb;

extern "C" int svoxb_out_data_dim(int32_t format, int32_t basis_dim, int32_t D) {
    if (format == SVOXB_FORMAT_RGBA) return D;
    return basis_dim > 0 ? (D - 1) / basis_dim + 1 : -1;
}

extern "C" int svoxb_motion_feature_render_fwd(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                               const svoxb_render_options* opt, const float* joint_features,
                                               const float* skinning_weights, const int32_t* joint_index, int32_t J,
                                               int32_t F, int32_t B, float* out, void* stream) {
    TreeArgs tr; JointArgs ja;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    SVOXB_REQUIRE(opt != nullptr, "render options are NULL");
    rc = make_joint_args(tree, joint_features, skinning_weights, joint_index, J, F, B, ja); if (rc) return rc;
    SVOXB_REQUIRE(Q >= 0 && Q < (1ll << 31) && (Q == 0 || (origins && dirs && out)), "bad ray batch");
    if (Q == 0) return 0;
    MarchOpts m{opt->step_size, opt->background_brightness, opt->sigma_thresh, opt->stop_thresh};
    RaySource src{}; src.origins = origins; src.dirs = dirs; src.total = Q; src.ndc_w = -1;
    cudaStream_t st = (cudaStream_t)stream;
    if (m.sigma_thresh < 0.0f) tr.acc_miss_mask = 0;          // the marks encode sigma > 0: too strict for this predicate
    if (mf_table_route(tr, ja, Q)) return mf_table_fwd(tr, ja, src, m, out, st);
    SVOXB_REQUIRE(F <= 32, "motion feature render: F=%d > 32 needs a large ray batch (Q * 32 >= M, the table form)", F);
    if (mf_staged_ok(tr, ja, 0)) {
        if (m.sigma_thresh < 0.0f) tr.acc_miss_mask = 0;      // the marks encode sigma > 0: too strict for this predicate
        const int top_words = tr.use_accel ? (1 << (3 * tr.acc.bits[0])) : 0;
        const size_t smem2 = mf_smem_bytes(0, top_words, J, F, B, WARPS);
        const bool nb4 = mf_nb4(ja);
        void (*k2)(TreeArgs, RaySource, MarchOpts, JointArgs, float*, unsigned long long*) =
            tr.use_accel ? (nb4 ? mf_fwd_kernel<true, true> : mf_fwd_kernel<true, false>)
                         : (nb4 ? mf_fwd_kernel<false, true> : mf_fwd_kernel<false, false>);
        int grid2 = 0;
        rc = persistent_grid(k2, smem2, Q, grid2); if (rc) return rc;
        unsigned long long* counter2 = work_counter(st);
        if (!counter2) return SVOXB_ECUDA;
        k2<<<grid2, BLOCK, smem2, st>>>(tr, src, m, ja, out, counter2);
        count_launch();
        return check_cuda(cudaGetLastError(), "mf_fwd_kernel launch");
    }
    tr.acc_miss_mask = 0;                                     // the one-hit-at-a-time kernels fetch sigma themselves
    const size_t smem = (tr.use_accel ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0) + sizeof(float) * WARPS * 32 * 32;
    void (*kern)(TreeArgs, RaySource, MarchOpts, JointArgs, float*, unsigned long long*) =
        tr.use_accel ? motion_feature_fwd_kernel<true> : motion_feature_fwd_kernel<false>;
    int grid = 0;
    rc = persistent_grid(kern, smem, Q, grid); if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, BLOCK, smem, st>>>(tr, src, m, ja, out, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "motion_feature_fwd_kernel launch");
}

extern "C" int svoxb_motion_feature_render_bwd(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                               const svoxb_render_options* opt, const float* joint_features,
                                               const float* skinning_weights, const int32_t* joint_index, int32_t J,
                                               int32_t F, int32_t B, const float* grad_out, float* grad_joint_features,
                                               void* stream) {
    TreeArgs tr; JointArgs ja;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    SVOXB_REQUIRE(opt != nullptr, "render options are NULL");
    rc = make_joint_args(tree, joint_features, skinning_weights, joint_index, J, F, B, ja); if (rc) return rc;
    SVOXB_REQUIRE(grad_joint_features != nullptr, "grad_joint_features is NULL");
    SVOXB_REQUIRE(Q >= 0 && Q < (1ll << 31) && (Q == 0 || (origins && dirs && grad_out)), "bad ray batch");
    cudaStream_t st = (cudaStream_t)stream;
    SVOXB_CUDA(cudaMemsetAsync(grad_joint_features, 0, sizeof(float) * (size_t)J * F, st));
    if (Q == 0) return 0;
    MarchOpts m{opt->step_size, opt->background_brightness, opt->sigma_thresh, opt->stop_thresh};
    RaySource src{}; src.origins = origins; src.dirs = dirs; src.total = Q; src.ndc_w = -1;
    if (mf_table_route(tr, ja, Q)) return mf_table_bwd(tr, ja, src, m, grad_out, grad_joint_features, st);
    SVOXB_REQUIRE(F <= 32, "motion feature render: F=%d > 32 needs a large ray batch (Q * 32 >= M, the table form)", F);
    if (mf_staged_ok(tr, ja, 1)) {
        const int top_words = tr.use_accel ? (1 << (3 * tr.acc.bits[0])) : 0;
        const bool nb4 = mf_nb4(ja);
        // hits served at once = private tables per warp: as many as still leave two CTAs per SM
        int hp = 1;
        if (nb4 && F > 16) {
            static const int want = getenv("SVOXB_MF_HP") ? atoi(getenv("SVOXB_MF_HP")) : 2;
            for (hp = want; hp > 1 && !mf_staged_ok(tr, ja, hp); hp >>= 1) {}
        }
        const size_t smem2 = mf_smem_bytes(hp, top_words, J, F, B, WARPS);
        void (*k2)(TreeArgs, RaySource, MarchOpts, JointArgs, const float*, float*, unsigned long long*);
        if (hp == 4) k2 = tr.use_accel ? mf_bwd_kernel<true, true, 4> : mf_bwd_kernel<false, true, 4>;
        else if (hp == 2) k2 = tr.use_accel ? mf_bwd_kernel<true, true, 2> : mf_bwd_kernel<false, true, 2>;
        else k2 = tr.use_accel ? (nb4 ? mf_bwd_kernel<true, true, 1> : mf_bwd_kernel<true, false, 1>)
                               : (nb4 ? mf_bwd_kernel<false, true, 1> : mf_bwd_kernel<false, false, 1>);
        int grid2 = 0;
        rc = persistent_grid(k2, smem2, Q, grid2); if (rc) return rc;
        unsigned long long* counter2 = work_counter(st);
        if (!counter2) return SVOXB_ECUDA;
        k2<<<grid2, BLOCK, smem2, st>>>(tr, src, m, ja, grad_out, grad_joint_features, counter2);
        count_launch();
        return check_cuda(cudaGetLastError(), "mf_bwd_kernel launch");
    }
    tr.acc_miss_mask = 0;
    const size_t table = sizeof(float) * (size_t)J * F;
    const int use_table = table <= 96 * 1024;
    const size_t smem = (tr.use_accel ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0) + sizeof(float) * WARPS * 32 * 32 +
                        (use_table ? table : 0);
    void (*kern)(TreeArgs, RaySource, MarchOpts, JointArgs, const float*, float*, int, unsigned long long*) =
        tr.use_accel ? motion_feature_bwd_kernel<true> : motion_feature_bwd_kernel<false>;
    int grid = 0;
    rc = persistent_grid(kern, smem, Q, grid); if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, BLOCK, smem, st>>>(tr, src, m, ja, grad_out, grad_joint_features, use_table, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "motion_feature_bwd_kernel launch");
}
