// svoxb_render.cu -- per-ray march of the octree: forward (features + opacity + first-hit depth), backward
// (gradient scatter into the leaf feature table) and depth-only, for explicit ray batches and pinhole images.
//
// Replaces (reference paths relative to /root/reference/svox_t/csrc):
//   trace_ray / render_ray_kernel / render_image_kernel           rt_kernel.cu:221-328, 654-671, 1193-1213
//   trace_ray_backward / render_ray_backward_kernel / image bwd   rt_kernel.cu:330-496, 674-694, 1215-1238
//   depth_trace_ray / depth_render_ray_kernel                     rt_kernel.cu:781-834, 865-882
//   cam2world_ray                                                 rt_kernel.cu:1152-1166
//
// B200 design (not a port -- the reference is one thread per ray doing scalar row reads and scalar atomics):
//   * persistent warps; every LANE owns one ray for the traversal (phase A: descend, slab test, sigma gather,
//     transmittance update -- thread-private, no divergence inside a sample);
//   * the feature ROW work is warp-cooperative (phase B): the hits of the 32 lanes are served one after another
//     by the whole warp, lane j <-> channel j (+32k), so every row read / output write / gradient reduction is
//     ONE coalesced 128-byte transaction and the sigmoid costs one MUFU pair per hit instead of D per thread;
//     the 32 x D partial outputs of a warp's rays live in registers (static indexing through full unrolling);
//   * finished lanes are refilled from a chunked global ray queue (ray compaction: no lane idles while its
//     neighbours march on);
//   * backward is a single re-march: sum_j g_j * out_j from the saved forward output replaces the reference's
//     first pass; the per-hit channel dot product is reduced across lanes with a transposing butterfly
//     (9 shuffles per 8 hits) and the row gradient leaves as one coalesced red.global.add per hit;
//   * N == 2 trees are walked through the packed grid+brick accelerator (top grid staged in shared memory),
//     any other N through the reference tensors.
#include <stdlib.h>
#include "svoxb_march.cuh"

namespace svoxb {

// ------------------------------------------------------------------------------------------------------------
// Forward: out[row, 0..D-2] = sum_i w_i * sigmoid(f_i) + T * bg ; out[row, D-1] = 1 - T ; depth[row] = first hit.
template <int K, bool ACCEL, bool IMAGE>
__global__ void __launch_bounds__(BLOCK)
march_fwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, float* __restrict__ out, float* __restrict__ depth,
                 unsigned long long* counter) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31;
    const int D = tr.D;
    const float* off = tr.offset;
    const float* scl = tr.scaling;

    const char* fbase = reinterpret_cast<const char*>(tr.features + lane);
    const unsigned row_bytes = (unsigned)D * 4u;
    bool chan_ok[K];
#pragma unroll
    for (int k = 0; k < K; ++k) chan_ok[k] = lane + 32 * k < D - 1;

    float acc[32][K];
#pragma unroll
    for (int r = 0; r < 32; ++r)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[r][k] = 0.0f;

    Ray ray;
    float T = 1.0f, depth_v = 0.0f;
    int row = 0;
    bool active = false, got_depth = false;
    Queue q{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (need) {
            const unsigned got = refill<IMAGE>(src, off, scl, counter, q, need, lane, ray, row);
            if ((got >> lane) & 1u) { active = true; T = 1.0f; got_depth = false; depth_v = 0.0f; }
            need = 0;
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        // ---- phase A: one sample per lane -----------------------------------------------------------------
        bool hit = false;
        float w = 0.0f;
        int hidx = 0;
        int fin = 0;                 // 1 = ray left the volume, 2 = stopped early (T <= stop_thresh)
        if (active) {
            if (!(ray.t < ray.tmax)) {
                fin = 1;
            } else {
                int64_t idx; float delta_t, sigma;
                sample<ACCEL>(tr, top, ray, opt.step, idx, delta_t, sigma);
                // idx >= 0: an EMPTY leaf counts as sigma = 0 in the reference, which then dereferences a null row when the
                // threshold is negative (rt_kernel.cu:278-304); here it is simply not a hit
                if (idx >= 0 && sigma > opt.sigma_thresh) {                                   // rt_kernel.cu:279-320
                    const float att = expf(-delta_t * ray.ds * sigma);
                    w = T * (1.0f - att);
                    hit = true; hidx = (int)idx;
                    if (!got_depth) { depth_v = ray.ds * ray.t; got_depth = true; } // rt_kernel.cu:826-830
                    T *= att;
                    if (T <= opt.stop_thresh) fin = 2;
                }
                ray.t += delta_t;
                if (fin == 0 && !(ray.t < ray.tmax)) fin = 1;
            }
        }

        // ---- phase B: the warp serves the hits, 8 at a time, lane <-> channel ---------------------------------
        const unsigned hm = __ballot_sync(FULL, hit);
        if (hm) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const unsigned gm = (hm >> (8 * g)) & 0xffu;
                if (gm) {
                    float x[8][K], wr[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 8 * g + i;
                        const int idx_r = __shfl_sync(FULL, hidx, r);      // 0 for lanes without a hit: a valid row
                        wr[i] = __shfl_sync(FULL, w, r);
                        const float* rowp = row_ptr(fbase, idx_r, row_bytes);
#pragma unroll
                        for (int k = 0; k < K; ++k) x[i][k] = chan_ok[k] ? __ldg(rowp + 32 * k) : 0.0f;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if ((gm >> i) & 1u) {
#pragma unroll
                            for (int k = 0; k < K; ++k)
                                acc[8 * g + i][k] = fmaf(wr[i], fast_sigmoid(x[i][k]), acc[8 * g + i][k]);
                        }
                    }
                }
            }
        }

        // ---- finished rays: coalesced row write, then the lane goes back to the queue --------------------------
        const unsigned fm = __ballot_sync(FULL, fin != 0);
        if (fm) {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                if (fm & (1u << r)) {
                    const float T_r = __shfl_sync(FULL, T, r);
                    const int fin_r = __shfl_sync(FULL, fin, r);
                    const int row_r = __shfl_sync(FULL, row, r);
                    float* o = out + (int64_t)row_r * D;
                    const float scale = (float)(1.0 / (1.0 - (double)T_r));        // rt_kernel.cu:315
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int c = lane + 32 * k;
                        const float v = (fin_r == 2) ? acc[r][k] * scale : acc[r][k] + T_r * opt.bg;
                        if (c < D - 1) o[c] = v;
                        else if (c == D - 1) o[c] = 1.0f - T_r;
                        acc[r][k] = 0.0f;
                    }
                }
            }
            if (fin != 0) {
                if (depth) depth[row] = depth_v;
                active = false;
            }
            need = fm;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Backward. Per ray (owner lane): accum starts at sum_{j<D-1} g_j*out_j (= reference pass 1 total, rt:428-436),
// T_end = 1 - out[D-1]. Per hit: grad[idx, j] += w s_j (1-s_j) g_j ; c = sum_j s_j g_j ; T *= att ; accum -= w c ;
// grad[idx, D-1] += dd (c T - accum) + dd g_{D-1} T_end, dd = delta_t * delta_scale (rt:479-490).
template <int K, bool ACCEL, bool IMAGE>
__global__ void __launch_bounds__(BLOCK)
march_bwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, const float* __restrict__ grad_out,
                 const float* __restrict__ saved_out, float* __restrict__ grad, unsigned long long* counter) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    const int top_words = ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0;
    uint32_t* top = smem_u32;
    float* gs_all = reinterpret_cast<float*>(smem_u32 + top_words);      // [WARPS][32][32*K]
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* gs = gs_all + (size_t)warp * 32 * (32 * K);
    const int D = tr.D;
    const float* off = tr.offset;
    const float* scl = tr.scaling;

    const char* fbase = reinterpret_cast<const char*>(tr.features + lane);
    char* gbase = reinterpret_cast<char*>(grad + lane);
    const unsigned row_bytes = (unsigned)D * 4u;
    bool chan_ok[K];
#pragma unroll
    for (int k = 0; k < K; ++k) chan_ok[k] = lane + 32 * k < D - 1;

    Ray ray;
    float T = 1.0f, accum = 0.0f, T_end = 0.0f, gop = 0.0f;
    int row = 0;
    bool active = false;
    Queue q{0, 0, false};
    unsigned need = FULL;

    while (true) {
        if (need) {
            unsigned got = refill<IMAGE>(src, off, scl, counter, q, need, lane, ray, row);
            if ((got >> lane) & 1u) { active = true; T = 1.0f; }
            need = 0;
            // cooperative per-ray set-up: stage grad_out row in shared memory, accum = <g, out>, T_end, g_opacity
            while (got) {
                const int r = __ffs(got) - 1;
                got &= got - 1;
                const int row_r = __shfl_sync(FULL, row, r);
                const float* g = grad_out + (int64_t)row_r * D;
                const float* so = saved_out + (int64_t)row_r * D;
                float part = 0.0f, g_last = 0.0f, o_last = 0.0f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int c = lane + 32 * k;
                    const float gv = (c < D) ? __ldg(g + c) : 0.0f;
                    const float ov = (c < D) ? __ldg(so + c) : 0.0f;
                    gs[r * (32 * K) + c] = gv;
                    if (c < D - 1) part = fmaf(gv, ov, part);
                    if (c == D - 1) { g_last = gv; o_last = ov; }
                }
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(FULL, part, s);
                const int lsrc = (D - 1) & 31;
                // the lane holding channel D-1 is the same for every k it can appear in
                g_last = __shfl_sync(FULL, g_last, lsrc);
                o_last = __shfl_sync(FULL, o_last, lsrc);
                if (lane == r) { accum = part; T_end = 1.0f - o_last; gop = g_last; }
            }
            __syncwarp();
        }
        if (__ballot_sync(FULL, active) == 0u) break;

        // ---- phase A ---------------------------------------------------------------------------------------
        bool hit = false;
        float w = 0.0f, dd = 0.0f;
        int hidx = 0;
        bool fin = false;
        if (active) {
            if (!(ray.t < ray.tmax)) {
                fin = true;
            } else {
                int64_t idx; float delta_t, sigma;
                sample<ACCEL>(tr, top, ray, opt.step, idx, delta_t, sigma);
                if (sigma > 0.0f) {                                              // rt_kernel.cu:382,456
                    const float att = expf(-delta_t * sigma * ray.ds);
                    w = T * (1.0f - att);
                    dd = delta_t * ray.ds;
                    hit = true; hidx = (int)idx;
                    T *= att;
                }
                ray.t += delta_t;
                if (!(ray.t < ray.tmax)) fin = true;
            }
        }

        // ---- phase B ---------------------------------------------------------------------------------------
        const unsigned hm = __ballot_sync(FULL, hit);
        if (hm) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const unsigned gm = (hm >> (8 * g)) & 0xffu;
                if (gm) {
                    float x[8][K], sv[8][K], cp[8];
                    int idxr[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        idxr[i] = __shfl_sync(FULL, hidx, 8 * g + i);     // 0 for lanes without a hit: a valid row
                        const float* rowp = row_ptr(fbase, idxr[i], row_bytes);
#pragma unroll
                        for (int k = 0; k < K; ++k) x[i][k] = chan_ok[k] ? __ldg(rowp + 32 * k) : 0.0f;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 8 * g + i;
                        const bool on = (gm >> i) & 1u;
                        cp[i] = 0.0f;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const float s = fast_sigmoid(x[i][k]);
                            const float gv = gs[r * (32 * K) + lane + 32 * k];
                            const float sg = (on && chan_ok[k]) ? s * gv : 0.0f;
                            cp[i] += sg;
                            sv[i][k] = sg * (1.0f - s);
                        }
                    }
                    // transposing butterfly: 8 per-lane partials -> totals; value i ends up in lanes 4i..4i+3
                    float a4[4], a2[2], a1;
                    {
                        const bool up = lane & 16;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float send = up ? cp[j] : cp[j + 4];
                            const float keep = up ? cp[j + 4] : cp[j];
                            a4[j] = keep + __shfl_xor_sync(FULL, send, 16);
                        }
                    }
                    {
                        const bool up = lane & 8;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float send = up ? a4[j] : a4[j + 2];
                            const float keep = up ? a4[j + 2] : a4[j];
                            a2[j] = keep + __shfl_xor_sync(FULL, send, 8);
                        }
                    }
                    {
                        const bool up = lane & 4;
                        const float send = up ? a2[0] : a2[1];
                        const float keep = up ? a2[1] : a2[0];
                        a1 = keep + __shfl_xor_sync(FULL, send, 4);
                    }
                    a1 += __shfl_xor_sync(FULL, a1, 2);
                    a1 += __shfl_xor_sync(FULL, a1, 1);
                    const float c_own = __shfl_sync(FULL, a1, 4 * (lane & 7));   // owner lane 8g+i reads value i
                    float sgrad = 0.0f;
                    if (hit && (lane >> 3) == g) {
                        accum -= w * c_own;
                        sgrad = dd * (c_own * T - accum) + dd * gop * T_end;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 8 * g + i;
                        const float w_i = __shfl_sync(FULL, w, r);
                        const float sg_i = __shfl_sync(FULL, sgrad, r);
                        if ((gm >> i) & 1u) {
                            float* grow = row_ptr(gbase, idxr[i], row_bytes);
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                if (chan_ok[k]) atomicAdd(grow + 32 * k, w_i * sv[i][k]);
                                else if (lane + 32 * k == D - 1) atomicAdd(grow + 32 * k, sg_i);
                            }
                        }
                    }
                }
            }
        }

        const unsigned fm = __ballot_sync(FULL, fin);
        if (fm) {
            if (fin) active = false;
            need = fm;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Depth only (rt_kernel.cu:781-834): one thread per ray, stops at the first sample with sigma > sigma_thresh.
template <bool ACCEL>
__global__ void __launch_bounds__(BLOCK)
depth_kernel(TreeArgs tr, const float* __restrict__ origins, const float* __restrict__ dirs, int64_t Q,
             MarchOpts opt, float* __restrict__ depth) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    const float* off = tr.offset;
    const float* scl = tr.scaling;
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < Q; id += (int64_t)gridDim.x * blockDim.x) {
        Ray ray;
        ray_setup(off, scl, __ldg(origins + 3 * id), __ldg(origins + 3 * id + 1), __ldg(origins + 3 * id + 2),
                  __ldg(dirs + 3 * id), __ldg(dirs + 3 * id + 1), __ldg(dirs + 3 * id + 2), ray);
        float d = 0.0f;
        while (ray.t < ray.tmax) {
            int64_t idx; float delta_t, sigma;
            sample<ACCEL>(tr, top, ray, opt.step, idx, delta_t, sigma);
            if (sigma > opt.sigma_thresh) { d = ray.ds * ray.t; break; }
            ray.t += delta_t;
        }
        depth[id] = d;
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
int make_tree_args(const svoxb_tree* t, TreeArgs& a, void* use_stream);   // svoxb_tree.cu

// svoxb_render_sh.cu: the view-dependent formats (SH / SG / ASG)
int fmt_render_fwd(const svoxb_tree* tree, const TreeArgs& tr, const RaySource& src, const MarchOpts& m,
                   const svoxb_render_options* opt, bool image, float* out, cudaStream_t st);
int fmt_render_bwd(const svoxb_tree* tree, const TreeArgs& tr, const RaySource& src, const MarchOpts& m,
                   const svoxb_render_options* opt, bool image, const float* go, const float* so, float* grad,
                   cudaStream_t st);

static int check_opts(const svoxb_render_options* opt, MarchOpts& m) {
    SVOXB_REQUIRE(opt != nullptr, "render options are NULL");
    SVOXB_REQUIRE(opt->format >= SVOXB_FORMAT_RGBA && opt->format <= SVOXB_FORMAT_ASG, "unknown data format %d",
                  opt->format);
    m.step = opt->step_size; m.bg = opt->background_brightness;
    m.sigma_thresh = opt->sigma_thresh; m.stop_thresh = opt->stop_thresh;
    return 0;
}

// NDC (rt_kernel.cu:1168-1191) applies to camera rays only: the reference's ray-batch kernels never call it.
static int make_source(const float* origins, const float* dirs, const float* vdirs, int64_t Q, const svoxb_camera* cam,
                       const svoxb_render_options* opt, RaySource& s) {
    s = RaySource{};
    s.ndc_w = -1;
    s.vdirs = vdirs;
    if (cam) {
        if (opt->ndc_width >= 0) {
            SVOXB_REQUIRE(opt->ndc_width > 0 && opt->ndc_height > 0, "bad NDC size %d x %d", opt->ndc_width, opt->ndc_height);
            s.ndc_w = opt->ndc_width; s.ndc_h = opt->ndc_height; s.ndc_focal = opt->ndc_focal;
        }
        SVOXB_REQUIRE(cam->c2w != nullptr && cam->width > 0 && cam->height > 0, "bad camera spec");
        SVOXB_REQUIRE((int64_t)cam->width * cam->height < (1ll << 31), "image too large");
        s.c2w = cam->c2w; s.fx = cam->fx; s.fy = cam->fy; s.width = cam->width; s.height = cam->height;
        s.row_begin = 0; s.row_end = cam->height;
        if (cam->row_end > 0) {
            SVOXB_REQUIRE(cam->row_begin >= 0 && cam->row_begin < cam->row_end && cam->row_end <= cam->height,
                          "bad image band [%d, %d) for height %d", cam->row_begin, cam->row_end, cam->height);
            s.row_begin = cam->row_begin; s.row_end = cam->row_end;
        }
        s.tiles_x = (cam->width + 7) / 8;
        s.total = (int64_t)s.tiles_x * ((s.row_end - s.row_begin + 7) / 8) * 64;
    } else {
        SVOXB_REQUIRE(Q >= 0 && Q < (1ll << 31), "ray count out of range");
        SVOXB_REQUIRE(Q == 0 || (origins && dirs), "origins/dirs are NULL");
        s.origins = origins; s.dirs = dirs; s.total = Q;
    }
    return 0;
}

template <int K, bool ACCEL, bool IMAGE>
static int launch_fwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, float* out, float* depth,
                      cudaStream_t st) {
    const size_t smem = ACCEL ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0;
    auto kern = march_fwd_kernel<K, ACCEL, IMAGE>;
    int grid = 0;
    int rc = persistent_grid(kern, smem, src.total, grid);
    if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, BLOCK, smem, st>>>(tr, src, m, out, depth, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "march_fwd_kernel launch");
}

template <int K, bool ACCEL, bool IMAGE>
static int launch_bwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, const float* grad_out,
                      const float* saved_out, float* grad, cudaStream_t st) {
    const size_t smem = (ACCEL ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0) + sizeof(float) * WARPS * 32 * 32 * K;
    auto kern = march_bwd_kernel<K, ACCEL, IMAGE>;
    int grid = 0;
    int rc = persistent_grid(kern, smem, src.total, grid);
    if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, BLOCK, smem, st>>>(tr, src, m, grad_out, saved_out, grad, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "march_bwd_kernel launch");
}

template <bool IMAGE>
static int dispatch_fwd(const TreeArgs& tr_in, const RaySource& src, const MarchOpts& m, float* out, float* depth,
                        cudaStream_t st) {
    if (quad_supported(tr_in)) return launch_fwd_quad(tr_in, src, m, IMAGE, out, depth, st);
    TreeArgs tr = tr_in;
    if (m.sigma_thresh < 0.0f) tr.acc_miss_mask = 0;     // the marks encode sigma > 0: too strict for this predicate
    const int K = (tr.D + 31) / 32;
#define SVOXB_FWD(KK)                                                                        \
    case KK:                                                                                 \
        return tr.use_accel ? launch_fwd<KK, true, IMAGE>(tr, src, m, out, depth, st)        \
                            : launch_fwd<KK, false, IMAGE>(tr, src, m, out, depth, st);
    switch (K) {
        SVOXB_FWD(1) SVOXB_FWD(2) SVOXB_FWD(3) SVOXB_FWD(4)
        default: break;
    }
#undef SVOXB_FWD
    return launch_fwd_wide(tr, src, m, IMAGE, out, depth, st);      // D > 128: svoxb_render_wide.cu
}

template <bool IMAGE>
static int dispatch_bwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, const float* go,
                        const float* so, float* grad, cudaStream_t st) {
    if (quad_supported(tr)) return launch_bwd_quad(tr, src, m, IMAGE, go, so, grad, st);
    const int K = (tr.D + 31) / 32;
#define SVOXB_BWD(KK)                                                                        \
    case KK:                                                                                 \
        return tr.use_accel ? launch_bwd<KK, true, IMAGE>(tr, src, m, go, so, grad, st)      \
                            : launch_bwd<KK, false, IMAGE>(tr, src, m, go, so, grad, st);
    switch (K) {
        SVOXB_BWD(1) SVOXB_BWD(2) SVOXB_BWD(3) SVOXB_BWD(4)
        default: break;
    }
#undef SVOXB_BWD
    return launch_bwd_wide(tr, src, m, IMAGE, go, so, grad, st);      // D > 128: svoxb_render_wide.cu
}

}  // namespace svoxb

using namespace svoxb;

static int render_rays_fwd_impl(const svoxb_tree* tree, const float* origins, const float* dirs,
                                const float* vdirs, int64_t Q, const svoxb_render_options* opt, float* out,
                                float* depth, int32_t* ray_cost, void* stream) {
    TreeArgs tr; MarchOpts m; RaySource src;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    rc = check_opts(opt, m); if (rc) return rc;
    rc = make_source(origins, dirs, vdirs, Q, nullptr, opt, src); if (rc) return rc;
    SVOXB_REQUIRE(Q == 0 || out != nullptr, "out is NULL");
    if (Q == 0) return 0;
    if (opt->format != SVOXB_FORMAT_RGBA) {
        SVOXB_REQUIRE(vdirs != nullptr, "view-dependent formats need vdirs");
        SVOXB_REQUIRE(depth == nullptr, "fused depth is only available for the RGBA format; call svoxb_render_depth");
    }
    // short batches: the march also writes every ray's iteration count, which the backward over the same batch is
    // ordered by (svoxb_order.cu); the forward itself runs in the caller's order
    cudaStream_t st = (cudaStream_t)stream;
    if (ray_cost != nullptr && want_ray_order(tr, Q)) {
        // ray_cost[0] = -1 until a kernel that counts overwrites it: the backward then keeps the caller's order
        SVOXB_CUDA(cudaMemsetAsync(ray_cost, 0xff, sizeof(int32_t), st));
        src.steps_out = ray_cost;
    }
    if (opt->format != SVOXB_FORMAT_RGBA) rc = fmt_render_fwd(tree, tr, src, m, opt, false, out, st);
    else rc = dispatch_fwd<false>(tr, src, m, out, depth, st);
    return rc;
}

extern "C" int svoxb_render_rays_fwd(const svoxb_tree* tree, const float* origins, const float* dirs,
                                     const float* vdirs, int64_t Q, const svoxb_render_options* opt, float* out,
                                     float* depth, void* stream) {
    return render_rays_fwd_impl(tree, origins, dirs, vdirs, Q, opt, out, depth, nullptr, stream);
}

extern "C" int svoxb_render_rays_fwd_cost(const svoxb_tree* tree, const float* origins, const float* dirs,
                                          const float* vdirs, int64_t Q, const svoxb_render_options* opt, float* out,
                                          float* depth, int32_t* ray_cost, void* stream) {
    return render_rays_fwd_impl(tree, origins, dirs, vdirs, Q, opt, out, depth, ray_cost, stream);
}

extern "C" int64_t svoxb_ray_order_max_rays(void) { return ray_order_max_rays(); }
extern "C" int64_t svoxb_ray_order_min_rays(void) { return ray_order_min_rays(); }

static int render_rays_bwd_impl(const svoxb_tree* tree, const float* origins, const float* dirs,
                                const float* vdirs, int64_t Q, const svoxb_render_options* opt,
                                const float* grad_out, const float* saved_out, float* grad_features,
                                const int32_t* ray_cost, void* stream) {
    TreeArgs tr; MarchOpts m; RaySource src;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    rc = check_opts(opt, m); if (rc) return rc;
    rc = make_source(origins, dirs, vdirs, Q, nullptr, opt, src); if (rc) return rc;
    SVOXB_REQUIRE(Q == 0 || (grad_out && saved_out && grad_features), "grad_out/saved_out/grad_features NULL");
    if (Q == 0) return 0;
    if (opt->format != SVOXB_FORMAT_RGBA) {
        SVOXB_REQUIRE(vdirs != nullptr, "view-dependent formats need vdirs");
    }
    cudaStream_t st = (cudaStream_t)stream;
    int* order = nullptr;
    if (ray_cost != nullptr && want_ray_order(tr, Q)) {
        rc = build_ray_order(ray_cost, Q, &order, st);
        if (rc) return rc;
    }
    src.order = order;
    if (opt->format != SVOXB_FORMAT_RGBA) rc = fmt_render_bwd(tree, tr, src, m, opt, false, grad_out, saved_out, grad_features, st);
    else rc = dispatch_bwd<false>(tr, src, m, grad_out, saved_out, grad_features, st);
    if (order) cudaFreeAsync(order, st);
    return rc;
}

extern "C" int svoxb_render_rays_bwd(const svoxb_tree* tree, const float* origins, const float* dirs,
                                     const float* vdirs, int64_t Q, const svoxb_render_options* opt,
                                     const float* grad_out, const float* saved_out, float* grad_features, void* stream) {
    return render_rays_bwd_impl(tree, origins, dirs, vdirs, Q, opt, grad_out, saved_out, grad_features, nullptr, stream);
}

extern "C" int svoxb_render_rays_bwd_cost(const svoxb_tree* tree, const float* origins, const float* dirs,
                                          const float* vdirs, int64_t Q, const svoxb_render_options* opt,
                                          const float* grad_out, const float* saved_out, float* grad_features,
                                          const int32_t* ray_cost, void* stream) {
    return render_rays_bwd_impl(tree, origins, dirs, vdirs, Q, opt, grad_out, saved_out, grad_features, ray_cost, stream);
}

extern "C" int svoxb_render_image_fwd(const svoxb_tree* tree, const svoxb_camera* cam,
                                      const svoxb_render_options* opt, float* out, float* depth, void* stream) {
    TreeArgs tr; MarchOpts m; RaySource src;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    rc = check_opts(opt, m); if (rc) return rc;
    SVOXB_REQUIRE(cam != nullptr && out != nullptr, "camera/out NULL");
    rc = make_source(nullptr, nullptr, nullptr, 0, cam, opt, src); if (rc) return rc;
    if (opt->format != SVOXB_FORMAT_RGBA) {
        SVOXB_REQUIRE(depth == nullptr, "fused depth is only available for the RGBA format");
        return fmt_render_fwd(tree, tr, src, m, opt, true, out, (cudaStream_t)stream);
    }
    return dispatch_fwd<true>(tr, src, m, out, depth, (cudaStream_t)stream);
}

extern "C" int svoxb_render_image_bwd(const svoxb_tree* tree, const svoxb_camera* cam,
                                      const svoxb_render_options* opt, const float* grad_out,
                                      const float* saved_out, float* grad_features, void* stream) {
    TreeArgs tr; MarchOpts m; RaySource src;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    rc = check_opts(opt, m); if (rc) return rc;
    SVOXB_REQUIRE(cam && grad_out && saved_out && grad_features, "camera/grad_out/saved_out/grad_features NULL");
    rc = make_source(nullptr, nullptr, nullptr, 0, cam, opt, src); if (rc) return rc;
    if (opt->format != SVOXB_FORMAT_RGBA)
        return fmt_render_bwd(tree, tr, src, m, opt, true, grad_out, saved_out, grad_features, (cudaStream_t)stream);
    return dispatch_bwd<true>(tr, src, m, grad_out, saved_out, grad_features, (cudaStream_t)stream);
}

extern "C" int svoxb_render_depth(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                  const svoxb_render_options* opt, float* depth, void* stream) {
    TreeArgs tr; MarchOpts m;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    rc = check_opts(opt, m); if (rc) return rc;
    if (m.sigma_thresh < 0.0f) tr.acc_miss_mask = 0;
    SVOXB_REQUIRE(Q >= 0 && (Q == 0 || (origins && dirs && depth)), "bad ray batch");
    if (Q == 0) return 0;
    const int grid = (int)min((Q + BLOCK - 1) / BLOCK, (int64_t)sm_count() * 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (tr.use_accel) {
        const size_t smem = sizeof(uint32_t) << (3 * tr.acc.bits[0]);
        if (smem > 48 * 1024)
            SVOXB_CUDA(cudaFuncSetAttribute(depth_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        depth_kernel<true><<<grid, BLOCK, smem, st>>>(tr, origins, dirs, Q, m, depth);
    } else {
        depth_kernel<false><<<grid, BLOCK, 0, st>>>(tr, origins, dirs, Q, m, depth);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "depth_kernel launch");
}
