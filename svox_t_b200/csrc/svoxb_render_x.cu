// svoxb_render_x.cu -- march variants next to the feature render: opacity-only render with a
// CORRECT backward, and the first-hit "motion" outputs. One thread per ray: these touch one float per sample
// (sigma), so there is no row work to share across a warp. sigma comes from the compact array when the caller attached
// one (svoxb_gather_sigma), and rows the hit marks flag as dead are not fetched at all (leaf_sigma, svoxb_march.cuh).
//
// Replaces (reference paths relative to /root/reference/svox_t/csrc):
//   opacity_trace_ray / opacity_render                 rt_kernel.cu:499-560, 1109-1126, 1574-1591
//   opacity_trace_ray_backward                         rt_kernel.cu:562-651  (never instantiated in the reference:
//                                                      its host wrapper launches render_ray_backward_kernel, :1607)
//   motion_trace_ray / motion_render                   rt_kernel.cu:698-778, 836-862, 1480-1504
#include <stdlib.h>
#include "svoxb_march.cuh"

namespace svoxb {

template <bool ACCEL>
__device__ __forceinline__ void sample_sigma(const TreeArgs& tr, const uint32_t* top, const Ray& r, float step,
                                             Leaf& lf, float& delta_t, float& sigma) {
    const float px = fmaf(r.t, r.dx, r.ox), py = fmaf(r.t, r.dy, r.oy), pz = fmaf(r.t, r.dz, r.oz);
    lf = locate<ACCEL>(tr, top, px, py, pz);
    float smin, smax;
    dda_unit(lf.rx, lf.ry, lf.rz, r.ix, r.iy, r.iz, smin, smax);
    const float tsub = ACCEL ? (smax - smin) * lf.inv_cube : (smax - smin) / lf.cube;
    delta_t = tsub + step;
    sigma = leaf_sigma(tr, lf);
}

__device__ __forceinline__ void load_ray(const TreeArgs& tr, const float* origins, const float* dirs, int64_t id, Ray& ray) {
    ray_setup(tr.offset, tr.scaling, __ldg(origins + 3 * id), __ldg(origins + 3 * id + 1), __ldg(origins + 3 * id + 2),
              __ldg(dirs + 3 * id), __ldg(dirs + 3 * id + 1), __ldg(dirs + 3 * id + 2), ray);
}

template <bool ACCEL>
__global__ void __launch_bounds__(BLOCK)
opacity_fwd_kernel(TreeArgs tr, const float* __restrict__ origins, const float* __restrict__ dirs, int64_t Q,
                   MarchOpts opt, float* __restrict__ out) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < Q; id += (int64_t)gridDim.x * blockDim.x) {
        Ray ray;
        load_ray(tr, origins, dirs, id, ray);
        float T = 1.0f;
        while (ray.t < ray.tmax) {
            Leaf lf; float delta_t, sigma;
            sample_sigma<ACCEL>(tr, top, ray, opt.step, lf, delta_t, sigma);
            if (sigma > opt.sigma_thresh) {                                  // rt_kernel.cu:547-555
                T *= expf(-delta_t * ray.ds * sigma);
                if (T <= opt.stop_thresh) break;
            }
            ray.t += delta_t;
        }
        out[id] = 1.0f - T;
    }
}

// The same march with persistent warps: a lane whose ray has ended takes the next ray from the global queue instead of
// idling until the longest ray of its warp is done (rays of one warp differ by 235 against 157 samples on average in the
// reference's headline scene). STEPS samples between two looks at the queue keep the bookkeeping off the sample loop.
// BWD: the scatter pass of the backward, T_end from the forward's saved output (svoxb_opacity_render_bwd_saved).
template <bool ACCEL, bool BWD>
__global__ void __launch_bounds__(BLOCK)
opacity_queue_kernel(TreeArgs tr, RaySource src, MarchOpts opt, float* __restrict__ out,
                     const float* __restrict__ grad_out, const float* __restrict__ saved_out, float* __restrict__ grad,
                     unsigned long long* counter) {
    constexpr int STEPS = 4;
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31;
    Ray ray;
    float T = 1.0f;                 // forward: transmittance so far; backward: grad_out * T_end
    int row = 0;
    bool active = false;
    Queue q{0, 0, false};
    unsigned need = FULL;
    while (true) {
        if (need) {
            const unsigned got = refill<false>(src, tr.offset, tr.scaling, counter, q, need, lane, ray, row);
            if ((got >> lane) & 1u) {
                active = true;
                T = BWD ? __ldg(grad_out + row) * (1.0f - __ldg(saved_out + row)) : 1.0f;
            }
            need = 0;
        }
        if (__ballot_sync(FULL, active) == 0u) break;
        bool fin = false;
        if (active) {
#pragma unroll 1
            for (int k = 0; k < STEPS; ++k) {
                if (!(ray.t < ray.tmax)) { fin = true; break; }
                Leaf lf; float delta_t, sigma;
                sample_sigma<ACCEL>(tr, top, ray, opt.step, lf, delta_t, sigma);
                if (BWD) {
                    if (sigma > 0.0f) atomicAdd(grad + lf.idx * tr.D + (tr.D - 1), delta_t * ray.ds * T);   // rt_kernel.cu:610-646
                } else if (sigma > opt.sigma_thresh) {                                  // rt_kernel.cu:547-555
                    T *= expf(-delta_t * ray.ds * sigma);
                    if (T <= opt.stop_thresh) { fin = true; break; }
                }
                ray.t += delta_t;
            }
            if (!fin && !(ray.t < ray.tmax)) fin = true;
            if (fin) {
                if (!BWD) out[row] = 1.0f - T;
                active = false;
            }
        }
        need = __ballot_sync(FULL, fin);
    }
}

// d(1 - T_end)/d sigma_i = delta_i * delta_scale * T_end for every sample with sigma_i > 0 (rt_kernel.cu:610-646).
template <bool ACCEL>
__global__ void __launch_bounds__(BLOCK)
opacity_bwd_kernel(TreeArgs tr, const float* __restrict__ origins, const float* __restrict__ dirs, int64_t Q,
                   MarchOpts opt, const float* __restrict__ grad_out, const float* __restrict__ saved_out,
                   float* __restrict__ grad) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < Q; id += (int64_t)gridDim.x * blockDim.x) {
        Ray ray;
        load_ray(tr, origins, dirs, id, ray);
        const float t0 = ray.t;
        float T = 1.0f;
        if (saved_out) {                                                     // T_end from the forward's own output
            T = 1.0f - __ldg(saved_out + id);
        } else {
            while (ray.t < ray.tmax) {                                       // pass 1: T_end
                Leaf lf; float delta_t, sigma;
                sample_sigma<ACCEL>(tr, top, ray, opt.step, lf, delta_t, sigma);
                if (sigma > 0.0f) T *= expf(-delta_t * sigma * ray.ds);
                ray.t += delta_t;
            }
        }
        const float gT = __ldg(grad_out + id) * T;
        ray.t = t0;
        while (ray.t < ray.tmax) {                                           // pass 2: scatter
            Leaf lf; float delta_t, sigma;
            sample_sigma<ACCEL>(tr, top, ray, opt.step, lf, delta_t, sigma);
            if (sigma > 0.0f) atomicAdd(grad + lf.idx * tr.D + (tr.D - 1), delta_t * ray.ds * gT);
            ray.t += delta_t;
        }
    }
}

// First hit: distances from the hit point to the J rows of extra_data, depth, hit point, data index.
// NOTE the hit point is what the reference computes (rt_kernel.cu:757-761): transform_coord_world applied to the
// IN-LEAF RELATIVE coordinates the descent leaves in `pos`, not to the sample position. Kept for parity.
template <bool ACCEL>
__global__ void __launch_bounds__(BLOCK)
motion_kernel(TreeArgs tr, const float* __restrict__ origins, const float* __restrict__ dirs, int64_t Q, MarchOpts opt,
              const float* __restrict__ extra, int J, float* __restrict__ out, float* __restrict__ depth,
              float* __restrict__ hit_point, int64_t* __restrict__ data_idx) {
    extern __shared__ __align__(128) uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    const float o0 = __ldg(tr.offset), o1 = __ldg(tr.offset + 1), o2 = __ldg(tr.offset + 2);
    const float s0 = __ldg(tr.scaling), s1 = __ldg(tr.scaling + 1), s2 = __ldg(tr.scaling + 2);
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < Q; id += (int64_t)gridDim.x * blockDim.x) {
        Ray ray;
        load_ray(tr, origins, dirs, id, ray);
        float d = 0.0f, hx = 0.0f, hy = 0.0f, hz = 0.0f;
        int64_t di = 0;
        bool hit = false;
        while (ray.t < ray.tmax) {
            Leaf lf; float delta_t, sigma;
            sample_sigma<ACCEL>(tr, top, ray, opt.step, lf, delta_t, sigma);
            if (sigma > opt.sigma_thresh) {
                hx = (lf.rx - o0) / s0; hy = (lf.ry - o1) / s1; hz = (lf.rz - o2) / s2;   // common.cuh:53-60
                d = ray.t * ray.ds;
                di = lf.idx;
                hit = true;
                break;
            }
            ray.t += delta_t;
        }
        for (int j = 0; j < J; ++j) {
            float v = 0.0f;
            if (hit) {
                const float a = hx - __ldg(extra + 3 * j), b = hy - __ldg(extra + 3 * j + 1), c = hz - __ldg(extra + 3 * j + 2);
                v = sqrtf(a * a + b * b + c * c);
            }
            out[id * J + j] = v;
        }
        depth[id] = d;
        hit_point[3 * id] = hx; hit_point[3 * id + 1] = hy; hit_point[3 * id + 2] = hz;
        data_idx[id] = di;
    }
}

// Per-leaf accumulation of the compositing weights (rt_kernel.cu:266-267, 308-310): weight_accum[slot] += T (1 - att) at
// every hit, slot = the leaf's packed id node*N^3 + u*N^2 + v*N + w. The packed accelerator does not carry slot ids,
// so this walks the reference tensors (any N); it runs only inside a `with tree.accumulate_weights()` block, next to
// the normal render. float atomics (the reference's plain += loses updates when rays collide).
template <bool IMAGE>
__global__ void __launch_bounds__(BLOCK)
weight_accum_kernel(TreeArgs tr, RaySource src, MarchOpts opt, float* __restrict__ weight_accum) {
    const int64_t total = IMAGE ? (int64_t)src.width * (src.row_end - src.row_begin) : src.total;
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
        float ox, oy, oz, dx, dy, dz;
        if (IMAGE) {
            camera_ray(src, (int)(id % src.width), src.row_begin + (int)(id / src.width), ox, oy, oz, dx, dy, dz);
            if (src.ndc_w >= 0) world2ndc(src, ox, oy, oz, dx, dy, dz);
        } else {
            ox = __ldg(src.origins + 3 * id); oy = __ldg(src.origins + 3 * id + 1); oz = __ldg(src.origins + 3 * id + 2);
            dx = __ldg(src.dirs + 3 * id); dy = __ldg(src.dirs + 3 * id + 1); dz = __ldg(src.dirs + 3 * id + 2);
        }
        Ray ray;
        ray_setup(tr.offset, tr.scaling, ox, oy, oz, dx, dy, dz, ray);
        float T = 1.0f;
        while (ray.t < ray.tmax) {
            const float px = fmaf(ray.t, ray.dx, ray.ox), py = fmaf(ray.t, ray.dy, ray.oy), pz = fmaf(ray.t, ray.dz, ray.oz);
            float rx, ry, rz, cube, smin, smax;
            const int64_t slot = descend_ref(tr.child, tr.N, px, py, pz, rx, ry, rz, cube);
            const int di = __ldg(tr.data + slot);
            dda_unit(rx, ry, rz, ray.ix, ray.iy, ray.iz, smin, smax);
            const float delta_t = (smax - smin) / cube + opt.step;
            const float sigma = ((int64_t)di < tr.M && di >= 0) ? __ldg(tr.features + (int64_t)di * tr.D + (tr.D - 1)) : 0.0f;
            if (sigma > opt.sigma_thresh) {
                const float att = expf(-delta_t * ray.ds * sigma);
                atomicAdd(weight_accum + slot, T * (1.0f - att));
                T *= att;
                if (T <= opt.stop_thresh) break;
            }
            ray.t += delta_t;
        }
    }
}

int make_tree_args(const svoxb_tree* t, TreeArgs& a, void* use_stream);   // svoxb_tree.cu

static int simple_opts(const svoxb_render_options* opt, MarchOpts& m) {
    SVOXB_REQUIRE(opt != nullptr, "render options are NULL");
    // ndc_* are ignored here exactly as in the reference: only its image kernels convert rays (rt_kernel.cu:1204)
    m.step = opt->step_size; m.bg = opt->background_brightness;
    m.sigma_thresh = opt->sigma_thresh; m.stop_thresh = opt->stop_thresh;
    return 0;
}

template <typename KA, typename KR, typename... Args>
static int launch_simple(const TreeArgs& tr, int64_t Q, cudaStream_t st, KA kacc, KR kref, Args... args) {
    const int grid = (int)min((Q + BLOCK - 1) / BLOCK, (int64_t)sm_count() * 8);
    if (tr.use_accel) {
        const size_t smem = sizeof(uint32_t) << (3 * tr.acc.bits[0]);
        kacc<<<grid, BLOCK, smem, st>>>(tr, args...);
    } else {
        kref<<<grid, BLOCK, 0, st>>>(tr, args...);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "ray kernel launch");
}

// Batches with at least two rays per resident lane take the persistent-queue kernel (SVOXB_OPACITY_QUEUE=0: never).
static int64_t opacity_queue_min_rays() {
    static const int on = getenv("SVOXB_OPACITY_QUEUE") ? atoi(getenv("SVOXB_OPACITY_QUEUE")) : 1;
    return on ? (int64_t)sm_count() * 2048 * 2 : ((int64_t)1 << 40);
}

template <bool BWD>
static int launch_opacity_queue(const TreeArgs& tr, const float* origins, const float* dirs, int64_t Q, const MarchOpts& m,
                                float* out, const float* grad_out, const float* saved_out, float* grad, cudaStream_t st) {
    SVOXB_REQUIRE(Q < (1ll << 31), "ray count out of range");
    RaySource src{};
    src.origins = origins; src.dirs = dirs; src.total = Q; src.ndc_w = -1;
    const size_t smem = tr.use_accel ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0;
    auto kern = tr.use_accel ? opacity_queue_kernel<true, BWD> : opacity_queue_kernel<false, BWD>;
    int grid = 0;
    int rc = persistent_grid(kern, smem, Q, grid);
    if (rc) return rc;
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, BLOCK, smem, st>>>(tr, src, m, out, grad_out, saved_out, grad, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "opacity_queue_kernel launch");
}

}  // namespace svoxb

using namespace svoxb;

extern "C" int svoxb_opacity_render_fwd(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                        const svoxb_render_options* opt, float* out, void* stream) {
    TreeArgs tr; MarchOpts m;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    rc = simple_opts(opt, m); if (rc) return rc;
    if (m.sigma_thresh < 0.0f) tr.acc_miss_mask = 0;     // the marks encode sigma > 0: too strict for this predicate
    SVOXB_REQUIRE(Q >= 0 && (Q == 0 || (origins && dirs && out)), "bad ray batch");
    if (Q == 0) return 0;
    if (Q >= opacity_queue_min_rays())
        return launch_opacity_queue<false>(tr, origins, dirs, Q, m, out, nullptr, nullptr, nullptr, (cudaStream_t)stream);
    return launch_simple(tr, Q, (cudaStream_t)stream, opacity_fwd_kernel<true>, opacity_fwd_kernel<false>, origins, dirs,
                         Q, m, out);
}

static int opacity_bwd_impl(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                            const svoxb_render_options* opt, const float* grad_out, const float* saved_out,
                            float* grad_features, void* stream) {
    TreeArgs tr; MarchOpts m;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    rc = simple_opts(opt, m); if (rc) return rc;
    if (m.sigma_thresh < 0.0f) tr.acc_miss_mask = 0;     // the marks encode sigma > 0: too strict for this predicate
    SVOXB_REQUIRE(Q >= 0 && (Q == 0 || (origins && dirs && grad_out && grad_features)), "bad arguments");
    if (Q == 0) return 0;
    if (saved_out != nullptr && Q >= opacity_queue_min_rays())
        return launch_opacity_queue<true>(tr, origins, dirs, Q, m, nullptr, grad_out, saved_out, grad_features,
                                          (cudaStream_t)stream);
    return launch_simple(tr, Q, (cudaStream_t)stream, opacity_bwd_kernel<true>, opacity_bwd_kernel<false>, origins, dirs,
                         Q, m, grad_out, saved_out, grad_features);
}

extern "C" int svoxb_opacity_render_bwd(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                        const svoxb_render_options* opt, const float* grad_out, float* grad_features,
                                        void* stream) {
    return opacity_bwd_impl(tree, origins, dirs, Q, opt, grad_out, nullptr, grad_features, stream);
}

extern "C" int svoxb_opacity_render_bwd_saved(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                              const svoxb_render_options* opt, const float* grad_out,
                                              const float* saved_out, float* grad_features, void* stream) {
    SVOXB_REQUIRE(Q == 0 || saved_out != nullptr, "saved_out is NULL");
    return opacity_bwd_impl(tree, origins, dirs, Q, opt, grad_out, saved_out, grad_features, stream);
}

extern "C" int svoxb_motion_render(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                   const svoxb_render_options* opt, const float* extra_data, int32_t J, float* out,
                                   float* depth, float* hit_point, int64_t* data_idx, void* stream) {
    TreeArgs tr; MarchOpts m;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    rc = simple_opts(opt, m); if (rc) return rc;
    if (m.sigma_thresh < 0.0f) tr.acc_miss_mask = 0;     // the marks encode sigma > 0: too strict for this predicate
    SVOXB_REQUIRE(J >= 0 && (J == 0 || extra_data), "extra_data is NULL");
    SVOXB_REQUIRE(Q >= 0 && (Q == 0 || (origins && dirs && depth && hit_point && data_idx && (J == 0 || out))), "bad arguments");
    if (Q == 0) return 0;
    return launch_simple(tr, Q, (cudaStream_t)stream, motion_kernel<true>, motion_kernel<false>, origins, dirs, Q, m,
                         extra_data, (int)J, out, depth, hit_point, data_idx);
}

extern "C" int svoxb_accumulate_weights(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                        const svoxb_camera* cam, const svoxb_render_options* opt, float* weight_accum,
                                        void* stream) {
    TreeArgs tr; MarchOpts m;
    int rc = make_tree_args(tree, tr, stream); if (rc) return rc;
    rc = simple_opts(opt, m); if (rc) return rc;
    SVOXB_REQUIRE(weight_accum != nullptr, "weight_accum is NULL");
    RaySource src{};
    src.ndc_w = -1;
    int64_t total;
    if (cam) {
        SVOXB_REQUIRE(cam->c2w != nullptr && cam->width > 0 && cam->height > 0, "bad camera spec");
        src.c2w = cam->c2w; src.fx = cam->fx; src.fy = cam->fy; src.width = cam->width; src.height = cam->height;
        src.row_begin = 0; src.row_end = cam->height;
        if (cam->row_end > 0) {
            SVOXB_REQUIRE(cam->row_begin >= 0 && cam->row_begin < cam->row_end && cam->row_end <= cam->height, "bad image band");
            src.row_begin = cam->row_begin; src.row_end = cam->row_end;
        }
        if (opt->ndc_width >= 0) { src.ndc_w = opt->ndc_width; src.ndc_h = opt->ndc_height; src.ndc_focal = opt->ndc_focal; }
        total = (int64_t)cam->width * (src.row_end - src.row_begin);
    } else {
        SVOXB_REQUIRE(Q >= 0 && (Q == 0 || (origins && dirs)), "bad ray batch");
        src.origins = origins; src.dirs = dirs; src.total = Q;
        total = Q;
    }
    if (total == 0) return 0;
    const int grid = (int)min((total + BLOCK - 1) / BLOCK, (int64_t)sm_count() * 8);
    if (cam) weight_accum_kernel<true><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(tr, src, m, weight_accum);
    else weight_accum_kernel<false><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(tr, src, m, weight_accum);
    count_launch();
    return check_cuda(cudaGetLastError(), "weight_accum_kernel launch");
}
