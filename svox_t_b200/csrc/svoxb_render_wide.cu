// svoxb_render_wide.cu -- the GENERAL instantiation of the march: any feature width, float32 or float64.
//
// The tuned kernels (svoxb_render_q.cu, svoxb_render.cu) keep a ray's partial output in registers / shared memory and
// stop at D = 128, in float32. The reference has neither limit: it dispatches AT_DISPATCH_FLOATING_TYPES on every entry
// point (rt_kernel.cu:1373, 1413, 1517; svox_kernel.cu:290) and loops over any out_data_dim (rt_kernel.cu:302-306).
// This file closes both gaps with one family of kernels templated on the arithmetic type R:
//   R = float  : svoxb_render_rays_* / svoxb_render_image_* route here for D > 128 (svoxb_render.cu dispatch);
//   R = double : the *_f64 entry points of include/svoxb.h (RGBA format; the reference's own double build evaluates
//                exp() in float -- expf(double) -- so it is only float-accurate; these kernels use true fp64).
// Replaces (paths relative to /root/reference/svox_t/csrc): trace_ray rt_kernel.cu:221-328, trace_ray_backward :330-496,
// depth_trace_ray :781-834, cam2world_ray :1152-1166, query_single_from_root include/common.cuh:62-100,
// query_single_kernel svox_kernel.cu:36-81.
//
// Design: every LANE owns one ray for the traversal (descent over the reference's child/data tensors, any N); the row
// work of a hit is done by the whole warp, lane <-> channel (+32k), so row reads, output updates and gradient
// reductions are coalesced. The partial output of a ray lives in its own row of `out` (zeroed, accumulated into and
// finalised by the same lane per channel, so plain loads/stores are ordered) -- that is what makes the width unbounded.
#include <type_traits>
#include "svoxb_march.cuh"

namespace svoxb {

template <typename R>
struct GTree {
    const R* features;
    int64_t M;
    int D, N;
    const int32_t* child;
    const int32_t* data;
    const R* offset;
    const R* scaling;
};

template <typename R>
struct GSource {
    const R* origins;
    const R* dirs;
    const R* c2w;
    R fx, fy;
    int width, height, row_begin, row_end;
    int64_t total;          // rays, or pixels of the band
};

template <typename R>
struct GOpts {
    R step, bg, sigma_thresh, stop_thresh;
};

template <typename R>
struct GRay {
    R ox, oy, oz, dx, dy, dz, ix, iy, iz, t, tmax, ds;
};

__device__ __forceinline__ float g_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double g_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float g_exp(float x) { return expf(x); }
__device__ __forceinline__ double g_exp(double x) { return exp(x); }
__device__ __forceinline__ float g_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double g_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float g_floor(float x) { return floorf(x); }
__device__ __forceinline__ double g_floor(double x) { return floor(x); }
__device__ __forceinline__ float g_sigmoid(float x) { return fast_sigmoid(x); }                 // as the float kernels
__device__ __forceinline__ double g_sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }          // rt_kernel.cu:304

// include/common.cuh:37-42
template <typename R>
__device__ __forceinline__ R g_clamp01(R q) {
    const R hi = (R)(1.0 - 1e-6);
    return max((R)0, min(hi, q));
}

// rt_kernel.cu:201-218
template <typename R>
__device__ __forceinline__ void g_dda(R cx, R cy, R cz, R ix, R iy, R iz, R& tmin, R& tmax) {
    R t1, t2;
    tmin = (R)0; tmax = (R)1e9f;
    t1 = -cx * ix; t2 = t1 + ix; tmin = max(tmin, min(t1, t2)); tmax = min(tmax, max(t1, t2));
    t1 = -cy * iy; t2 = t1 + iy; tmin = max(tmin, min(t1, t2)); tmax = min(tmax, max(t1, t2));
    t1 = -cz * iz; t2 = t1 + iz; tmin = max(tmin, min(t1, t2)); tmax = min(tmax, max(t1, t2));
}

// include/common.cuh:62-100 -- descent over child, any N; p unclamped. Returns the packed leaf slot.
template <typename R>
__device__ __forceinline__ int64_t g_descend(const int32_t* __restrict__ child, int N, R px, R py, R pz,
                                             R& rx, R& ry, R& rz, R& cube) {
    const R fN = (R)N;
    const int N2 = N * N;
    const int64_t N3 = (int64_t)N2 * N;
    px = g_clamp01(px); py = g_clamp01(py); pz = g_clamp01(pz);
    int64_t node = 0;
    cube = fN;
    while (true) {
        px *= fN; py *= fN; pz *= fN;
        const R fu = g_floor(px), fv = g_floor(py), fw = g_floor(pz);
        px -= fu; py -= fv; pz -= fw;
        const int64_t slot = node * N3 + (int)fu * N2 + (int)fv * N + (int)fw;
        const int skip = __ldg(child + slot);
        if (skip == 0) { rx = px; ry = py; rz = pz; return slot; }
        cube *= fN;
        node += skip;
    }
}

// rt_kernel.cu:663-665 (transform_coord) + 187-199 (delta scale) + 227-247 (invdir through double, slab test)
template <typename R>
__device__ __forceinline__ void g_ray_setup(const GTree<R>& tr, R owx, R owy, R owz, R dwx, R dwy, R dwz, GRay<R>& r) {
    const R s0 = __ldg(tr.scaling), s1 = __ldg(tr.scaling + 1), s2 = __ldg(tr.scaling + 2);
    r.ox = g_fma(s0, owx, __ldg(tr.offset));
    r.oy = g_fma(s1, owy, __ldg(tr.offset + 1));
    r.oz = g_fma(s2, owz, __ldg(tr.offset + 2));
    R dx = dwx * s0, dy = dwy * s1, dz = dwz * s2;
    const R nrm = g_sqrt(dx * dx + dy * dy + dz * dz);
    r.ds = (R)1 / nrm;
    dx *= r.ds; dy *= r.ds; dz *= r.ds;
    r.dx = dx; r.dy = dy; r.dz = dz;
    r.ix = (R)(1.0 / ((double)dx + 1e-9));
    r.iy = (R)(1.0 / ((double)dy + 1e-9));
    r.iz = (R)(1.0 / ((double)dz + 1e-9));
    R tmin, tmax;
    g_dda(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, tmin, tmax);
    if (tmax < (R)0 || tmin > tmax) { tmin = (R)0; tmax = (R)0; }    // misses the cube: no samples, T stays 1
    r.t = tmin; r.tmax = tmax;
}

// rt_kernel.cu:1152-1166
template <typename R>
__device__ __forceinline__ void g_camera_ray(const GSource<R>& s, int px, int py, R& ox, R& oy, R& oz, R& dx, R& dy, R& dz) {
    R x = (R)(((double)px - 0.5 * (double)s.width) / (double)s.fx);
    R y = (R)(-((double)py - 0.5 * (double)s.height) / (double)s.fy);
    R z = g_sqrt((R)((double)(x * x + y * y) + 1.0));
    x /= z; y /= z; z = (R)-1 / z;
    const R* c = s.c2w;
    dx = __ldg(c + 0) * x + __ldg(c + 1) * y + __ldg(c + 2) * z;
    dy = __ldg(c + 4) * x + __ldg(c + 5) * y + __ldg(c + 6) * z;
    dz = __ldg(c + 8) * x + __ldg(c + 9) * y + __ldg(c + 10) * z;
    ox = __ldg(c + 3); oy = __ldg(c + 7); oz = __ldg(c + 11);
}

// The lane's ray of this round: ray `id` of the batch, or pixel `id` of the band (row-major). False past the end.
template <typename R, bool IMAGE>
__device__ __forceinline__ bool g_fetch(const GTree<R>& tr, const GSource<R>& src, int64_t id, GRay<R>& ray) {
    if (id >= src.total) return false;
    R ox, oy, oz, dx, dy, dz;
    if (IMAGE) {
        const int py = src.row_begin + (int)(id / src.width), px = (int)(id % src.width);
        g_camera_ray(src, px, py, ox, oy, oz, dx, dy, dz);
    } else {
        const R* o = src.origins + id * 3;
        const R* d = src.dirs + id * 3;
        ox = __ldg(o); oy = __ldg(o + 1); oz = __ldg(o + 2);
        dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
    }
    g_ray_setup(tr, ox, oy, oz, dx, dy, dz, ray);
    return true;
}

// One sample (rt_kernel.cu:261-277): leaf row (or -1 for an empty leaf), delta_t, sigma.
template <typename R>
__device__ __forceinline__ void g_sample(const GTree<R>& tr, const GRay<R>& r, R step, int& idx, R& delta_t, R& sigma) {
    const R px = g_fma(r.t, r.dx, r.ox), py = g_fma(r.t, r.dy, r.oy), pz = g_fma(r.t, r.dz, r.oz);
    R rx, ry, rz, cube, smin, smax;
    const int64_t slot = g_descend(tr.child, tr.N, px, py, pz, rx, ry, rz, cube);
    SVOXB_DBG(slot >= 0);
    const int di = __ldg(tr.data + slot);
    idx = ((int64_t)di >= tr.M || di < 0) ? -1 : di;
    g_dda(rx, ry, rz, r.ix, r.iy, r.iz, smin, smax);
    delta_t = (smax - smin) / cube + step;
    sigma = idx >= 0 ? __ldg(tr.features + (int64_t)idx * tr.D + (tr.D - 1)) : (R)0;
}

template <typename R>
__device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(FULL, v, s);
    return v;
}

constexpr int WIDE_BLOCK = 256;

// ------------------------------------------------------------------------------------------------------------
// Forward: out[row, 0..D-2] = sum_i w_i sigmoid(f_i) + T bg ; out[row, D-1] = 1 - T ; depth[row] = first hit.
template <typename R, bool IMAGE>
__global__ void __launch_bounds__(WIDE_BLOCK)
wide_fwd_kernel(GTree<R> tr, GSource<R> src, GOpts<R> opt, R* out, R* depth) {
    const int lane = threadIdx.x & 31;
    const int D = tr.D;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t base = warp * 32; base < src.total; base += nwarps * 32) {
        GRay<R> ray;
        bool active = g_fetch<R, IMAGE>(tr, src, base + lane, ray);
        const int n_rows = (int)min((int64_t)32, src.total - base);
        for (int r = 0; r < n_rows; ++r) {
            R* o = out + (base + r) * D;
            for (int c = lane; c < D; c += 32) o[c] = (R)0;
        }
        R T = (R)1, depth_v = (R)0;
        bool got_depth = false;
        while (__ballot_sync(FULL, active)) {
            bool hit = false;
            R w = (R)0;
            int hidx = 0, fin = 0;            // fin: 1 = the ray left the volume, 2 = stopped early (T <= stop_thresh)
            if (active) {
                if (!(ray.t < ray.tmax)) {
                    fin = 1;
                } else {
                    int idx; R delta_t, sigma;
                    g_sample(tr, ray, opt.step, idx, delta_t, sigma);
                    if (idx >= 0 && sigma > opt.sigma_thresh) {                      // rt_kernel.cu:279-320
                        const R att = g_exp(-delta_t * ray.ds * sigma);
                        w = T * ((R)1 - att);
                        hit = true; hidx = idx;
                        if (!got_depth) { depth_v = ray.ds * ray.t; got_depth = true; }   // rt_kernel.cu:826-830
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
                const R w_r = __shfl_sync(FULL, w, r);
                SVOXB_DBG(idx_r >= 0 && (int64_t)idx_r < tr.M && base + r < src.total);
                const R* f = tr.features + (int64_t)idx_r * D;
                R* o = out + (base + r) * D;
                for (int c = lane; c < D - 1; c += 32) o[c] = g_fma(w_r, g_sigmoid(__ldg(f + c)), o[c]);
            }
            unsigned fm = __ballot_sync(FULL, fin != 0);
            if (fin != 0) {
                if (depth) depth[base + lane] = depth_v;
                active = false;
            }
            while (fm) {
                const int r = __ffs(fm) - 1;
                fm &= fm - 1;
                const R T_r = __shfl_sync(FULL, T, r);
                const int fin_r = __shfl_sync(FULL, fin, r);
                R* o = out + (base + r) * D;
                const R scale = (R)(1.0 / (1.0 - (double)T_r));                      // rt_kernel.cu:315
                for (int c = lane; c < D; c += 32) {
                    if (c == D - 1) o[c] = (R)1 - T_r;
                    else o[c] = (fin_r == 2) ? o[c] * scale : o[c] + T_r * opt.bg;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Backward, one re-march (see svoxb_render.cu): accum starts at sum_{j<D-1} g_j out_j, T_end = 1 - out[D-1]. Per hit:
// grad[idx, j] += w s_j (1 - s_j) g_j ; c = sum_j s_j g_j ; T *= att ; accum -= w c ;
// grad[idx, D-1] += dd (c T - accum) + dd g_{D-1} T_end, dd = delta_t delta_scale (rt_kernel.cu:479-490).
template <typename R, bool IMAGE>
__global__ void __launch_bounds__(WIDE_BLOCK)
wide_bwd_kernel(GTree<R> tr, GSource<R> src, GOpts<R> opt, const R* __restrict__ grad_out,
                const R* __restrict__ saved_out, R* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const int D = tr.D;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t base = warp * 32; base < src.total; base += nwarps * 32) {
        GRay<R> ray;
        bool active = g_fetch<R, IMAGE>(tr, src, base + lane, ray);
        const int n_rows = (int)min((int64_t)32, src.total - base);
        R accum = (R)0, T_end = (R)0, gop = (R)0, T = (R)1;
        for (int r = 0; r < n_rows; ++r) {
            const R* g = grad_out + (base + r) * D;
            const R* so = saved_out + (base + r) * D;
            R part = (R)0;
            for (int c = lane; c < D - 1; c += 32) part = g_fma(__ldg(g + c), __ldg(so + c), part);
            part = warp_sum(part);
            if (lane == r) { accum = part; T_end = (R)1 - __ldg(so + D - 1); gop = __ldg(g + D - 1); }
        }
        while (__ballot_sync(FULL, active)) {
            bool hit = false;
            R w = (R)0, dd = (R)0;
            int hidx = 0;
            if (active) {
                if (!(ray.t < ray.tmax)) {
                    active = false;
                } else {
                    int idx; R delta_t, sigma;
                    g_sample(tr, ray, opt.step, idx, delta_t, sigma);
                    if (sigma > (R)0) {                                             // rt_kernel.cu:382,456
                        const R att = g_exp(-delta_t * sigma * ray.ds);
                        w = T * ((R)1 - att);
                        dd = delta_t * ray.ds;
                        hit = true; hidx = idx;
                        T *= att;
                    }
                    ray.t += delta_t;
                    if (!(ray.t < ray.tmax)) active = false;
                }
            }
            unsigned hm = __ballot_sync(FULL, hit);
            while (hm) {
                const int r = __ffs(hm) - 1;
                hm &= hm - 1;
                const int idx_r = __shfl_sync(FULL, hidx, r);
                const R w_r = __shfl_sync(FULL, w, r);
                const R* f = tr.features + (int64_t)idx_r * D;
                SVOXB_DBG(idx_r >= 0 && (int64_t)idx_r < tr.M && base + r < src.total);
                const R* g = grad_out + (base + r) * D;
                R* grow = grad + (int64_t)idx_r * D;
                R cp = (R)0;
                for (int c = lane; c < D - 1; c += 32) {
                    const R s = g_sigmoid(__ldg(f + c));
                    const R sg = s * __ldg(g + c);
                    cp += sg;
                    atomicAdd(grow + c, w_r * (sg * ((R)1 - s)));
                }
                cp = warp_sum(cp);
                if (lane == r) {
                    accum -= w * cp;
                    atomicAdd(grow + (D - 1), dd * (cp * T - accum) + dd * gop * T_end);
                }
            }
        }
    }
}

// Depth only (rt_kernel.cu:781-834): thread per ray, stops at the first sample with sigma > sigma_thresh.
template <typename R>
__global__ void __launch_bounds__(WIDE_BLOCK)
wide_depth_kernel(GTree<R> tr, GSource<R> src, GOpts<R> opt, R* __restrict__ depth) {
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < src.total; id += (int64_t)gridDim.x * blockDim.x) {
        GRay<R> ray;
        g_fetch<R, false>(tr, src, id, ray);
        R d = (R)0;
        while (ray.t < ray.tmax) {
            int idx; R delta_t, sigma;
            g_sample(tr, ray, opt.step, idx, delta_t, sigma);
            if (sigma > opt.sigma_thresh) { d = ray.ds * ray.t; break; }
            ray.t += delta_t;
        }
        depth[id] = d;
    }
}

// query_single_kernel (svox_kernel.cu:36-81): lane per point, warp-cooperative row copy.
template <typename R>
__global__ void __launch_bounds__(WIDE_BLOCK)
wide_query_kernel(GTree<R> tr, const R* __restrict__ pts, int64_t Q, R* __restrict__ values,
                  int64_t* __restrict__ node_ids, int64_t* __restrict__ data_ids, uint8_t* __restrict__ slot_mask) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t base = warp * 32; base < Q; base += nwarps * 32) {
        const int64_t q = base + lane;
        int idx = -1;
        if (q < Q) {
            const R px = g_fma(__ldg(tr.scaling), __ldg(pts + 3 * q), __ldg(tr.offset));
            const R py = g_fma(__ldg(tr.scaling + 1), __ldg(pts + 3 * q + 1), __ldg(tr.offset + 1));
            const R pz = g_fma(__ldg(tr.scaling + 2), __ldg(pts + 3 * q + 2), __ldg(tr.offset + 2));
            R rx, ry, rz, cube;
            const int64_t slot = g_descend(tr.child, tr.N, px, py, pz, rx, ry, rz, cube);
            node_ids[q] = slot;
            if (slot_mask) slot_mask[slot] = 1;
            const int di = __ldg(tr.data + slot);
            if (di >= 0 && (int64_t)di < tr.M) {
                idx = di;
                if (data_ids) data_ids[q] = di;
            }
        }
        if (values) {
            unsigned vm = __ballot_sync(FULL, idx >= 0);
            while (vm) {
                const int r = __ffs(vm) - 1;
                vm &= vm - 1;
                const int idx_r = __shfl_sync(FULL, idx, r);
                const R* s = tr.features + (int64_t)idx_r * tr.D;
                R* dst = values + (base + r) * tr.D;
                for (int c = lane; c < tr.D; c += 32) dst[c] = __ldg(s + c);
            }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
template <typename R>
static GOpts<R> g_opts(const svoxb_render_options* o) {
    // the option fields are float in the reference as well (data_spec.hpp:129-145): promoted, not re-parsed
    return GOpts<R>{(R)o->step_size, (R)o->background_brightness, (R)o->sigma_thresh, (R)o->stop_thresh};
}

// One ray (point) per thread and round; grid-stride beyond 16 CTAs per SM.
static int g_grid(int64_t total) {
    const int64_t want = (total + WIDE_BLOCK - 1) / WIDE_BLOCK;
    return (int)max((int64_t)1, min(want, (int64_t)sm_count() * 16));
}

template <typename R>
static int g_launch_fwd(const GTree<R>& tr, const GSource<R>& src, const GOpts<R>& m, bool image, R* out, R* depth,
                        cudaStream_t st) {
    const int grid = g_grid(src.total);
    if (image) wide_fwd_kernel<R, true><<<grid, WIDE_BLOCK, 0, st>>>(tr, src, m, out, depth);
    else wide_fwd_kernel<R, false><<<grid, WIDE_BLOCK, 0, st>>>(tr, src, m, out, depth);
    count_launch();
    return check_cuda(cudaGetLastError(), "wide_fwd_kernel launch");
}

template <typename R>
static int g_launch_bwd(const GTree<R>& tr, const GSource<R>& src, const GOpts<R>& m, bool image, const R* go,
                        const R* so, R* grad, cudaStream_t st) {
    const int grid = g_grid(src.total);
    if (image) wide_bwd_kernel<R, true><<<grid, WIDE_BLOCK, 0, st>>>(tr, src, m, go, so, grad);
    else wide_bwd_kernel<R, false><<<grid, WIDE_BLOCK, 0, st>>>(tr, src, m, go, so, grad);
    count_launch();
    return check_cuda(cudaGetLastError(), "wide_bwd_kernel launch");
}

// float32, D > 128: called by the dispatchers of svoxb_render.cu with the source they already validated.
static GSource<float> g_source_f32(const RaySource& s) {
    GSource<float> g{};
    g.origins = s.origins; g.dirs = s.dirs; g.c2w = s.c2w; g.fx = s.fx; g.fy = s.fy;
    g.width = s.width; g.height = s.height; g.row_begin = s.row_begin; g.row_end = s.row_end;
    g.total = s.c2w ? (int64_t)(s.row_end - s.row_begin) * s.width : s.total;
    return g;
}

static GTree<float> g_tree_f32(const TreeArgs& t) {
    return GTree<float>{t.features, t.M, t.D, t.N, t.child, t.data, t.offset, t.scaling};
}

int launch_fwd_wide(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, bool image, float* out, float* depth,
                    cudaStream_t st) {
    SVOXB_REQUIRE(src.ndc_w < 0, "NDC cameras are not available for feature widths above 128");
    const GOpts<float> o{m.step, m.bg, m.sigma_thresh, m.stop_thresh};
    return g_launch_fwd<float>(g_tree_f32(tr), g_source_f32(src), o, image, out, depth, st);
}

int launch_bwd_wide(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, bool image, const float* go,
                    const float* so, float* grad, cudaStream_t st) {
    SVOXB_REQUIRE(src.ndc_w < 0, "NDC cameras are not available for feature widths above 128");
    const GOpts<float> o{m.step, m.bg, m.sigma_thresh, m.stop_thresh};
    return g_launch_bwd<float>(g_tree_f32(tr), g_source_f32(src), o, image, go, so, grad, st);
}

static int g_tree_f64(const svoxb_tree_f64* t, GTree<double>& g) {
    SVOXB_REQUIRE(t != nullptr, "tree is NULL");
    SVOXB_REQUIRE(t->features && t->child && t->data && t->offset && t->scaling, "tree has NULL tensors");
    SVOXB_REQUIRE(t->M > 0 && t->M < (1ll << 31) && t->D >= 2 && t->N >= 2 && t->n_nodes > 0, "bad tree shape");
    g = GTree<double>{t->features, t->M, t->D, t->N, t->child, t->data, t->offset, t->scaling};
    return 0;
}

static int g_check_opts_f64(const svoxb_render_options* opt) {
    SVOXB_REQUIRE(opt != nullptr, "render options are NULL");
    SVOXB_REQUIRE(opt->format == SVOXB_FORMAT_RGBA, "the float64 entry points implement the RGBA format only");
    SVOXB_REQUIRE(opt->ndc_width < 0, "the float64 entry points have no NDC conversion");
    return 0;
}

static int g_rays_f64(const double* origins, const double* dirs, int64_t Q, GSource<double>& s) {
    SVOXB_REQUIRE(Q >= 0 && Q < (1ll << 31), "ray count out of range");
    SVOXB_REQUIRE(Q == 0 || (origins && dirs), "origins/dirs are NULL");
    s = GSource<double>{};
    s.origins = origins; s.dirs = dirs; s.total = Q;
    return 0;
}

static int g_camera_f64(const svoxb_camera_f64* cam, GSource<double>& s) {
    SVOXB_REQUIRE(cam != nullptr && cam->c2w != nullptr && cam->width > 0 && cam->height > 0, "bad camera spec");
    SVOXB_REQUIRE((int64_t)cam->width * cam->height < (1ll << 31), "image too large");
    s = GSource<double>{};
    s.c2w = cam->c2w; s.fx = cam->fx; s.fy = cam->fy; s.width = cam->width; s.height = cam->height;
    s.row_begin = 0; s.row_end = cam->height;
    if (cam->row_end > 0) {
        SVOXB_REQUIRE(cam->row_begin >= 0 && cam->row_begin < cam->row_end && cam->row_end <= cam->height,
                      "bad image band [%d, %d) for height %d", cam->row_begin, cam->row_end, cam->height);
        s.row_begin = cam->row_begin; s.row_end = cam->row_end;
    }
    s.total = (int64_t)(s.row_end - s.row_begin) * s.width;
    return 0;
}

}  // namespace svoxb

using namespace svoxb;

extern "C" int svoxb_query_f64(const svoxb_tree_f64* tree, const double* pts, int64_t Q, double* values,
                               int64_t* node_ids, int64_t* data_ids, uint8_t* slot_mask, void* stream) {
    GTree<double> tr;
    int rc = g_tree_f64(tree, tr); if (rc) return rc;
    SVOXB_REQUIRE(Q >= 0 && (Q == 0 || (pts && node_ids)), "pts/node_ids NULL");
    if (Q == 0) return 0;
    wide_query_kernel<double><<<g_grid(Q), WIDE_BLOCK, 0, (cudaStream_t)stream>>>(tr, pts, Q, values, node_ids,
                                                                                        data_ids, slot_mask);
    count_launch();
    return check_cuda(cudaGetLastError(), "wide_query_kernel launch");
}

extern "C" int svoxb_render_rays_fwd_f64(const svoxb_tree_f64* tree, const double* origins, const double* dirs,
                                         int64_t Q, const svoxb_render_options* opt, double* out, double* depth,
                                         void* stream) {
    GTree<double> tr; GSource<double> src;
    int rc = g_tree_f64(tree, tr); if (rc) return rc;
    rc = g_check_opts_f64(opt); if (rc) return rc;
    rc = g_rays_f64(origins, dirs, Q, src); if (rc) return rc;
    SVOXB_REQUIRE(Q == 0 || out != nullptr, "out is NULL");
    if (Q == 0) return 0;
    return g_launch_fwd<double>(tr, src, g_opts<double>(opt), false, out, depth, (cudaStream_t)stream);
}

extern "C" int svoxb_render_rays_bwd_f64(const svoxb_tree_f64* tree, const double* origins, const double* dirs,
                                         int64_t Q, const svoxb_render_options* opt, const double* grad_out,
                                         const double* saved_out, double* grad_features, void* stream) {
    GTree<double> tr; GSource<double> src;
    int rc = g_tree_f64(tree, tr); if (rc) return rc;
    rc = g_check_opts_f64(opt); if (rc) return rc;
    rc = g_rays_f64(origins, dirs, Q, src); if (rc) return rc;
    SVOXB_REQUIRE(Q == 0 || (grad_out && saved_out && grad_features), "grad_out/saved_out/grad_features NULL");
    if (Q == 0) return 0;
    return g_launch_bwd<double>(tr, src, g_opts<double>(opt), false, grad_out, saved_out, grad_features,
                                (cudaStream_t)stream);
}

extern "C" int svoxb_render_image_fwd_f64(const svoxb_tree_f64* tree, const svoxb_camera_f64* cam,
                                          const svoxb_render_options* opt, double* out, double* depth, void* stream) {
    GTree<double> tr; GSource<double> src;
    int rc = g_tree_f64(tree, tr); if (rc) return rc;
    rc = g_check_opts_f64(opt); if (rc) return rc;
    rc = g_camera_f64(cam, src); if (rc) return rc;
    SVOXB_REQUIRE(out != nullptr, "out is NULL");
    return g_launch_fwd<double>(tr, src, g_opts<double>(opt), true, out, depth, (cudaStream_t)stream);
}

extern "C" int svoxb_render_image_bwd_f64(const svoxb_tree_f64* tree, const svoxb_camera_f64* cam,
                                          const svoxb_render_options* opt, const double* grad_out,
                                          const double* saved_out, double* grad_features, void* stream) {
    GTree<double> tr; GSource<double> src;
    int rc = g_tree_f64(tree, tr); if (rc) return rc;
    rc = g_check_opts_f64(opt); if (rc) return rc;
    rc = g_camera_f64(cam, src); if (rc) return rc;
    SVOXB_REQUIRE(grad_out && saved_out && grad_features, "grad_out/saved_out/grad_features NULL");
    return g_launch_bwd<double>(tr, src, g_opts<double>(opt), true, grad_out, saved_out, grad_features,
                                (cudaStream_t)stream);
}

extern "C" int svoxb_render_depth_f64(const svoxb_tree_f64* tree, const double* origins, const double* dirs, int64_t Q,
                                      const svoxb_render_options* opt, double* depth, void* stream) {
    GTree<double> tr; GSource<double> src;
    int rc = g_tree_f64(tree, tr); if (rc) return rc;
    rc = g_check_opts_f64(opt); if (rc) return rc;
    rc = g_rays_f64(origins, dirs, Q, src); if (rc) return rc;
    SVOXB_REQUIRE(Q == 0 || depth != nullptr, "depth is NULL");
    if (Q == 0) return 0;
    wide_depth_kernel<double><<<g_grid(Q), WIDE_BLOCK, 0, (cudaStream_t)stream>>>(tr, src, g_opts<double>(opt), depth);
    count_launch();
    return check_cuda(cudaGetLastError(), "wide_depth_kernel launch");
}
