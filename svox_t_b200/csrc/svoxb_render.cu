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
#include <math.h>
#include "svoxb_common.cuh"

namespace svoxb {

constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;
constexpr int CHUNK = 64;          // rays fetched from the global queue per atomic; one 8x8 pixel tile for images

struct RaySource {
    const float* origins;          // explicit rays: [Q,3] world space
    const float* dirs;
    const float* c2w;              // camera rays: row-major [>=3,4]
    float fx, fy;
    int width, height;
    int tiles_x;
    int64_t total;                 // queue length: Q, or n_tiles * 64
};

struct MarchOpts {
    float step, bg, sigma_thresh, stop_thresh;
};

struct Ray {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz, t, tmax, ds;
};

// rt_kernel.cu:663-665 (transform_coord, one FFMA per axis) + 227-247 (delta scale, invdir in double, slab test).
__device__ __forceinline__ void ray_setup(const float* __restrict__ off, const float* __restrict__ scl,
                                          float owx, float owy, float owz, float dwx, float dwy, float dwz, Ray& r) {
    r.ox = fmaf(scl[0], owx, off[0]);
    r.oy = fmaf(scl[1], owy, off[1]);
    r.oz = fmaf(scl[2], owz, off[2]);
    float dx = dwx * scl[0], dy = dwy * scl[1], dz = dwz * scl[2];
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    r.ds = 1.0f / nrm;
    dx *= r.ds; dy *= r.ds; dz *= r.ds;
    r.dx = dx; r.dy = dy; r.dz = dz;
    r.ix = (float)(1.0 / ((double)dx + 1e-9));
    r.iy = (float)(1.0 / ((double)dy + 1e-9));
    r.iz = (float)(1.0 / ((double)dz + 1e-9));
    float tmin, tmax;
    dda_unit(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, tmin, tmax);
    if (tmax < 0.0f || tmin > tmax) { tmin = 0.0f; tmax = 0.0f; }   // misses the cube: zero samples, T stays 1
    r.t = tmin; r.tmax = tmax;
}

// rt_kernel.cu:1152-1166 -- pinhole ray of pixel (px, py); double sub-expressions as in the reference.
__device__ __forceinline__ void camera_ray(const RaySource& s, int px, int py,
                                           float& ox, float& oy, float& oz, float& dx, float& dy, float& dz) {
    float x = (float)(((double)px - 0.5 * (double)s.width) / (double)s.fx);
    float y = (float)(-((double)py - 0.5 * (double)s.height) / (double)s.fy);
    float z = sqrtf((float)((double)(x * x + y * y) + 1.0));
    x /= z; y /= z; z = -1.0f / z;
    const float* c = s.c2w;
    dx = __ldg(c + 0) * x + __ldg(c + 1) * y + __ldg(c + 2) * z;
    dy = __ldg(c + 4) * x + __ldg(c + 5) * y + __ldg(c + 6) * z;
    dz = __ldg(c + 8) * x + __ldg(c + 9) * y + __ldg(c + 10) * z;
    ox = __ldg(c + 3); oy = __ldg(c + 7); oz = __ldg(c + 11);
}

// Per-warp view of the global ray queue.
struct Queue {
    int64_t next, end;
    bool exhausted;
};

// Hands queue entries to the lanes in `need`; returns the mask of lanes that now own a fresh ray.
// row = output row of the ray (ray index, or iy*W+ix).
template <bool IMAGE>
__device__ __forceinline__ unsigned refill(const RaySource& src, const float* off, const float* scl,
                                           unsigned long long* counter, Queue& q, unsigned need, int lane,
                                           Ray& ray, int& row) {
    unsigned got = 0;
    while (need) {
        if (q.next >= q.end) {
            if (q.exhausted) break;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)CHUNK);
            base = __shfl_sync(FULL, base, 0);
            if ((int64_t)base >= src.total) { q.exhausted = true; break; }
            q.next = (int64_t)base;
            q.end = min((int64_t)base + CHUNK, src.total);
        }
        const int avail = (int)(q.end - q.next);
        const int rank = __popc(need & ((1u << lane) - 1u));
        const bool take = ((need >> lane) & 1u) && rank < avail;
        bool valid = false;
        if (take) {
            const int64_t id = q.next + rank;
            float ox, oy, oz, dx, dy, dz;
            if (IMAGE) {
                const int tile = (int)(id >> 6), in = (int)(id & 63);
                const int px = (tile % src.tiles_x) * 8 + (in & 7);
                const int py = (tile / src.tiles_x) * 8 + (in >> 3);
                valid = px < src.width && py < src.height;
                if (valid) {
                    camera_ray(src, px, py, ox, oy, oz, dx, dy, dz);
                    row = py * src.width + px;
                }
            } else {
                valid = true;
                const float* o = src.origins + id * 3;
                const float* d = src.dirs + id * 3;
                ox = __ldg(o); oy = __ldg(o + 1); oz = __ldg(o + 2);
                dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
                row = (int)id;
            }
            if (valid) ray_setup(off, scl, ox, oy, oz, dx, dy, dz, ray);
        }
        const unsigned tm = __ballot_sync(FULL, take);
        const unsigned vm = __ballot_sync(FULL, valid);
        q.next += __popc(tm);
        need &= ~vm;
        got |= vm;
        // lanes that drew an out-of-image pixel stay in `need` and draw again
    }
    return got;
}

__device__ __forceinline__ void load_top(const TreeArgs& tr, uint32_t* top) {
    const int n = 1 << (3 * tr.acc.bits[0]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) top[i] = __ldg(tr.acc.cells[0] + i);
    __syncthreads();
}

// One march sample of the lane's ray (rt_kernel.cu:261-277): returns the leaf row (or -1), delta_t and sigma.
template <bool ACCEL>
__device__ __forceinline__ void sample(const TreeArgs& tr, const uint32_t* top, const Ray& r, float step,
                                       int64_t& idx, float& delta_t, float& sigma) {
    const float px = fmaf(r.t, r.dx, r.ox), py = fmaf(r.t, r.dy, r.oy), pz = fmaf(r.t, r.dz, r.oz);
    const Leaf lf = locate<ACCEL>(tr, top, px, py, pz);
    float smin, smax;
    dda_unit(lf.rx, lf.ry, lf.rz, r.ix, r.iy, r.iz, smin, smax);
    const float tsub = ACCEL ? (smax - smin) * lf.inv_cube : (smax - smin) / lf.cube;
    delta_t = tsub + step;
    idx = lf.idx;
    sigma = 0.0f;
    if (idx >= 0) sigma = __ldg(tr.features + idx * tr.D + (tr.D - 1));
}

// ------------------------------------------------------------------------------------------------------------
// Forward: out[row, 0..D-2] = sum_i w_i * sigmoid(f_i) + T * bg ; out[row, D-1] = 1 - T ; depth[row] = first hit.
template <int K, bool ACCEL, bool IMAGE>
__global__ void __launch_bounds__(BLOCK)
march_fwd_kernel(TreeArgs tr, RaySource src, MarchOpts opt, float* __restrict__ out, float* __restrict__ depth,
                 unsigned long long* counter) {
    extern __shared__ uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31;
    const int D = tr.D;
    const float off[3] = {__ldg(tr.offset), __ldg(tr.offset + 1), __ldg(tr.offset + 2)};
    const float scl[3] = {__ldg(tr.scaling), __ldg(tr.scaling + 1), __ldg(tr.scaling + 2)};

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
                if (sigma > opt.sigma_thresh) {                                   // rt_kernel.cu:279-320
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
    extern __shared__ uint32_t smem_u32[];
    const int top_words = ACCEL ? (1 << (3 * tr.acc.bits[0])) : 0;
    uint32_t* top = smem_u32;
    float* gs_all = reinterpret_cast<float*>(smem_u32 + top_words);      // [WARPS][32][32*K]
    if (ACCEL) load_top(tr, top);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* gs = gs_all + (size_t)warp * 32 * (32 * K);
    const int D = tr.D;
    const float off[3] = {__ldg(tr.offset), __ldg(tr.offset + 1), __ldg(tr.offset + 2)};
    const float scl[3] = {__ldg(tr.scaling), __ldg(tr.scaling + 1), __ldg(tr.scaling + 2)};

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
    extern __shared__ uint32_t smem_u32[];
    uint32_t* top = smem_u32;
    if (ACCEL) load_top(tr, top);
    const float off[3] = {__ldg(tr.offset), __ldg(tr.offset + 1), __ldg(tr.offset + 2)};
    const float scl[3] = {__ldg(tr.scaling), __ldg(tr.scaling + 1), __ldg(tr.scaling + 2)};
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
int make_tree_args(const svoxb_tree* t, TreeArgs& a);   // svoxb_tree.cu

static int check_opts(const svoxb_render_options* opt, MarchOpts& m) {
    SVOXB_REQUIRE(opt != nullptr, "render options are NULL");
    if (opt->format != SVOXB_FORMAT_RGBA) {
        set_error("data format %d (SH/SG/ASG) is not implemented: only the feature-level RGBA format is", opt->format);
        return SVOXB_EUNSUPPORTED;
    }
    if (opt->ndc_width >= 0) {
        set_error("NDC ray conversion (ndc_width >= 0) is not implemented");
        return SVOXB_EUNSUPPORTED;
    }
    m.step = opt->step_size; m.bg = opt->background_brightness;
    m.sigma_thresh = opt->sigma_thresh; m.stop_thresh = opt->stop_thresh;
    return 0;
}

static int make_source(const float* origins, const float* dirs, int64_t Q, const svoxb_camera* cam, RaySource& s) {
    s = RaySource{};
    if (cam) {
        SVOXB_REQUIRE(cam->c2w != nullptr && cam->width > 0 && cam->height > 0, "bad camera spec");
        SVOXB_REQUIRE((int64_t)cam->width * cam->height < (1ll << 31), "image too large");
        s.c2w = cam->c2w; s.fx = cam->fx; s.fy = cam->fy; s.width = cam->width; s.height = cam->height;
        s.tiles_x = (cam->width + 7) / 8;
        s.total = (int64_t)s.tiles_x * ((cam->height + 7) / 8) * 64;
    } else {
        SVOXB_REQUIRE(Q >= 0 && Q < (1ll << 31), "ray count out of range");
        SVOXB_REQUIRE(Q == 0 || (origins && dirs), "origins/dirs are NULL");
        s.origins = origins; s.dirs = dirs; s.total = Q;
    }
    return 0;
}

template <typename Kern>
static int persistent_grid(Kern kern, size_t smem, int& grid) {
    if (smem > 48 * 1024)
        SVOXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SVOXB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem));
    SVOXB_REQUIRE(per_sm > 0, "kernel does not fit on an SM (smem %zu)", smem);
    grid = per_sm * sm_count();
    return 0;
}

template <int K, bool ACCEL, bool IMAGE>
static int launch_fwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, float* out, float* depth,
                      cudaStream_t st) {
    const size_t smem = ACCEL ? sizeof(uint32_t) << (3 * tr.acc.bits[0]) : 0;
    auto kern = march_fwd_kernel<K, ACCEL, IMAGE>;
    int grid = 0;
    int rc = persistent_grid(kern, smem, grid);
    if (rc) return rc;
    const int64_t warps_needed = (src.total + CHUNK - 1) / CHUNK;
    grid = (int)max((int64_t)1, min((int64_t)grid, (warps_needed + WARPS - 1) / WARPS));
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
    int rc = persistent_grid(kern, smem, grid);
    if (rc) return rc;
    const int64_t warps_needed = (src.total + CHUNK - 1) / CHUNK;
    grid = (int)max((int64_t)1, min((int64_t)grid, (warps_needed + WARPS - 1) / WARPS));
    unsigned long long* counter = work_counter(st);
    if (!counter) return SVOXB_ECUDA;
    kern<<<grid, BLOCK, smem, st>>>(tr, src, m, grad_out, saved_out, grad, counter);
    count_launch();
    return check_cuda(cudaGetLastError(), "march_bwd_kernel launch");
}

template <bool IMAGE>
static int dispatch_fwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, float* out, float* depth,
                        cudaStream_t st) {
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
    set_error("feature width D=%d not supported (max 128)", tr.D);
    return SVOXB_EINVAL;
}

template <bool IMAGE>
static int dispatch_bwd(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, const float* go,
                        const float* so, float* grad, cudaStream_t st) {
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
    set_error("feature width D=%d not supported (max 128)", tr.D);
    return SVOXB_EINVAL;
}

}  // namespace svoxb

using namespace svoxb;

extern "C" int svoxb_render_rays_fwd(const svoxb_tree* tree, const float* origins, const float* dirs,
                                     const float* vdirs, int64_t Q, const svoxb_render_options* opt, float* out,
                                     float* depth, void* stream) {
    (void)vdirs;
    TreeArgs tr; MarchOpts m; RaySource src;
    int rc = make_tree_args(tree, tr); if (rc) return rc;
    rc = check_opts(opt, m); if (rc) return rc;
    rc = make_source(origins, dirs, Q, nullptr, src); if (rc) return rc;
    SVOXB_REQUIRE(Q == 0 || out != nullptr, "out is NULL");
    if (Q == 0) return 0;
    return dispatch_fwd<false>(tr, src, m, out, depth, (cudaStream_t)stream);
}

extern "C" int svoxb_render_rays_bwd(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                     const svoxb_render_options* opt, const float* grad_out, const float* saved_out,
                                     float* grad_features, void* stream) {
    TreeArgs tr; MarchOpts m; RaySource src;
    int rc = make_tree_args(tree, tr); if (rc) return rc;
    rc = check_opts(opt, m); if (rc) return rc;
    rc = make_source(origins, dirs, Q, nullptr, src); if (rc) return rc;
    SVOXB_REQUIRE(Q == 0 || (grad_out && saved_out && grad_features), "grad_out/saved_out/grad_features NULL");
    if (Q == 0) return 0;
    return dispatch_bwd<false>(tr, src, m, grad_out, saved_out, grad_features, (cudaStream_t)stream);
}

extern "C" int svoxb_render_image_fwd(const svoxb_tree* tree, const svoxb_camera* cam,
                                      const svoxb_render_options* opt, float* out, float* depth, void* stream) {
    TreeArgs tr; MarchOpts m; RaySource src;
    int rc = make_tree_args(tree, tr); if (rc) return rc;
    rc = check_opts(opt, m); if (rc) return rc;
    SVOXB_REQUIRE(cam != nullptr && out != nullptr, "camera/out NULL");
    rc = make_source(nullptr, nullptr, 0, cam, src); if (rc) return rc;
    return dispatch_fwd<true>(tr, src, m, out, depth, (cudaStream_t)stream);
}

extern "C" int svoxb_render_image_bwd(const svoxb_tree* tree, const svoxb_camera* cam,
                                      const svoxb_render_options* opt, const float* grad_out,
                                      const float* saved_out, float* grad_features, void* stream) {
    TreeArgs tr; MarchOpts m; RaySource src;
    int rc = make_tree_args(tree, tr); if (rc) return rc;
    rc = check_opts(opt, m); if (rc) return rc;
    SVOXB_REQUIRE(cam && grad_out && saved_out && grad_features, "camera/grad_out/saved_out/grad_features NULL");
    rc = make_source(nullptr, nullptr, 0, cam, src); if (rc) return rc;
    return dispatch_bwd<true>(tr, src, m, grad_out, saved_out, grad_features, (cudaStream_t)stream);
}

extern "C" int svoxb_render_depth(const svoxb_tree* tree, const float* origins, const float* dirs, int64_t Q,
                                  const svoxb_render_options* opt, float* depth, void* stream) {
    TreeArgs tr; MarchOpts m;
    int rc = make_tree_args(tree, tr); if (rc) return rc;
    rc = check_opts(opt, m); if (rc) return rc;
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
