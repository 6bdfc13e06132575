// svoxb_march.cuh -- device code shared by the march kernels: ray sources (explicit batches, pinhole images), per-ray
// set-up, the chunked global ray queue that keeps every lane busy (ray compaction), and one traversal step.
// Reference semantics: rt_kernel.cu:187-247 (set-up), 261-277 (sample), 1152-1166 (camera rays).
#pragma once
#include <math.h>
#include "svoxb_common.cuh"

namespace svoxb {

#ifndef SVOXB_BLOCK
#define SVOXB_BLOCK 256
#endif
constexpr int BLOCK = SVOXB_BLOCK;
constexpr int WARPS = BLOCK / 32;
constexpr int CHUNK = 64;          // image queue entry: one 8x8 pixel tile (coherent lanes)
#ifndef SVOXB_RAY_CHUNK
#define SVOXB_RAY_CHUNK 32
#endif
constexpr int RAY_CHUNK = SVOXB_RAY_CHUNK;      // explicit-ray queue entry: one ray per lane (measured against 8 / 16 / 64, svoxb_render_q.cu)

struct RaySource {
    const float* origins;          // explicit rays: [Q,3] world space
    const float* dirs;
    const float* c2w;              // camera rays: row-major [>=3,4]
    float fx, fy;
    int width, height;
    int tiles_x;
    int row_begin, row_end;        // camera rays: image rows [row_begin, row_end) are rendered (a band of the image)
    int64_t total;                 // queue length: Q, or n_tiles * 64
    const float* vdirs;            // explicit rays: [Q,3] view directions (view-dependent formats only)
    int ndc_w, ndc_h;              // camera rays: NDC conversion when ndc_w >= 0 (rt_kernel.cu:1168-1191)
    float ndc_focal;
    int chunk;                     // explicit rays: rays fetched from the queue per atomic (0 = RAY_CHUNK)
    const int* order;              // explicit rays, optional: queue position -> ray index (svoxb_order.cu: longest first)
    int* steps_out;                // explicit rays, optional [Q]: march iterations of each ray, written by the forward
};

struct ViewDir {
    float x, y, z;
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
    // offset / scaling are read here (L1-resident, once per ray) rather than held in registers by every lane
    const float s0 = __ldg(scl), s1 = __ldg(scl + 1), s2 = __ldg(scl + 2);
    r.ox = fmaf(s0, owx, __ldg(off));
    r.oy = fmaf(s1, owy, __ldg(off + 1));
    r.oz = fmaf(s2, owz, __ldg(off + 2));
    float dx = dwx * s0, dy = dwy * s1, dz = dwz * s2;
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

// rt_kernel.cu:1168-1191 (maybe_world2ndc, near = 1): world-space pinhole ray -> NDC ray, direction re-normalised.
__device__ __forceinline__ void world2ndc(const RaySource& s, float& ox, float& oy, float& oz,
                                          float& dx, float& dy, float& dz) {
    const float t = -(1.0f + oz) / dz;
    ox = ox + t * dx; oy = oy + t * dy; oz = oz + t * dz;
    const float kx = (2.0f * s.ndc_focal) / (float)s.ndc_w, ky = (2.0f * s.ndc_focal) / (float)s.ndc_h;
    dx = -kx * (dx / dz - ox / oz);
    dy = -ky * (dy / dz - oy / oz);
    dz = -2.0f / oz;
    ox = -kx * (ox / oz);
    oy = -ky * (oy / oz);
    oz = 1.0f + 2.0f / oz;
    const float n = sqrtf(dx * dx + dy * dy + dz * dz);
    dx /= n; dy /= n; dz /= n;
}

// Per-warp view of the global ray queue.
struct Queue {
    int next, end;          // queue positions fit 31 bits (checked on the host)
    bool exhausted;
};

// Hands queue entries to the lanes in `need`; returns the mask of lanes that now own a fresh ray.
// row = output row of the ray (ray index, or iy*W+ix). VDIR: also hand back the ray's view direction (explicit rays:
// vdirs[id]; camera rays: the world-space direction before any NDC conversion, rt_kernel.cu:1203).
template <bool IMAGE, bool VDIR = false>
__device__ __forceinline__ unsigned refill(const RaySource& src, const float* off, const float* scl,
                                           unsigned long long* counter, Queue& q, unsigned need, int lane,
                                           Ray& ray, int& row, ViewDir* vd = nullptr) {
    unsigned got = 0;
    while (need) {
        if (q.next >= q.end) {
            if (q.exhausted) break;
            unsigned long long base = 0;
            const int chunk = IMAGE ? CHUNK : (src.chunk > 0 ? src.chunk : RAY_CHUNK);
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)chunk);
            base = __shfl_sync(FULL, base, 0);
            if ((int64_t)base >= src.total) { q.exhausted = true; break; }
            q.next = (int)base;
            q.end = (int)min((int64_t)base + chunk, src.total);
        }
        const int avail = q.end - q.next;
        const int rank = __popc(need & ((1u << lane) - 1u));
        const bool take = ((need >> lane) & 1u) && rank < avail;
        bool valid = false;
        if (take) {
            const int pos = q.next + rank;
            float ox, oy, oz, dx, dy, dz;
            if (IMAGE) {
                const int id = pos;
                const int tile = (int)(id >> 6), in = (int)(id & 63);
                const int px = (tile % src.tiles_x) * 8 + (in & 7);
                const int py = src.row_begin + (tile / src.tiles_x) * 8 + (in >> 3);
                valid = px < src.width && py < src.row_end;
                if (valid) {
                    camera_ray(src, px, py, ox, oy, oz, dx, dy, dz);
                    row = (py - src.row_begin) * src.width + px;      // output rows start at the band
                    if (VDIR) { vd->x = dx; vd->y = dy; vd->z = dz; }
                    if (src.ndc_w >= 0) world2ndc(src, ox, oy, oz, dx, dy, dz);
                }
            } else {
                valid = true;
                const int id = src.order ? __ldg(src.order + pos) : pos;
                SVOXB_DBG(id >= 0 && (int64_t)id < src.total);
                const float* o = src.origins + (int64_t)id * 3;
                const float* d = src.dirs + (int64_t)id * 3;
                ox = __ldg(o); oy = __ldg(o + 1); oz = __ldg(o + 2);
                dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
                row = id;
                if (VDIR) {
                    const float* v = src.vdirs + (int64_t)id * 3;
                    vd->x = __ldg(v); vd->y = __ldg(v + 1); vd->z = __ldg(v + 2);
                }
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

// Stage the accelerator's top grid (8^bits[0] tagged words, <= 16 KB) in shared memory with ONE bulk asynchronous copy
// (TMA, cp.async.bulk global -> shared, completion on an mbarrier) issued by thread 0; every thread then waits on
// the barrier's phase. `top` must be the 16-byte aligned start of the dynamic shared memory; the grid is a
// stream-ordered pool allocation (256-byte aligned) whose size is a multiple of 32 bytes.
#ifndef SVOXB_TOP_TMA
#define SVOXB_TOP_TMA 1
#endif
__device__ __forceinline__ void load_top(const TreeArgs& tr, uint32_t* top) {
#if !SVOXB_TOP_TMA
    const int n = 1 << (3 * tr.acc.bits[0]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) top[i] = __ldg(tr.acc.cells[0] + i);
    __syncthreads();
    return;
#endif
    __shared__ __align__(8) unsigned long long top_bar;
    const uint32_t bytes = 4u << (3 * tr.acc.bits[0]);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&top_bar), dst = (uint32_t)__cvta_generic_to_shared(top);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(tr.acc.cells[0]), "r"(bytes), "r"(bar) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar) : "memory");
}

// sigma of a located leaf for the kernels that need nothing else of the row (depth, opacity, motion): 0 for empty
// leaves and for rows the hit marks flag as sigma <= 0 (no fetch); from the compact sigma array when the caller attached
// one (svoxb_tree.features_sigma: 4 M bytes, L2-resident) instead of one 32-byte sector of the [M, D] table per sample.
__device__ __forceinline__ float leaf_sigma(const TreeArgs& tr, const Leaf& lf) {
    if (lf.idx < 0 || lf.miss) return 0.0f;
    return tr.sigma_c ? __ldg(tr.sigma_c + lf.idx) : __ldg(tr.features + lf.idx * tr.D + (tr.D - 1));
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
    sigma = leaf_sigma(tr, lf);
}


// One traversal step without touching the feature table: leaf row (or -1) and delta_t of the sample at ray.t.
template <bool ACCEL>
__device__ __forceinline__ void traverse(const TreeArgs& tr, const uint32_t* top, const Ray& r, float step,
                                         int& idx, float& delta_t) {
    const float px = fmaf(r.t, r.dx, r.ox), py = fmaf(r.t, r.dy, r.oy), pz = fmaf(r.t, r.dz, r.oz);
    const Leaf lf = locate<ACCEL>(tr, top, px, py, pz);
    float smin, smax;
    dda_unit(lf.rx, lf.ry, lf.rz, r.ix, r.iy, r.iz, smin, smax);
    const float tsub = ACCEL ? (smax - smin) * lf.inv_cube : (smax - smin) / lf.cube;
    delta_t = tsub + step;
    idx = (int)lf.idx;
}


// Split traversal step for software pipelining: probe_begin computes the sample position and ISSUES the first
// brick lookup (accelerator path); probe_end consumes it -- deeper stages if any, leaf decode, slab test, delta_t.
// Whatever the caller runs in between hides the lookup latency.
struct Probe {
    float px, py, pz;
    uint32_t cell;
};

template <bool ACCEL>
__device__ __forceinline__ void probe_begin(const TreeArgs& tr, const uint32_t* __restrict__ top, const Ray& r,
                                            Probe& pb) {
    pb.px = fmaf(r.t, r.dx, r.ox); pb.py = fmaf(r.t, r.dy, r.oy); pb.pz = fmaf(r.t, r.dz, r.oz);
    pb.cell = 0;
    if (ACCEL) {
        const AccelView& a = tr.acc;
        pb.px = clamp01(pb.px); pb.py = clamp01(pb.py); pb.pz = clamp01(pb.pz);
        const float s = __int_as_float((127 + a.lmax) << 23);
        const int Ix = (int)(pb.px * s), Iy = (int)(pb.py * s), Iz = (int)(pb.pz * s);
        const int b0 = a.bits[0], s0 = a.shift[0];
        uint32_t cell = top[(((Ix >> s0) << b0 | (Iy >> s0)) << b0) | (Iz >> s0)];
        if (cell & ACC_PTR) {
            const int b = a.bits[1], sh = a.shift[1], m = (1 << b) - 1;
            const uint32_t lin = (((((Ix >> sh) & m) << b) | ((Iy >> sh) & m)) << b) | ((Iz >> sh) & m);
            cell = __ldg(a.cells[1] + (((size_t)(cell & 0x7fffffffu)) << (3 * b)) + lin);
        }
        pb.cell = cell;
    }
}

// Trees deeper than two stages (depth > 8): consume the stage-1 word and ISSUE the stage-2 lookup. Stage 1 is small
// (L2-resident); stage 2 of a depth-10 scene is hundreds of MB, its lookup goes to DRAM -- called between the row
// requests and the compositing, so that it is in flight during the compositing instead of being waited for in
// probe_end (measured on the C5 scene: no change, 5.60 vs 5.57 ms -- the lookup is not the exposed latency there; kept
// because it costs nothing). A no-op (no wait on the stage-1 word) for <= 2 stages.
template <bool ACCEL>
__device__ __forceinline__ void probe_mid(const TreeArgs& tr, Probe& pb) {
    if (ACCEL) {
        const AccelView& a = tr.acc;
        if (a.n_stages > 2 && (pb.cell & ACC_PTR)) {
            const float s = __int_as_float((127 + a.lmax) << 23);
            const int Ix = (int)(pb.px * s), Iy = (int)(pb.py * s), Iz = (int)(pb.pz * s);
            const int b = a.bits[2], sh = a.shift[2], m = (1 << b) - 1;
            const uint32_t lin = (((((Ix >> sh) & m) << b) | ((Iy >> sh) & m)) << b) | ((Iz >> sh) & m);
            pb.cell = __ldg(a.cells[2] + (((size_t)(pb.cell & 0x7fffffffu)) << (3 * b)) + lin);
        }
    }
}

// FIRST: the first stage probe_end may still have to resolve (2; 3 when the caller ran probe_mid).
template <bool ACCEL, int FIRST = 2>
__device__ __forceinline__ void probe_end(const TreeArgs& tr, const Probe& pb, const Ray& r, float step,
                                          int& idx, float& delta_t) {
    float rx, ry, rz, smin, smax;
    const float px = pb.px, py = pb.py, pz = pb.pz;
    if (ACCEL) {
        const AccelView& a = tr.acc;
        uint32_t cell = pb.cell;
        if (cell & ACC_PTR) {                                   // trees deeper than two stages (depth > 8)
            const float s = __int_as_float((127 + a.lmax) << 23);
            const int Ix = (int)(px * s), Iy = (int)(py * s), Iz = (int)(pz * s);
#pragma unroll
            for (int st = FIRST; st < MAX_STAGES; ++st) {
                if (cell & ACC_PTR) {
                    const int b = a.bits[st], sh = a.shift[st], m = (1 << b) - 1;
                    const uint32_t lin = (((((Ix >> sh) & m) << b) | ((Iy >> sh) & m)) << b) | ((Iz >> sh) & m);
                    cell = __ldg(a.cells[st] + (((size_t)(cell & 0x7fffffffu)) << (3 * b)) + lin);
                }
            }
        }
        const int d = (int)(cell >> ACC_DEPTH_SHIFT) & 0xf;
        const uint32_t ci = cell & ACC_IDX_MASK;
        SVOXB_DBG(!(cell & ACC_PTR) && (ci == ACC_EMPTY || (int64_t)ci < tr.M));
        // rows marked "sigma <= 0" are not candidates: no row fetch, exactly what the hit predicate would decide
        idx = (ci == ACC_EMPTY || (cell & tr.acc_miss_mask)) ? -1 : (int)ci;
        const float sc = __int_as_float((127 + d) << 23);
        const float qx = px * sc, qy = py * sc, qz = pz * sc;
        rx = qx - floorf(qx); ry = qy - floorf(qy); rz = qz - floorf(qz);
        dda_unit(rx, ry, rz, r.ix, r.iy, r.iz, smin, smax);
        delta_t = (smax - smin) * __int_as_float((127 - d) << 23) + step;
    } else {
        float cube;
        const int64_t slot = descend_ref(tr.child, tr.N, px, py, pz, rx, ry, rz, cube);
        const int di = __ldg(tr.data + slot);
        idx = ((int64_t)di >= tr.M || di < 0) ? -1 : di;
        dda_unit(rx, ry, rz, r.ix, r.iy, r.iz, smin, smax);
        delta_t = (smax - smin) / cube + step;
    }
}

// Host-side launch helpers shared by the scalar-lane and quad-lane kernels. `carveout_kb` > 0: ask for exactly that
// much shared memory per SM, so that everything else of the 228 KB stays L1 (see svoxb_render_q.cu).
template <typename Kern>
static int persistent_grid(Kern kern, size_t smem, int64_t queue_len, int& grid, int threads = BLOCK,
                           int carveout_kb = 0, int chunk = RAY_CHUNK) {
    if (smem + 1024 > 48 * 1024)      // static shared memory (the top grid's mbarrier) counts against the 48 KB default
        SVOXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (carveout_kb > 0)
        SVOXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        min(100, (carveout_kb * 100 + 227) / 228)));
    int per_sm = 0;
    SVOXB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    SVOXB_REQUIRE(per_sm > 0, "kernel does not fit on an SM (smem %zu)", smem);
    const int warps = threads / 32;
    const int64_t warps_needed = (queue_len + chunk - 1) / chunk;
    const int64_t want = (warps_needed + warps - 1) / warps;
    grid = (int)max((int64_t)1, min((int64_t)per_sm * sm_count(), want));
    return 0;
}

// Quad-lane kernels (svoxb_render_q.cu): D % 4 == 0 (4 <= D <= 128), or any D <= 128 with the padded activated table.
bool quad_supported(const TreeArgs& tr);
// svoxb_order.cu: longest-first order of an explicit ray batch for the backward, from the forward's exact per-ray
// iteration counts (stream-ordered scratch, released by the caller)
bool want_ray_order(const TreeArgs& tr, int64_t Q);
int64_t ray_order_max_rays();
int64_t ray_order_min_rays();
int build_ray_order(const int* cost, int64_t Q, int** order, cudaStream_t st);
int launch_fwd_quad(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, bool image, float* out,
                    float* depth, cudaStream_t st);
int launch_bwd_quad(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, bool image, const float* grad_out,
                    const float* saved_out, float* grad, cudaStream_t st);
// General kernels (svoxb_render_wide.cu): any D, walk over child/data; the float32 route for D > 128.
int launch_fwd_wide(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, bool image, float* out,
                    float* depth, cudaStream_t st);
int launch_bwd_wide(const TreeArgs& tr, const RaySource& src, const MarchOpts& m, bool image, const float* grad_out,
                    const float* saved_out, float* grad, cudaStream_t st);

}  // namespace svoxb
