// svoxb_common.cuh -- shared host/device helpers of libsvoxb (sm_100a only).
//
// Semantics follow the reference's device helpers (paths relative to /root/reference/svox_t/csrc):
// clamp_coord / transform_coord / query_single_from_root (include/common.cuh:37-100), _get_delta_scale and
// _dda_unit (rt_kernel.cu:187-218). Nothing is copied: the descent is restated for two layouts -- the
// reference's own child/data tensors (any N) and the packed grid+brick accelerator (N == 2, bit-sliced,
// exact because *2 / floor / subtract are exact in binary floating point).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../../include/svoxb.h"

namespace svoxb {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_STAGES = 4;

// ---- error plumbing (host) -----------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch(int n = 1);
unsigned long long* work_counter(cudaStream_t stream);   // zeroed 64-bit device counter for the next launch on `stream`
int sm_count();

#define SVOXB_CUDA(expr)                                                   \
    do {                                                                   \
        int _rc = ::svoxb::check_cuda((expr), #expr);                      \
        if (_rc != 0) return _rc;                                          \
    } while (0)
#define SVOXB_REQUIRE(cond, ...)                                           \
    do {                                                                   \
        if (!(cond)) { ::svoxb::set_error(__VA_ARGS__); return SVOXB_EINVAL; } \
    } while (0)

// ---- packed accelerator ------------------------------------------------------------------------------------------
// One 32-bit word per cell:
//   bit 31 set  : pointer -- low 31 bits = brick index in the next stage
//   bit 31 clear: leaf    -- bits 27..30 = leaf depth d (cube_sz = 2^d), bits 0..25 = feature-row index,
//                            ACC_EMPTY in the index field = empty leaf (reference: data idx >= M);
//                            bit 26 (ACC_MISS) = "this row's sigma was <= 0 when svoxb_accel_mark_hits last ran":
//                            kernels told that the marks are current skip such rows without fetching them
constexpr uint32_t ACC_PTR = 0x80000000u;
constexpr uint32_t ACC_IDX_MASK = 0x03ffffffu;
constexpr uint32_t ACC_EMPTY = 0x03ffffffu;
constexpr uint32_t ACC_MISS = 0x04000000u;
constexpr int ACC_DEPTH_SHIFT = 27;
constexpr int ACC_MAX_DEPTH = 15;

struct AccelView {
    int n_stages;                 // 1..MAX_STAGES
    int lmax;                     // sum of bits; integer coordinates carry lmax bits per axis
    int bits[MAX_STAGES];         // levels resolved by stage s
    int shift[MAX_STAGES];        // lmax - (levels resolved up to and including stage s)
    const uint32_t* cells[MAX_STAGES];  // stage 0: one grid of 8^bits[0] cells; stage s: bricks of 8^bits[s] cells
};

struct TreeView {
    const float* features;
    int64_t M;
    int D;
    int N;
    const int32_t* child;
    const int32_t* data;
    float off[3];   // filled in-kernel from tree->offset / scaling (device pointers)
    float scl[3];
};

// Kernel-side view of svoxb_tree (passed by value).
struct TreeArgs {
    const float* features;
    int64_t M;
    int D;
    int N;
    const int32_t* child;
    const int32_t* data;
    const float* offset;
    const float* scaling;
    AccelView acc;
    int use_accel;
    const float* feat_act;   // optional pre-activated table (sigmoid applied to channels 0..D-2), or nullptr
    int act_stride;          // its row stride in floats: D (D % 4 == 0), else D-1 rounded up to a multiple of 4 (payload only)
    const float* sigma_c;    // D % 4 != 0: compact sigma[M] that goes with the payload-only activated table
    uint32_t acc_miss_mask;  // ACC_MISS when the accelerator's hit marks are current for `features`, else 0
};

// ---- device math ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Debug build (make debug / -DSVOXB_DEBUG): index checks that trap the kernel. compute-sanitizer is not available on
// the target pool, so the test-suite is run once against this build instead (tests/tools/README in DESIGN.md section 2).
#ifdef SVOXB_DEBUG
#define SVOXB_DBG(cond) do { if (!(cond)) { printf("svoxb: check failed %s:%d: %s\n", __FILE__, __LINE__, #cond); __trap(); } } while (0)
#else
#define SVOXB_DBG(cond) do { } while (0)
#endif

// include/common.cuh:37-42. The reference clamps in double against 1.0 - 1e-6 and rounds back to float;
// for float inputs that is exactly min(q, (float)(1.0 - 1e-6)) followed by max(0, .).
__device__ __forceinline__ float clamp01(float q) {
    const float hi = (float)(1.0 - 1e-6);
    return fmaxf(0.0f, fminf(hi, q));
}

// rt_kernel.cu:201-218 -- slab test against the unit cube.
__device__ __forceinline__ void dda_unit(float cx, float cy, float cz, float ix, float iy, float iz,
                                         float& tmin, float& tmax) {
    float t1, t2;
    tmin = 0.0f; tmax = 1e9f;
    t1 = -cx * ix; t2 = t1 + ix; tmin = fmaxf(tmin, fminf(t1, t2)); tmax = fminf(tmax, fmaxf(t1, t2));
    t1 = -cy * iy; t2 = t1 + iy; tmin = fmaxf(tmin, fminf(t1, t2)); tmax = fminf(tmax, fmaxf(t1, t2));
    t1 = -cz * iz; t2 = t1 + iz; tmin = fmaxf(tmin, fminf(t1, t2)); tmax = fminf(tmax, fmaxf(t1, t2));
}

// sigmoid(x) = 1 / (1 + e^-x) = rcp(1 + ex2(-x * log2 e)): one FMUL, MUFU.EX2, FADD, MUFU.RCP (abs error ~2e-7;
// the reference evaluates the same expression with expf and a double divide, rt_kernel.cu:304). The .ftz forms
// skip the denormal range fix-up __expf would add (3 extra instructions per call).
__device__ __forceinline__ float fast_sigmoid(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

// Address of channel `lane` (+32k) of feature row idx: lane_base + idx * row_bytes, one IMAD.WIDE.U32.
__device__ __forceinline__ const float* row_ptr(const char* lane_base, int idx, unsigned row_bytes) {
    return reinterpret_cast<const float*>(lane_base + (size_t)(unsigned)idx * (size_t)row_bytes);
}
__device__ __forceinline__ float* row_ptr(char* lane_base, int idx, unsigned row_bytes) {
    return reinterpret_cast<float*>(lane_base + (size_t)(unsigned)idx * (size_t)row_bytes);
}

struct Leaf {
    int64_t idx;      // feature row, or -1 when the leaf is empty (data idx >= M)
    float rx, ry, rz; // position inside the leaf, [0,1)^3
    float inv_cube;   // 1 / cube_sz  (cube_sz = N^levels)
    float cube;       // cube_sz
    bool miss;        // accelerator hit marks current and this row's sigma <= 0: the sample cannot be a hit (no fetch)
};

// Descent over the reference tensors, any N (include/common.cuh:62-100). p in [0,1]^3 (unclamped).
__device__ __forceinline__ int64_t descend_ref(const int32_t* __restrict__ child, int N,
                                               float px, float py, float pz,
                                               float& rx, float& ry, float& rz, float& cube) {
    const float fN = (float)N;
    const int N2 = N * N;
    const int64_t N3 = (int64_t)N2 * N;
    px = clamp01(px); py = clamp01(py); pz = clamp01(pz);
    int64_t node = 0;
    cube = fN;
    while (true) {
        px *= fN; py *= fN; pz *= fN;
        const float fu = floorf(px), fv = floorf(py), fw = floorf(pz);
        px -= fu; py -= fv; pz -= fw;
        const int64_t slot = node * N3 + (int)fu * N2 + (int)fv * N + (int)fw;
        const int skip = __ldg(child + slot);
        if (skip == 0) { rx = px; ry = py; rz = pz; return slot; }
        cube *= fN;
        node += skip;
    }
}

template <bool ACCEL>
__device__ __forceinline__ Leaf locate(const TreeArgs& tr, const uint32_t* __restrict__ top_smem,
                                       float px, float py, float pz) {
    Leaf lf;
    if (ACCEL) {
        const AccelView& a = tr.acc;
        px = clamp01(px); py = clamp01(py); pz = clamp01(pz);
        const float s = __int_as_float((127 + a.lmax) << 23);       // 2^lmax, exact scaling
        const int Ix = (int)(px * s), Iy = (int)(py * s), Iz = (int)(pz * s);
        const int b0 = a.bits[0], s0 = a.shift[0];
        uint32_t cell = top_smem[(((Ix >> s0) << b0 | (Iy >> s0)) << b0) | (Iz >> s0)];
#pragma unroll 1
        for (int st = 1; (cell & ACC_PTR) && st < MAX_STAGES; ++st) {
            const int b = a.bits[st], sh = a.shift[st], m = (1 << b) - 1;
            const uint32_t lin = (((((Ix >> sh) & m) << b) | ((Iy >> sh) & m)) << b) | ((Iz >> sh) & m);
            cell = __ldg(a.cells[st] + (((size_t)(cell & 0x7fffffffu)) << (3 * b)) + lin);
        }
        const int d = (int)(cell >> ACC_DEPTH_SHIFT) & 0xf;
        const uint32_t idx = cell & ACC_IDX_MASK;
        lf.idx = (idx == ACC_EMPTY) ? -1 : (int64_t)idx;
        lf.miss = (cell & tr.acc_miss_mask) != 0;
        const float sc = __int_as_float((127 + d) << 23);           // cube_sz = 2^d
        lf.cube = sc;
        lf.inv_cube = __int_as_float((127 - d) << 23);
        const float qx = px * sc, qy = py * sc, qz = pz * sc;       // exact; identical to d rounds of *2,floor,-
        lf.rx = qx - floorf(qx); lf.ry = qy - floorf(qy); lf.rz = qz - floorf(qz);
    } else {
        const int64_t slot = descend_ref(tr.child, tr.N, px, py, pz, lf.rx, lf.ry, lf.rz, lf.cube);
        const int idx = __ldg(tr.data + slot);
        lf.idx = ((int64_t)idx >= tr.M || idx < 0) ? -1 : (int64_t)idx;
        lf.miss = false;
        lf.inv_cube = 1.0f / lf.cube;
    }
    return lf;
}

#endif  // __CUDACC__

}  // namespace svoxb
