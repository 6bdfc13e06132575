// svoxb_tree.cu -- octree descent outside the march: batched point query + unique-leaf set, construct_tree, and the
// packed grid+brick accelerator the march kernels walk. Also the library's housekeeping (errors, launch counter).
//
// Replaces (reference paths relative to /root/reference/svox_t/csrc):
//   query_single_from_root                                   include/common.cuh:62-100
//   query_single_kernel / query_vertical                     svox_kernel.cu:44-81, 274-324
//   generate_index_kernel / unpack_mask_kernel               svox_kernel.cu:239-269
//   construct_tree_kernel / construct_tree                   svox_kernel.cu:110-121, 341-352
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <stdlib.h>
#include "svoxb_common.cuh"

namespace svoxb {

// ---- housekeeping -------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    return SVOXB_ECUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// A ring of 64-bit work counters per (device, stream); each launch gets the next slot of ITS stream's ring, zeroed in
// that stream's order. A slot is reused after N_COUNTERS later launches on the same stream, which stream order places
// after the kernel that owned it; launches on other streams never touch it (a ring shared by all streams would let 256
// short launches on stream B reset the queue of a long persistent march on stream A). All rings of a device come from
// one allocation made at the first launch (no cudaMalloc later: launches inside a stream capture stay legal); streams
// beyond N_RINGS share rings by hash.
static constexpr int N_COUNTERS = 256;
static constexpr int N_RINGS = 32;
struct DeviceRings {
    unsigned long long* base = nullptr;
    cudaStream_t owner[N_RINGS] = {};
    bool used[N_RINGS] = {};
    unsigned next[N_RINGS] = {};
};
static DeviceRings g_rings[64];
static std::mutex g_counter_mu;

unsigned long long* work_counter(cudaStream_t stream) {
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return nullptr;
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return nullptr; }
    unsigned long long* c = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_counter_mu);
        DeviceRings& d = g_rings[dev];
        if (!d.base &&
            check_cuda(cudaMalloc(&d.base, sizeof(unsigned long long) * N_COUNTERS * N_RINGS), "cudaMalloc(work counters)"))
            return nullptr;
        int ring = -1;
        for (int i = 0; i < N_RINGS && ring < 0; ++i)
            if (d.used[i] && d.owner[i] == stream) ring = i;
        for (int i = 0; i < N_RINGS && ring < 0; ++i)
            if (!d.used[i]) { d.used[i] = true; d.owner[i] = stream; ring = i; }
        if (ring < 0) ring = (int)(((uintptr_t)stream >> 4) % N_RINGS);
        c = d.base + (size_t)ring * N_COUNTERS + (d.next[ring]++ % N_COUNTERS);
    }
    if (check_cuda(cudaMemsetAsync(c, 0, sizeof(unsigned long long), stream), "cudaMemsetAsync(work counter)"))
        return nullptr;
    return c;
}

}  // namespace svoxb

// The accelerator object behind the opaque handle.
struct svoxb_accel {
    svoxb::AccelView view;
    uint32_t* cells[svoxb::MAX_STAGES];
    int64_t n_bricks[svoxb::MAX_STAGES];
    int64_t bytes;
    int64_t M;
    const int32_t* child;   // identity of the tensors it was built from (sanity check only)
    const int32_t* data;
    cudaStream_t stream;    // creation stream: the stream-ordered allocations are released on it
    const float* marks_features;   // table the ACC_MISS bits were last computed from (svoxb_accel_mark_hits), or NULL
    int marks_D;
    uint32_t* row_cell;     // [M] inverse map: feature row -> the leaf cell that holds it (stage << 30 | cell index), so
                            // that a pass over the ROWS can refresh the hit marks (svoxb_prepare_step); 0xFFFFFFFF = none
    int rows_shared;        // a row is held by several leaf cells (children of a refined leaf inherit its row,
                            // svox.py:539-540) or a stage has >= 2^30 cells: the inverse map is not usable
    int* d_flags;           // device scalars of the build: [1] overflow, [2 + s] bricks of stage s + 1, [6] rows shared
    int lazy;               // built without read-backs: brick counts / flags are only known on the device (d_flags)
    int inplace;            // built by the no-read-back path: svoxb_accel_rebuild can refill it without any allocation
    int32_t* roots;         // no-read-back path: brick roots of stage 1 (kept for the next rebuild)
    int64_t roots_cap, row_cap;
    cudaStream_t used[4];   // streams other than `stream` that kernels reading the cells were launched on (the most
    int n_used;             // recent four): svoxb_accel_destroy orders the release after their work
};

namespace svoxb {

int make_tree_args(const svoxb_tree* t, TreeArgs& a);

// Entry points that launch on `use_stream` kernels reading the accelerator: also notes the stream in the accelerator,
// so that its stream-ordered release waits for them.
int make_tree_args(const svoxb_tree* t, TreeArgs& a, void* use_stream) {
    const int rc = make_tree_args(t, a);
    if (rc == 0 && t->accel) {
        svoxb_accel* acc = const_cast<svoxb_accel*>(t->accel);
        cudaStream_t st = static_cast<cudaStream_t>(use_stream);
        if (st != acc->stream) {
            bool seen = false;
            for (int i = 0; i < acc->n_used; ++i) seen |= acc->used[i] == st;
            if (!seen) {
                if (acc->n_used == 4) { for (int i = 0; i < 3; ++i) acc->used[i] = acc->used[i + 1]; acc->n_used = 3; }
                acc->used[acc->n_used++] = st;
            }
        }
    }
    return rc;
}

int make_tree_args(const svoxb_tree* t, TreeArgs& a) {
    SVOXB_REQUIRE(t != nullptr, "tree is NULL");
    SVOXB_REQUIRE((t->features || t->M == 0) && t->child && t->data && t->offset && t->scaling, "tree has NULL tensors");
    SVOXB_REQUIRE(t->N >= 2 && t->N <= 16, "branching factor N=%d out of range [2,16]", t->N);
    SVOXB_REQUIRE(t->D >= 2, "feature width D=%d must be >= 2 (payload + sigma)", t->D);
    SVOXB_REQUIRE(t->M >= 0 && t->M < (1ll << 31), "M=%lld out of range", (long long)t->M);
    SVOXB_REQUIRE(t->n_internal >= 1 && t->n_internal <= t->n_nodes, "n_internal=%lld out of range",
                  (long long)t->n_internal);
    a.features = t->features; a.M = t->M;
    if (t->M == 0) {                       // no rows at all: the kernels still prefetch "row 0" -> point at zeros
        static float* dummy[64] = {nullptr};
        int dev = 0;
        SVOXB_CUDA(cudaGetDevice(&dev));
        if (!dummy[dev & 63]) {
            SVOXB_CUDA(cudaMalloc(&dummy[dev & 63], 2048));
            SVOXB_CUDA(cudaMemset(dummy[dev & 63], 0, 2048));
        }
        a.features = dummy[dev & 63];
    } a.D = t->D; a.N = t->N;
    a.child = t->child; a.data = t->data; a.offset = t->offset; a.scaling = t->scaling;
    a.use_accel = 0;
    a.feat_act = t->M > 0 ? t->features_act : nullptr;
    a.act_stride = t->features_act_stride > 0 ? t->features_act_stride : t->D;
    a.sigma_c = t->M > 0 ? t->features_sigma : nullptr;
    if (a.feat_act != nullptr) {
        if (t->D % 4 == 0)
            SVOXB_REQUIRE(a.act_stride == t->D, "features_act_stride=%d must be D=%d", a.act_stride, t->D);
        else
            SVOXB_REQUIRE(a.act_stride == (t->D - 1 + 3) / 4 * 4 && a.sigma_c != nullptr,
                          "D %% 4 != 0: features_act must be the payload-only table (stride %d) with features_sigma",
                          (t->D - 1 + 3) / 4 * 4);
    }
    a.acc_miss_mask = 0;
    memset(&a.acc, 0, sizeof(a.acc));
    if (t->accel) {
        SVOXB_REQUIRE(t->N == 2, "accelerator requires N == 2");
        SVOXB_REQUIRE(t->accel->child == t->child && t->accel->data == t->data,
                      "accelerator was built from different child/data tensors");
        SVOXB_REQUIRE(t->accel->M == t->M, "accelerator was built for M=%lld, tree has M=%lld",
                      (long long)t->accel->M, (long long)t->M);
        a.acc = t->accel->view;
        a.use_accel = 1;
        if (t->accel_marks_current) {
            SVOXB_REQUIRE(t->accel->marks_features == t->features && t->accel->marks_D == t->D,
                          "accel_marks_current is set but svoxb_accel_mark_hits was last run on a different table");
            a.acc_miss_mask = ACC_MISS;
        }
    }
    return 0;
}

// ---- accelerator build ----------------------------------------------------------------------------------------------
__global__ void max_depth_kernel(const int32_t* __restrict__ parent_depth, int64_t n, int* __restrict__ out) {
    int m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, __ldg(parent_depth + 2 * i + 1));
    for (int s = 16; s > 0; s >>= 1) m = max(m, __shfl_xor_sync(FULL, m, s));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// One thread per cell of one stage. roots == nullptr: stage 0 (single grid rooted at node 0).
__global__ void accel_stage_kernel(const int32_t* __restrict__ child, const int32_t* __restrict__ data, int64_t M,
                                   const int32_t* __restrict__ roots, int64_t n_bricks, int bits, int base_depth,
                                   uint32_t* __restrict__ cells, int32_t* __restrict__ next_roots,
                                   int* __restrict__ next_count, int is_last, int* __restrict__ overflow,
                                   uint32_t* __restrict__ row_cell, uint32_t stage_tag, int* __restrict__ rows_shared,
                                   const int* __restrict__ n_bricks_dev) {
    const int64_t per = 1ll << (3 * bits);
    if (n_bricks_dev) n_bricks = min(n_bricks, (int64_t)*n_bricks_dev);     // built without a read-back: the count is here
    const int64_t total = n_bricks * per;
    for (int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
         gid += (int64_t)gridDim.x * blockDim.x) {
        const int64_t brick = gid >> (3 * bits);
        const int lin = (int)(gid & (per - 1));
        const int m = (1 << bits) - 1;
        const int x = (lin >> (2 * bits)) & m, y = (lin >> bits) & m, z = lin & m;
        int64_t node = roots ? (int64_t)__ldg(roots + brick) : 0;
        uint32_t cell = 0;
        bool done = false;
        for (int l = 0; l < bits; ++l) {
            const int sh = bits - 1 - l;
            const int64_t slot = node * 8 + (((x >> sh) & 1) << 2) + (((y >> sh) & 1) << 1) + ((z >> sh) & 1);
            const int skip = __ldg(child + slot);
            if (skip == 0) {
                const int idx = __ldg(data + slot);
                const uint32_t enc = (idx < 0 || (int64_t)idx >= M) ? ACC_EMPTY : (uint32_t)idx;
                cell = ((uint32_t)(base_depth + l + 1) << ACC_DEPTH_SHIFT) | enc;
                if (row_cell && enc != ACC_EMPTY && atomicExch(row_cell + enc, stage_tag | (uint32_t)gid) != 0xffffffffu)
                    *rows_shared = 1;
                done = true;
                break;
            }
            node += skip;
        }
        if (!done) {
            if (is_last) {
                *overflow = 1;
                cell = ((uint32_t)(base_depth + bits) << ACC_DEPTH_SHIFT) | ACC_EMPTY;
            } else {
                const int id = atomicAdd(next_count, 1);
                next_roots[id] = (int32_t)node;
                cell = ACC_PTR | (uint32_t)id;
            }
        }
        cells[gid] = cell;
    }
}

// Hit marks: one thread per cell; leaf cells that hold a row get ACC_MISS iff !(sigma > 0) -- the negation of the
// hit predicate of the march (rt_kernel.cu:279 with the default threshold, :382/:456 always), NaN included.
__global__ void __launch_bounds__(256)
accel_mark_kernel(uint32_t* __restrict__ cells, int64_t n, const float* __restrict__ features, int D,
                  const int* __restrict__ n_bricks_dev, int64_t per_brick, const int* __restrict__ only_if) {
    if (only_if && !*only_if) return;
    if (n_bricks_dev) n = min(n, (int64_t)*n_bricks_dev * per_brick);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t cell = cells[i];
        if (cell & ACC_PTR) continue;
        const uint32_t ci = cell & ACC_IDX_MASK;
        if (ci == ACC_EMPTY) continue;
        const float sigma = __ldg(features + (size_t)ci * D + (D - 1));
        const uint32_t marked = (sigma > 0.0f) ? (cell & ~ACC_MISS) : (cell | ACC_MISS);
        if (marked != cell) cells[i] = marked;
    }
}

// ---- per-row activation -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
activate_kernel(const float* __restrict__ f, int64_t n, int D, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(f + i);
        out[i] = ((int)(i % D) == D - 1) ? v : fast_sigmoid(v);
    }
}

// Payload-only rows (stride S = D-1 rounded up to a multiple of 4, zeros in the padding) + the compact sigma array:
// one thread per output float.
__global__ void __launch_bounds__(256)
activate_padded_kernel(const float* __restrict__ f, int64_t M, int D, int S, float* __restrict__ out,
                       float* __restrict__ sigma) {
    const int64_t n = M * S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / S;
        const int c = (int)(i - r * S);
        out[i] = c < D - 1 ? fast_sigmoid(__ldg(f + r * D + c)) : 0.0f;
        if (c == 0) sigma[r] = __ldg(f + r * D + (D - 1));
    }
}

__global__ void __launch_bounds__(256)
activate4_kernel(const float4* __restrict__ f, int64_t n4, int D4, float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(f + i);
        float4 r = make_float4(fast_sigmoid(v.x), fast_sigmoid(v.y), fast_sigmoid(v.z), fast_sigmoid(v.w));
        if ((int)(i % D4) == D4 - 1) r.w = v.w;           // the sigma channel stays raw
        out[i] = r;
    }
}

// ---- per-step table pass (svoxb_prepare_step) ---------------------------------------------------------------------------
// ONE streaming pass over the rows does everything a training step needs before its marches, because the features
// changed: the activated table (as activate4_kernel), the accelerator's hit marks -- through the inverse map
// row -> leaf cell, so no second pass over the cells and no strided sigma gather -- and, optionally, the zero-fill of
// the gradient table the backward reduces into (the reference's zeros_like(features), rt_kernel.cu:1415).
// Reads 4 M D + 4 M bytes, writes 4 M D (+ 4 M D) bytes: C3 0.11 ms against 0.19 ms for the three separate passes.
struct CellPtrs {
    uint32_t* c[MAX_STAGES];
};
// The grid is a multiple of D4 blocks, so the grid stride is a multiple of the row length: a thread keeps its column
// and steps through the rows -- no division in the loop. (Unrolling by four independent loads measured slower:
// 0.144 -> 0.165 ms at C3.)
__global__ void __launch_bounds__(256)
prepare4_kernel(const float4* __restrict__ f, int64_t n4, int D4, float4* __restrict__ act, float4* __restrict__ zero,
                const uint32_t* __restrict__ row_cell, CellPtrs cells, const int* __restrict__ rows_shared_dev) {
    if (rows_shared_dev && *rows_shared_dev) row_cell = nullptr;      // the pass over the cells marks instead
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool is_sigma = (int)(tid % D4) == D4 - 1;
    const int64_t row_step = stride / D4;
    int64_t row = tid / D4;
    for (int64_t i = tid; i < n4; i += stride, row += row_step) {
        const float4 v = __ldcs(f + i);
        float4 r = make_float4(fast_sigmoid(v.x), fast_sigmoid(v.y), fast_sigmoid(v.z), fast_sigmoid(v.w));
        if (is_sigma) {
            r.w = v.w;                                         // the sigma channel stays raw
            if (row_cell) {
                const uint32_t rc = __ldg(row_cell + row);
                if (rc != 0xffffffffu) {
                    uint32_t* cp = cells.c[rc >> 30] + (rc & 0x3fffffffu);
                    const uint32_t cell = *cp;
                    const uint32_t marked = (v.w > 0.0f) ? (cell & ~ACC_MISS) : (cell | ACC_MISS);
                    if (marked != cell) *cp = marked;
                }
            }
        }
        act[i] = r;
        if (zero) __stcs(zero + i, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

// ---- point query ------------------------------------------------------------------------------------------------------
// One lane per point for the descent; the row copy is served by the whole warp (coalesced 4*D-byte rows).
__global__ void __launch_bounds__(256)
query_kernel(TreeArgs tr, const float* __restrict__ pts, int64_t Q, float* __restrict__ values,
             int64_t* __restrict__ node_ids, int64_t* __restrict__ data_ids, uint8_t* __restrict__ slot_mask) {
    const int lane = threadIdx.x & 31;
    const float off[3] = {__ldg(tr.offset), __ldg(tr.offset + 1), __ldg(tr.offset + 2)};
    const float scl[3] = {__ldg(tr.scaling), __ldg(tr.scaling + 1), __ldg(tr.scaling + 2)};
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t base = warp_id * 32; base < Q; base += warps_total * 32) {
        const int64_t q = base + lane;
        int idx = -1;
        if (q < Q) {
            // common.cuh:44-51: one FFMA per axis
            const float px = fmaf(scl[0], __ldg(pts + 3 * q), off[0]);
            const float py = fmaf(scl[1], __ldg(pts + 3 * q + 1), off[1]);
            const float pz = fmaf(scl[2], __ldg(pts + 3 * q + 2), off[2]);
            float rx, ry, rz, cube;
            const int64_t slot = descend_ref(tr.child, tr.N, px, py, pz, rx, ry, rz, cube);
            node_ids[q] = slot;
            if (slot_mask) slot_mask[slot] = 1;                       // svox_kernel.cu:57-58 (empty leaves too)
            const int di = __ldg(tr.data + slot);
            if (di >= 0 && (int64_t)di < tr.M) {                       // svox_kernel.cu:61
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
                const float* src = tr.features + (int64_t)idx_r * tr.D;
                float* dst = values + (base + r) * tr.D;
                for (int c = lane; c < tr.D; c += 32) dst[c] = __ldg(src + c);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
construct_mark_kernel(TreeArgs tr, int32_t* __restrict__ data_mut, const float* __restrict__ pts, int64_t P,
                      int phase) {
    const float off[3] = {__ldg(tr.offset), __ldg(tr.offset + 1), __ldg(tr.offset + 2)};
    const float scl[3] = {__ldg(tr.scaling), __ldg(tr.scaling + 1), __ldg(tr.scaling + 2)};
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < P; q += (int64_t)gridDim.x * blockDim.x) {
        const float px = fmaf(scl[0], __ldg(pts + 3 * q), off[0]);
        const float py = fmaf(scl[1], __ldg(pts + 3 * q + 1), off[1]);
        const float pz = fmaf(scl[2], __ldg(pts + 3 * q + 2), off[2]);
        float rx, ry, rz, cube;
        const int64_t slot = descend_ref(tr.child, tr.N, px, py, pz, rx, ry, rz, cube);
        if (phase == 0) data_mut[slot] = -1;                 // every leaf that receives a point
        else atomicMax(data_mut + slot, (int)q);             // deterministic winner: the largest point index
    }
}

// ---- unique-leaf compaction -------------------------------------------------------------------------------------------------
constexpr int LS_THREADS = 256;
constexpr int LS_PER_THREAD = 16;
constexpr int LS_PER_BLOCK = LS_THREADS * LS_PER_THREAD;

__device__ __forceinline__ unsigned load_mask16(const uint8_t* __restrict__ mask, int64_t n, int64_t start) {
    unsigned bits = 0;
#pragma unroll
    for (int i = 0; i < LS_PER_THREAD; ++i) {
        const int64_t s = start + i;
        if (s < n && mask[s]) bits |= 1u << i;
    }
    return bits;
}

__global__ void __launch_bounds__(LS_THREADS)
leafset_count_kernel(const uint8_t* __restrict__ mask, int64_t n, int64_t* __restrict__ block_counts) {
    __shared__ int wsum[LS_THREADS / 32];
    const int64_t start = ((int64_t)blockIdx.x * LS_THREADS + threadIdx.x) * LS_PER_THREAD;
    int c = __popc(load_mask16(mask, n, start));
    for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(FULL, c, s);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < LS_THREADS / 32; ++i) t += wsum[i];
        block_counts[blockIdx.x] = t;
    }
}

// Single block: exclusive scan of the block counts in place; total -> *n_hit. Thread t owns a contiguous chunk.
__global__ void __launch_bounds__(1024)
leafset_scan_kernel(int64_t* __restrict__ block_counts, int64_t n_blocks, int64_t* __restrict__ n_hit) {
    __shared__ int64_t wtot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (n_blocks + 1023) / 1024;
    const int64_t lo = min(n_blocks, (int64_t)threadIdx.x * per), hi = min(n_blocks, lo + per);
    int64_t sum = 0;
    for (int64_t i = lo; i < hi; ++i) sum += block_counts[i];
    int64_t inc = sum;
    for (int s = 1; s < 32; s <<= 1) {
        const int64_t o = __shfl_up_sync(FULL, inc, s);
        if (lane >= s) inc += o;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int64_t w = wtot[lane];
        int64_t winc = w;
        for (int s = 1; s < 32; s <<= 1) {
            const int64_t o = __shfl_up_sync(FULL, winc, s);
            if (lane >= s) winc += o;
        }
        wtot[lane] = winc - w;            // exclusive offset of each warp
        if (lane == 31) *n_hit = winc;
    }
    __syncthreads();
    int64_t run = wtot[warp] + inc - sum; // exclusive offset of this thread's chunk
    for (int64_t i = lo; i < hi; ++i) {
        const int64_t c = block_counts[i];
        block_counts[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(LS_THREADS)
leafset_emit_kernel(const uint8_t* __restrict__ mask, int64_t n, int N, const int64_t* __restrict__ block_offsets,
                    int64_t* __restrict__ leaf_node) {
    __shared__ int woff[LS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t start = ((int64_t)blockIdx.x * LS_THREADS + threadIdx.x) * LS_PER_THREAD;
    const unsigned bits = load_mask16(mask, n, start);
    const int c = __popc(bits);
    int inc = c;
    for (int s = 1; s < 32; s <<= 1) {
        const int o = __shfl_up_sync(FULL, inc, s);
        if (lane >= s) inc += o;
    }
    if (lane == 31) woff[warp] = inc;
    __syncthreads();
    int wbase = 0;
    for (int i = 0; i < warp; ++i) wbase += woff[i];
    int64_t pos = block_offsets[blockIdx.x] + wbase + inc - c;
    unsigned b = bits;
    while (b) {
        const int i = __ffs(b) - 1;
        b &= b - 1;
        int64_t v = start + i;                       // packed slot id node*N^3 + u*N^2 + v*N + w
        int64_t* o = leaf_node + pos * 4;
        o[3] = v % N; v /= N;
        o[2] = v % N; v /= N;
        o[1] = v % N; v /= N;
        o[0] = v;
        ++pos;
    }
}

}  // namespace svoxb

using namespace svoxb;

// ---- C ABI --------------------------------------------------------------------------------------------------------
extern "C" int svoxb_abi_version(void) { return SVOXB_ABI_VERSION; }
extern "C" const char* svoxb_last_error(void) { return g_err; }
extern "C" int64_t svoxb_launch_count(void) { return g_launches.load(); }

extern "C" int svoxb_device_info(int* sm, int* major, int* minor) {
    int dev = 0;
    SVOXB_CUDA(cudaGetDevice(&dev));
    int a = 0, b = 0, c = 0;
    SVOXB_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
    SVOXB_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
    SVOXB_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm) *sm = a;
    if (major) *major = b;
    if (minor) *minor = c;
    return 0;
}

// Stream-ordered allocation from the device's default pool, which is told to keep its memory: per-frame rebuilds
// then recycle the same blocks instead of paying cudaMalloc/cudaFree (and their device-wide syncs) every frame.
static int pool_alloc(void** p, size_t bytes, cudaStream_t st) {
    static bool tuned[64] = {false};
    int dev = 0;
    SVOXB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !tuned[dev]) {
        cudaMemPool_t pool;
        SVOXB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = ~0ull;
        SVOXB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        tuned[dev] = true;
    }
    return check_cuda(cudaMallocAsync(p, bytes ? bytes : 4, st), "cudaMallocAsync");
}

namespace svoxb {
int scratch_alloc(void** p, size_t bytes, cudaStream_t st) { return pool_alloc(p, bytes, st); }
}  // namespace svoxb

static int* pinned_scalars() {
    static thread_local int* h = nullptr;
    if (!h && cudaHostAlloc(&h, sizeof(int) * 8, cudaHostAllocDefault) != cudaSuccess) h = nullptr;
    return h;
}

extern "C" void svoxb_accel_destroy(svoxb_accel* a) {
    if (!a) return;
    // released in stream order on the creation stream: work that still reads the accelerator on that stream
    // finishes first; work launched through this library on other streams is waited for with an event each.
    for (int i = 0; i < a->n_used; ++i) {
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) continue;
        if (cudaEventRecord(ev, a->used[i]) == cudaSuccess) cudaStreamWaitEvent(a->stream, ev, 0);
        else cudaGetLastError();        // the stream no longer exists: its work has finished
        cudaEventDestroy(ev);
    }
    for (int s = 0; s < MAX_STAGES; ++s)
        if (a->cells[s]) cudaFreeAsync(a->cells[s], a->stream);
    if (a->row_cell) cudaFreeAsync(a->row_cell, a->stream);
    if (a->d_flags) cudaFreeAsync(a->d_flags, a->stream);
    if (a->roots) cudaFreeAsync(a->roots, a->stream);
    delete a;
}

extern "C" int64_t svoxb_accel_bytes(const svoxb_accel* a) { return a ? a->bytes : 0; }

// Built without read-backs: fetch the brick counts and flags now (synchronises the creation stream).
static int accel_resolve(svoxb_accel* a) {
    if (!a->lazy) return 0;
    int h[8];
    SVOXB_CUDA(cudaMemcpyAsync(h, a->d_flags, sizeof(h), cudaMemcpyDeviceToHost, a->stream));
    SVOXB_CUDA(cudaStreamSynchronize(a->stream));
    a->lazy = 0;
    for (int s = 1; s < a->view.n_stages; ++s) a->n_bricks[s] = min(a->n_bricks[s], (int64_t)h[2 + s - 1]);
    if (h[6]) a->rows_shared = 1;
    SVOXB_REQUIRE(!h[1], "tree is deeper than max_depth=%d", a->view.lmax);
    return 0;
}

extern "C" int svoxb_accel_describe(const svoxb_accel* a_in, int* n_stages, int* bits, int64_t* bricks) {
    SVOXB_REQUIRE(a_in != nullptr, "accel is NULL");
    svoxb_accel* a = const_cast<svoxb_accel*>(a_in);
    if (int rc = accel_resolve(a)) return rc;
    if (n_stages) *n_stages = a->view.n_stages;
    for (int s = 0; s < MAX_STAGES; ++s) {
        if (bits) bits[s] = s < a->view.n_stages ? a->view.bits[s] : 0;
        if (bricks) bricks[s] = s < a->view.n_stages ? a->n_bricks[s] : 0;
    }
    return 0;
}

// (Re)fill an accelerator of at most two stages without any read-back: stage 1 is sized for the most bricks the top
// grid can point to, its kernel takes the actual count from device memory, and the flags (overflow, shared rows) stay
// on the device (svoxb_accel_describe fetches them on demand). Buffers already present and large enough are reused.
static int accel_fill_nosync(svoxb_accel* a, const svoxb_tree* tree, cudaStream_t st) {
    AccelView& v = a->view;
    int rc;
    a->M = tree->M; a->child = tree->child; a->data = tree->data;
    a->lazy = 1; a->rows_shared = 0; a->marks_features = nullptr;
    SVOXB_CUDA(cudaMemsetAsync(a->d_flags, 0, sizeof(int) * 8, st));
    if (tree->M > a->row_cap) {
        if (a->row_cell) { cudaFreeAsync(a->row_cell, st); a->bytes -= (int64_t)sizeof(uint32_t) * a->row_cap; a->row_cell = nullptr; }
        if ((rc = pool_alloc((void**)&a->row_cell, sizeof(uint32_t) * (size_t)tree->M, st))) return rc;
        a->row_cap = tree->M;
        a->bytes += (int64_t)sizeof(uint32_t) * tree->M;
    }
    if (tree->M > 0) SVOXB_CUDA(cudaMemsetAsync(a->row_cell, 0xff, sizeof(uint32_t) * (size_t)tree->M, st));
    if (v.n_stages > 1 && tree->n_internal > a->roots_cap) {
        if (a->roots) cudaFreeAsync(a->roots, st);
        a->roots = nullptr;
        if ((rc = pool_alloc((void**)&a->roots, sizeof(int32_t) * (size_t)tree->n_internal, st))) return rc;
        a->roots_cap = tree->n_internal;
    }
    int64_t bound = 1;
    int base = 0;
    for (int s = 0; s < v.n_stages; ++s) {
        const int64_t words = bound << (3 * v.bits[s]);
        if (!a->cells[s] || a->n_bricks[s] < bound) {
            if (a->cells[s]) { cudaFreeAsync(a->cells[s], st); a->bytes -= (int64_t)sizeof(uint32_t) * (a->n_bricks[s] << (3 * v.bits[s])); }
            a->cells[s] = nullptr;
            if ((rc = pool_alloc((void**)&a->cells[s], sizeof(uint32_t) * (size_t)words, st))) return rc;
            a->bytes += (int64_t)sizeof(uint32_t) * words;
        }
        a->n_bricks[s] = bound;
        v.cells[s] = a->cells[s];
        const int is_last = (s == v.n_stages - 1);
        const int grid = (int)min((words + 255) / 256, (int64_t)sm_count() * 16);
        accel_stage_kernel<<<grid, 256, 0, st>>>(tree->child, tree->data, tree->M, s == 0 ? nullptr : a->roots, bound, v.bits[s],
                                                 base, a->cells[s], is_last ? nullptr : a->roots, a->d_flags + 2 + s, is_last,
                                                 a->d_flags + 1, tree->M > 0 ? a->row_cell : nullptr, (uint32_t)s << 30,
                                                 a->d_flags + 6, s == 0 ? nullptr : a->d_flags + 2 + s - 1);
        count_launch();
        if ((rc = check_cuda(cudaGetLastError(), "accel_stage_kernel launch"))) return rc;
        bound = min((int64_t)1 << (3 * v.bits[s]), tree->n_internal);      // pointers a grid of 8^bits cells can hold
        base += v.bits[s];
    }
    return 0;
}

extern "C" int svoxb_accel_rebuild(svoxb_accel* a, const svoxb_tree* tree, int max_depth, void* stream) {
    SVOXB_REQUIRE(a != nullptr && tree != nullptr && tree->child && tree->data, "NULL argument");
    if (!a->inplace || tree->N != 2 || max_depth != a->view.lmax || tree->M >= (int64_t)ACC_EMPTY ||
        (cudaStream_t)stream != a->stream) {
        set_error("this accelerator cannot be refilled in place for this tree (create a new one)");
        return SVOXB_EUNSUPPORTED;
    }
    return accel_fill_nosync(a, tree, (cudaStream_t)stream);
}

extern "C" int svoxb_accel_create(const svoxb_tree* tree, int max_depth, void* stream, svoxb_accel** out) {
    SVOXB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    SVOXB_REQUIRE(tree != nullptr && tree->child && tree->data, "tree has NULL tensors");
    if (tree->N != 2) { set_error("accelerator requires N == 2 (got %d)", tree->N); return SVOXB_EUNSUPPORTED; }
    if (tree->M >= (int64_t)ACC_EMPTY) {
        set_error("accelerator supports M < %u feature rows (got %lld)", ACC_EMPTY, (long long)tree->M);
        return SVOXB_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int* h_scalars = pinned_scalars();
    SVOXB_REQUIRE(h_scalars != nullptr, "cudaHostAlloc failed");
    int* d_scalars = nullptr;     // [0] max depth, [1] overflow, [2..] brick counts per stage
    int rc = pool_alloc((void**)&d_scalars, sizeof(int) * 8, st);
    if (rc) return rc;
    SVOXB_CUDA(cudaMemsetAsync(d_scalars, 0, sizeof(int) * 8, st));
    int lmax = max_depth;
    if (lmax <= 0) {
        if (!tree->parent_depth) {
            cudaFreeAsync(d_scalars, st);
            set_error("max_depth <= 0 needs tree->parent_depth");
            return SVOXB_EINVAL;
        }
        const int grid = (int)min((tree->n_internal + 255) / 256, (int64_t)1024);
        max_depth_kernel<<<grid, 256, 0, st>>>(tree->parent_depth, tree->n_internal, d_scalars);
        count_launch();
        SVOXB_CUDA(cudaMemcpyAsync(h_scalars, d_scalars, sizeof(int), cudaMemcpyDeviceToHost, st));
        SVOXB_CUDA(cudaStreamSynchronize(st));
        lmax = h_scalars[0] + 1;
    }
    if (lmax > ACC_MAX_DEPTH) {
        cudaFreeAsync(d_scalars, st);
        set_error("accelerator supports depth <= %d (tree depth %d)", ACC_MAX_DEPTH, lmax);
        return SVOXB_EUNSUPPORTED;
    }
    svoxb_accel* a = new svoxb_accel();
    memset(a, 0, sizeof(*a));
    a->M = tree->M; a->child = tree->child; a->data = tree->data; a->stream = st;
    // stage split: top grid of <= 4 levels (16 KB of shared memory), the rest in bricks of <= 4 levels
    AccelView& v = a->view;
    v.lmax = lmax;
    const int b0 = lmax < 4 ? lmax : 4;
    const int rem = lmax - b0;
    const int more = (rem + 3) / 4;
    v.n_stages = 1 + more;
    v.bits[0] = b0;
    for (int s = 0; s < more; ++s) v.bits[1 + s] = rem / more + (s < rem % more ? 1 : 0);
    int resolved = 0;
    for (int s = 0; s < v.n_stages; ++s) { resolved += v.bits[s]; v.shift[s] = lmax - resolved; }

    // Known depth and at most two stages (depth <= 8): NO read-back (accel_fill_nosync). A per-frame rebuild then never
    // synchronises, and svoxb_accel_rebuild refills the same allocations frame after frame.
    if (max_depth > 0 && v.n_stages <= 2 && !getenv("SVOXB_ACCEL_SYNC")) {
        a->d_flags = d_scalars;
        a->inplace = 1;
        rc = accel_fill_nosync(a, tree, st);
        if (rc) { svoxb_accel_destroy(a); return rc; }
        *out = a;
        return 0;
    }
    int32_t* roots[2] = {nullptr, nullptr};
    auto fail = [&](int code) {
        if (roots[0]) cudaFreeAsync(roots[0], st);
        if (roots[1]) cudaFreeAsync(roots[1], st);
        cudaFreeAsync(d_scalars, st);
        svoxb_accel_destroy(a);
        return code;
    };
    if (v.n_stages > 1) {
        for (int i = 0; i < 2; ++i)
            if ((rc = pool_alloc((void**)&roots[i], sizeof(int32_t) * (size_t)tree->n_internal, st))) return fail(rc);
    }
    if (tree->M > 0) {      // inverse map row -> cell, filled by the stage kernels
        if ((rc = pool_alloc((void**)&a->row_cell, sizeof(uint32_t) * (size_t)tree->M, st))) return fail(rc);
        if ((rc = check_cuda(cudaMemsetAsync(a->row_cell, 0xff, sizeof(uint32_t) * (size_t)tree->M, st), "memset"))) return fail(rc);
        a->bytes += (int64_t)sizeof(uint32_t) * tree->M;
    }
    int64_t n_bricks = 1;
    int base_depth = 0;
    for (int s = 0; s < v.n_stages; ++s) {
        const int64_t words = n_bricks << (3 * v.bits[s]);
        if (words >= (1ll << 30)) a->rows_shared = 1;
        a->n_bricks[s] = n_bricks;
        if ((rc = pool_alloc((void**)&a->cells[s], sizeof(uint32_t) * (size_t)max(words, (int64_t)1), st))) return fail(rc);
        a->bytes += (int64_t)sizeof(uint32_t) * words;
        v.cells[s] = a->cells[s];
        const int is_last = (s == v.n_stages - 1);
        if (words > 0) {
            const int grid = (int)min((words + 255) / 256, (int64_t)sm_count() * 16);
            accel_stage_kernel<<<grid, 256, 0, st>>>(tree->child, tree->data, tree->M, s == 0 ? nullptr : roots[(s - 1) & 1],
                                                     n_bricks, v.bits[s], base_depth, a->cells[s],
                                                     is_last ? nullptr : roots[s & 1], d_scalars + 2 + s, is_last,
                                                     d_scalars + 1, words < (1ll << 30) ? a->row_cell : nullptr,
                                                     (uint32_t)s << 30, d_scalars + 6, nullptr);
            count_launch();
            if ((rc = check_cuda(cudaGetLastError(), "accel_stage_kernel launch"))) return fail(rc);
        }
        // the brick count of the next stage sizes its allocation: one small pinned read-back per stage
        if ((rc = check_cuda(cudaMemcpyAsync(h_scalars, d_scalars, sizeof(int) * 8, cudaMemcpyDeviceToHost, st), "memcpy")))
            return fail(rc);
        if ((rc = check_cuda(cudaStreamSynchronize(st), "accel build sync"))) return fail(rc);
        n_bricks = is_last ? 0 : h_scalars[2 + s];
        base_depth += v.bits[s];
    }
    if (h_scalars[1]) {
        set_error("tree is deeper than max_depth=%d", lmax);
        return fail(SVOXB_EINVAL);
    }
    if (h_scalars[6]) a->rows_shared = 1;
    if (roots[0]) cudaFreeAsync(roots[0], st);
    if (roots[1]) cudaFreeAsync(roots[1], st);
    cudaFreeAsync(d_scalars, st);
    *out = a;
    return 0;
}

// The pass over the cells. `only_if_shared` (accelerators built without read-backs): the kernels run only if the
// device-side flag says a row is held by several cells -- otherwise the table pass has marked through the inverse map.
static int mark_cells(svoxb_accel* a, const float* features, int64_t M, int32_t D, bool only_if_shared, cudaStream_t st) {
    for (int s = 0; s < a->view.n_stages && M > 0; ++s) {
        const int64_t per = (int64_t)1 << (3 * a->view.bits[s]);
        const int64_t words = a->n_bricks[s] * per;
        if (words <= 0) continue;
        const int grid = (int)min((words + 255) / 256, (int64_t)sm_count() * 16);
        accel_mark_kernel<<<grid, 256, 0, st>>>(a->cells[s], words, features, D,
                                                (a->lazy && s > 0) ? a->d_flags + 2 + s - 1 : nullptr, per,
                                                only_if_shared ? a->d_flags + 6 : nullptr);
        count_launch();
        SVOXB_CUDA(cudaGetLastError());
    }
    a->marks_features = features; a->marks_D = D;
    return 0;
}

extern "C" int svoxb_accel_mark_hits(svoxb_accel* a, const float* features, int64_t M, int32_t D, void* stream) {
    SVOXB_REQUIRE(a != nullptr, "accel is NULL");
    SVOXB_REQUIRE(M == a->M, "accelerator was built for M=%lld, got M=%lld", (long long)a->M, (long long)M);
    SVOXB_REQUIRE(D >= 2 && (features != nullptr || M == 0), "bad feature table");
    return mark_cells(a, features, M, D, false, (cudaStream_t)stream);
}

extern "C" int svoxb_activate_features(const float* features, int64_t M, int32_t D, float* out, int32_t out_stride,
                                       float* sigma_out, void* stream) {
    SVOXB_REQUIRE(M >= 0 && D >= 2, "bad sizes");
    if (out_stride <= 0) out_stride = D;
    const int payload = (D - 1 + 3) / 4 * 4;
    SVOXB_REQUIRE(out_stride == D || (out_stride == payload && sigma_out != nullptr),
                  "out_stride=%d must be D, or %d (payload only) together with sigma_out", out_stride, payload);
    if (M == 0) return 0;
    SVOXB_REQUIRE(features && out, "NULL tensor");
    const int64_t n = M * D;
    cudaStream_t st = (cudaStream_t)stream;
    if (sigma_out != nullptr && out_stride == payload) {
        const int grid = (int)min((M * out_stride + 255) / 256, (int64_t)sm_count() * 16);
        activate_padded_kernel<<<grid, 256, 0, st>>>(features, M, D, out_stride, out, sigma_out);
    } else if (D % 4 == 0 && (((uintptr_t)features | (uintptr_t)out) & 15) == 0) {
        const int grid = (int)min((n / 4 + 255) / 256, (int64_t)sm_count() * 16);
        activate4_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(features), n / 4, D / 4,
                                              reinterpret_cast<float4*>(out));
    } else {
        const int grid = (int)min((n + 255) / 256, (int64_t)sm_count() * 16);
        activate_kernel<<<grid, 256, 0, st>>>(features, n, D, out);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "activate_kernel launch");
}

__global__ void __launch_bounds__(256)
gather_sigma_kernel(const float* __restrict__ f, int64_t M, int D, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __ldg(f + i * D + (D - 1));
}

extern "C" int svoxb_gather_sigma(const float* features, int64_t M, int32_t D, float* sigma_out, void* stream) {
    SVOXB_REQUIRE(M >= 0 && D >= 2, "bad sizes");
    if (M == 0) return 0;
    SVOXB_REQUIRE(features && sigma_out, "NULL tensor");
    const int grid = (int)min((M + 255) / 256, (int64_t)sm_count() * 16);
    gather_sigma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(features, M, D, sigma_out);
    count_launch();
    return check_cuda(cudaGetLastError(), "gather_sigma_kernel launch");
}

extern "C" int svoxb_prepare_step(svoxb_accel* a, const float* features, int64_t M, int32_t D, float* act,
                                  int32_t act_stride, float* sigma_out, float* zero_table, void* stream) {
    SVOXB_REQUIRE(M >= 0 && D >= 2, "bad sizes");
    SVOXB_REQUIRE(a == nullptr || M == a->M, "accelerator was built for another M");
    if (M == 0) return 0;
    SVOXB_REQUIRE(features && act, "NULL tensor");
    cudaStream_t st = (cudaStream_t)stream;
    if (act_stride <= 0) act_stride = D;
    const bool fused = D % 4 == 0 && act_stride == D &&
                       (((uintptr_t)features | (uintptr_t)act | (uintptr_t)zero_table) & 15) == 0;
    if (!fused) {            // odd widths / unaligned tables: the separate passes
        int rc = svoxb_activate_features(features, M, D, act, act_stride, sigma_out, stream);
        if (rc == 0 && a) rc = svoxb_accel_mark_hits(a, features, M, D, stream);
        if (rc == 0 && zero_table) rc = check_cuda(cudaMemsetAsync(zero_table, 0, sizeof(float) * (size_t)M * D, st), "memset");
        return rc;
    }
    const bool marks_here = a != nullptr && a->row_cell != nullptr && !a->rows_shared;      // (lazy: decided on the device)
    CellPtrs cp;
    for (int s = 0; s < MAX_STAGES; ++s) cp.c[s] = a ? a->cells[s] : nullptr;
    const int64_t n4 = M * D / 4;
    int grid = (int)min((n4 + 255) / 256, (int64_t)sm_count() * 16);
    grid = (grid + D / 4 - 1) / (D / 4) * (D / 4);          // grid stride = a whole number of rows (see the kernel)
    prepare4_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(features), n4, D / 4,
                                          reinterpret_cast<float4*>(act), reinterpret_cast<float4*>(zero_table),
                                          marks_here ? a->row_cell : nullptr, cp, (marks_here && a->lazy) ? a->d_flags + 6 : nullptr);
    count_launch();
    int rc = check_cuda(cudaGetLastError(), "prepare4_kernel launch");
    if (rc) return rc;
    if (a) {
        if (!marks_here) rc = svoxb_accel_mark_hits(a, features, M, D, stream);
        else if (a->lazy) rc = mark_cells(a, features, M, D, true, st);        // no-ops unless rows turn out to be shared
        else { a->marks_features = features; a->marks_D = D; }
    }
    return rc;
}

extern "C" int svoxb_query(const svoxb_tree* tree, const float* pts, int64_t Q, float* values, int64_t* node_ids,
                           int64_t* data_ids, uint8_t* slot_mask, void* stream) {
    TreeArgs tr;
    int rc = make_tree_args(tree, tr); if (rc) return rc;
    SVOXB_REQUIRE(Q >= 0 && (Q == 0 || (pts && node_ids)), "pts/node_ids NULL");
    if (Q == 0) return 0;
    const int grid = (int)min((Q + 255) / 256, (int64_t)sm_count() * 8);
    query_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tr, pts, Q, values, node_ids, data_ids, slot_mask);
    count_launch();
    return check_cuda(cudaGetLastError(), "query_kernel launch");
}

extern "C" size_t svoxb_leafset_scratch_bytes(int64_t n_slots) {
    const int64_t n_blocks = (n_slots + LS_PER_BLOCK - 1) / LS_PER_BLOCK;
    return sizeof(int64_t) * (size_t)(n_blocks + 1);
}

extern "C" int svoxb_leafset_scan(const uint8_t* slot_mask, int64_t n_slots, void* scratch, int64_t* n_hit_dev,
                                  void* stream) {
    SVOXB_REQUIRE(slot_mask && scratch && n_hit_dev && n_slots > 0, "leafset_scan: bad arguments");
    const int64_t n_blocks = (n_slots + LS_PER_BLOCK - 1) / LS_PER_BLOCK;
    cudaStream_t st = (cudaStream_t)stream;
    leafset_count_kernel<<<(int)n_blocks, LS_THREADS, 0, st>>>(slot_mask, n_slots, (int64_t*)scratch);
    leafset_scan_kernel<<<1, 1024, 0, st>>>((int64_t*)scratch, n_blocks, n_hit_dev);
    count_launch(2);
    return check_cuda(cudaGetLastError(), "leafset_scan launch");
}

extern "C" int svoxb_leafset_emit(const uint8_t* slot_mask, int64_t n_slots, int32_t N, const void* scratch,
                                  int64_t* leaf_node, void* stream) {
    SVOXB_REQUIRE(slot_mask && scratch && leaf_node && n_slots > 0 && N >= 2, "leafset_emit: bad arguments");
    const int64_t n_blocks = (n_slots + LS_PER_BLOCK - 1) / LS_PER_BLOCK;
    leafset_emit_kernel<<<(int)n_blocks, LS_THREADS, 0, (cudaStream_t)stream>>>(slot_mask, n_slots, N,
                                                                                 (const int64_t*)scratch, leaf_node);
    count_launch();
    return check_cuda(cudaGetLastError(), "leafset_emit launch");
}

extern "C" int svoxb_construct_tree(const svoxb_tree* tree, int32_t* data_mut, const float* pts, int64_t P,
                                    void* stream) {
    TreeArgs tr;
    int rc = make_tree_args(tree, tr); if (rc) return rc;
    SVOXB_REQUIRE(data_mut != nullptr && (P == 0 || pts != nullptr), "data/pts NULL");
    SVOXB_REQUIRE(P >= 0 && P < (1ll << 31), "point count out of range");
    if (P == 0) return 0;
    const int grid = (int)min((P + 255) / 256, (int64_t)sm_count() * 8);
    cudaStream_t st = (cudaStream_t)stream;
    construct_mark_kernel<<<grid, 256, 0, st>>>(tr, data_mut, pts, P, 0);
    construct_mark_kernel<<<grid, 256, 0, st>>>(tr, data_mut, pts, P, 1);
    count_launch(2);
    return check_cuda(cudaGetLastError(), "construct_tree launch");
}
