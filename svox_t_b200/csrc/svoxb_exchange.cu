// svoxb_exchange.cu -- the one exchange step of the path: sum of the leaf-gradient table grad[M, D] over the GPUs of a
// node (the table is the reference's zeros_like(features), rt_kernel.cu:1415 -- the reference itself is
// single-GPU, so there is nothing to cite for the collective).
//
// One kernel per rank, in place on a SYMMETRIC buffer (the same allocation mapped on every GPU of the node; the
// caller sets it up -- torch.distributed._symmetric_memory in svox_t_b200/dist.py -- and passes raw addresses):
//   1. flag barrier with the peers (every rank's backward has finished writing its table);
//   2. rank r owns the r-th slice of the table. NVLS form (multicast address given): one
//      multimem.ld_reduce.add.v4.f32 per 16 bytes pulls that piece from ALL ranks and adds it INSIDE the NVSwitch, one
//      multimem.st.v4.f32 pushes the sum back to ALL ranks -- each GPU sends and receives the table once (2 x 243 MB
//      over its 900 GB/s links at C3), against (W-1)/W x 4 tables for a ring. Peer-to-peer form (no multicast): the
//      owner loads the piece from each peer's mapping, adds in rank order and stores it to each peer;
//   3. flag barrier again (every peer's stores into this rank's table have landed) -- the kernel's end is the
//      "gradient is summed" event of the stream.
// The flags live behind the table in the same symmetric allocation and carry a monotonically increasing epoch, so
// they never need resetting; a spin that exceeds its time budget records a status code and gives up (no hang).
#include <type_traits>
#include "svoxb_common.cuh"

namespace svoxb {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ float4 multimem_ld_sum(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_peer(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer(float* p, const float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

constexpr int XCH_MAX_WORLD = 16;
constexpr int XCH_THREADS = 512;
constexpr int XCH_UNROLL = 4;
constexpr unsigned long long XCH_SPIN_BUDGET_NS = 4000000000ull;   // 4 s: a peer that never arrives

struct ExchangeArgs {
    char* peer[XCH_MAX_WORLD];      // this rank's mappings of every rank's buffer (peer[rank] = the local one)
    char* multicast;                // multicast mapping of the buffer, or nullptr
    int64_t table_off;              // byte offset of the table inside the buffer
    int64_t flags_off;              // byte offset of the flag words (gridDim.x * world u32) inside the buffer
    int64_t status_off;             // byte offset of one u32 status word (local use)
    int64_t n_vec;                  // table length in float4
    int rank, world;
    uint32_t epoch;                 // this call's first barrier uses `epoch`, the second `epoch + 1`
    const float* live;              // ROWS form: sigma of row r at live[r * live_stride] (this rank's own feature table)
    int64_t live_stride;            // in floats
    int row_vec;                    // float4 per table row (D / 4)
};

// Block b of every rank meets block b of every other rank: thread t < world publishes `epoch` into slot
// [b][rank] of rank t's flags and waits for slot [b][t] of its own flags to reach it.
__device__ __forceinline__ bool peer_barrier(const ExchangeArgs& a, uint32_t epoch) {
    __shared__ int ok_s;
    __syncthreads();
    if (threadIdx.x == 0) ok_s = 1;
    __syncthreads();
    const int t = threadIdx.x;
    if (t < a.world) {
        __threadfence_system();
        uint32_t* theirs = reinterpret_cast<uint32_t*>(a.peer[t] + a.flags_off) + (size_t)blockIdx.x * a.world + a.rank;
        st_release_sys(theirs, epoch);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(a.peer[a.rank] + a.flags_off) + (size_t)blockIdx.x * a.world + t;
        const unsigned long long t0 = global_ns();
        // epochs only grow: ">=" in wrap-around arithmetic lets a fast peer run one barrier ahead
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            if (global_ns() - t0 > XCH_SPIN_BUDGET_NS) {
                atomicExch(reinterpret_cast<uint32_t*>(a.peer[a.rank] + a.status_off), 1u + (uint32_t)t);
                ok_s = 0;
                break;
            }
        }
    }
    __syncthreads();
    return ok_s != 0;
}

// ROWS: the table is the gradient table grad[M, D] a backward produced for the feature table `live` points into. The
// backward's hit predicate is sigma > 0 (rt_kernel.cu:382, 456), so a row with !(sigma > 0) has received no gradient
// on ANY rank (the features are replicated): it is zero everywhere already and is neither reduced nor stored -- at C3
// 20 % of the rows, i.e. 20 % less NVLink traffic. The sigma reads are local (one 32-byte sector per row).
template <bool NVLS, bool ROWS>
__global__ void __launch_bounds__(XCH_THREADS) exchange_sum_kernel(ExchangeArgs a) {
    if (!peer_barrier(a, a.epoch)) return;
    // this rank's slice [lo, hi) of the float4 indices, split evenly over the blocks
    const int64_t per = (a.n_vec + a.world - 1) / a.world;
    const int64_t lo = min(a.n_vec, per * a.rank), hi = min(a.n_vec, lo + per);
    const int64_t stride = (int64_t)gridDim.x * XCH_THREADS;
    int64_t i = lo + (int64_t)blockIdx.x * XCH_THREADS + threadIdx.x;
    auto live = [&](int64_t idx) -> bool {
        if constexpr (!ROWS) return true;
        else return __ldg(a.live + (idx / a.row_vec) * a.live_stride) > 0.0f;
    };
    if constexpr (NVLS) {
        float* mc = reinterpret_cast<float*>(a.multicast + a.table_off);
        for (; i + (XCH_UNROLL - 1) * stride < hi; i += XCH_UNROLL * stride) {
            float4 v[XCH_UNROLL];
            bool on[XCH_UNROLL];
#pragma unroll
            for (int u = 0; u < XCH_UNROLL; ++u) on[u] = live(i + u * stride);
#pragma unroll
            for (int u = 0; u < XCH_UNROLL; ++u)
                if (on[u]) v[u] = multimem_ld_sum(mc + 4 * (i + u * stride));
#pragma unroll
            for (int u = 0; u < XCH_UNROLL; ++u)
                if (on[u]) multimem_st(mc + 4 * (i + u * stride), v[u]);
        }
        for (; i < hi; i += stride)
            if (live(i)) multimem_st(mc + 4 * i, multimem_ld_sum(mc + 4 * i));
    } else {
        // U pieces x W peers: all loads of a trip are issued before the first add (the loop is bound by the bytes in
        // flight per SM over the ~2 us NVLink round trip), sums in rank order: every rank receives the same bits
        auto trip = [&](auto U_tag, int64_t i0) {
            constexpr int U = decltype(U_tag)::value;
            float4 s[U];
            bool on[U];
#pragma unroll
            for (int u = 0; u < U; ++u) on[u] = live(i0 + u * stride);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (on[u]) s[u] = ld_peer(reinterpret_cast<const float*>(a.peer[0] + a.table_off) + 4 * (i0 + u * stride));
#pragma unroll 4
            for (int p = 1; p < a.world; ++p) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (on[u]) v[u] = ld_peer(reinterpret_cast<const float*>(a.peer[p] + a.table_off) + 4 * (i0 + u * stride));
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (on[u]) { s[u].x += v[u].x; s[u].y += v[u].y; s[u].z += v[u].z; s[u].w += v[u].w; }
            }
            for (int p = 0; p < a.world; ++p)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (on[u]) st_peer(reinterpret_cast<float*>(a.peer[p] + a.table_off) + 4 * (i0 + u * stride), s[u]);
        };
        for (; i + (XCH_UNROLL - 1) * stride < hi; i += XCH_UNROLL * stride) trip(std::integral_constant<int, XCH_UNROLL>{}, i);
        for (; i < hi; i += stride) trip(std::integral_constant<int, 1>{}, i);
    }
    peer_barrier(a, a.epoch + 1u);
}

}  // namespace svoxb

using namespace svoxb;

extern "C" {

SVOXB_API int svoxb_exchange_max_blocks(void) { return sm_count(); }

static int exchange_launch(const svoxb_peer_group* g, int64_t n_floats, const float* live, int64_t live_stride,
                           int row_vec, void* stream) {
    SVOXB_REQUIRE(g != nullptr && g->buffers != nullptr, "peer group is NULL");
    SVOXB_REQUIRE(g->world >= 1 && g->world <= XCH_MAX_WORLD && g->rank >= 0 && g->rank < g->world,
                  "peer group: rank %d / world %d out of range (world <= %d)", g->rank, g->world, XCH_MAX_WORLD);
    SVOXB_REQUIRE(n_floats >= 0 && n_floats % 4 == 0 && g->table_offset % 16 == 0,
                  "the table must be a whole number of 16-byte pieces at a 16-byte aligned offset");
    SVOXB_REQUIRE(g->blocks >= 1 && g->blocks <= sm_count(), "peer group: blocks=%d must be in [1, %d] (all blocks must be co-resident)",
                  g->blocks, sm_count());
    SVOXB_REQUIRE(g->flags_offset % 4 == 0 && g->status_offset % 4 == 0, "flag / status words must be 4-byte aligned");
    if (g->world == 1 || n_floats == 0) return SVOXB_OK;
    ExchangeArgs a;
    for (int p = 0; p < g->world; ++p) {
        SVOXB_REQUIRE(g->buffers[p] != nullptr, "peer group: buffer of rank %d is NULL", p);
        a.peer[p] = static_cast<char*>(g->buffers[p]);
    }
    a.multicast = static_cast<char*>(g->multicast);
    a.table_off = g->table_offset; a.flags_off = g->flags_offset; a.status_off = g->status_offset;
    a.n_vec = n_floats / 4; a.rank = g->rank; a.world = g->world; a.epoch = g->epoch;
    a.live = live; a.live_stride = live_stride; a.row_vec = row_vec;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (a.multicast) {
        if (live) exchange_sum_kernel<true, true><<<g->blocks, XCH_THREADS, 0, st>>>(a);
        else exchange_sum_kernel<true, false><<<g->blocks, XCH_THREADS, 0, st>>>(a);
    } else {
        if (live) exchange_sum_kernel<false, true><<<g->blocks, XCH_THREADS, 0, st>>>(a);
        else exchange_sum_kernel<false, false><<<g->blocks, XCH_THREADS, 0, st>>>(a);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "exchange_sum_kernel launch");
}

SVOXB_API int svoxb_exchange_sum(const svoxb_peer_group* g, int64_t n_floats, void* stream) {
    return exchange_launch(g, n_floats, nullptr, 0, 1, stream);
}

SVOXB_API int svoxb_exchange_sum_rows(const svoxb_peer_group* g, int64_t M, int32_t D, const float* features, void* stream) {
    SVOXB_REQUIRE(M >= 0 && D >= 2, "bad table shape");
    if (D % 4 != 0 || features == nullptr) return exchange_launch(g, (M * D + 3) / 4 * 4, nullptr, 0, 1, stream);
    return exchange_launch(g, M * D, features + (D - 1), D, D / 4, stream);
}

}  // extern "C"
