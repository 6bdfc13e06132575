// svoxb_build_dense.cu -- sort-free one-shot octree build for the per-frame rebuild (depth L <= 10), with no host
// read-back anywhere: the whole frame (warp -> splat -> rebuild -> accelerator -> render) can be queued without a sync.
//
// Replaces the same reference loop as svoxb_build.cu (depth-1 rounds of query_vertical + N3Tree.refine, then
// construct_tree: svox_kernel.cu:274-324, 110-121; svox_t/svox.py:488-560; helpers.py:38-109) and produces the SAME
// tensors as that file's sort-based build, bit for bit (breadth-first numbering, Morton order within a level).
//
// Idea: the finest level has only 8^L cells (2^24 at L = 8), so occupancy fits a BITMAP (2 MB at L = 8), and with
// Morton keys the eight children of a cell are eight consecutive bits -- one byte:
//   1. every point sets its finest cell's bit (atomicOr into an L2-resident bitmap);
//   2. pyramid: bit c of level l = (byte c of level l+1 != 0), level by level; while a level is produced its words
//      are pop-counted and prefix-summed inside 1024-word chunks (rank directory);
//   3. one single-CTA kernel scans the chunk totals, finishes the small top levels (<= 8192 words) in one go and
//      derives the per-level node counts, the breadth-first bases, the node total and the overflow flag ON THE DEVICE;
//   4. emit: node of an occupied level-l cell = base[l] + rank_l(cell) (rank = chunk base + in-chunk prefix + popcount
//      of the lower bits of its word) -> child / parent_depth in the reference's tensor format;
//   5. every point finds its leaf's parent node by the same rank query and atomicMax-es its index into the data slot
//      (construct_tree with "largest index wins", as svoxb_construct_tree), then empty slots get the reference's
//      sentinel.
// No sort, no scan library: everything is hand-written. The caller passes a node CAPACITY (tensors sized for it);
// the number of nodes needed and an overflow flag stay on the device (read them whenever convenient).
#include "svoxb_common.cuh"

namespace svoxb {

constexpr int DENSE_MAX_L = 10;
constexpr int32_t DENSE_EMPTY = 1410065408;       // int(1e10) wrapped to int32 (svox_t/svox.py:124)
constexpr int CHUNK_WORDS = 1024;                  // rank directory granularity = one CTA of the pyramid kernel
constexpr int SMALL_WORDS = 8192;                  // levels up to this many words are finished by the top kernel

struct DenseLevels {
    // level l (1..L): bitmap of 8^l bits; in-chunk exclusive prefix of the word popcounts; exclusive chunk bases
    uint32_t* bits[DENSE_MAX_L + 1];
    uint32_t* pref[DENSE_MAX_L + 1];
    uint32_t* cbase[DENSE_MAX_L + 1];
    int64_t words[DENSE_MAX_L + 1];
    int64_t* meta;      // [0] nodes needed, [1] overflow, [2 + l] base[l] (first node id of level l; base[0] = 0 = root)
    int L;
};

static inline size_t up256(size_t x) { return (x + 255) & ~size_t(255); }

static inline int64_t level_words(int l) { return max((int64_t)1, ((int64_t)1 << (3 * l)) / 32); }

static size_t dense_layout(int L, char* base, DenseLevels* lv) {
    size_t o = 0;
    if (lv) { lv->L = L; lv->meta = reinterpret_cast<int64_t*>(base + o); }
    o = up256(o + sizeof(int64_t) * (DENSE_MAX_L + 4));
    for (int l = 1; l <= L; ++l) {
        const int64_t w = level_words(l);
        const int64_t chunks = (w + CHUNK_WORDS - 1) / CHUNK_WORDS;
        if (lv) { lv->words[l] = w; lv->bits[l] = reinterpret_cast<uint32_t*>(base + o); }
        o = up256(o + 4 * (size_t)w);
        if (lv) lv->pref[l] = reinterpret_cast<uint32_t*>(base + o);
        o = up256(o + 4 * (size_t)w);
        if (lv) lv->cbase[l] = reinterpret_cast<uint32_t*>(base + o);
        o = up256(o + 4 * (size_t)chunks);
    }
    return o;
}

__device__ __forceinline__ uint32_t spread3_30(uint32_t v) {      // 10 bits -> every third bit
    uint32_t x = v & 0x3ffu;
    x = (x | x << 16) & 0x030000ffu;
    x = (x | x << 8) & 0x0300f00fu;
    x = (x | x << 4) & 0x030c30c3u;
    x = (x | x << 2) & 0x09249249u;
    return x;
}

// Finest-level cell of a point with the descent's own arithmetic (common.cuh:44-51 FFMA, 37-42 clamp, exact
// power-of-two scaling), as a Morton key whose 3-bit groups are the slots u*4 + v*2 + w of successive levels.
__device__ __forceinline__ uint32_t point_key(const float* __restrict__ pts, int64_t i, int L, float o0, float o1, float o2,
                                              float s0, float s1, float s2, float sc) {
    const float x = clamp01(fmaf(s0, __ldg(pts + 3 * i), o0));
    const float y = clamp01(fmaf(s1, __ldg(pts + 3 * i + 1), o1));
    const float z = clamp01(fmaf(s2, __ldg(pts + 3 * i + 2), o2));
    return (spread3_30((uint32_t)(x * sc)) << 2) | (spread3_30((uint32_t)(y * sc)) << 1) | spread3_30((uint32_t)(z * sc));
}

__global__ void __launch_bounds__(256)
dense_setbits_kernel(const float* __restrict__ pts, int64_t P, int L, const float* __restrict__ offset,
                     const float* __restrict__ scaling, uint32_t* __restrict__ bits) {
    const float o0 = __ldg(offset), o1 = __ldg(offset + 1), o2 = __ldg(offset + 2);
    const float s0 = __ldg(scaling), s1 = __ldg(scaling + 1), s2 = __ldg(scaling + 2);
    const float sc = __int_as_float((127 + L) << 23);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = point_key(pts, i, L, o0, o1, o2, s0, s1, s2, sc);
        atomicOr(bits + (k >> 5), 1u << (k & 31));
    }
}

// Exclusive prefix of one value per thread over a CTA of 1024 threads; returns the prefix, *total = the CTA's sum.
__device__ __forceinline__ uint32_t cta_exclusive_scan_1024(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, inc, s);
        if (lane >= s) inc += t;
    }
    __syncthreads();                      // warp_sum may still be read by the previous call
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sum[lane], wi = w;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, wi, s);
            if (lane >= s) wi += t;
        }
        warp_sum[lane] = wi - w;          // exclusive warp offsets
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    return warp_sum[warp] + inc - v;
}

// One CTA = one chunk of 1024 words of level l: word w <- 32 bytes of level l+1; in-chunk prefix of the popcounts;
// chunk total (turned into a chunk base by the top kernel).
__global__ void __launch_bounds__(CHUNK_WORDS)
dense_pyramid_kernel(const uint32_t* __restrict__ finer, uint32_t* __restrict__ bits, uint32_t* __restrict__ pref,
                     uint32_t* __restrict__ csum, int64_t words) {
    __shared__ uint32_t total;
    const int64_t w = (int64_t)blockIdx.x * CHUNK_WORDS + threadIdx.x;
    uint32_t word = 0;
    if (w < words) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(finer) + 2 * w);
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(finer) + 2 * w + 1);
        const uint32_t src[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            // bit 4j+t of the word = byte t of src[j] != 0
            const uint32_t s = src[j];
            const uint32_t nz = ((s & 0xffu) ? 1u : 0u) | ((s & 0xff00u) ? 2u : 0u) | ((s & 0xff0000u) ? 4u : 0u) |
                                ((s & 0xff000000u) ? 8u : 0u);
            word |= nz << (4 * j);
        }
        bits[w] = word;
    }
    const uint32_t p = cta_exclusive_scan_1024((uint32_t)__popc(word), &total);
    if (w < words) pref[w] = p;
    if (threadIdx.x == 0) csum[blockIdx.x] = total;
}

// Single CTA: chunk bases of the large levels, the small top levels (pyramid + rank directory), node counts, bases.
__global__ void __launch_bounds__(1024)
dense_top_kernel(DenseLevels lv, int first_small /* finest level handled here, 0 = none */, int64_t cap_nodes) {
    __shared__ uint32_t total;
    __shared__ int64_t count[DENSE_MAX_L + 1];
    const int L = lv.L;
    // (a) small levels, finest first: level l from level l+1 (a large level the pyramid kernel produced, or level L
    // itself, whose bits the points have set), then its rank directory
    for (int l = first_small; l >= 1; --l) {
        const int64_t words = lv.words[l];
        if (l < L) {
            for (int64_t w = threadIdx.x; w < words; w += blockDim.x) {
                uint32_t word = 0;
                if (l >= 2) {                  // 32 cells <- 32 bytes = 8 words of level l+1
                    const uint32_t* src = lv.bits[l + 1] + 8 * w;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t s = src[j];
                        const uint32_t nz = ((s & 0xffu) ? 1u : 0u) | ((s & 0xff00u) ? 2u : 0u) | ((s & 0xff0000u) ? 4u : 0u) |
                                            ((s & 0xff000000u) ? 8u : 0u);
                        word |= nz << (4 * j);
                    }
                } else {                       // level 1: 8 cells <- the 8 bytes of level 2
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t byte = (lv.bits[2][j >> 2] >> (8 * (j & 3))) & 0xffu;
                        word |= (byte ? 1u : 0u) << j;
                    }
                }
                lv.bits[l][w] = word;
            }
            __syncthreads();
        }
        uint32_t running = 0;
        for (int64_t w0 = 0; w0 < words; w0 += blockDim.x) {
            const int64_t w = w0 + threadIdx.x;
            const uint32_t v = w < words ? (uint32_t)__popc(lv.bits[l][w]) : 0u;
            const uint32_t p = cta_exclusive_scan_1024(v, &total);
            if (w < words) lv.pref[l][w] = running + p;
            running += total;
            __syncthreads();
        }
        // small levels: one logical chunk per 1024 words with prefix already global -> chunk bases are zero
        for (int64_t c = threadIdx.x; c < (words + CHUNK_WORDS - 1) / CHUNK_WORDS; c += blockDim.x) lv.cbase[l][c] = 0;
        if (threadIdx.x == 0) count[l] = running;
        __syncthreads();
    }
    // (b) large levels: exclusive scan of the chunk totals the pyramid kernel left in cbase
    for (int l = first_small + 1; l <= L - 1; ++l) {
        const int64_t chunks = (lv.words[l] + CHUNK_WORDS - 1) / CHUNK_WORDS;
        uint32_t running = 0;
        for (int64_t c0 = 0; c0 < chunks; c0 += blockDim.x) {
            const int64_t c = c0 + threadIdx.x;
            const uint32_t v = c < chunks ? lv.cbase[l][c] : 0u;
            const uint32_t p = cta_exclusive_scan_1024(v, &total);
            if (c < chunks) lv.cbase[l][c] = running + p;
            running += total;
            __syncthreads();
        }
        if (threadIdx.x == 0) count[l] = running;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int64_t n = 1;                       // the root
        lv.meta[2] = 0;
        for (int l = 1; l <= L - 1; ++l) { lv.meta[2 + l] = n; n += count[l]; }
        lv.meta[0] = n;
        lv.meta[1] = n > cap_nodes ? 1 : 0;
    }
}

__device__ __forceinline__ int64_t dense_rank(const DenseLevels& lv, int l, uint32_t cell) {
    const uint32_t w = cell >> 5;
    return (int64_t)__ldg(lv.cbase[l] + (w / CHUNK_WORDS)) + __ldg(lv.pref[l] + w) +
           __popc(__ldg(lv.bits[l] + w) & ((1u << (cell & 31)) - 1u));
}

__global__ void __launch_bounds__(256)
dense_init_kernel(int32_t* __restrict__ child, int32_t* __restrict__ data, int32_t* __restrict__ parent_depth,
                  int64_t n_nodes) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes * 8; i += (int64_t)gridDim.x * blockDim.x) {
        child[i] = 0;
        data[i] = -1;                        // "no point yet" for the atomicMax of the leaf pass; fixed up at the end
        if (i < n_nodes * 2) parent_depth[i] = 0;
    }
}

// One thread per bitmap word of levels 1 .. L-1 (concatenated): every set bit is an internal node.
__global__ void __launch_bounds__(256)
dense_emit_kernel(DenseLevels lv, int64_t total_words, int64_t cap_nodes, int32_t* __restrict__ child,
                  int32_t* __restrict__ parent_depth) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_words; g += (int64_t)gridDim.x * blockDim.x) {
        int l = lv.L - 1;
        int64_t w = g;                       // finest internal level first (it holds most of the words)
        while (l > 1 && w >= lv.words[l]) { w -= lv.words[l]; --l; }
        uint32_t word = __ldg(lv.bits[l] + w);
        if (!word) continue;
        const int64_t base_l = lv.meta[2 + l], base_p = lv.meta[2 + l - 1];
        int64_t node = base_l + (int64_t)__ldg(lv.cbase[l] + (w / CHUNK_WORDS)) + __ldg(lv.pref[l] + w);
        while (word) {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            const uint32_t cell = (uint32_t)(w * 32 + b);
            const int64_t parent = l == 1 ? 0 : base_p + dense_rank(lv, l - 1, cell >> 3);
            const int slot = (int)(cell & 7u);
            SVOXB_DBG(parent >= 0 && parent < node);
            if (node < cap_nodes) {
                child[parent * 8 + slot] = (int32_t)(node - parent);
                parent_depth[2 * node] = (int32_t)(parent * 8 + slot);
                parent_depth[2 * node + 1] = l;
            }
            ++node;
        }
    }
}

__global__ void __launch_bounds__(256)
dense_leaf_kernel(DenseLevels lv, const float* __restrict__ pts, int64_t P, const float* __restrict__ offset,
                  const float* __restrict__ scaling, int64_t cap_nodes, int32_t* __restrict__ data) {
    const int L = lv.L;
    const float o0 = __ldg(offset), o1 = __ldg(offset + 1), o2 = __ldg(offset + 2);
    const float s0 = __ldg(scaling), s1 = __ldg(scaling + 1), s2 = __ldg(scaling + 2);
    const float sc = __int_as_float((127 + L) << 23);
    const int64_t base_p = lv.meta[2 + L - 1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = point_key(pts, i, L, o0, o1, o2, s0, s1, s2, sc);
        const int64_t parent = L == 1 ? 0 : base_p + dense_rank(lv, L - 1, k >> 3);
        SVOXB_DBG(parent >= 0 && parent < lv.meta[0]);
        if (parent < cap_nodes) atomicMax(data + parent * 8 + (k & 7u), (int32_t)i);     // the largest point index wins
    }
}

__global__ void __launch_bounds__(256)
dense_fix_kernel(int32_t* __restrict__ data, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (data[i] < 0) data[i] = DENSE_EMPTY;
}

}  // namespace svoxb

using namespace svoxb;

extern "C" int32_t svoxb_build_dense_max_depth(void) { return DENSE_MAX_L; }

extern "C" size_t svoxb_build_dense_work_bytes(int32_t L) {
    if (L < 1 || L > DENSE_MAX_L) return 0;
    return dense_layout(L, nullptr, nullptr);
}

extern "C" int svoxb_build_dense(const float* pts, int64_t P, int32_t L, const float* offset, const float* scaling,
                                 void* work, int64_t cap_nodes, int32_t* child, int32_t* data, int32_t* parent_depth,
                                 int64_t* status_dev, void* stream) {
    SVOXB_REQUIRE(L >= 1 && L <= DENSE_MAX_L, "depth L=%d out of range [1,%d] for the bitmap build", L, DENSE_MAX_L);
    SVOXB_REQUIRE(P >= 0 && P < (1ll << 31), "point count out of range");
    SVOXB_REQUIRE(work && offset && scaling && child && data && parent_depth && (P == 0 || pts), "NULL argument");
    SVOXB_REQUIRE(cap_nodes >= 1, "cap_nodes must be >= 1");
    SVOXB_REQUIRE(((uintptr_t)work & 255) == 0, "work must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    DenseLevels lv{};
    dense_layout(L, static_cast<char*>(work), &lv);
    SVOXB_CUDA(cudaMemsetAsync(lv.bits[L], 0, 4 * (size_t)lv.words[L], st));
    int launches = 0;
    if (P > 0) {
        const int grid = (int)min((P + 255) / 256, (int64_t)sm_count() * 8);
        dense_setbits_kernel<<<grid, 256, 0, st>>>(pts, P, L, offset, scaling, lv.bits[L]);
        ++launches;
    }
    // large levels (more than SMALL_WORDS words), finest first; the rest in the single-CTA top kernel
    int l = L - 1;
    for (; l >= 1 && lv.words[l] > SMALL_WORDS; --l) {
        const int chunks = (int)((lv.words[l] + CHUNK_WORDS - 1) / CHUNK_WORDS);
        dense_pyramid_kernel<<<chunks, CHUNK_WORDS, 0, st>>>(lv.bits[l + 1], lv.bits[l], lv.pref[l], lv.cbase[l], lv.words[l]);
        ++launches;
    }
    dense_top_kernel<<<1, 1024, 0, st>>>(lv, l, cap_nodes);
    const int g0 = (int)min((cap_nodes * 8 + 255) / 256, (int64_t)sm_count() * 8);
    dense_init_kernel<<<g0, 256, 0, st>>>(child, data, parent_depth, cap_nodes);
    launches += 2;
    if (L > 1) {
        int64_t total_words = 0;
        for (int k = 1; k <= L - 1; ++k) total_words += lv.words[k];
        const int g1 = (int)min((total_words + 255) / 256, (int64_t)sm_count() * 8);
        dense_emit_kernel<<<g1, 256, 0, st>>>(lv, total_words, cap_nodes, child, parent_depth);
        ++launches;
    }
    if (P > 0) {
        const int grid = (int)min((P + 255) / 256, (int64_t)sm_count() * 8);
        dense_leaf_kernel<<<grid, 256, 0, st>>>(lv, pts, P, offset, scaling, cap_nodes, data);
        ++launches;
    }
    dense_fix_kernel<<<g0, 256, 0, st>>>(data, cap_nodes * 8);
    ++launches;
    if (status_dev) SVOXB_CUDA(cudaMemcpyAsync(status_dev, lv.meta, 2 * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    count_launch(launches);
    return check_cuda(cudaGetLastError(), "svoxb_build_dense launch");
}
