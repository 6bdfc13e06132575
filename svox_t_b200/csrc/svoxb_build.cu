// svoxb_build.cu -- one-shot octree build from points for the per-frame rebuild of animated scenes.
//
// Replaces the reference's rebuild loop: depth-1 rounds of query_vertical (3 kernels + a host sync each,
// svox_kernel.cu:274-324) + N3Tree.refine (about ten small torch kernels and up to three reallocations per round,
// svox_t/svox.py:488-560, helpers.py:38-109) followed by construct_tree (svox_kernel.cu:110-121).
// Here: Morton keys of the points at the finest level -> one radix sort -> per-level cell ranks by prefix sums ->
// a single emit kernel that writes child / data / parent_depth in the reference tensor format. The result is
// isomorphic to what the reference's loop produces (same leaf set, same point -> leaf map); node numbering is
// breadth-first and, within a level, by Morton key, i.e. deterministic (the reference's depends on atomic order).
// The radix sort and the scans are CUB device primitives (plumbing); key generation, level detection and emit are
// hand-written.
#include <cub/cub.cuh>
#include "svoxb_common.cuh"

namespace svoxb {

constexpr int BUILD_MAX_L = 16;
constexpr int32_t BUILD_EMPTY = 1410065408;      // int(1e10) wrapped to int32 (svox_t/svox.py:124)

struct BuildLayout {
    size_t keys_in, keys_out, vals_in, vals_out, lev, cellidx, counts, base, cub_temp, cub_bytes, total;
};

static inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

static size_t cub_temp_bytes(int64_t P, int L) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)P, 0, 3 * L);
    cub::DeviceScan::InclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, (int)P);
    size_t m = a > b ? a : b;
    if (m == 0) m = (size_t)16 * (size_t)P + (16u << 20);     // no device to ask: generous upper bound
    return m + 1024;
}

static BuildLayout build_layout(int64_t P, int L) {
    BuildLayout l{};
    size_t o = 0;
    const size_t n = (size_t)(P > 0 ? P : 1);
    l.keys_in = o; o = align_up(o + 8 * n);
    l.keys_out = o; o = align_up(o + 8 * n);
    l.vals_in = o; o = align_up(o + 4 * n);
    l.vals_out = o; o = align_up(o + 4 * n);
    l.lev = o; o = align_up(o + 4 * n);                        // int per point: first level at which a new cell starts
    l.cellidx = o; o = align_up(o + 4 * n * (size_t)(L > 1 ? L - 1 : 1));
    l.counts = o; o = align_up(o + 8 * (BUILD_MAX_L + 1));
    l.base = o; o = align_up(o + 8 * (BUILD_MAX_L + 1));
    l.cub_bytes = cub_temp_bytes(P, L);
    l.cub_temp = o; o = align_up(o + l.cub_bytes);
    l.total = o;
    return l;
}

__device__ __forceinline__ uint64_t spread3(uint32_t v) {      // 21 bits -> every third bit
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// Finest-level cell of each point, with the descent's own arithmetic (common.cuh:44-51 FFMA, 37-42 clamp, exact
// power-of-two scaling), as a Morton key whose 3-bit groups are the slot u*4 + v*2 + w of successive levels.
__global__ void build_keys_kernel(const float* __restrict__ pts, int64_t P, int L, const float* __restrict__ offset,
                                  const float* __restrict__ scaling, uint64_t* __restrict__ keys,
                                  uint32_t* __restrict__ vals) {
    const float o0 = __ldg(offset), o1 = __ldg(offset + 1), o2 = __ldg(offset + 2);
    const float s0 = __ldg(scaling), s1 = __ldg(scaling + 1), s2 = __ldg(scaling + 2);
    const float sc = __int_as_float((127 + L) << 23);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = clamp01(fmaf(s0, __ldg(pts + 3 * i), o0));
        const float y = clamp01(fmaf(s1, __ldg(pts + 3 * i + 1), o1));
        const float z = clamp01(fmaf(s2, __ldg(pts + 3 * i + 2), o2));
        const uint32_t ix = (uint32_t)(x * sc), iy = (uint32_t)(y * sc), iz = (uint32_t)(z * sc);
        keys[i] = (spread3(ix) << 2) | (spread3(iy) << 1) | spread3(iz);
        vals[i] = (uint32_t)i;
    }
}

// lev[i] = first level (1..L) at which sorted point i starts a new cell; L+1 if it shares its finest cell with i-1.
__global__ void build_levels_kernel(const uint64_t* __restrict__ keys, int64_t P, int L, int* __restrict__ lev) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        int l = 1;
        if (i > 0) {
            const uint64_t d = keys[i] ^ keys[i - 1];
            l = d == 0 ? L + 1 : L - (63 - __clzll((long long)d)) / 3;
        }
        lev[i] = l;
    }
}

struct LevelFlag {
    const int* lev;
    int l;
    __host__ __device__ int operator()(int i) const { return lev[i] <= l ? 1 : 0; }
};

__global__ void build_last_kernel(const int* __restrict__ cellidx_rows, int64_t P, int n_levels,
                                  int64_t* __restrict__ counts) {
    const int l = threadIdx.x;                   // level l+1 lives in row l
    if (l < n_levels) counts[l + 1] = P > 0 ? (int64_t)cellidx_rows[(size_t)l * P + (P - 1)] : 0;
}

__global__ void build_init_kernel(int32_t* __restrict__ child, int32_t* __restrict__ data,
                                  int32_t* __restrict__ parent_depth, int64_t n_nodes) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes * 8; i += (int64_t)gridDim.x * blockDim.x) {
        child[i] = 0;
        data[i] = BUILD_EMPTY;
        if (i < n_nodes * 2) parent_depth[i] = 0;
    }
}

__global__ void build_emit_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                  const int* __restrict__ lev, const int* __restrict__ cellidx /* inclusive sums */,
                                  const int64_t* __restrict__ base, int64_t P, int L, int32_t* __restrict__ child,
                                  int32_t* __restrict__ data, int32_t* __restrict__ parent_depth) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[i];
        const int l0 = lev[i];
        // internal nodes that start at this point: levels l0 .. L-1
        for (int l = l0; l < L; ++l) {
            const int64_t node = base[l] + cellidx[(size_t)(l - 1) * P + i] - 1;
            const int64_t parent = l == 1 ? 0 : base[l - 1] + cellidx[(size_t)(l - 2) * P + i] - 1;
            const int slot = (int)((key >> (3 * (L - l))) & 7);
            child[parent * 8 + slot] = (int32_t)(node - parent);
            parent_depth[2 * node] = (int32_t)(parent * 8 + slot);
            parent_depth[2 * node + 1] = l;
        }
        // the finest cell's row: the last point of a run of equal keys carries the largest point index
        const bool last = (i == P - 1) || (lev[i + 1] <= L);
        if (last) {
            const int64_t parent = L == 1 ? 0 : base[L - 1] + cellidx[(size_t)(L - 2) * P + i] - 1;
            data[parent * 8 + (int)(key & 7)] = (int32_t)vals[i];
        }
    }
}

}  // namespace svoxb

using namespace svoxb;

extern "C" size_t svoxb_build_work_bytes(int64_t P, int32_t L) {
    if (P < 0 || L < 1 || L > BUILD_MAX_L) return 0;
    return build_layout(P, L).total;
}

extern "C" int svoxb_build_octree_count(const float* pts, int64_t P, int32_t L, const float* offset,
                                        const float* scaling, void* work, int64_t* n_nodes_host, void* stream) {
    SVOXB_REQUIRE(L >= 1 && L <= BUILD_MAX_L, "depth L=%d out of range [1,%d]", L, BUILD_MAX_L);
    SVOXB_REQUIRE(P >= 0 && P < (1ll << 31), "point count out of range");
    SVOXB_REQUIRE(work && n_nodes_host && offset && scaling && (P == 0 || pts), "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const BuildLayout lay = build_layout(P, L);
    char* w = (char*)work;
    uint64_t* keys_in = (uint64_t*)(w + lay.keys_in);
    uint64_t* keys_out = (uint64_t*)(w + lay.keys_out);
    uint32_t* vals_in = (uint32_t*)(w + lay.vals_in);
    uint32_t* vals_out = (uint32_t*)(w + lay.vals_out);
    int* lev = (int*)(w + lay.lev);
    int* cellidx = (int*)(w + lay.cellidx);
    int64_t* counts = (int64_t*)(w + lay.counts);
    int64_t* base = (int64_t*)(w + lay.base);
    int64_t h_counts[BUILD_MAX_L + 1] = {0}, h_base[BUILD_MAX_L + 1] = {0};
    if (P > 0) {
        const int grid = (int)min((P + 255) / 256, (int64_t)sm_count() * 8);
        build_keys_kernel<<<grid, 256, 0, st>>>(pts, P, L, offset, scaling, keys_in, vals_in);
        size_t tb = lay.cub_bytes;
        SVOXB_CUDA(cub::DeviceRadixSort::SortPairs(w + lay.cub_temp, tb, keys_in, keys_out, vals_in, vals_out, (int)P, 0,
                                                   3 * L, st));
        build_levels_kernel<<<grid, 256, 0, st>>>(keys_out, P, L, lev);
        for (int l = 1; l < L; ++l) {
            cub::TransformInputIterator<int, LevelFlag, cub::CountingInputIterator<int>> it(
                cub::CountingInputIterator<int>(0), LevelFlag{lev, l});
            tb = lay.cub_bytes;
            SVOXB_CUDA(cub::DeviceScan::InclusiveSum(w + lay.cub_temp, tb, it, cellidx + (size_t)(l - 1) * P, (int)P, st));
        }
        if (L > 1) build_last_kernel<<<1, 32, 0, st>>>(cellidx, P, L - 1, counts);
        count_launch(3 + 2 * (L - 1));
        SVOXB_CUDA(cudaGetLastError());
        if (L > 1) SVOXB_CUDA(cudaMemcpyAsync(h_counts, counts, sizeof(int64_t) * (L + 1), cudaMemcpyDeviceToHost, st));
    }
    SVOXB_CUDA(cudaStreamSynchronize(st));
    int64_t n = 1;
    for (int l = 1; l < L; ++l) { h_base[l] = n; n += h_counts[l]; }
    SVOXB_CUDA(cudaMemcpyAsync(base, h_base, sizeof(h_base), cudaMemcpyHostToDevice, st));
    SVOXB_CUDA(cudaStreamSynchronize(st));
    *n_nodes_host = n;
    return 0;
}

extern "C" int svoxb_build_octree_emit(int64_t P, int32_t L, const void* work, int64_t n_nodes, int32_t* child,
                                       int32_t* data, int32_t* parent_depth, void* stream) {
    SVOXB_REQUIRE(L >= 1 && L <= BUILD_MAX_L && P >= 0 && n_nodes >= 1, "bad sizes");
    SVOXB_REQUIRE(work && child && data && parent_depth, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const BuildLayout lay = build_layout(P, L);
    const char* w = (const char*)work;
    const int g0 = (int)min((n_nodes * 8 + 255) / 256, (int64_t)sm_count() * 8);
    build_init_kernel<<<g0, 256, 0, st>>>(child, data, parent_depth, n_nodes);
    count_launch();
    if (P > 0) {
        const int grid = (int)min((P + 255) / 256, (int64_t)sm_count() * 8);
        build_emit_kernel<<<grid, 256, 0, st>>>((const uint64_t*)(w + lay.keys_out), (const uint32_t*)(w + lay.vals_out),
                                               (const int*)(w + lay.lev), (const int*)(w + lay.cellidx),
                                               (const int64_t*)(w + lay.base), P, L, child, data, parent_depth);
        count_launch();
    }
    return check_cuda(cudaGetLastError(), "build_octree_emit launch");
}
