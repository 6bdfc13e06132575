/*
 * svox_oracle.c -- CPU oracle for the svox_t octree volume-rendering hot path.
 *
 * TEST INFRASTRUCTURE ONLY. This library is the checker the parity tests, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs compare against or time. The product
 * (svox_t_b200/) never imports, links or executes it; the product path fails loudly when its CUDA
 * library is missing instead of falling back to this code.
 *
 * Parity pin: the reference repository has no tests, golden vectors or fixtures for this path
 * (SURVEY.md 8c). The oracle is therefore pinned against outputs of the reference's own CUDA kernels
 * (oracle/_ref, built by oracle/build_ref.sh) run on a B200: see tests/golden/ and
 * tests/golden/make_golden.py; until those fixtures are committed the status is "parity unpinned".
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off [-mfma] -shared -fPIC svox_oracle.c -o _build/libsvox_oracle.so -lm
 * (-ffp-contract=off: only the FMAs the reference's SASS is known to contain are written explicitly.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define FN(name) name##_f32
#define R_SQRT(x) sqrtf(x)
#define R_EXP(x) expf(x)
#define R_FMA(a, b, c) fmaf((a), (b), (c))
#include "svox_oracle_impl.inc"
#undef REAL
#undef FN
#undef R_SQRT
#undef R_EXP
#undef R_FMA

#define REAL double
#define FN(name) name##_f64
#define R_SQRT(x) sqrt(x)
#define R_EXP(x) exp(x)
#define R_FMA(a, b, c) fma((a), (b), (c))
#include "svox_oracle_impl.inc"
#undef REAL
#undef FN
#undef R_SQRT
#undef R_EXP
#undef R_FMA

static int cmp_i64(const void* a, const void* b) {
    const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

/* svox_kernel.cu:239-269,291-320 -- the set of unique leaf slots hit by a query batch (empty leaves
 * included, svox_kernel.cu:57-58), returned sorted as rows [node, i, j, k]. The reference's order is
 * nondeterministic (float atomic counter); compare as a sorted set. Returns n_hit. */
int64_t orc_leafset(const int64_t* node_ids, int64_t Q, int N, int64_t* leaf_node /* [Q,4] capacity */) {
    int64_t* tmp = (int64_t*)malloc(sizeof(int64_t) * (size_t)(Q > 0 ? Q : 1));
    memcpy(tmp, node_ids, sizeof(int64_t) * (size_t)Q);
    qsort(tmp, (size_t)Q, sizeof(int64_t), cmp_i64);
    int64_t n = 0;
    for (int64_t i = 0; i < Q; ++i) {
        if (i && tmp[i] == tmp[i - 1]) continue;
        int64_t v = tmp[i];
        for (int k = 3; k > 0; --k) { leaf_node[n * 4 + k] = v % N; v /= N; }
        leaf_node[n * 4] = v;
        ++n;
    }
    free(tmp);
    return n;
}

int orc_abi_version(void) { return 1; }

/* Thread count of the OpenMP loops over rays / points. Launchers such as torchrun export OMP_NUM_THREADS=1, which the
 * runtime reads once at load time; bench.py's CPU legs set the count explicitly through this call instead. Returns
 * the count now in effect. */
#ifdef _OPENMP
#include <omp.h>
int orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int orc_set_threads(int n) { (void)n; return 1; }
#endif
