/* C helpers of the CPU oracle — TEST INFRASTRUCTURE, not product code.
 *
 * Restates the two third-party (qiskit 1.2.4, Rust) functions the reference calls, so that the CPU
 * baseline is not dominated by a Python stand-in:
 *   orc_unordered_unique  <- qiskit._accelerate.sparse_pauli_op.unordered_unique
 *                            (reference call site symmer/operators/utils.py:271)
 *   orc_to_matrix_sparse  <- qiskit._accelerate.sparse_pauli_op.to_matrix_sparse
 *                            (reference call site symmer/operators/base.py:1500-1510)
 *   orc_add_at            <- np.add.at(reduced_coeff_vec, inverse_map, coeff_vec), utils.py:274
 * Built by oracle/Makefile into oracle/_build/liboracle.so. Only tests/, smoke() and bench.py's
 * cpu_baseline / --impl reference legs load it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t fnv_row(const unsigned char *p, int64_t nbytes) {
    uint64_t h = 0xcbf29ce484222325ull;
    int64_t i = 0;
    for (; i + 8 <= nbytes; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 0x100000001b3ull;
        h ^= h >> 29;
    }
    for (; i < nbytes; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
    h ^= h >> 32;
    return h;
}

/* Sequential first-occurrence unique of T rows of `row_bytes` bytes each (contiguous rows).
 * first[u] = index of the first row with the u-th distinct value (scan order), inv[t] = u.
 * Returns U. Single-threaded like the original hash-map loop. */
int64_t orc_unordered_unique(const void *rows_, int64_t T, int64_t row_bytes, int64_t *first,
                             int64_t *inv) {
    const unsigned char *rows = (const unsigned char *)rows_;
    int64_t cap = 16;
    while (cap < 2 * T) cap <<= 1;
    int64_t *slot = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    if (!slot) return -1;
    for (int64_t i = 0; i < cap; ++i) slot[i] = -1;
    int64_t U = 0;
    for (int64_t t = 0; t < T; ++t) {
        const unsigned char *r = rows + t * row_bytes;
        uint64_t h = fnv_row(r, row_bytes) & (uint64_t)(cap - 1);
        for (;;) {
            int64_t u = slot[h];
            if (u < 0) {
                slot[h] = U;
                first[U] = t;
                inv[t] = U;
                ++U;
                break;
            }
            if (memcmp(rows + first[u] * row_bytes, r, (size_t)row_bytes) == 0) {
                inv[t] = u;
                break;
            }
            h = (h + 1) & (uint64_t)(cap - 1);
        }
    }
    free(slot);
    return U;
}

/* acc[inv[t]] += c[t] for t in input order; complex128 stored as (re, im) pairs. */
void orc_add_at(double *acc, const int64_t *inv, const double *c, int64_t T) {
    for (int64_t t = 0; t < T; ++t) {
        acc[2 * inv[t]] += c[2 * t];
        acc[2 * inv[t] + 1] += c[2 * t + 1];
    }
}

typedef struct { int64_t col; double re, im; } entry_t;

static int cmp_entry(const void *a, const void *b) {
    int64_t ca = ((const entry_t *)a)->col, cb = ((const entry_t *)b)->col;
    return (ca > cb) - (ca < cb);
}

static int cmp_i64(const void *a, const void *b) {
    int64_t ca = *(const int64_t *)a, cb = *(const int64_t *)b;
    return (ca > cb) - (ca < cb);
}

/* CSR of sum_t c[t] * P_t where entry (r, r ^ x[t]) gets c[t] * (-1)^popcount(r & z[t]); the
 * (-i)^Y factor is already folded into c by the caller. Per row the entries are sorted by column
 * and equal columns summed (in term order). nnz_expected = 2^n * (#distinct x). Serial here (the
 * original is rayon-parallel over rows; this image has no libgomp). Returns nnz written or -1. */
int64_t orc_to_matrix_sparse(const int64_t *x, const int64_t *z, const double *c, int64_t M,
                             int32_t n, double *data, int64_t *indices, int64_t *indptr,
                             int64_t nnz_expected) {
    int64_t side = (int64_t)1 << n;
    int64_t *ux = (int64_t *)malloc(sizeof(int64_t) * (size_t)M);
    memcpy(ux, x, sizeof(int64_t) * (size_t)M);
    qsort(ux, (size_t)M, sizeof(int64_t), cmp_i64);
    int64_t G = 0;
    for (int64_t t = 0; t < M; ++t)
        if (t == 0 || ux[t] != ux[t - 1]) ux[G++] = ux[t];
    if (side * G != nnz_expected) { free(ux); return -1; }
    /* group id of every term */
    int64_t *gid = (int64_t *)malloc(sizeof(int64_t) * (size_t)M);
    for (int64_t t = 0; t < M; ++t) {
        int64_t lo = 0, hi = G - 1;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (ux[mid] < x[t]) lo = mid + 1; else hi = mid; }
        gid[t] = lo;
    }
    {
        entry_t *e = (entry_t *)malloc(sizeof(entry_t) * (size_t)G);
        for (int64_t r = 0; r < side; ++r) {
            for (int64_t g = 0; g < G; ++g) { e[g].col = r ^ ux[g]; e[g].re = 0.0; e[g].im = 0.0; }
            for (int64_t t = 0; t < M; ++t) {
                double s = (__builtin_popcountll((uint64_t)(r & z[t])) & 1) ? -1.0 : 1.0;
                e[gid[t]].re += s * c[2 * t];
                e[gid[t]].im += s * c[2 * t + 1];
            }
            qsort(e, (size_t)G, sizeof(entry_t), cmp_entry);
            int64_t base = r * G;
            for (int64_t g = 0; g < G; ++g) {
                indices[base + g] = e[g].col;
                data[2 * (base + g)] = e[g].re;
                data[2 * (base + g) + 1] = e[g].im;
            }
            indptr[r] = base;
        }
        free(e);
    }
    indptr[side] = side * G;
    free(ux);
    free(gid);
    return side * G;
}
