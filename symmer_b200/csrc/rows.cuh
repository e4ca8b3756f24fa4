// Row accessors shared by the dedup pipeline: where a "term" t gets its packed row and coefficient.
#pragma once
#include "common.cuh"

namespace symb {

// Term t is the cross term (p, q) of a product A*B in the reference's flattened order t = q*M + p
// (base.py:783-792); its row is A[p] ^ B[q] and is never stored. The record key carries the phase
// exponent of the pair in its two low bits.
struct ProductRows {
    const uint64_t *__restrict__ A;
    const uint64_t *__restrict__ B;
    const double *__restrict__ Ac;
    const double *__restrict__ Bc;
    uint32_t M;  // rows of A (divisor of t)
    int words;   // 2*W

    __device__ __forceinline__ void split(uint32_t t, uint32_t &p, uint32_t &q) const {
        q = t / M;
        p = t - q * M;
    }
    __device__ __forceinline__ uint64_t word(uint32_t t, int k) const {
        uint32_t p, q;
        split(t, p, q);
        return A[(size_t)p * words + k] ^ B[(size_t)q * words + k];
    }
    __device__ __forceinline__ bool equal(uint32_t t1, uint32_t t2) const {
        uint32_t p1, q1, p2, q2;
        split(t1, p1, q1);
        split(t2, p2, q2);
        const uint64_t *a1 = A + (size_t)p1 * words, *b1 = B + (size_t)q1 * words;
        const uint64_t *a2 = A + (size_t)p2 * words, *b2 = B + (size_t)q2 * words;
        for (int k = 0; k < words; ++k)
            if ((a1[k] ^ b1[k]) != (a2[k] ^ b2[k])) return false;
        return true;
    }
    __device__ __forceinline__ void coeff(uint32_t t, uint64_t key, double &re, double &im) const {
        uint32_t p, q;
        split(t, p, q);
        cmul(Ac[2 * (size_t)p], Ac[2 * (size_t)p + 1], Bc[2 * (size_t)q], Bc[2 * (size_t)q + 1], re, im);
        mul_i_pow(re, im, (int)(key & 3ull));
    }
};

// Term t is row t of a stored operator.
struct PlainRows {
    const uint64_t *__restrict__ X;
    const double *__restrict__ C;
    int words;

    __device__ __forceinline__ uint64_t word(uint32_t t, int k) const { return X[(size_t)t * words + k]; }
    __device__ __forceinline__ bool equal(uint32_t t1, uint32_t t2) const {
        const uint64_t *r1 = X + (size_t)t1 * words, *r2 = X + (size_t)t2 * words;
        for (int k = 0; k < words; ++k)
            if (r1[k] != r2[k]) return false;
        return true;
    }
    __device__ __forceinline__ void coeff(uint32_t t, uint64_t, double &re, double &im) const {
        re = C[2 * (size_t)t];
        im = C[2 * (size_t)t + 1];
    }
};

// dedup driver (dedup.cu), split at the point where the survivor count is known so that the caller
// can allocate exact-size outputs. by_t: vals are a permutation of 0..T-1 and the output is written
// in increasing-t (first occurrence) order; otherwise output is in sorted-key order.
size_t dedup_ws_bytes(int64_t T);
int dedup_product_plan(uint64_t *keys, uint32_t *vals, bool vals_iota, int64_t T, const ProductRows &rows, bool by_t,
                       double thr, int64_t *n_out, int64_t *n_out_host, void *ws, size_t ws_bytes, cudaStream_t st);
int dedup_product_emit(const uint32_t *vals, int64_t T, const ProductRows &rows, bool by_t, int64_t U, uint64_t *out_xz,
                       double *out_c, void *ws, size_t ws_bytes, cudaStream_t st);
int dedup_plain_plan(uint64_t *keys, uint32_t *vals, int64_t T, const PlainRows &rows, double thr, int64_t *n_out,
                     int64_t *n_out_host, void *ws, size_t ws_bytes, cudaStream_t st);
int dedup_plain_emit(int64_t T, const PlainRows &rows, int64_t U, uint64_t *out_xz, double *out_c, void *ws,
                     size_t ws_bytes, cudaStream_t st);

}  // namespace symb
