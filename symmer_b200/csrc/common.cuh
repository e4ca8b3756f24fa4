// Shared device/host helpers for the symmer_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/symmer_b200.h"

namespace symb {

// ------------------------------------------------------------------ host-side error plumbing
void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;
extern uint64_t g_key_mask;

#define SYM_CUDA_OK(expr)                                                                    \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            symb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                       \
            return SYM_E_CUDA;                                                               \
        }                                                                                    \
    } while (0)

#define SYM_LAUNCH_OK()                                                                      \
    do {                                                                                     \
        symb::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            symb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),      \
                            __FILE__, __LINE__);                                             \
            return SYM_E_CUDA;                                                               \
        }                                                                                    \
    } while (0)

#define SYM_REQUIRE(cond, msg)                                     \
    do {                                                           \
        if (!(cond)) {                                             \
            symb::set_error("invalid argument: %s (%s)", msg, #cond); \
            return SYM_E_INVALID;                                  \
        }                                                          \
    } while (0)

#define SYM_TRY(expr)              \
    do {                           \
        int _rc = (expr);          \
        if (_rc != SYM_OK) return _rc; \
    } while (0)

// Bump allocator over the caller's workspace (256-byte aligned slices).
struct Arena {
    char *base;
    size_t cap;
    size_t off;
    Arena(void *p, size_t n) : base(static_cast<char *>(p)), cap(n), off(0) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        if (off + bytes > cap) return nullptr;
        T *r = reinterpret_cast<T *>(base + off);
        off += bytes;
        return r;
    }
};
static inline size_t arena_need(size_t count, size_t elem) { return (count * elem + 255) & ~size_t(255); }

static inline int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// ------------------------------------------------------------------ device helpers
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    // murmur3 fmix64: a bijection on 64-bit words, so equal sketches <=> equal keys
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// Lane-specific invertible GF(2)-linear map of one 64-bit word: two xorshift rounds whose shift
// triple depends on the lane. Linear => sketch(r1 ^ r2) = sketch(r1) ^ sketch(r2).
__host__ __device__ __forceinline__ uint64_t lane_linear(uint64_t w, int lane) {
    const int a = 5 + lane;                 // 5..36
    const int b = 3 + ((lane * 5) % 29);    // 3..31
    const int c = 11 + ((lane * 11) % 37);  // 11..47
    w ^= w << a;
    w ^= w >> b;
    w ^= w << c;
    w ^= w >> (a + 7);
    w ^= w << (b + 2);
    w ^= w >> (c - 4);
    return w;
}

__host__ __device__ __forceinline__ uint64_t chain_linear(uint64_t h) {
    // xorshift64 step (13,7,17): invertible linear map used to chain 32-word groups
    h ^= h << 13;
    h ^= h >> 7;
    h ^= h << 17;
    return h;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t warp_xor(uint64_t v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// Sketch of a packed row of `words` uint64, computed cooperatively by one warp; every lane returns
// the same value.
__device__ __forceinline__ uint64_t warp_sketch_row(const uint64_t *__restrict__ row, int words, int lane) {
    uint64_t h = 0;
    for (int base = 0; base < words; base += 32) {
        int k = base + lane;
        uint64_t w = (k < words) ? row[k] : 0ull;
        uint64_t g = warp_xor(lane_linear(w, lane));
        h = chain_linear(h) ^ g;
    }
    return h;
}

// Same sketch for rows of at most 32 words with an even word count per block (W even, W <= 16),
// computed by a group of 8 lanes with 16-byte loads: 4 rows per warp, 4x the bytes in flight of the
// warp-per-row form. g = lane within the group. Every lane of the group returns the sketch.
__device__ __forceinline__ uint64_t group8_sketch_row(const uint64_t *__restrict__ row, int words, int g) {
    const uint4 *r4 = reinterpret_cast<const uint4 *>(row);
    const int chunks = words >> 1;
    uint64_t h = 0;
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
        const int c = g + 8 * rep;
        if (c < chunks) {
            const uint4 v = r4[c];
            const uint64_t w0 = ((uint64_t)v.y << 32) | v.x, w1 = ((uint64_t)v.w << 32) | v.z;
            h ^= lane_linear(w0, 2 * c) ^ lane_linear(w1, 2 * c + 1);
        }
    }
    h ^= __shfl_xor_sync(0xffffffffu, h, 1);
    h ^= __shfl_xor_sync(0xffffffffu, h, 2);
    h ^= __shfl_xor_sync(0xffffffffu, h, 4);
    return h;
}

__host__ __device__ __forceinline__ bool group8_ok(int W) { return W <= 16 && (W % 2 == 0); }

// sketches and Y counts of both operands of a product, one launch when the rows fit the 8-lane form (layout.cu)
int operand_tables(const uint64_t *a_xz, int64_t M, const uint64_t *b_xz, int64_t N, int W, uint64_t *a_sk, int32_t *a_y,
                   uint64_t *b_sk, int32_t *b_y, cudaStream_t st);

// i^k applied to a complex number (k mod 4), exact.
__device__ __forceinline__ void mul_i_pow(double &re, double &im, int k) {
    double r = re, i = im;
    switch (k & 3) {
        case 0: break;
        case 1: re = -i; im = r; break;
        case 2: re = -r; im = -i; break;
        default: re = i; im = -r; break;
    }
}

// Commutative complex product (no FMA contraction): a*b == b*a bit for bit, so the two orderings of
// an anticommuting pair cancel exactly in a square.
__device__ __forceinline__ void cmul(double ar, double ai, double br, double bi, double &re, double &im) {
    re = __dsub_rn(__dmul_rn(ar, br), __dmul_rn(ai, bi));
    im = __dadd_rn(__dmul_rn(ar, bi), __dmul_rn(ai, br));
}
#endif

}  // namespace symb
