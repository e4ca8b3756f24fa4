// Dedup of terms by 64-bit key with exact row verification, segmented coefficient reduction in input
// order, threshold filter and row emission. Replaces qiskit's `unordered_unique` + `np.add.at`
// (symmer/operators/utils.py:271-278) — HBM-bound integer/byte work, no tensor cores.
//
// Pipeline (all on one stream):
//   1. stable LSD radix sort of (key, t) on the top K key bits   -> equal rows become neighbours
//   2. link: every sorted position decides head / same-as-predecessor / irregular
//   3. sum: each head adds the coefficients of its chain in input (t) order
//   4. irregular chains (key collisions inside a sort bucket; rare) are folded in with atomics
//   5. keep = head && |sum| > threshold ; exclusive scan -> output slots ; compact ; emit rows
#include "rows.cuh"
#include "sort.cuh"

namespace symb {

constexpr uint8_t FLAG_HEAD = 0, FLAG_PREV = 1, FLAG_LINK = 2;

template <class Rows>
__global__ void __launch_bounds__(256) link_kernel(Rows rows, const uint64_t *__restrict__ sk, const uint32_t *__restrict__ st,
                                                    int64_t T, int sort_shift, uint8_t *__restrict__ flag,
                                                    uint32_t *__restrict__ link) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    uint8_t f = FLAG_HEAD;
    if (i > 0) {
        const uint64_t ki = sk[i], kp = sk[i - 1];
        if ((ki >> sort_shift) == (kp >> sort_shift)) {
            const uint32_t ti = st[i];
            if (((ki ^ kp) >> 2) == 0 && rows.equal(ti, st[i - 1])) {
                f = FLAG_PREV;
            } else {
                // irregular: a different row shares this sort bucket; look further back for a twin
                for (int64_t j = i - 2; j >= 0; --j) {
                    const uint64_t kj = sk[j];
                    if ((kj >> sort_shift) != (ki >> sort_shift)) break;
                    if (((ki ^ kj) >> 2) == 0 && rows.equal(ti, st[j])) {
                        f = FLAG_LINK;
                        link[i] = (uint32_t)j;
                        break;
                    }
                }
            }
        }
    }
    flag[i] = f;
}

// heads: sequential sum over the chain of FLAG_PREV successors (input order, like np.add.at)
template <class Rows, bool BY_T>
__global__ void __launch_bounds__(256) sum_kernel(Rows rows, const uint64_t *__restrict__ sk, const uint32_t *__restrict__ st,
                                                   int64_t T, const uint8_t *__restrict__ flag, double2 *__restrict__ acc) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    if (flag[i] != FLAG_HEAD) return;
    double re, im;
    const uint32_t t0 = st[i];
    rows.coeff(t0, sk[i], re, im);
    for (int64_t j = i + 1; j < T && flag[j] == FLAG_PREV; ++j) {
        double r2, i2;
        rows.coeff(st[j], sk[j], r2, i2);
        re += r2;
        im += i2;
    }
    acc[BY_T ? (int64_t)t0 : i] = make_double2(re, im);
}

template <class Rows, bool BY_T>
__global__ void __launch_bounds__(256) sum_irregular_kernel(Rows rows, const uint64_t *__restrict__ sk,
                                                             const uint32_t *__restrict__ st, int64_t T,
                                                             const uint8_t *__restrict__ flag, const uint32_t *__restrict__ link,
                                                             double2 *__restrict__ acc) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    if (flag[i] != FLAG_LINK) return;
    double re, im;
    rows.coeff(st[i], sk[i], re, im);
    for (int64_t j = i + 1; j < T && flag[j] == FLAG_PREV; ++j) {
        double r2, i2;
        rows.coeff(st[j], sk[j], r2, i2);
        re += r2;
        im += i2;
    }
    int64_t r = link[i];
    while (flag[r] != FLAG_HEAD) r = (flag[r] == FLAG_PREV) ? r - 1 : (int64_t)link[r];
    double2 *dst = acc + (BY_T ? (int64_t)st[r] : r);
    atomicAdd(&dst->x, re);
    atomicAdd(&dst->y, im);
}

template <bool BY_T>
__global__ void __launch_bounds__(256) keep_kernel(const uint32_t *__restrict__ st, int64_t T, const uint8_t *__restrict__ flag,
                                                    const double2 *__restrict__ acc, double thr, uint8_t *__restrict__ keep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const int64_t d = BY_T ? (int64_t)st[i] : i;
    uint8_t k = 0;
    if (flag[i] == FLAG_HEAD) {
        if (thr < 0.0) {
            k = 1;
        } else {
            double2 a = acc[d];
            k = hypot(a.x, a.y) > thr ? 1 : 0;
        }
    }
    keep[d] = k;
}

__global__ void __launch_bounds__(256) compact_kernel(const uint8_t *__restrict__ keep, const uint32_t *__restrict__ slot,
                                                       int64_t T, uint32_t *__restrict__ kept) {
    int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= T) return;
    if (keep[d]) kept[slot[d]] = (uint32_t)d;
}

__global__ void total_to_i64_kernel(const uint32_t *__restrict__ total, int64_t *__restrict__ n_out) { *n_out = (int64_t)*total; }

// one thread per (kept record, word): coalesced 8-byte stores of the surviving rows
template <class Rows, bool BY_T>
__global__ void __launch_bounds__(256) emit_kernel(Rows rows, const uint32_t *__restrict__ kept, int64_t U,
                                                    const uint32_t *__restrict__ st, const double2 *__restrict__ acc,
                                                    uint64_t *__restrict__ out_xz, double2 *__restrict__ out_c) {
    const int words = rows.words;
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t rec = g / words;
    if (rec >= U) return;
    int k = (int)(g - rec * words);
    const uint32_t d = kept[rec];
    const uint32_t t = BY_T ? d : st[d];
    out_xz[g] = rows.word(t, k);
    if (k == 0) out_c[rec] = acc[d];
}

size_t dedup_ws_bytes(int64_t T) {
    if (T < 1) T = 1;
    size_t n = (size_t)T;
    return arena_need(n, 8)                        // keys_alt
           + arena_need(n, 4)                      // vals_alt
           + arena_need(sort_hist_elems(T), 4)     // radix histograms + scan scratch
           + arena_need(n, 1)                      // flag
           + arena_need(n, 4)                      // link
           + arena_need(n, 16)                     // acc
           + arena_need(n, 1)                      // keep
           + arena_need(n, 4)                      // slot
           + arena_need(n, 4)                      // kept
           + arena_need(scan_scratch_elems(T), 4)  // scan scratch
           + arena_need(4, 4) + 4096;
}

static int sort_bits_for(int64_t T) {
    int lg = 0;
    while ((int64_t(1) << lg) < T) ++lg;
    int k = ((lg + 8 + 7) / 8) * 8;
    if (k > 56) k = 56;
    if (k < 8) k = 8;
    return k;
}

struct DedupLayout {
    uint64_t *keys_alt;
    uint32_t *vals_alt;
    uint32_t *hist;
    uint8_t *flag;
    uint32_t *link;
    double2 *acc;
    uint8_t *keep;
    uint32_t *slot;
    uint32_t *kept;
    uint32_t *scratch;
    uint32_t *total;
    bool ok;
};

static DedupLayout dedup_layout(void *ws, size_t ws_bytes, int64_t T) {
    Arena ar(ws, ws_bytes);
    DedupLayout L;
    L.keys_alt = ar.take<uint64_t>((size_t)T);
    L.vals_alt = ar.take<uint32_t>((size_t)T);
    L.hist = ar.take<uint32_t>(sort_hist_elems(T));
    L.flag = ar.take<uint8_t>((size_t)T);
    L.link = ar.take<uint32_t>((size_t)T);
    L.acc = ar.take<double2>((size_t)T);
    L.keep = ar.take<uint8_t>((size_t)T);
    L.slot = ar.take<uint32_t>((size_t)T);
    L.kept = ar.take<uint32_t>((size_t)T);
    L.scratch = ar.take<uint32_t>(scan_scratch_elems(T));
    L.total = ar.take<uint32_t>(4);
    L.ok = L.total != nullptr;
    return L;
}

// Phase 1: everything up to the survivor count (synchronises the stream once to read it).
template <class Rows, bool BY_T>
static int dedup_plan(uint64_t *keys, uint32_t *vals, bool vals_iota, int64_t T, const Rows &rows, double thr,
                      int64_t *n_out, int64_t *n_out_host, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (T == 0) {
        if (n_out) SYM_CUDA_OK(cudaMemsetAsync(n_out, 0, sizeof(int64_t), st));
        if (n_out_host) *n_out_host = 0;
        return SYM_OK;
    }
    if (ws_bytes < dedup_ws_bytes(T)) {
        set_error("workspace too small: need %zu bytes, got %zu", dedup_ws_bytes(T), ws_bytes);
        return SYM_E_WORKSPACE;
    }
    DedupLayout L = dedup_layout(ws, ws_bytes, T);
    if (!L.ok) {
        set_error("workspace arena exhausted");
        return SYM_E_WORKSPACE;
    }
    const int sort_bits = sort_bits_for(T);
    const int sort_shift = 64 - sort_bits;
    SYM_TRY(radix_sort_pairs(keys, vals, L.keys_alt, L.vals_alt, T, sort_shift, vals_iota, L.hist, st));

    const unsigned nb = (unsigned)((T + 255) / 256);
    link_kernel<Rows><<<nb, 256, 0, st>>>(rows, keys, vals, T, sort_shift, L.flag, L.link);
    SYM_LAUNCH_OK();
    sum_kernel<Rows, BY_T><<<nb, 256, 0, st>>>(rows, keys, vals, T, L.flag, L.acc);
    SYM_LAUNCH_OK();
    sum_irregular_kernel<Rows, BY_T><<<nb, 256, 0, st>>>(rows, keys, vals, T, L.flag, L.link, L.acc);
    SYM_LAUNCH_OK();
    keep_kernel<BY_T><<<nb, 256, 0, st>>>(vals, T, L.flag, L.acc, thr, L.keep);
    SYM_LAUNCH_OK();
    SYM_TRY(scan_exclusive_u8(L.keep, L.slot, T, L.total, L.scratch, st));
    compact_kernel<<<nb, 256, 0, st>>>(L.keep, L.slot, T, L.kept);
    SYM_LAUNCH_OK();
    if (n_out) {
        total_to_i64_kernel<<<1, 1, 0, st>>>(L.total, n_out);
        SYM_LAUNCH_OK();
    }
    uint32_t U32 = 0;
    SYM_CUDA_OK(cudaMemcpyAsync(&U32, L.total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SYM_CUDA_OK(cudaStreamSynchronize(st));
    if (n_out_host) *n_out_host = (int64_t)U32;
    return SYM_OK;
}

// Phase 2: write the U surviving rows + coefficients (asynchronous). `vals` must be the sorted
// values left by the plan phase (only read when !BY_T).
template <class Rows, bool BY_T>
static int dedup_emit(const uint32_t *vals, int64_t T, const Rows &rows, int64_t U, uint64_t *out_xz, double *out_c,
                      void *ws, size_t ws_bytes, cudaStream_t st) {
    if (T == 0 || U == 0) return SYM_OK;
    DedupLayout L = dedup_layout(ws, ws_bytes, T);
    if (!L.ok) {
        set_error("workspace arena exhausted");
        return SYM_E_WORKSPACE;
    }
    int64_t threads = U * rows.words;
    emit_kernel<Rows, BY_T><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(rows, L.kept, U, vals, L.acc, out_xz,
                                                                              reinterpret_cast<double2 *>(out_c));
    SYM_LAUNCH_OK();
    return SYM_OK;
}

int dedup_product_plan(uint64_t *keys, uint32_t *vals, bool vals_iota, int64_t T, const ProductRows &rows, bool by_t,
                       double thr, int64_t *n_out, int64_t *n_out_host, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (by_t) return dedup_plan<ProductRows, true>(keys, vals, vals_iota, T, rows, thr, n_out, n_out_host, ws, ws_bytes, st);
    return dedup_plan<ProductRows, false>(keys, vals, vals_iota, T, rows, thr, n_out, n_out_host, ws, ws_bytes, st);
}

int dedup_product_emit(const uint32_t *vals, int64_t T, const ProductRows &rows, bool by_t, int64_t U, uint64_t *out_xz,
                       double *out_c, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (by_t) return dedup_emit<ProductRows, true>(vals, T, rows, U, out_xz, out_c, ws, ws_bytes, st);
    return dedup_emit<ProductRows, false>(vals, T, rows, U, out_xz, out_c, ws, ws_bytes, st);
}

int dedup_plain_plan(uint64_t *keys, uint32_t *vals, int64_t T, const PlainRows &rows, double thr, int64_t *n_out,
                     int64_t *n_out_host, void *ws, size_t ws_bytes, cudaStream_t st) {
    return dedup_plan<PlainRows, true>(keys, vals, true, T, rows, thr, n_out, n_out_host, ws, ws_bytes, st);
}

int dedup_plain_emit(int64_t T, const PlainRows &rows, int64_t U, uint64_t *out_xz, double *out_c, void *ws,
                     size_t ws_bytes, cudaStream_t st) {
    return dedup_emit<PlainRows, true>(nullptr, T, rows, U, out_xz, out_c, ws, ws_bytes, st);
}

// keys of stored rows: mix64(sketch) with the two low bits cleared (no phase for plain rows)
__global__ void __launch_bounds__(256) plain_keys_kernel(const uint64_t *__restrict__ xz, int64_t T, int words, uint64_t mask,
                                                          uint64_t *__restrict__ keys) {
    const int lane = threadIdx.x & 31;
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= T) return;
    uint64_t h = warp_sketch_row(xz + row * words, words, lane);
    if (lane == 0) keys[row] = (mix64(h) & mask) & ~3ull;
}

}  // namespace symb

using namespace symb;

extern "C" size_t sym_cleanup_ws_bytes(int64_t T, int32_t W) {
    (void)W;
    if (T < 1) T = 1;
    return dedup_ws_bytes(T) + arena_need((size_t)T, 8) + arena_need((size_t)T, 4) + 1024;
}

static int cleanup_check(int64_t T, int32_t W, size_t ws_bytes) {
    SYM_REQUIRE(T >= 0 && T < (int64_t)4000000000LL, "T out of range");
    SYM_REQUIRE(W >= 1, "W must be >= 1");
    if (T > 0 && ws_bytes < sym_cleanup_ws_bytes(T, W)) {
        set_error("workspace too small");
        return SYM_E_WORKSPACE;
    }
    return SYM_OK;
}

extern "C" int sym_cleanup_count(const uint64_t *xz, const double *c, int64_t T, int32_t W, double zero_threshold,
                                 int64_t *n_out, int64_t *n_out_host, void *ws, size_t ws_bytes, void *stream) {
    SYM_TRY(cleanup_check(T, W, ws_bytes));
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) {
        if (n_out) SYM_CUDA_OK(cudaMemsetAsync(n_out, 0, sizeof(int64_t), st));
        if (n_out_host) *n_out_host = 0;
        return SYM_OK;
    }
    Arena ar(ws, ws_bytes);
    uint64_t *keys = ar.take<uint64_t>((size_t)T);
    uint32_t *vals = ar.take<uint32_t>((size_t)T);
    plain_keys_kernel<<<(unsigned)((T * 32 + 255) / 256), 256, 0, st>>>(xz, T, 2 * W, g_key_mask, keys);
    SYM_LAUNCH_OK();
    PlainRows rows{xz, c, 2 * W};
    return dedup_plain_plan(keys, vals, T, rows, zero_threshold, n_out, n_out_host, ar.base + ar.off, ws_bytes - ar.off,
                            st);
}

extern "C" int sym_cleanup_emit(const uint64_t *xz, const double *c, int64_t T, int32_t W, int64_t U, uint64_t *out_xz,
                                double *out_c, void *ws, size_t ws_bytes, void *stream) {
    SYM_TRY(cleanup_check(T, W, ws_bytes));
    if (T == 0 || U == 0) return SYM_OK;
    Arena ar(ws, ws_bytes);
    ar.take<uint64_t>((size_t)T);
    ar.take<uint32_t>((size_t)T);
    PlainRows rows{xz, c, 2 * W};
    return dedup_plain_emit(T, rows, U, out_xz, out_c, ar.base + ar.off, ws_bytes - ar.off, (cudaStream_t)stream);
}

extern "C" int sym_cleanup(const uint64_t *xz, const double *c, int64_t T, int32_t W, double zero_threshold,
                           uint64_t *out_xz, double *out_c, int64_t out_capacity, int64_t *n_out, int64_t *n_out_host,
                           void *ws, size_t ws_bytes, void *stream) {
    int64_t U = 0;
    SYM_TRY(sym_cleanup_count(xz, c, T, W, zero_threshold, n_out, &U, ws, ws_bytes, stream));
    if (n_out_host) *n_out_host = U;
    if (U > out_capacity) {
        set_error("output capacity %lld < %lld surviving terms", (long long)out_capacity, (long long)U);
        return SYM_E_CAPACITY;
    }
    return sym_cleanup_emit(xz, c, T, W, U, out_xz, out_c, ws, ws_bytes, stream);
}
