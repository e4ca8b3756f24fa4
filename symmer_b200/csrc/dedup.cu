// Dedup of terms by hashed 64-bit records with exact row verification, segmented coefficient
// reduction in input order, threshold filter and row emission. Replaces qiskit's `unordered_unique`
// + `np.add.at` (symmer/operators/utils.py:271-278) — HBM-bound integer/byte work, no tensor cores.
//
// Pipeline (all on one stream):
//   1. stable LSD radix sort of the records on their top hash bits -> equal rows become neighbours,
//      ordered by t inside a run
//   2. link: every sorted position finds its nearest earlier twin inside its sort bucket (exact row compare)
//   3. sum: each head adds the coefficients of its group in input (t) order and applies |c| > thr
//   4. exclusive scan of keep flags -> output slots ; compact ; emit rows with 16-byte stores
#include <type_traits>

#include "rows.cuh"
#include "sort.cuh"

namespace symb {

constexpr uint8_t FLAG_HEAD = 0, FLAG_PREV = 1, FLAG_LINK = 2;

// A "bucket" is a maximal run of sorted records sharing the sorted hash prefix (bits >= sort_shift).
// The sort leaves each bucket in input (t) order. On average a bucket holds <= 8 records
// (sort_begin_bit), so the scans below are a handful of cached 8-byte reads per record.
//
// link: every sorted position finds the nearest EARLIER record of its bucket with the same row
// (full hash match, then exact word-by-word compare): none -> HEAD, the immediate predecessor ->
// PREV, further back -> LINK (position stored in link[]).
template <class Rows>
__global__ void __launch_bounds__(256) link_kernel(Rows rows, RecFmt fmt, const uint64_t *__restrict__ sr, int64_t T,
                                                    int sort_shift, uint8_t *__restrict__ flag, uint32_t *__restrict__ link) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    uint8_t f = FLAG_HEAD;
    if (i > 0) {
        const uint64_t ri = sr[i];
        const uint32_t ti = fmt.t(ri);
        for (int64_t j = i - 1; j >= 0; --j) {
            const uint64_t rj = sr[j];
            if (((ri ^ rj) >> sort_shift) != 0) break;   // left the bucket
            if (fmt.same_hash(ri, rj) && rows.equal(ti, fmt.t(rj))) {
                if (j == i - 1) {
                    f = FLAG_PREV;
                } else {
                    f = FLAG_LINK;
                    link[i] = (uint32_t)j;
                }
                break;
            }
        }
    }
    flag[i] = f;
}

__device__ __forceinline__ uint8_t keep_test(double re, double im, double thr) {
    return (thr < 0.0) ? 1 : (hypot(re, im) > thr ? 1 : 0);
}

__device__ __forceinline__ int64_t chain_root(const uint8_t *flag, const uint32_t *link, int64_t i) {
    while (flag[i] != FLAG_HEAD) i = (flag[i] == FLAG_PREV) ? i - 1 : (int64_t)link[i];
    return i;
}

// sum: each HEAD walks forward through its bucket and adds the coefficients of the records whose
// chain leads back to it, in input (t) order like np.add.at — deterministic, no atomics.
template <class Rows, bool BY_T, bool DIRECT>
__global__ void __launch_bounds__(256) sum_kernel(Rows rows, RecFmt fmt, const uint64_t *__restrict__ sr, int64_t T,
                                                   int sort_shift, const uint8_t *__restrict__ flag,
                                                   const uint32_t *__restrict__ link, double thr, double2 *__restrict__ acc,
                                                   uint8_t *__restrict__ keep, uint8_t *__restrict__ multi) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint64_t r0 = sr[i];
    const int64_t d = BY_T ? (int64_t)fmt.t(r0) : i;
    if (flag[i] != FLAG_HEAD) {
        keep[d] = 0;
        return;
    }
    double re = 0.0, im = 0.0;
    bool have = false, is_multi = false;
    bool prev_mine = true;   // is record j-1 a member of this head's group?
    for (int64_t j = i + 1; j < T; ++j) {
        const uint64_t rj = sr[j];
        if (((r0 ^ rj) >> sort_shift) != 0) break;       // end of the bucket
        const uint8_t fj = flag[j];
        bool mine;
        if (fj == FLAG_PREV) mine = prev_mine;
        else if (fj == FLAG_LINK) mine = fmt.same_hash(r0, rj) && chain_root(flag, link, (int64_t)link[j]) == i;
        else mine = false;
        if (mine) {
            if (!have) {
                rows.coeff(fmt.t(r0), fmt.e(r0), re, im);
                have = true;
            }
            double r2, i2;
            rows.coeff(fmt.t(rj), fmt.e(rj), r2, i2);
            re += r2;
            im += i2;
            is_multi = true;
        }
        prev_mine = mine;
    }
    if (is_multi || !DIRECT) {
        if (!have) rows.coeff(fmt.t(r0), fmt.e(r0), re, im);
        acc[d] = make_double2(re, im);
        multi[d] = 1;
        keep[d] = keep_test(re, im, thr);
    } else if (thr < 0.0) {
        keep[d] = 1;         // singleton, no threshold: the coefficient is recomputed at compaction
    } else {
        rows.coeff(fmt.t(r0), fmt.e(r0), re, im);
        keep[d] = keep_test(re, im, thr);
    }
}

// Compaction (emit phase): kept_t[slot] = term index of the survivor, out_c[slot] = its coefficient,
// so that the row-emission kernel has a two-step dependency chain (kept_t -> rows) only.
template <class Rows, bool BY_T>
__global__ void __launch_bounds__(256) compact_kernel(Rows rows, RecFmt fmt, const uint64_t *__restrict__ sr,
                                                       const uint8_t *__restrict__ keep, const uint8_t *__restrict__ multi,
                                                       const uint32_t *__restrict__ slot, const double2 *__restrict__ acc,
                                                       int64_t T, uint32_t *__restrict__ kept_t, double2 *__restrict__ out_c) {
    int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= T) return;
    if (keep[d]) {
        const uint32_t s = slot[d];
        uint32_t t;
        int e = 0;
        if (BY_T) {
            t = (uint32_t)d;
        } else {
            const uint64_t rec = sr[d];
            t = fmt.t(rec);
            e = fmt.e(rec);
        }
        kept_t[s] = t;
        if (multi[d]) {
            out_c[s] = acc[d];
        } else {   // singleton survivor: its coefficient was never materialised
            double re, im;
            rows.coeff(t, e, re, im);
            out_c[s] = make_double2(re, im);
        }
    }
}

__global__ void total_to_i64_kernel(const uint32_t *__restrict__ total, int64_t *__restrict__ n_out) { *n_out = (int64_t)*total; }

// Row emission. One thread per (kept record, 16-byte chunk), EMIT_UN records per thread so that
// 2*EMIT_UN independent 16-byte loads are in flight per thread (the kernel is latency-bound
// otherwise). Stores are streaming (st.global.cs): the output is never re-read, and A/B must stay
// L2-resident. LW: chunks per row = 1 << LW.
int g_emit_variant = 1;  // tuning knob 1: 8 records per thread, plain stores (measured best on B200)

__device__ __forceinline__ void store_streaming(uint4 *p, const uint4 &v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <class Rows, int LW, int EMIT_UN, bool CS>
__global__ void __launch_bounds__(256) emit_kernel(Rows rows, const uint32_t *__restrict__ kept_t, uint32_t U,
                                                    uint4 *__restrict__ out_xz) {
    constexpr uint32_t ROWS_PP = 256u >> LW;
    const uint32_t r_in = threadIdx.x >> LW;
    const uint32_t c = threadIdx.x & ((1u << LW) - 1u);
    uint32_t rec[EMIT_UN], t[EMIT_UN];
    uint4 v[EMIT_UN];
#pragma unroll
    for (int u = 0; u < EMIT_UN; ++u) {
        rec[u] = (blockIdx.x * EMIT_UN + u) * ROWS_PP + r_in;
        t[u] = rec[u] < U ? kept_t[rec[u]] : 0u;
    }
#pragma unroll
    for (int u = 0; u < EMIT_UN; ++u) v[u] = rows.chunk(t[u], (int)c);
#pragma unroll
    for (int u = 0; u < EMIT_UN; ++u) {
        if (rec[u] < U) {
            if (CS) store_streaming(out_xz + (((size_t)rec[u]) << LW) + c, v[u]);
            else out_xz[(((size_t)rec[u]) << LW) + c] = v[u];
        }
    }
}

// generic chunk count (W not a power of two, or W > 16)
template <class Rows>
__global__ void __launch_bounds__(256) emit_generic_kernel(Rows rows, const uint32_t *__restrict__ kept_t, uint32_t U,
                                                            uint32_t chunks, uint4 *__restrict__ out_xz) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t rec = (uint32_t)(g / chunks);
    const uint32_t c = (uint32_t)(g - (size_t)rec * chunks);
    if (rec >= U) return;
    out_xz[g] = rows.chunk(kept_t[rec], (int)c);
}

size_t dedup_ws_bytes(int64_t T) {
    if (T < 1) T = 1;
    size_t n = (size_t)T;
    return arena_need(n, 8)                          // alt record buffer
           + arena_need(record_hist_elems(T), 4)     // radix histograms + scan scratch
           + arena_need(n, 1)                        // flag
           + arena_need(n, 4)                        // link
           + arena_need(n, 16)                       // acc
           + arena_need(n, 1)                        // keep
           + arena_need(n, 1)                        // multi
           + arena_need(n, 4)                        // slot
           + arena_need(n, 4)                        // kept
           + arena_need(scan_scratch_elems(T), 4)    // scan scratch
           + arena_need(4, 4) + 4096;
}

// Sort on log2(T) + g_sort_extra_bits hash bits, rounded up to whole 8-bit passes. A sort bucket then
// holds T / 2^bits records on average, which link/sum resolve exactly with in-bucket scans: fewer
// bits save radix passes (24 B of HBM traffic per record each) but lengthen those scans.
int g_sort_extra_bits = 5;  // tuning knob 3
static int sort_begin_bit(int64_t T, RecFmt fmt) {
    int lg = t_bits_for(T);
    int want = ((lg + g_sort_extra_bits + 7) / 8) * 8;
    if (want < 8) want = 8;
    int begin = 64 - want;
    if (begin < fmt.tb + 2) begin = fmt.tb + 2;
    return begin;
}

static bool sorted_in_alt(int begin_bit) {
    int passes = (64 - begin_bit + 7) / 8;
    return (passes & 1) != 0;
}

struct DedupLayout {
    uint64_t *alt;
    uint32_t *hist;
    uint8_t *flag;
    uint32_t *link;
    double2 *acc;
    uint8_t *keep;
    uint8_t *multi;
    uint32_t *slot;
    uint32_t *kept;
    uint32_t *scratch;
    uint32_t *total;
    bool ok;
};

static DedupLayout dedup_layout(void *ws, size_t ws_bytes, int64_t T) {
    Arena ar(ws, ws_bytes);
    DedupLayout L;
    L.alt = ar.take<uint64_t>((size_t)T);
    L.hist = ar.take<uint32_t>(record_hist_elems(T));
    L.flag = ar.take<uint8_t>((size_t)T);
    L.link = ar.take<uint32_t>((size_t)T);
    L.acc = ar.take<double2>((size_t)T);
    L.keep = ar.take<uint8_t>((size_t)T);
    L.multi = ar.take<uint8_t>((size_t)T);
    L.slot = ar.take<uint32_t>((size_t)T);
    L.kept = ar.take<uint32_t>((size_t)T);
    L.scratch = ar.take<uint32_t>(scan_scratch_elems(T));
    L.total = ar.take<uint32_t>(4);
    L.ok = L.total != nullptr;
    return L;
}

// Phase 1: everything up to the survivor count (synchronises the stream once to read it).
template <class Rows, bool BY_T>
static int dedup_plan(uint64_t *recs, int64_t T, RecFmt fmt, const Rows &rows, double thr, int64_t *n_out,
                      int64_t *n_out_host, void *ws, size_t ws_bytes, cudaStream_t st) {
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
    const int begin = sort_begin_bit(T, fmt);
    uint64_t *sr = nullptr;
    SYM_TRY(radix_sort_records(recs, L.alt, T, begin, L.hist, &sr, st));

    const unsigned nb = (unsigned)((T + 255) / 256);
    link_kernel<Rows><<<nb, 256, 0, st>>>(rows, fmt, sr, T, begin, L.flag, L.link);
    SYM_LAUNCH_OK();
    // singleton survivors skip the acc[] round trip when their coefficient can be recomputed at
    // compaction: always in sorted order (the record carries t and the phase), and for stored rows
    constexpr bool DIRECT = !BY_T || std::is_same<Rows, PlainRows>::value;
    SYM_CUDA_OK(cudaMemsetAsync(L.multi, 0, (size_t)T, st));
    sum_kernel<Rows, BY_T, DIRECT><<<nb, 256, 0, st>>>(rows, fmt, sr, T, begin, L.flag, L.link, thr, L.acc, L.keep, L.multi);
    SYM_LAUNCH_OK();
    SYM_TRY(scan_exclusive_u8(L.keep, L.slot, T, L.total, L.scratch, st));
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

// Phase 2: write the U surviving rows + coefficients (asynchronous). `recs` is the same buffer
// that was given to the plan phase.
template <class Rows, bool BY_T>
static int dedup_emit(const uint64_t *recs, int64_t T, RecFmt fmt, const Rows &rows, int64_t U, uint64_t *out_xz,
                      double *out_c, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (T == 0 || U == 0) return SYM_OK;
    DedupLayout L = dedup_layout(ws, ws_bytes, T);
    if (!L.ok) {
        set_error("workspace arena exhausted");
        return SYM_E_WORKSPACE;
    }
    const uint64_t *sr = (T > 1 && sorted_in_alt(sort_begin_bit(T, fmt))) ? L.alt : recs;
    const uint32_t chunks = (uint32_t)(rows.words / 2);
    uint4 *o = reinterpret_cast<uint4 *>(out_xz);
    double2 *oc = reinterpret_cast<double2 *>(out_c);
    compact_kernel<Rows, BY_T><<<(unsigned)((T + 255) / 256), 256, 0, st>>>(rows, fmt, sr, L.keep, L.multi, L.slot, L.acc, T,
                                                                           L.kept, oc);
    SYM_LAUNCH_OK();
#define EMIT_LAUNCH(LW, UN, CS)                                                                   \
    {                                                                                             \
        const uint32_t rows_per_block = (256u >> LW) * UN;                                        \
        const unsigned nbe = (unsigned)((U + rows_per_block - 1) / rows_per_block);               \
        emit_kernel<Rows, LW, UN, CS><<<nbe, 256, 0, st>>>(rows, L.kept, (uint32_t)U, o);         \
    }
#define EMIT_CASE(LW)                                       \
    switch (g_emit_variant) {                               \
        case 1: EMIT_LAUNCH(LW, 8, false); break;           \
        case 2: EMIT_LAUNCH(LW, 8, true); break;            \
        case 3: EMIT_LAUNCH(LW, 4, false); break;           \
        case 4: EMIT_LAUNCH(LW, 2, false); break;           \
        case 5: EMIT_LAUNCH(LW, 1, true); break;            \
        default: EMIT_LAUNCH(LW, 1, false); break;          \
    }
    switch (chunks) {
        case 1: EMIT_CASE(0); break;
        case 2: EMIT_CASE(1); break;
        case 4: EMIT_CASE(2); break;
        case 8: EMIT_CASE(3); break;
        case 16: EMIT_CASE(4); break;
        default: {
            const size_t threads = (size_t)U * chunks;
            emit_generic_kernel<Rows><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(rows, L.kept, (uint32_t)U, chunks, o);
        } break;
    }
#undef EMIT_CASE
#undef EMIT_LAUNCH
    SYM_LAUNCH_OK();
    return SYM_OK;
}

int dedup_product_plan(uint64_t *recs, int64_t T, RecFmt fmt, const ProductRows &rows, bool by_t, double thr,
                       int64_t *n_out, int64_t *n_out_host, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (by_t) return dedup_plan<ProductRows, true>(recs, T, fmt, rows, thr, n_out, n_out_host, ws, ws_bytes, st);
    return dedup_plan<ProductRows, false>(recs, T, fmt, rows, thr, n_out, n_out_host, ws, ws_bytes, st);
}

int dedup_product_emit(const uint64_t *recs, int64_t T, RecFmt fmt, const ProductRows &rows, bool by_t, int64_t U,
                       uint64_t *out_xz, double *out_c, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (by_t) return dedup_emit<ProductRows, true>(recs, T, fmt, rows, U, out_xz, out_c, ws, ws_bytes, st);
    return dedup_emit<ProductRows, false>(recs, T, fmt, rows, U, out_xz, out_c, ws, ws_bytes, st);
}

int dedup_plain_plan(uint64_t *recs, int64_t T, RecFmt fmt, const PlainRows &rows, double thr, int64_t *n_out,
                     int64_t *n_out_host, void *ws, size_t ws_bytes, cudaStream_t st) {
    return dedup_plan<PlainRows, true>(recs, T, fmt, rows, thr, n_out, n_out_host, ws, ws_bytes, st);
}

int dedup_plain_emit(const uint64_t *recs, int64_t T, RecFmt fmt, const PlainRows &rows, int64_t U, uint64_t *out_xz,
                     double *out_c, void *ws, size_t ws_bytes, cudaStream_t st) {
    return dedup_emit<PlainRows, true>(recs, T, fmt, rows, U, out_xz, out_c, ws, ws_bytes, st);
}

// records of stored rows: hash of the row sketch, t = row index, e = 0
__global__ void __launch_bounds__(256) plain_records_kernel(const uint64_t *__restrict__ xz, int64_t T, int words, uint64_t mask,
                                                             RecFmt fmt, uint64_t *__restrict__ recs) {
    const int lane = threadIdx.x & 31;
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= T) return;
    uint64_t h = warp_sketch_row(xz + row * words, words, lane);
    if (lane == 0) recs[row] = fmt.make(mix64(h) & mask, (uint64_t)row, 0);
}

__global__ void __launch_bounds__(256) plain_records8_kernel(const uint64_t *__restrict__ xz, int64_t T, int words, uint64_t mask,
                                                              RecFmt fmt, uint64_t *__restrict__ recs) {
    const int lane = threadIdx.x & 31;
    int64_t row = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 2) + (lane >> 3);
    const bool ok = row < T;
    uint64_t h = group8_sketch_row(xz + (ok ? row : 0) * words, words, lane & 7);
    if (ok && (lane & 7) == 0) recs[row] = fmt.make(mix64(h) & mask, (uint64_t)row, 0);
}

}  // namespace symb

using namespace symb;

extern "C" size_t sym_cleanup_ws_bytes(int64_t T, int32_t W) {
    (void)W;
    if (T < 1) T = 1;
    return dedup_ws_bytes(T) + arena_need((size_t)T, 8) + 1024;
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
    uint64_t *recs = ar.take<uint64_t>((size_t)T);
    RecFmt fmt{t_bits_for(T)};
    if (group8_ok(W))
        plain_records8_kernel<<<(unsigned)((((T + 3) / 4) * 32 + 255) / 256), 256, 0, st>>>(xz, T, 2 * W, g_key_mask, fmt, recs);
    else
        plain_records_kernel<<<(unsigned)((T * 32 + 255) / 256), 256, 0, st>>>(xz, T, 2 * W, g_key_mask, fmt, recs);
    SYM_LAUNCH_OK();
    PlainRows rows{xz, c, 2 * W};
    return dedup_plain_plan(recs, T, fmt, rows, zero_threshold, n_out, n_out_host, ar.base + ar.off, ws_bytes - ar.off, st);
}

extern "C" int sym_cleanup_emit(const uint64_t *xz, const double *c, int64_t T, int32_t W, int64_t U, uint64_t *out_xz,
                                double *out_c, void *ws, size_t ws_bytes, void *stream) {
    SYM_TRY(cleanup_check(T, W, ws_bytes));
    if (T == 0 || U == 0) return SYM_OK;
    Arena ar(ws, ws_bytes);
    uint64_t *recs = ar.take<uint64_t>((size_t)T);
    RecFmt fmt{t_bits_for(T)};
    PlainRows rows{xz, c, 2 * W};
    return dedup_plain_emit(recs, T, fmt, rows, U, out_xz, out_c, ar.base + ar.off, ws_bytes - ar.off,
                            (cudaStream_t)stream);
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
