// PauliwordOp multiplication (symmer/operators/base.py:764-794): all-pairs cross terms as XOR of
// packed rows with the i-phase from popcounts of X&Z overlaps.
//
// Two forms:
//   * sym_cross_mul     — materialises the M*N cross terms (parity checks, small products);
//   * pair records      — never materialises rows: for every pair (p, q) emit one 64-bit record
//                         key = (mix64(sketch(A[p]) ^ sketch(B[q])) & ~3) | phase_exponent(p, q)
//                         which is all the dedup needs; rows are rebuilt only for survivors.
// Integer/bit work bound by the LOP3 pipe and by the 8 B/pair store; no tensor cores.
#include "rows.cuh"
#include "sort.cuh"

namespace symb {

// phase exponent e with coefficient factor i^e = (-1)^{|xa&zb|} * i^{(3(Ya+Yb)+Yout) mod 4}  (base.py:785-788)
__device__ __forceinline__ int finish_phase(int ya, int yb, uint64_t s, uint64_t c0, uint64_t c1) {
    int yout = __popcll(c0) + 2 * __popcll(c1);
    int sgn = __popcll(s) & 1;
    return (3 * (ya + yb) + yout + 2 * sgn) & 3;
}

// ---------------------------------------------------------------------------------------------
// tiled pair kernel: thread = one A row held in registers, CTA streams QCH B rows through smem
// ---------------------------------------------------------------------------------------------
constexpr int PAIR_THREADS = 256;
constexpr int PAIR_QCH = 64;

template <int WT>
__global__ void __launch_bounds__(PAIR_THREADS) pair_records_kernel(
    const uint64_t *__restrict__ a_xz, const uint64_t *__restrict__ a_sk, const int32_t *__restrict__ a_y, uint32_t M_total,
    uint32_t p_begin, uint32_t p_end, const uint64_t *__restrict__ b_xz, const uint64_t *__restrict__ b_sk,
    const int32_t *__restrict__ b_y, uint32_t N, uint32_t q_off, int W, uint64_t key_mask, RecFmt fmt,
    uint64_t *__restrict__ recs) {
    // B tile: QCH rows, (x_w, z_w) interleaved so one 16-byte shared load feeds a word step
    __shared__ ulonglong2 sb[PAIR_QCH][WT];
    __shared__ uint64_t sb_sk[PAIR_QCH];
    __shared__ int sb_y[PAIR_QCH];

    const uint32_t q0 = blockIdx.y * PAIR_QCH;
    const uint32_t nq = min((uint32_t)PAIR_QCH, N - q0);
    for (int i = threadIdx.x; i < PAIR_QCH * WT; i += PAIR_THREADS) {
        int qi = i / WT, w = i % WT;
        ulonglong2 v = make_ulonglong2(0ull, 0ull);
        if ((uint32_t)qi < nq && w < W) {
            const uint64_t *row = b_xz + (size_t)(q0 + qi) * 2 * W;
            v.x = row[w];
            v.y = row[W + w];
        }
        sb[qi][w] = v;
    }
    for (int i = threadIdx.x; i < PAIR_QCH; i += PAIR_THREADS) {
        sb_sk[i] = ((uint32_t)i < nq) ? b_sk[q0 + i] : 0ull;
        sb_y[i] = ((uint32_t)i < nq) ? b_y[q0 + i] : 0;
    }

    const uint32_t p = p_begin + blockIdx.x * PAIR_THREADS + threadIdx.x;
    const bool active = p < p_end;
    uint64_t xa[WT], za[WT];
    uint64_t ska = 0;
    int ya = 0;
    if (active) {
        const uint64_t *row = a_xz + (size_t)p * 2 * W;
#pragma unroll
        for (int w = 0; w < WT; ++w) {
            xa[w] = (w < W) ? row[w] : 0ull;
            za[w] = (w < W) ? row[W + w] : 0ull;
        }
        ska = a_sk[p];
        ya = a_y[p];
    } else {
#pragma unroll
        for (int w = 0; w < WT; ++w) xa[w] = za[w] = 0ull;
    }
    __syncthreads();
    if (!active) return;

    const uint32_t m_blk = p_end - p_begin;
    const uint32_t p_local = p - p_begin;
    for (uint32_t qi = 0; qi < nq; ++qi) {
        uint64_t s = 0, c0 = 0, c1 = 0;
#pragma unroll
        for (int w = 0; w < WT; ++w) {
            const ulonglong2 b = sb[qi][w];  // broadcast read
            s ^= xa[w] & b.y;
            const uint64_t v = (xa[w] ^ b.x) & (za[w] ^ b.y);
            c1 ^= c0 & v;
            c0 ^= v;
        }
        const int e = finish_phase(ya, sb_y[qi], s, c0, c1);
        const size_t j = (size_t)(q0 + qi) * m_blk + p_local;
        recs[j] = fmt.make(mix64(ska ^ sb_sk[qi]) & key_mask, (uint64_t)(q_off + q0 + qi) * M_total + p, e);
    }
}

// generic fallback for W > 16: one thread per pair, rows read through L1/L2
__global__ void __launch_bounds__(256) pair_records_generic_kernel(
    const uint64_t *__restrict__ a_xz, const uint64_t *__restrict__ a_sk, const int32_t *__restrict__ a_y, uint32_t M_total,
    uint32_t p_begin, uint32_t p_end, const uint64_t *__restrict__ b_xz, const uint64_t *__restrict__ b_sk,
    const int32_t *__restrict__ b_y, uint32_t N, uint32_t q_off, int W, uint64_t key_mask, RecFmt fmt,
    uint64_t *__restrict__ recs) {
    const uint32_t m_blk = p_end - p_begin;
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (size_t)m_blk * N) return;
    uint32_t q = (uint32_t)(j / m_blk);
    uint32_t p = p_begin + (uint32_t)(j - (size_t)q * m_blk);
    const uint64_t *ra = a_xz + (size_t)p * 2 * W, *rb = b_xz + (size_t)q * 2 * W;
    uint64_t s = 0, c0 = 0, c1 = 0;
    for (int w = 0; w < W; ++w) {
        uint64_t xa = ra[w], za = ra[W + w], xb = rb[w], zb = rb[W + w];
        s ^= xa & zb;
        uint64_t v = (xa ^ xb) & (za ^ zb);
        c1 ^= c0 & v;
        c0 ^= v;
    }
    int e = finish_phase(a_y[p], b_y[q], s, c0, c1);
    recs[j] = fmt.make(mix64(a_sk[p] ^ b_sk[q]) & key_mask, (uint64_t)(q_off + q) * M_total + p, e);
}

static int launch_pair_records(const uint64_t *a_xz, const uint64_t *a_sk, const int32_t *a_y, int64_t M_total,
                               int64_t p_begin, int64_t p_end, const uint64_t *b_xz, const uint64_t *b_sk,
                               const int32_t *b_y, int64_t N, int W, RecFmt fmt, uint64_t *recs, cudaStream_t st,
                               int64_t q_off = 0) {
    // b_xz / b_sk / b_y point at B row q_off; N rows from there. t uses the global index q_off + q.
    const int64_t m_blk = p_end - p_begin;
    if (m_blk <= 0 || N <= 0) return SYM_OK;
    dim3 grid((unsigned)((m_blk + PAIR_THREADS - 1) / PAIR_THREADS), (unsigned)((N + PAIR_QCH - 1) / PAIR_QCH));
#define PAIR_CASE(WT)                                                                                               \
    pair_records_kernel<WT><<<grid, PAIR_THREADS, 0, st>>>(a_xz, a_sk, a_y, (uint32_t)M_total, (uint32_t)p_begin,   \
                                                           (uint32_t)p_end, b_xz, b_sk, b_y, (uint32_t)N,           \
                                                           (uint32_t)q_off, W, g_key_mask, fmt, recs)
    if (grid.y > 65535) {
        set_error("too many B rows for one launch (N=%lld)", (long long)N);
        return SYM_E_UNSUPPORTED;
    }
    if (W <= 1) PAIR_CASE(1);
    else if (W <= 2) PAIR_CASE(2);
    else if (W <= 4) PAIR_CASE(4);
    else if (W <= 8) PAIR_CASE(8);
    else if (W <= 16) PAIR_CASE(16);
    else {
        int64_t total = m_blk * N;
        pair_records_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
            a_xz, a_sk, a_y, (uint32_t)M_total, (uint32_t)p_begin, (uint32_t)p_end, b_xz, b_sk, b_y, (uint32_t)N,
            (uint32_t)q_off, W, g_key_mask, fmt, recs);
    }
#undef PAIR_CASE
    SYM_LAUNCH_OK();
    return SYM_OK;
}

// ---------------------------------------------------------------------------------------------
// materialised cross terms
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cross_rows_kernel(const uint64_t *__restrict__ a, uint32_t M, const uint64_t *__restrict__ b,
                                                          int64_t total_words, int words, uint64_t *__restrict__ out) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_words) return;
    int64_t t = g / words;
    int k = (int)(g - t * words);
    uint32_t q = (uint32_t)(t / M), p = (uint32_t)(t - (int64_t)q * M);
    out[g] = a[(size_t)p * words + k] ^ b[(size_t)q * words + k];
}

__global__ void __launch_bounds__(256) cross_coeff_kernel(const uint64_t *__restrict__ a, const double *__restrict__ ac, uint32_t M,
                                                           const uint64_t *__restrict__ b, const double *__restrict__ bc,
                                                           int64_t T, int W, double2 *__restrict__ out_c) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    uint32_t q = (uint32_t)(t / M), p = (uint32_t)(t - (int64_t)q * M);
    const uint64_t *ra = a + (size_t)p * 2 * W, *rb = b + (size_t)q * 2 * W;
    uint64_t s = 0, c0 = 0, c1 = 0;
    int ya = 0, yb = 0;
    for (int w = 0; w < W; ++w) {
        uint64_t xa = ra[w], za = ra[W + w], xb = rb[w], zb = rb[W + w];
        ya += __popcll(xa & za);
        yb += __popcll(xb & zb);
        s ^= xa & zb;
        uint64_t v = (xa ^ xb) & (za ^ zb);
        c1 ^= c0 & v;
        c0 ^= v;
    }
    int e = finish_phase(ya, yb, s, c0, c1);
    double re, im;
    cmul(ac[2 * (size_t)p], ac[2 * (size_t)p + 1], bc[2 * (size_t)q], bc[2 * (size_t)q + 1], re, im);
    mul_i_pow(re, im, e);
    out_c[t] = make_double2(re, im);
}

}  // namespace symb

using namespace symb;

extern "C" int sym_cross_mul(const uint64_t *a_xz, const double *a_c, int64_t M, const uint64_t *b_xz, const double *b_c,
                             int64_t N, int32_t W, uint64_t *out_xz, double *out_c, void *stream) {
    SYM_REQUIRE(M >= 0 && N >= 0 && W >= 1, "bad size");
    SYM_REQUIRE(M < (int64_t)1 << 31 && N < (int64_t)1 << 31, "operand too large");
    int64_t T = M * N;
    if (T == 0) return SYM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t total_words = T * 2 * W;
    cross_rows_kernel<<<(unsigned)((total_words + 255) / 256), 256, 0, st>>>(a_xz, (uint32_t)M, b_xz, total_words, 2 * W,
                                                                            out_xz);
    SYM_LAUNCH_OK();
    cross_coeff_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(a_xz, a_c, (uint32_t)M, b_xz, b_c, T, W,
                                                                   reinterpret_cast<double2 *>(out_c));
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" size_t sym_pair_records_ws_bytes(int64_t M_total, int64_t N, int32_t W) {
    (void)W;
    return arena_need((size_t)(M_total > 0 ? M_total : 1), 8) + arena_need((size_t)(N > 0 ? N : 1), 8) +
           arena_need((size_t)(M_total > 0 ? M_total : 1), 4) + arena_need((size_t)(N > 0 ? N : 1), 4) + 1024;
}

static int prepare_operand_tables(const uint64_t *a_xz, int64_t M, const uint64_t *b_xz, int64_t N, int W, Arena &ar,
                                  uint64_t *&a_sk, uint64_t *&b_sk, int32_t *&a_y, int32_t *&b_y, cudaStream_t st) {
    a_sk = ar.take<uint64_t>((size_t)(M > 0 ? M : 1));
    b_sk = ar.take<uint64_t>((size_t)(N > 0 ? N : 1));
    a_y = ar.take<int32_t>((size_t)(M > 0 ? M : 1));
    b_y = ar.take<int32_t>((size_t)(N > 0 ? N : 1));
    if (!b_y) {
        set_error("workspace arena exhausted");
        return SYM_E_WORKSPACE;
    }
    SYM_TRY(sym_sketch_rows(a_xz, M, W, a_sk, st));
    SYM_TRY(sym_sketch_rows(b_xz, N, W, b_sk, st));
    SYM_TRY(sym_ycount(a_xz, M, W, a_y, st));
    SYM_TRY(sym_ycount(b_xz, N, W, b_y, st));
    return SYM_OK;
}

extern "C" int sym_pair_records(const uint64_t *a_xz, int64_t M_total, int64_t p_begin, int64_t p_end,
                                const uint64_t *b_xz, int64_t N, int32_t W, uint64_t *recs, void *ws, size_t ws_bytes,
                                void *stream) {
    SYM_REQUIRE(M_total >= 0 && N >= 0 && W >= 1, "bad size");
    SYM_REQUIRE(0 <= p_begin && p_begin <= p_end && p_end <= M_total, "bad row block");
    SYM_REQUIRE(M_total * N < (int64_t)4000000000LL, "M_total*N must be < 4e9");
    if (ws_bytes < sym_pair_records_ws_bytes(M_total, N, W)) {
        set_error("workspace too small");
        return SYM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(ws, ws_bytes);
    uint64_t *a_sk, *b_sk;
    int32_t *a_y, *b_y;
    SYM_TRY(prepare_operand_tables(a_xz, M_total, b_xz, N, W, ar, a_sk, b_sk, a_y, b_y, st));
    RecFmt fmt{t_bits_for(M_total * N)};
    return launch_pair_records(a_xz, a_sk, a_y, M_total, p_begin, p_end, b_xz, b_sk, b_y, N, W, fmt, recs, st);
}

extern "C" int sym_pair_records_blocks(const uint64_t *a_xz, int64_t M_total, const uint64_t *b_xz, int64_t N, int32_t W,
                                       const int64_t *blocks_host, int32_t nblk, uint64_t *recs, void *ws,
                                       size_t ws_bytes, void *stream) {
    SYM_REQUIRE(M_total >= 0 && N >= 0 && W >= 1 && nblk >= 0, "bad size");
    SYM_REQUIRE(M_total * N < (int64_t)4000000000LL, "M_total*N must be < 4e9");
    for (int b = 0; b < nblk; ++b) {
        const int64_t *q = blocks_host + 4 * b;
        SYM_REQUIRE(0 <= q[0] && q[0] <= q[1] && q[1] <= M_total, "bad A row block");
        SYM_REQUIRE(0 <= q[2] && q[2] <= q[3] && q[3] <= N, "bad B row block");
    }
    if (ws_bytes < sym_pair_records_ws_bytes(M_total, N, W)) {
        set_error("workspace too small");
        return SYM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(ws, ws_bytes);
    uint64_t *a_sk, *b_sk;
    int32_t *a_y, *b_y;
    SYM_TRY(prepare_operand_tables(a_xz, M_total, b_xz, N, W, ar, a_sk, b_sk, a_y, b_y, st));
    RecFmt fmt{t_bits_for(M_total * N)};
    size_t off = 0;
    for (int b = 0; b < nblk; ++b) {
        const int64_t p0 = blocks_host[4 * b], p1 = blocks_host[4 * b + 1], q0 = blocks_host[4 * b + 2],
                      q1 = blocks_host[4 * b + 3];
        SYM_TRY(launch_pair_records(a_xz, a_sk, a_y, M_total, p0, p1, b_xz + (size_t)q0 * 2 * W, b_sk + q0, b_y + q0,
                                    q1 - q0, W, fmt, recs + off, st, q0));
        off += (size_t)(p1 - p0) * (size_t)(q1 - q0);
    }
    return SYM_OK;
}

extern "C" size_t sym_partition_ws_bytes(int64_t T) {
    if (T < 1) T = 1;
    return arena_need(record_hist_elems(T), 4) + 1024;
}

extern "C" int sym_partition_records(const uint64_t *recs, int64_t T, int32_t log2_parts, uint64_t *out_recs,
                                     int64_t *counts, void *ws, size_t ws_bytes, void *stream) {
    SYM_REQUIRE(T >= 0 && T < (int64_t)4000000000LL, "T out of range");
    SYM_REQUIRE(log2_parts >= 0 && log2_parts <= 8, "log2_parts must be in [0,8]");
    if (ws_bytes < sym_partition_ws_bytes(T)) {
        set_error("workspace too small");
        return SYM_E_WORKSPACE;
    }
    Arena ar(ws, ws_bytes);
    uint32_t *hist = ar.take<uint32_t>(record_hist_elems(T));
    return radix_partition_records(recs, out_recs, T, log2_parts, counts, hist, (cudaStream_t)stream);
}

extern "C" size_t sym_dedup_records_ws_bytes(int64_t T, int32_t W) {
    (void)W;
    return dedup_ws_bytes(T);
}

extern "C" int sym_dedup_records_count(uint64_t *recs, int64_t T, const uint64_t *a_xz, const double *a_c,
                                       int64_t M_total, const uint64_t *b_xz, const double *b_c, int64_t N, int32_t W,
                                       double zero_threshold, int64_t *n_out, int64_t *n_out_host, void *ws,
                                       size_t ws_bytes, void *stream) {
    SYM_REQUIRE(T >= 0 && T < (int64_t)4000000000LL, "T out of range");
    SYM_REQUIRE(M_total >= 1 && N >= 1 && W >= 1, "bad size");
    SYM_REQUIRE(M_total * N < (int64_t)4000000000LL, "M_total*N must be < 4e9");
    ProductRows rows{a_xz, b_xz, a_c, b_c, (uint32_t)M_total, 2 * W, (uint32_t)N};
    RecFmt fmt{t_bits_for(M_total * N)};
    return dedup_product_plan(recs, T, fmt, rows, false, zero_threshold, n_out, n_out_host, ws, ws_bytes,
                              (cudaStream_t)stream);
}

extern "C" int sym_dedup_records_emit(const uint64_t *recs, int64_t T, const uint64_t *a_xz, const double *a_c,
                                      int64_t M_total, const uint64_t *b_xz, const double *b_c, int64_t N, int32_t W,
                                      int64_t U, uint64_t *out_xz, double *out_c, void *ws, size_t ws_bytes, void *stream) {
    SYM_REQUIRE(T >= 0 && U >= 0 && U <= T, "bad counts");
    SYM_REQUIRE(M_total >= 1 && N >= 1 && W >= 1, "bad size");
    ProductRows rows{a_xz, b_xz, a_c, b_c, (uint32_t)M_total, 2 * W, (uint32_t)N};
    RecFmt fmt{t_bits_for(M_total * N)};
    return dedup_product_emit(recs, T, fmt, rows, false, U, out_xz, out_c, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int sym_dedup_records(uint64_t *recs, int64_t T, const uint64_t *a_xz, const double *a_c, int64_t M_total,
                                 const uint64_t *b_xz, const double *b_c, int64_t N, int32_t W, double zero_threshold,
                                 uint64_t *out_xz, double *out_c, int64_t out_capacity, int64_t *n_out,
                                 int64_t *n_out_host, void *ws, size_t ws_bytes, void *stream) {
    int64_t U = 0;
    SYM_TRY(sym_dedup_records_count(recs, T, a_xz, a_c, M_total, b_xz, b_c, N, W, zero_threshold, n_out, &U, ws, ws_bytes,
                                    stream));
    if (n_out_host) *n_out_host = U;
    if (U > out_capacity) {
        set_error("output capacity %lld < %lld surviving terms", (long long)out_capacity, (long long)U);
        return SYM_E_CAPACITY;
    }
    return sym_dedup_records_emit(recs, T, a_xz, a_c, M_total, b_xz, b_c, N, W, U, out_xz, out_c, ws, ws_bytes, stream);
}

extern "C" size_t sym_mul_cleanup_ws_bytes(int64_t M, int64_t N, int32_t W) {
    int64_t T = M * N;
    if (T < 1) T = 1;
    return sym_pair_records_ws_bytes(M, N, W) + arena_need((size_t)T, 8) + dedup_ws_bytes(T) + 1024;
}

// Below this many cross terms the product is emitted in first-occurrence (reference) order; above,
// in sorted-hash order, which keeps every pass of the dedup streaming (no T-sized scatters).
static int64_t g_by_t_limit = (int64_t)1 << 22;
namespace symb { extern int g_emit_variant; extern int g_scatter_variant; extern int g_sort_extra_bits; extern int g_apply_variant; extern int g_rref_variant; }

extern "C" int sym_set_tuning(int32_t which, int64_t value) {
    if (which == 0) {
        g_by_t_limit = value;
        return SYM_OK;
    }
    if (which == 1) {
        symb::g_emit_variant = (int)value;
        return SYM_OK;
    }
    if (which == 2) {
        symb::g_scatter_variant = (int)value;
        return SYM_OK;
    }
    if (which == 3) {
        symb::g_sort_extra_bits = (int)value;
        return SYM_OK;
    }
    if (which == 4) {
        symb::g_apply_variant = (int)value;
        return SYM_OK;
    }
    if (which == 5) {
        symb::g_rref_variant = (int)value;
        return SYM_OK;
    }
    set_error("unknown tuning knob %d", which);
    return SYM_E_INVALID;
}

extern "C" int sym_set_emit_events(void *before_event, void *after_event) {
    symb::set_emit_events((cudaEvent_t)before_event, (cudaEvent_t)after_event);
    return SYM_OK;
}

// workspace layout shared by the count and emit phases
struct MulPlan {
    uint64_t *a_sk, *b_sk;
    int32_t *a_y, *b_y;
    uint64_t *recs;
    void *rest;
    size_t rest_bytes;
    bool ok;
};

static MulPlan mul_plan_layout(void *ws, size_t ws_bytes, int64_t M, int64_t N) {
    Arena ar(ws, ws_bytes);
    MulPlan P;
    const int64_t T = M * N;
    P.a_sk = ar.take<uint64_t>((size_t)(M > 0 ? M : 1));
    P.b_sk = ar.take<uint64_t>((size_t)(N > 0 ? N : 1));
    P.a_y = ar.take<int32_t>((size_t)(M > 0 ? M : 1));
    P.b_y = ar.take<int32_t>((size_t)(N > 0 ? N : 1));
    P.recs = ar.take<uint64_t>((size_t)(T > 0 ? T : 1));
    P.ok = P.recs != nullptr;
    P.rest = ar.base + ar.off;
    P.rest_bytes = ws_bytes > ar.off ? ws_bytes - ar.off : 0;
    return P;
}

static int mul_check(int64_t M, int64_t N, int32_t W, size_t ws_bytes) {
    SYM_REQUIRE(M >= 0 && N >= 0 && W >= 1, "bad size");
    SYM_REQUIRE(M * N < (int64_t)4000000000LL, "M*N must be < 4e9 cross terms per call");
    if (M * N > 0 && ws_bytes < sym_mul_cleanup_ws_bytes(M, N, W)) {
        set_error("workspace too small: need %zu", sym_mul_cleanup_ws_bytes(M, N, W));
        return SYM_E_WORKSPACE;
    }
    return SYM_OK;
}

extern "C" int sym_mul_cleanup_count(const uint64_t *a_xz, const double *a_c, int64_t M, const uint64_t *b_xz,
                                     const double *b_c, int64_t N, int32_t W, double zero_threshold, int64_t *n_out,
                                     int64_t *n_out_host, void *ws, size_t ws_bytes, void *stream) {
    SYM_TRY(mul_check(M, N, W, ws_bytes));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t T = M * N;
    if (T == 0) {
        if (n_out) SYM_CUDA_OK(cudaMemsetAsync(n_out, 0, sizeof(int64_t), st));
        if (n_out_host) *n_out_host = 0;
        return SYM_OK;
    }
    MulPlan P = mul_plan_layout(ws, ws_bytes, M, N);
    if (!P.ok) {
        set_error("workspace arena exhausted");
        return SYM_E_WORKSPACE;
    }
    SYM_TRY(sym_sketch_rows(a_xz, M, W, P.a_sk, st));
    SYM_TRY(sym_sketch_rows(b_xz, N, W, P.b_sk, st));
    SYM_TRY(sym_ycount(a_xz, M, W, P.a_y, st));
    SYM_TRY(sym_ycount(b_xz, N, W, P.b_y, st));
    RecFmt fmt{t_bits_for(T)};
    SYM_TRY(launch_pair_records(a_xz, P.a_sk, P.a_y, M, 0, M, b_xz, P.b_sk, P.b_y, N, W, fmt, P.recs, st));
    ProductRows rows{a_xz, b_xz, a_c, b_c, (uint32_t)M, 2 * W, (uint32_t)N};
    return dedup_product_plan(P.recs, T, fmt, rows, T <= g_by_t_limit, zero_threshold, n_out, n_out_host, P.rest,
                              P.rest_bytes, st);
}

extern "C" int sym_mul_cleanup_emit(const uint64_t *a_xz, const double *a_c, int64_t M, const uint64_t *b_xz,
                                    const double *b_c, int64_t N, int32_t W, int64_t U, uint64_t *out_xz, double *out_c,
                                    void *ws, size_t ws_bytes, void *stream) {
    SYM_TRY(mul_check(M, N, W, ws_bytes));
    const int64_t T = M * N;
    if (T == 0 || U == 0) return SYM_OK;
    MulPlan P = mul_plan_layout(ws, ws_bytes, M, N);
    ProductRows rows{a_xz, b_xz, a_c, b_c, (uint32_t)M, 2 * W, (uint32_t)N};
    RecFmt fmt{t_bits_for(T)};
    return dedup_product_emit(P.recs, T, fmt, rows, T <= g_by_t_limit, U, out_xz, out_c, P.rest, P.rest_bytes,
                              (cudaStream_t)stream);
}

extern "C" int sym_mul_cleanup(const uint64_t *a_xz, const double *a_c, int64_t M, const uint64_t *b_xz, const double *b_c,
                               int64_t N, int32_t W, double zero_threshold, uint64_t *out_xz, double *out_c,
                               int64_t out_capacity, int64_t *n_out, int64_t *n_out_host, void *ws, size_t ws_bytes,
                               void *stream) {
    int64_t U = 0;
    SYM_TRY(sym_mul_cleanup_count(a_xz, a_c, M, b_xz, b_c, N, W, zero_threshold, n_out, &U, ws, ws_bytes, stream));
    if (n_out_host) *n_out_host = U;
    if (U > out_capacity) {
        set_error("output capacity %lld < %lld surviving terms", (long long)out_capacity, (long long)U);
        return SYM_E_CAPACITY;
    }
    return sym_mul_cleanup_emit(a_xz, a_c, M, b_xz, b_c, N, W, U, out_xz, out_c, ws, ws_bytes, stream);
}
