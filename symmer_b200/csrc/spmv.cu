// Matrix-free to_sparse_matrix / expval (symmer/operators/base.py:1458-1510; qiskit to_matrix_sparse):
//   (H psi)[r] = sum_t c_t (-i)^{Y_t} (-1)^{popcount(r & z_t)} psi[r ^ x_t],   qubit 0 = MSB of r.
// column = row XOR x, phase = popcount(row & z). Terms arrive sorted by x mask so that consecutive
// terms with the same x share one gather of psi; one thread per basis row, terms streamed through
// shared memory (broadcast reads). Bound by the integer + FP64 issue rate, not HBM.
#include "common.cuh"

namespace symb {

constexpr int APPLY_THREADS = 256;
constexpr int APPLY_TERMS = 512;  // terms per shared-memory tile

struct TermTile {
    int64_t x[APPLY_TERMS];
    int64_t z[APPLY_TERMS];
    double2 c[APPLY_TERMS];
};

template <bool EXPVAL>
__global__ void __launch_bounds__(APPLY_THREADS) apply_kernel(const int64_t *__restrict__ xm, const int64_t *__restrict__ zm,
                                                               const double2 *__restrict__ cp, int64_t M,
                                                               const double2 *__restrict__ psi, double2 *__restrict__ y,
                                                               int64_t row_begin, int64_t row_end, double *__restrict__ partial) {
    __shared__ TermTile tile;
    __shared__ double red[2][APPLY_THREADS / 32];
    const int64_t r = row_begin + (int64_t)blockIdx.x * APPLY_THREADS + threadIdx.x;
    const bool active = r < row_end;
    double accr = 0.0, acci = 0.0;  // finished groups
    double wr = 0.0, wi = 0.0;      // weight of the current x group
    int64_t xcur = -1;
    for (int64_t base = 0; base < M; base += APPLY_TERMS) {
        const int nt = (int)min((int64_t)APPLY_TERMS, M - base);
        __syncthreads();
        for (int i = threadIdx.x; i < nt; i += APPLY_THREADS) {
            tile.x[i] = xm[base + i];
            tile.z[i] = zm[base + i];
            tile.c[i] = cp[base + i];
        }
        __syncthreads();
        if (active) {
            for (int i = 0; i < nt; ++i) {
                const int64_t x = tile.x[i];
                if (x != xcur) {  // uniform across the CTA
                    if (xcur >= 0) {
                        const double2 p = psi[r ^ xcur];
                        accr += wr * p.x - wi * p.y;
                        acci += wr * p.y + wi * p.x;
                    }
                    xcur = x;
                    wr = 0.0;
                    wi = 0.0;
                }
                const double2 c = tile.c[i];
                const bool neg = __popcll((uint64_t)(r & tile.z[i])) & 1;
                wr += neg ? -c.x : c.x;
                wi += neg ? -c.y : c.y;
            }
        }
    }
    if (active && xcur >= 0) {
        const double2 p = psi[r ^ xcur];
        accr += wr * p.x - wi * p.y;
        acci += wr * p.y + wi * p.x;
    }
    if (!EXPVAL) {
        if (active) y[r - row_begin] = make_double2(accr, acci);
        return;
    }
    // conj(psi[r]) * (H psi)[r]
    double er = 0.0, ei = 0.0;
    if (active) {
        const double2 p = psi[r];
        er = p.x * accr + p.y * acci;
        ei = p.x * acci - p.y * accr;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        er += __shfl_xor_sync(0xffffffffu, er, o);
        ei += __shfl_xor_sync(0xffffffffu, ei, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][wid] = er;
        red[1][wid] = ei;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int w = 0; w < APPLY_THREADS / 32; ++w) {
            sr += red[0][w];
            si += red[1][w];
        }
        atomicAdd(&partial[0], sr);
        atomicAdd(&partial[1], si);
    }
}

// Four basis rows per thread (r0 + k*256, k = 0..3, block base aligned to 1024 so the rows of a thread
// differ only in bits 8 and 9): one AND + POPC per term gives the sign of row 0 and the other three
// follow from bits 8/9 of z; the sign is applied by XOR on the sign bit. Per (row, term) this leaves
// ~2 LOP + 2 DADD (1 DADD when all coefficients are real): the FP64 add rate is the bound.
template <bool EXPVAL, bool REAL>
__global__ void __launch_bounds__(APPLY_THREADS) apply4_kernel(const int64_t *__restrict__ xm, const int64_t *__restrict__ zm,
                                                                const double2 *__restrict__ cp, int64_t M,
                                                                const double2 *__restrict__ psi, double2 *__restrict__ y,
                                                                int64_t row_begin, double *__restrict__ partial) {
    __shared__ TermTile tile;
    __shared__ double red[2][APPLY_THREADS / 32];
    const int64_t r0 = row_begin + (int64_t)blockIdx.x * (4 * APPLY_THREADS) + threadIdx.x;
    double ar[4] = {0, 0, 0, 0}, ai[4] = {0, 0, 0, 0};
    double wr[4] = {0, 0, 0, 0}, wi[4] = {0, 0, 0, 0};
    int64_t xcur = -1;
    for (int64_t base = 0; base < M; base += APPLY_TERMS) {
        const int nt = (int)min((int64_t)APPLY_TERMS, M - base);
        __syncthreads();
        for (int i = threadIdx.x; i < nt; i += APPLY_THREADS) {
            tile.x[i] = xm[base + i];
            tile.z[i] = zm[base + i];
            tile.c[i] = cp[base + i];
        }
        __syncthreads();
        for (int i = 0; i < nt; ++i) {
            const int64_t x = tile.x[i];
            if (x != xcur) {  // uniform across the CTA
                if (xcur >= 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double2 p = psi[(r0 + k * APPLY_THREADS) ^ xcur];
                        if (REAL) {
                            ar[k] += wr[k] * p.x;
                            ai[k] += wr[k] * p.y;
                        } else {
                            ar[k] += wr[k] * p.x - wi[k] * p.y;
                            ai[k] += wr[k] * p.y + wi[k] * p.x;
                        }
                        wr[k] = 0.0;
                        wi[k] = 0.0;
                    }
                }
                xcur = x;
            }
            const uint64_t z = (uint64_t)tile.z[i];
            const double2 c = tile.c[i];
            const uint32_t p0 = __popcll((uint64_t)r0 & z) & 1u;
            const uint32_t z8 = (uint32_t)(z >> 8) & 1u, z9 = (uint32_t)(z >> 9) & 1u;
            const uint32_t par[4] = {p0, p0 ^ z8, p0 ^ z9, p0 ^ z8 ^ z9};
            const int cxh = __double2hiint(c.x), cxl = __double2loint(c.x);
            const int cyh = __double2hiint(c.y), cyl = __double2loint(c.y);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int flip = (int)(par[k] << 31);
                wr[k] += __hiloint2double(cxh ^ flip, cxl);
                if (!REAL) wi[k] += __hiloint2double(cyh ^ flip, cyl);
            }
        }
    }
    if (xcur >= 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double2 p = psi[(r0 + k * APPLY_THREADS) ^ xcur];
            if (REAL) {
                ar[k] += wr[k] * p.x;
                ai[k] += wr[k] * p.y;
            } else {
                ar[k] += wr[k] * p.x - wi[k] * p.y;
                ai[k] += wr[k] * p.y + wi[k] * p.x;
            }
        }
    }
    if (!EXPVAL) {
#pragma unroll
        for (int k = 0; k < 4; ++k) y[r0 + k * APPLY_THREADS - row_begin] = make_double2(ar[k], ai[k]);
        return;
    }
    double er = 0.0, ei = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double2 p = psi[r0 + k * APPLY_THREADS];
        er += p.x * ar[k] + p.y * ai[k];
        ei += p.x * ai[k] - p.y * ar[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        er += __shfl_xor_sync(0xffffffffu, er, o);
        ei += __shfl_xor_sync(0xffffffffu, ei, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][wid] = er;
        red[1][wid] = ei;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int w = 0; w < APPLY_THREADS / 32; ++w) {
            sr += red[0][w];
            si += red[1][w];
        }
        atomicAdd(&partial[0], sr);
        atomicAdd(&partial[1], si);
    }
}

// 2^NB basis rows per thread (r0 + k*TH, k = 0 .. 2^NB-1, TH = block size; the block base is aligned to
// TH*2^NB so the rows of a thread differ only in bits SH .. SH+NB-1, SH = log2 TH). For a term with z mask
// z the sign of row k is
//   s0 * (-1)^{popcount(k & zeta)},   s0 = (-1)^{popcount(r0 & z)},   zeta = (z >> SH) & (2^NB - 1).
// zeta is uniform across the CTA, so a uniform switch selects one of 2^NB straight-line blocks in which
// the sign of every row is a compile-time constant: the block is 2^NB FP64 adds with the negation folded
// into the instruction (DADD with a negated operand) — no per-row sign arithmetic at all. One AND + POPC
// + shared-memory fetch per (thread, term) is spread over 2^NB rows, so the kernel sits on the FP64 add
// rate: 1 DADD per (row, term) for real coefficients, 2 for complex ones. psi gathers are coalesced.
template <int NB>
__device__ __forceinline__ void signed_add(double (&w)[1 << NB], uint32_t zeta, double v) {
    switch (zeta) {
#define SIGN_CASE(Z)                                                       \
    case Z:                                                                \
        if (Z < (1 << NB)) {                                               \
            _Pragma("unroll") for (int k = 0; k < (1 << NB); ++k) {        \
                if (__builtin_popcount(k & Z) & 1) w[k] -= v;              \
                else w[k] += v;                                            \
            }                                                              \
        }                                                                  \
        break;
        SIGN_CASE(0) SIGN_CASE(1) SIGN_CASE(2) SIGN_CASE(3) SIGN_CASE(4) SIGN_CASE(5) SIGN_CASE(6) SIGN_CASE(7)
        SIGN_CASE(8) SIGN_CASE(9) SIGN_CASE(10) SIGN_CASE(11) SIGN_CASE(12) SIGN_CASE(13) SIGN_CASE(14) SIGN_CASE(15)
#undef SIGN_CASE
        default: break;
    }
}

template <bool EXPVAL, bool REAL, int NB, int TH, int MINB, bool SYM, bool NARROW>
__global__ void __launch_bounds__(TH, MINB) applyw_kernel(const int64_t *__restrict__ xm, const int64_t *__restrict__ zm,
                                                                   const double2 *__restrict__ cp, int64_t M,
                                                                   const double2 *__restrict__ psi, double2 *__restrict__ y,
                                                                   int64_t row_begin, double *__restrict__ partial) {
    constexpr int R = 1 << NB;
    constexpr int SH = TH == 128 ? 7 : 8;   // rows of a thread differ in bits SH .. SH+NB-1
    static_assert(TH == 128 || TH == 256, "row stride = block size");
    static_assert(NB >= 1 && NB <= 4, "bins live in registers");
    __shared__ TermTile tile;
    __shared__ double red[2][TH / 32];
    const int64_t r0 = row_begin + (int64_t)blockIdx.x * (R * TH) + threadIdx.x;
    double ar[R], ai[R], Br[R], Bi[R];   // Bi is dead (and removed by the compiler) when REAL
#pragma unroll
    for (int k = 0; k < R; ++k) {
        ar[k] = 0.0;
        ai[k] = 0.0;
        Br[k] = 0.0;
        Bi[k] = 0.0;
    }
    int64_t xcur = -1;
    // SYM (Hermitian operator with real phased coefficients, expectation value only): the pair (r, r^x)
    // contributes a complex-conjugate pair, so a group whose x has its highest set bit h above the
    // CTA's row span is evaluated only by the CTAs whose rows have bit h equal to the group's side bit,
    // with doubled coefficients (sym_expval_prepare_sym stores h+1 in bits 56..62 of z, the group length
    // in bits 40..55, the side in bit 63, and doubles c).
    const uint64_t block_rows = (uint64_t)(row_begin + (int64_t)blockIdx.x * (R * TH));
    bool dirty = false;   // the current group has received at least one term
    int64_t skip_until = 0;   // SYM: global index of the first term after the group being skipped
    auto finish_group = [&]() {   // Br/Bi hold the row weights w_g(r) of the finished x group
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const double2 p = psi[(r0 + (int64_t)k * TH) ^ xcur];
            if constexpr (REAL) {
                ar[k] += Br[k] * p.x;
                ai[k] += Br[k] * p.y;
            } else {
                ar[k] += Br[k] * p.x - Bi[k] * p.y;
                ai[k] += Br[k] * p.y + Bi[k] * p.x;
                Bi[k] = 0.0;
            }
            Br[k] = 0.0;
        }
    };
    for (int64_t base = 0; base < M; base += APPLY_TERMS) {
        const int nt = (int)min((int64_t)APPLY_TERMS, M - base);
        __syncthreads();
        for (int i = threadIdx.x; i < nt; i += TH) {
            tile.x[i] = xm[base + i];
            tile.z[i] = zm[base + i];
            tile.c[i] = cp[base + i];
        }
        __syncthreads();
        int i = 0;
        if constexpr (SYM) {   // a skipped group may continue into this tile
            const int64_t rem = skip_until - base;
            i = rem <= 0 ? 0 : (rem >= nt ? nt : (int)rem);
        }
        for (; i < nt; ++i) {
            const int64_t x = tile.x[i];
            const uint64_t z = (uint64_t)tile.z[i];
            if (x != xcur) {  // uniform across the CTA
                if (dirty) finish_group();
                dirty = false;
                xcur = x;
                if constexpr (SYM) {
                    // first term of a group: bits 56..62 of z hold h+1, bits 40..55 the group length
                    const uint32_t hb = (uint32_t)(z >> 56) & 0x7fu;
                    // bit 63 of z says which half of the pairs (bit h of the row clear or set) evaluates this
                    // group; it alternates pseudo-randomly between groups so CTAs and row shards stay balanced
                    if (hb != 0u && (((block_rows >> (hb - 1u)) & 1ull) != (z >> 63))) {   // the partner rows own this group
                        const int len = (int)((z >> 40) & 0xffffu);
                        skip_until = base + i + len;
                        i += len - 1;
                        xcur = -1;           // re-evaluate at the next term (a capped group continues)
                        continue;
                    }
                }
            }
            const double2 c = tile.c[i];
            const uint32_t par = NARROW ? (uint32_t)__popc((uint32_t)r0 & (uint32_t)z) : (uint32_t)__popcll((uint64_t)r0 & z);
            const int flip = (int)((par & 1u) << 31);
            const uint32_t zeta = (uint32_t)(z >> SH) & (uint32_t)(R - 1);
            dirty = true;
            signed_add<NB>(Br, zeta, __hiloint2double(__double2hiint(c.x) ^ flip, __double2loint(c.x)));
            if constexpr (!REAL)
                signed_add<NB>(Bi, zeta, __hiloint2double(__double2hiint(c.y) ^ flip, __double2loint(c.y)));
        }
    }
    if (dirty) finish_group();
    if (!EXPVAL) {
#pragma unroll
        for (int k = 0; k < R; ++k) y[r0 + (int64_t)k * TH - row_begin] = make_double2(ar[k], ai[k]);
        return;
    }
    double er = 0.0, ei = 0.0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const double2 p = psi[r0 + (int64_t)k * TH];
        er += p.x * ar[k] + p.y * ai[k];
        ei += p.x * ai[k] - p.y * ar[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        er += __shfl_xor_sync(0xffffffffu, er, o);
        ei += __shfl_xor_sync(0xffffffffu, ei, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][wid] = er;
        red[1][wid] = ei;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int w = 0; w < TH / 32; ++w) {
            sr += red[0][w];
            si += red[1][w];
        }
        atomicAdd(&partial[0], sr);
        if (!SYM) atomicAdd(&partial[1], si);   // Hermitian: the conjugate pairs cancel the imaginary part exactly
    }
}

// z' = z | (h+1) << 56 and c' = 2c for the groups whose x has its highest set bit h >= min_bit (the
// symmetric expectation-value mode of applyw_kernel); other terms are copied.
__global__ void __launch_bounds__(256) prepare_sym_kernel(const int64_t *__restrict__ xm, const int64_t *__restrict__ zm,
                                                           const double2 *__restrict__ cp, int64_t M, int min_bit,
                                                           int64_t *__restrict__ z_out, double2 *__restrict__ c_out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    const uint64_t x = (uint64_t)xm[t];
    uint64_t z = (uint64_t)zm[t];
    double2 c = cp[t];
    if (x != 0ull) {
        const int h = 63 - __clzll((long long)x);
        if (h >= min_bit) {
            // remaining terms of the group from here (capped; a longer group is skipped in several hops)
            int64_t len = 1;
            while (t + len < M && len < 0xffff && (uint64_t)xm[t + len] == x) ++len;
            z |= (uint64_t)(h + 1) << 56 | (uint64_t)len << 40 | ((mix64(x) >> 17) & 1ull) << 63;
            c.x *= 2.0;
            c.y *= 2.0;
        }
    }
    z_out[t] = (int64_t)z;
    c_out[t] = c;
}

// CSR emitter for small n: thread per (row, group); value = sum over the group's terms, position =
// rank of the column among the row's columns (counting), so rows come out sorted by column.
__global__ void __launch_bounds__(256) csr_kernel(const int64_t *__restrict__ zm, const double2 *__restrict__ cp, int64_t M,
                                                   int n, const int64_t *__restrict__ xg, int64_t G,
                                                   const int32_t *__restrict__ group_start, double2 *__restrict__ data,
                                                   int64_t *__restrict__ indices, int64_t *__restrict__ indptr) {
    const int64_t side = (int64_t)1 << n;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) indptr[side] = side * G;
    if (idx >= side * G) return;
    const int64_t r = idx / G;
    const int64_t g = idx - r * G;
    const int64_t col = r ^ xg[g];
    int64_t rank = 0;
    for (int64_t h = 0; h < G; ++h) rank += ((r ^ xg[h]) < col) ? 1 : 0;
    double sr = 0.0, si = 0.0;
    for (int t = group_start[g]; t < group_start[g + 1]; ++t) {
        const double2 c = cp[t];
        const bool neg = __popcll((uint64_t)(r & zm[t])) & 1;
        sr += neg ? -c.x : c.x;
        si += neg ? -c.y : c.y;
    }
    data[r * G + rank] = make_double2(sr, si);
    indices[r * G + rank] = col;
    if (g == 0) indptr[r] = r * G;
}

int g_apply_variant = 1;  // tuning knob 4: 1 = 16/8-row static-sign kernel (default), 0 = 4-row kernel

}  // namespace symb

using namespace symb;

static int apply_common(const int64_t *x_masks, const int64_t *z_masks, const double *c_phased, int64_t M, int32_t n,
                        const double *psi, double *y, double *partial, int64_t row_begin, int64_t row_end, bool expval,
                        bool real_coeffs, cudaStream_t st, bool sym_mode = false) {
    SYM_REQUIRE(n >= 1 && n <= 40, "n_qubits out of range for the dense-state path");
    SYM_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= ((int64_t)1 << n), "bad row range");
    const int64_t rows = row_end - row_begin;
    if (rows == 0) return SYM_OK;
    const double2 *c2 = reinterpret_cast<const double2 *>(c_phased);
    const double2 *p2 = reinterpret_cast<const double2 *>(psi);
    double2 *y2 = reinterpret_cast<double2 *>(y);
    // static-sign kernel: 16 rows per thread for real coefficients, 8 for complex ones (register budget)
    const int64_t span_w = real_coeffs ? 16 * 128 : 8 * 256;
    const bool narrow = n <= 32;   // basis indices fit 32 bits: one POPC per sign instead of two
    if (sym_mode) {
        SYM_REQUIRE(expval && real_coeffs, "symmetric mode is for expectation values with real phased coefficients");
        SYM_REQUIRE(rows % span_w == 0 && row_begin % span_w == 0, "symmetric mode needs row ranges aligned to 2048");
        const unsigned nbs = (unsigned)(rows / span_w);
        if (narrow) applyw_kernel<true, true, 4, 128, 3, true, true><<<nbs, 128, 0, st>>>(x_masks, z_masks, c2, M, p2, nullptr, row_begin, partial);
        else applyw_kernel<true, true, 4, 128, 3, true, false><<<nbs, 128, 0, st>>>(x_masks, z_masks, c2, M, p2, nullptr, row_begin, partial);
        SYM_LAUNCH_OK();
        return SYM_OK;
    }
    if (g_apply_variant == 1 && rows % span_w == 0 && row_begin % span_w == 0) {
        const unsigned nbw = (unsigned)(rows / span_w);
#define APPLYW(E, RL, NBITS, THR, MB, OUT, PART)                                                                         \
    if (narrow) applyw_kernel<E, RL, NBITS, THR, MB, false, true><<<nbw, THR, 0, st>>>(x_masks, z_masks, c2, M, p2, OUT, row_begin, PART); \
    else applyw_kernel<E, RL, NBITS, THR, MB, false, false><<<nbw, THR, 0, st>>>(x_masks, z_masks, c2, M, p2, OUT, row_begin, PART)
        if (expval) {
            if (real_coeffs) { APPLYW(true, true, 4, 128, 3, nullptr, partial); }
            else { APPLYW(true, false, 3, 256, 2, nullptr, partial); }
        } else {
            if (real_coeffs) { APPLYW(false, true, 4, 128, 3, y2, nullptr); }
            else { APPLYW(false, false, 3, 256, 2, y2, nullptr); }
        }
#undef APPLYW
        SYM_LAUNCH_OK();
        return SYM_OK;
    }
    if (rows % (4 * APPLY_THREADS) == 0 && row_begin % (4 * APPLY_THREADS) == 0) {
        const unsigned nb4 = (unsigned)(rows / (4 * APPLY_THREADS));
        if (expval) {
            if (real_coeffs) apply4_kernel<true, true><<<nb4, APPLY_THREADS, 0, st>>>(x_masks, z_masks, c2, M, p2, nullptr, row_begin, partial);
            else apply4_kernel<true, false><<<nb4, APPLY_THREADS, 0, st>>>(x_masks, z_masks, c2, M, p2, nullptr, row_begin, partial);
        } else {
            if (real_coeffs) apply4_kernel<false, true><<<nb4, APPLY_THREADS, 0, st>>>(x_masks, z_masks, c2, M, p2, y2, row_begin, nullptr);
            else apply4_kernel<false, false><<<nb4, APPLY_THREADS, 0, st>>>(x_masks, z_masks, c2, M, p2, y2, row_begin, nullptr);
        }
        SYM_LAUNCH_OK();
        return SYM_OK;
    }
    const unsigned nb = (unsigned)((rows + APPLY_THREADS - 1) / APPLY_THREADS);
    if (expval)
        apply_kernel<true><<<nb, APPLY_THREADS, 0, st>>>(x_masks, z_masks, c2, M, p2, nullptr, row_begin, row_end, partial);
    else
        apply_kernel<false><<<nb, APPLY_THREADS, 0, st>>>(x_masks, z_masks, c2, M, p2, y2, row_begin, row_end, nullptr);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_apply(const int64_t *x_masks, const int64_t *z_masks, const double *c_phased, int64_t M,
                         int32_t n_qubits, const double *psi, double *y, int64_t row_begin, int64_t row_end,
                         int32_t real_coeffs, void *stream) {
    return apply_common(x_masks, z_masks, c_phased, M, n_qubits, psi, y, nullptr, row_begin, row_end, false,
                        real_coeffs != 0, (cudaStream_t)stream);
}

extern "C" int sym_expval(const int64_t *x_masks, const int64_t *z_masks, const double *c_phased, int64_t M,
                          int32_t n_qubits, const double *psi, double *partial, int64_t row_begin, int64_t row_end,
                          int32_t real_coeffs, void *stream) {
    return apply_common(x_masks, z_masks, c_phased, M, n_qubits, psi, nullptr, partial, row_begin, row_end, true,
                        real_coeffs != 0, (cudaStream_t)stream, real_coeffs == 2);
}

extern "C" int sym_expval_prepare_sym(const int64_t *x_masks, const int64_t *z_masks, const double *c_phased, int64_t M,
                                      int64_t *z_sym, double *c_sym, void *stream) {
    SYM_REQUIRE(M >= 0, "bad size");
    if (M == 0) return SYM_OK;
    prepare_sym_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x_masks, z_masks, reinterpret_cast<const double2 *>(c_phased), M, 11, z_sym, reinterpret_cast<double2 *>(c_sym));
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_to_csr(const int64_t *z_masks, const double *c_phased, int64_t M, int32_t n_qubits,
                          const int64_t *x_groups, int64_t G, const int32_t *group_start, double *data, int64_t *indices,
                          int64_t *indptr, void *stream) {
    SYM_REQUIRE(n_qubits >= 1 && n_qubits <= 24, "CSR emitter is for small n_qubits");
    SYM_REQUIRE(G >= 1 && M >= 1, "empty operator");
    const int64_t total = ((int64_t)1 << n_qubits) * G;
    csr_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        z_masks, reinterpret_cast<const double2 *>(c_phased), M, n_qubits, x_groups, G, group_start,
        reinterpret_cast<double2 *>(data), indices, indptr);
    SYM_LAUNCH_OK();
    return SYM_OK;
}
