// Layout kernels: bool <-> bit-packed rows, Y counts, GF(2)-linear row sketches, basis-index masks.
// Also owns the library-wide globals (error string, launch counter, debug key mask).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace symb {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
uint64_t g_key_mask = ~0ull;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// one warp per (row, block, word): two coalesced 32-byte reads + ballots
__global__ void pack_kernel(const uint8_t *__restrict__ symp, int64_t M, int n, int W, uint64_t *__restrict__ xz) {
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t total = M * 2 * W;
    if (warp >= total) return;
    int64_t row = warp / (2 * W);
    int rem = (int)(warp % (2 * W));
    int blk = rem / W, word = rem % W;
    const uint8_t *src = symp + row * (2 * (int64_t)n) + (int64_t)blk * n;
    int q0 = word * 64 + lane, q1 = q0 + 32;
    uint32_t b0 = (q0 < n) ? (src[q0] != 0) : 0u;
    uint32_t b1 = (q1 < n) ? (src[q1] != 0) : 0u;
    uint32_t lo = __ballot_sync(0xffffffffu, b0);
    uint32_t hi = __ballot_sync(0xffffffffu, b1);
    if (lane == 0) xz[warp] = ((uint64_t)hi << 32) | lo;
}

__global__ void unpack_kernel(const uint64_t *__restrict__ xz, int64_t M, int n, int W, uint8_t *__restrict__ symp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = M * 2 * (int64_t)n;
    if (i >= total) return;
    int64_t row = i / (2 * n);
    int col = (int)(i % (2 * n));
    int blk = col / n, q = col % n;
    uint64_t w = xz[row * 2 * W + (int64_t)blk * W + (q >> 6)];
    symp[i] = (uint8_t)((w >> (q & 63)) & 1ull);
}

// one warp per row
__global__ void ycount_kernel(const uint64_t *__restrict__ xz, int64_t M, int W, int32_t *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    const uint64_t *r = xz + row * 2 * W;
    int cnt = 0;
    for (int k = lane; k < W; k += 32) cnt += __popcll(r[k] & r[W + k]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    if (lane == 0) y[row] = cnt;
}

__global__ void sketch8_kernel(const uint64_t *__restrict__ xz, int64_t M, int W, uint64_t *__restrict__ sk) {
    const int lane = threadIdx.x & 31;
    int64_t row = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 2) + (lane >> 3);
    const bool ok = row < M;
    uint64_t h = group8_sketch_row(xz + (ok ? row : 0) * 2 * W, 2 * W, lane & 7);
    if (ok && (lane & 7) == 0) sk[row] = h;
}

// Sketch AND Y count of every row of two operands in one launch (what a product needs of A and B before it can
// generate records): 8 lanes per row, lane g holds X chunk g and Z chunk g. same != 0: B is A (a square): the A
// rows are read once and both tables written. A small product is a chain of short kernels, and four of them
// (sketch, Y count, twice) were these.
__global__ void __launch_bounds__(256) operand_tables8_kernel(const uint64_t *__restrict__ a, int64_t M, const uint64_t *__restrict__ b,
                                                              int64_t N, int W, uint64_t *__restrict__ a_sk, int32_t *__restrict__ a_y,
                                                              uint64_t *__restrict__ b_sk, int32_t *__restrict__ b_y, int same) {
    const int lane = threadIdx.x & 31, g = lane & 7;
    const int64_t r = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 2) + (lane >> 3);
    const int64_t total = same ? M : M + N;
    const bool ok = r < total;
    const bool in_a = r < M;
    const uint64_t *row = !ok ? a : (in_a ? a + r * 2 * W : b + (r - M) * 2 * W);
    const uint4 *r4 = reinterpret_cast<const uint4 *>(row);
    const int chunks = W >> 1;   // 16-byte chunks per block (X or Z)
    uint64_t h = 0;
    int y = 0;
    if (g < chunks) {
        const uint4 xv = r4[g], zv = r4[chunks + g];
        const uint64_t x0 = ((uint64_t)xv.y << 32) | xv.x, x1 = ((uint64_t)xv.w << 32) | xv.z;
        const uint64_t z0 = ((uint64_t)zv.y << 32) | zv.x, z1 = ((uint64_t)zv.w << 32) | zv.z;
        y = __popcll(x0 & z0) + __popcll(x1 & z1);
        // word index = lane index of the sketch: X words 2g, 2g+1; Z words 2(chunks+g), 2(chunks+g)+1
        h = lane_linear(x0, 2 * g) ^ lane_linear(x1, 2 * g + 1) ^ lane_linear(z0, 2 * (chunks + g)) ^
            lane_linear(z1, 2 * (chunks + g) + 1);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        h ^= __shfl_xor_sync(0xffffffffu, h, o);
        y += __shfl_xor_sync(0xffffffffu, y, o);
    }
    if (ok && g == 0) {
        if (in_a) {
            a_sk[r] = h;
            a_y[r] = y;
            if (same) {
                b_sk[r] = h;
                b_y[r] = y;
            }
        } else {
            b_sk[r - M] = h;
            b_y[r - M] = y;
        }
    }
}

__global__ void sketch_kernel(const uint64_t *__restrict__ xz, int64_t M, int W, uint64_t *__restrict__ sk) {
    const int lane = threadIdx.x & 31;
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    uint64_t h = warp_sketch_row(xz + row * 2 * W, 2 * W, lane);
    if (lane == 0) sk[row] = h;
}

// out[i] = words[perm[i] * pitch + column] (perm == nullptr: identity): one word column of a row-major matrix in a given
// row order — the key of one pass of the word-by-word lexicographic sort (base.py:469-470)
__global__ void gather_column_kernel(const uint64_t *__restrict__ words, int64_t M, int64_t pitch, int64_t column,
                                     const uint32_t *__restrict__ perm, uint64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t row = perm ? (int64_t)perm[i] : i;
    out[i] = words[row * pitch + column];
}

// Exact join of two row sets on equal rows (bra * ket, base.py:1781-1830: the reference joins two Python dicts of bit
// strings): match[i] = index of the row of the right set equal to left row i, or -1. The right keys (64-bit sketches)
// are sorted; a left row binary-searches its key and compares the packed rows word by word over the WHOLE run of
// equal keys, so sketch collisions between different rows cannot hide a match.
__global__ void join_rows_kernel(const uint64_t *__restrict__ keys_l, const uint64_t *__restrict__ rows_l, int64_t M,
                                 const uint64_t *__restrict__ keys_r, const uint32_t *__restrict__ perm_r,
                                 const uint64_t *__restrict__ rows_r, int64_t N, int words, int32_t *__restrict__ match) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint64_t k = keys_l[i];
    int64_t lo = 0, hi = N;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys_r[mid] < k) lo = mid + 1;
        else hi = mid;
    }
    int32_t found = -1;
    for (int64_t j = lo; j < N && keys_r[j] == k && found < 0; ++j) {
        const uint32_t r = perm_r[j];
        bool eq = true;
        for (int w = 0; w < words && eq; ++w) eq = rows_l[i * words + w] == rows_r[(int64_t)r * words + w];
        if (eq) found = (int32_t)r;
    }
    match[i] = found;
}

// basis-index masks (qubit 0 = most significant bit), n <= 62, and coefficient * (-i)^Y
__global__ void term_masks_kernel(const uint64_t *__restrict__ xz, const double *__restrict__ c, int64_t M, int n,
                                  int64_t *__restrict__ xm, int64_t *__restrict__ zm, double *__restrict__ cp) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    uint64_t x = xz[2 * t], z = xz[2 * t + 1];  // W == 1
    // packed bit q <-> basis bit (n-1-q): reverse the low n bits
    uint64_t xr = __brevll(x) >> (64 - n), zr = __brevll(z) >> (64 - n);
    xm[t] = (int64_t)xr;
    zm[t] = (int64_t)zr;
    double re = c[2 * t], im = c[2 * t + 1];
    int k = __popcll(x & z) & 3;
    mul_i_pow(re, im, (4 - k) & 3);  // (-i)^k = i^(-k)
    cp[2 * t] = re;
    cp[2 * t + 1] = im;
}

// generic GF(2) matrix packing (any column count): one warp per output word
__global__ void pack_matrix_kernel(const uint8_t *__restrict__ m, int64_t R, int64_t C, uint64_t *__restrict__ bits,
                                   int64_t Cw) {
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= R * Cw) return;
    int64_t row = warp / Cw, word = warp % Cw;
    const uint8_t *src = m + row * C;
    int64_t q0 = word * 64 + lane, q1 = q0 + 32;
    uint32_t b0 = (q0 < C) ? (src[q0] != 0) : 0u;
    uint32_t b1 = (q1 < C) ? (src[q1] != 0) : 0u;
    uint32_t lo = __ballot_sync(0xffffffffu, b0);
    uint32_t hi = __ballot_sync(0xffffffffu, b1);
    if (lane == 0) bits[warp] = ((uint64_t)hi << 32) | lo;
}

__global__ void unpack_matrix_kernel(const uint64_t *__restrict__ bits, int64_t R, int64_t C, int64_t Cw,
                                     uint8_t *__restrict__ m) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * C) return;
    int64_t row = i / C, col = i % C;
    m[i] = (uint8_t)((bits[row * Cw + (col >> 6)] >> (col & 63)) & 1ull);
}

}  // namespace symb

using namespace symb;

// Qubit gather: bit k of the output X (Z) block = bit src[k] of the input X (Z) block, 0 where src[k] < 0.
// Serves PauliwordOp.reindex (base.py:493-521: a permutation) and PauliwordOp.tensor (base.py:1188-1204:
// each factor embedded into the wider register before the product). thread = (row, output word).
__global__ void __launch_bounds__(256) gather_qubits_kernel(const uint64_t *__restrict__ xz, int64_t M, int W_in,
                                                             const int32_t *__restrict__ src, int n_out, int W_out,
                                                             uint64_t *__restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = g / (2 * W_out);
    if (row >= M) return;
    const int w = (int)(g - row * (2 * W_out));
    const int blk = w / W_out, wo = w - blk * W_out;   // blk 0 = X, 1 = Z
    const uint64_t *in = xz + (size_t)row * 2 * W_in + (size_t)blk * W_in;
    uint64_t v = 0;
    const int k0 = wo * 64;
    const int k1 = min(k0 + 64, n_out);
    for (int k = k0; k < k1; ++k) {
        const int q = src[k];
        if (q >= 0) v |= ((in[q >> 6] >> (q & 63)) & 1ull) << (k - k0);
    }
    out[g] = v;
}

static inline unsigned blocks_for(int64_t threads, int per_block) {
    int64_t b = (threads + per_block - 1) / per_block;
    return (unsigned)(b < 1 ? 1 : b);
}

extern "C" int sym_abi_version(void) { return SYM_ABI_VERSION; }
extern "C" const char *sym_last_error(void) { return symb::g_err; }
extern "C" int64_t sym_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" int sym_debug_set_key_mask(uint64_t mask) {
    g_key_mask = mask;
    return SYM_OK;
}

extern "C" int sym_check_device(int *sm_count_host, int *cc_major_host, int *cc_minor_host) {
    int dev = 0, major = 0, minor = 0, sms = 0;
    SYM_CUDA_OK(cudaGetDevice(&dev));
    SYM_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    SYM_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    SYM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sm_count_host) *sm_count_host = sms;
    if (cc_major_host) *cc_major_host = major;
    if (cc_minor_host) *cc_minor_host = minor;
    if (major != 10) {
        set_error("symmer_b200 is built for sm_100a only; device is sm_%d%d", major, minor);
        return SYM_E_UNSUPPORTED;
    }
    return SYM_OK;
}

extern "C" int sym_pack(const uint8_t *symp, int64_t M, int32_t n, uint64_t *xz, void *stream) {
    SYM_REQUIRE(M >= 0 && n >= 0, "negative size");
    int W = n > 0 ? (n + 63) / 64 : 1;
    if (M == 0) return SYM_OK;
    if (n == 0) {
        SYM_CUDA_OK(cudaMemsetAsync(xz, 0, sizeof(uint64_t) * 2 * (size_t)M, (cudaStream_t)stream));
        return SYM_OK;
    }
    int64_t warps = M * 2 * W;
    pack_kernel<<<blocks_for(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(symp, M, n, W, xz);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_unpack(const uint64_t *xz, int64_t M, int32_t n, uint8_t *symp, void *stream) {
    SYM_REQUIRE(M >= 0 && n >= 0, "negative size");
    if (M == 0 || n == 0) return SYM_OK;
    int W = (n + 63) / 64;
    unpack_kernel<<<blocks_for(M * 2 * (int64_t)n, 256), 256, 0, (cudaStream_t)stream>>>(xz, M, n, W, symp);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_ycount(const uint64_t *xz, int64_t M, int32_t W, int32_t *y, void *stream) {
    SYM_REQUIRE(M >= 0 && W >= 1, "bad size");
    if (M == 0) return SYM_OK;
    ycount_kernel<<<blocks_for(M * 32, 256), 256, 0, (cudaStream_t)stream>>>(xz, M, W, y);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_sketch_rows(const uint64_t *xz, int64_t M, int32_t W, uint64_t *sketch, void *stream) {
    SYM_REQUIRE(M >= 0 && W >= 1, "bad size");
    if (M == 0) return SYM_OK;
    if (group8_ok(W))
        sketch8_kernel<<<blocks_for(((M + 3) / 4) * 32, 256), 256, 0, (cudaStream_t)stream>>>(xz, M, W, sketch);
    else
        sketch_kernel<<<blocks_for(M * 32, 256), 256, 0, (cudaStream_t)stream>>>(xz, M, W, sketch);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

// sketches and Y counts of both operands of a product (internal; declared in rows.cuh)
int symb::operand_tables(const uint64_t *a_xz, int64_t M, const uint64_t *b_xz, int64_t N, int W, uint64_t *a_sk, int32_t *a_y,
                         uint64_t *b_sk, int32_t *b_y, cudaStream_t st) {
    if (group8_ok(W) && M > 0 && N > 0) {
        const int same = (a_xz == b_xz && M == N) ? 1 : 0;
        const int64_t total = same ? M : M + N;
        symb::operand_tables8_kernel<<<blocks_for(((total + 3) / 4) * 32, 256), 256, 0, st>>>(a_xz, M, b_xz, N, W, a_sk, a_y, b_sk,
                                                                                             b_y, same);
        SYM_LAUNCH_OK();
        return SYM_OK;
    }
    SYM_TRY(sym_sketch_rows(a_xz, M, W, a_sk, st));
    SYM_TRY(sym_sketch_rows(b_xz, N, W, b_sk, st));
    SYM_TRY(sym_ycount(a_xz, M, W, a_y, st));
    SYM_TRY(sym_ycount(b_xz, N, W, b_y, st));
    return SYM_OK;
}

extern "C" int sym_term_masks(const uint64_t *xz, const double *c, int64_t M, int32_t n, int64_t *x_masks,
                              int64_t *z_masks, double *c_phased, void *stream) {
    SYM_REQUIRE(M >= 0 && n >= 1 && n <= 62, "term masks need 1 <= n_qubits <= 62");
    if (M == 0) return SYM_OK;
    term_masks_kernel<<<blocks_for(M, 256), 256, 0, (cudaStream_t)stream>>>(xz, c, M, n, x_masks, z_masks, c_phased);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_pack_matrix(const uint8_t *m, int64_t R, int64_t C, uint64_t *bits, int64_t Cw, void *stream) {
    SYM_REQUIRE(R >= 0 && C >= 0 && Cw * 64 >= C, "bad matrix shape");
    if (R == 0 || Cw == 0) return SYM_OK;
    pack_matrix_kernel<<<blocks_for(R * Cw * 32, 256), 256, 0, (cudaStream_t)stream>>>(m, R, C, bits, Cw);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_unpack_matrix(const uint64_t *bits, int64_t R, int64_t C, int64_t Cw, uint8_t *m, void *stream) {
    SYM_REQUIRE(R >= 0 && C >= 0 && Cw * 64 >= C, "bad matrix shape");
    if (R == 0 || C == 0) return SYM_OK;
    unpack_matrix_kernel<<<blocks_for(R * C, 256), 256, 0, (cudaStream_t)stream>>>(bits, R, C, Cw, m);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_gather_column(const uint64_t *words, int64_t M, int64_t pitch_words, int64_t column, const uint32_t *perm,
                                 uint64_t *out, void *stream) {
    SYM_REQUIRE(M >= 0 && pitch_words >= 1 && column >= 0 && column < pitch_words, "bad size");
    if (M == 0) return SYM_OK;
    gather_column_kernel<<<blocks_for(M, 256), 256, 0, (cudaStream_t)stream>>>(words, M, pitch_words, column, perm, out);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_join_rows(const uint64_t *keys_l, const uint64_t *rows_l, int64_t M, const uint64_t *keys_r_sorted,
                             const uint32_t *perm_r, const uint64_t *rows_r, int64_t N, int32_t words, int32_t *match,
                             void *stream) {
    SYM_REQUIRE(M >= 0 && N >= 0 && words >= 1, "bad size");
    if (M == 0) return SYM_OK;
    join_rows_kernel<<<blocks_for(M, 256), 256, 0, (cudaStream_t)stream>>>(keys_l, rows_l, M, keys_r_sorted, perm_r, rows_r, N,
                                                                          words, match);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_gather_qubits(const uint64_t *xz, int64_t M, int32_t W_in, const int32_t *src, int32_t n_out,
                                 uint64_t *out_xz, void *stream) {
    SYM_REQUIRE(M >= 0 && W_in >= 1 && n_out >= 0, "bad size");
    if (M == 0) return SYM_OK;
    const int W_out = n_out > 0 ? (n_out + 63) / 64 : 1;
    gather_qubits_kernel<<<blocks_for(M * 2 * (int64_t)W_out, 256), 256, 0, (cudaStream_t)stream>>>(xz, M, W_in, src, n_out,
                                                                                                    W_out, out_xz);
    SYM_LAUNCH_OK();
    return SYM_OK;
}
