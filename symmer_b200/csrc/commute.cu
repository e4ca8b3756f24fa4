// commutes_termwise / adjacency_matrix (symmer/operators/base.py:938-971, utils.py:9-78):
// symplectic inner product (A.x . B.z^T + A.z . B.x^T) mod 2 on bit-packed rows, True = commute.
// Bit-packed AND/XOR accumulate + one popcount parity per pair. Thread = one B row in registers
// (consecutive threads -> consecutive output columns, coalesced byte stores); the CTA streams a
// chunk of A rows through shared memory (broadcast reads).
#include "common.cuh"

namespace symb {

constexpr int COM_THREADS = 256;
constexpr int COM_ICH = 64;

template <int WT, bool BITS>
__global__ void __launch_bounds__(COM_THREADS) commute_kernel(const uint64_t *__restrict__ a_xz, uint32_t M,
                                                               const uint64_t *__restrict__ b_xz, uint32_t N, int W,
                                                               uint8_t *__restrict__ out, uint32_t *__restrict__ out_bits,
                                                               uint32_t bit_stride) {
    // A tile with (z_w, x_w) swapped per word so that acc ^= (xb & sa.x) ^ (zb & sa.y)
    __shared__ ulonglong2 sa[COM_ICH][WT];
    const uint32_t i0 = blockIdx.y * COM_ICH;
    const uint32_t ni = min((uint32_t)COM_ICH, M - i0);
    for (int i = threadIdx.x; i < COM_ICH * WT; i += COM_THREADS) {
        int ii = i / WT, w = i % WT;
        ulonglong2 v = make_ulonglong2(0ull, 0ull);
        if ((uint32_t)ii < ni && w < W) {
            const uint64_t *row = a_xz + (size_t)(i0 + ii) * 2 * W;
            v.x = row[W + w];  // z of A pairs with x of B
            v.y = row[w];      // x of A pairs with z of B
        }
        sa[ii][w] = v;
    }
    const uint32_t j = blockIdx.x * COM_THREADS + threadIdx.x;
    const bool active = j < N;
    uint64_t xb[WT], zb[WT];
#pragma unroll
    for (int w = 0; w < WT; ++w) {
        xb[w] = (active && w < W) ? b_xz[(size_t)j * 2 * W + w] : 0ull;
        zb[w] = (active && w < W) ? b_xz[(size_t)j * 2 * W + W + w] : 0ull;
    }
    __syncthreads();
    for (uint32_t ii = 0; ii < ni; ++ii) {
        uint64_t acc = 0;
#pragma unroll
        for (int w = 0; w < WT; ++w) {
            const ulonglong2 a = sa[ii][w];
            acc ^= (xb[w] & a.x) ^ (zb[w] & a.y);
        }
        const uint32_t commute = ((__popcll(acc) & 1) == 0) ? 1u : 0u;
        if (BITS) {
            uint32_t word = __ballot_sync(0xffffffffu, active && commute);
            if ((threadIdx.x & 31) == 0 && active) out_bits[(size_t)(i0 + ii) * bit_stride + (j >> 5)] = word;
        } else if (active) {
            out[(size_t)(i0 + ii) * N + j] = (uint8_t)commute;
        }
    }
}

template <bool BITS>
__global__ void __launch_bounds__(256) commute_generic_kernel(const uint64_t *__restrict__ a_xz, uint32_t M,
                                                               const uint64_t *__restrict__ b_xz, uint32_t N, int W,
                                                               uint8_t *__restrict__ out, uint32_t *__restrict__ out_bits,
                                                               uint32_t bit_stride) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = blockIdx.y;
    const bool active = j < N;
    uint64_t acc = 0;
    if (active) {
        const uint64_t *ra = a_xz + (size_t)i * 2 * W, *rb = b_xz + (size_t)j * 2 * W;
        for (int w = 0; w < W; ++w) acc ^= (ra[w] & rb[W + w]) ^ (ra[W + w] & rb[w]);
    }
    const uint32_t commute = ((__popcll(acc) & 1) == 0) ? 1u : 0u;
    if (BITS) {
        uint32_t word = __ballot_sync(0xffffffffu, active && commute);
        if ((threadIdx.x & 31) == 0 && active) out_bits[(size_t)i * bit_stride + (j >> 5)] = word;
    } else if (active) {
        out[(size_t)i * N + j] = (uint8_t)commute;
    }
}

template <bool BITS>
static int commute_launch(const uint64_t *a_xz, int64_t M, const uint64_t *b_xz, int64_t N, int W, uint8_t *out,
                          uint32_t *out_bits, cudaStream_t st) {
    if (M == 0 || N == 0) return SYM_OK;
    const uint32_t stride = (uint32_t)((N + 31) / 32);
    if (W <= 16) {
        int64_t done = 0;
        // grid.y is limited to 65535 chunks of COM_ICH rows: loop over slabs of A if needed
        while (done < M) {
            int64_t rows = M - done;
            const int64_t max_rows = (int64_t)65535 * COM_ICH;
            if (rows > max_rows) rows = max_rows;
            dim3 grid((unsigned)((N + COM_THREADS - 1) / COM_THREADS), (unsigned)((rows + COM_ICH - 1) / COM_ICH));
            const uint64_t *a = a_xz + (size_t)done * 2 * W;
            uint8_t *o = BITS ? nullptr : out + (size_t)done * N;
            uint32_t *ob = BITS ? out_bits + (size_t)done * stride : nullptr;
#define COM_CASE(WT) commute_kernel<WT, BITS><<<grid, COM_THREADS, 0, st>>>(a, (uint32_t)rows, b_xz, (uint32_t)N, W, o, ob, stride)
            if (W <= 1) COM_CASE(1);
            else if (W <= 2) COM_CASE(2);
            else if (W <= 4) COM_CASE(4);
            else if (W <= 8) COM_CASE(8);
            else COM_CASE(16);
#undef COM_CASE
            SYM_LAUNCH_OK();
            done += rows;
        }
    } else {
        int64_t done = 0;
        while (done < M) {
            int64_t rows = M - done;
            if (rows > 65535) rows = 65535;
            dim3 grid((unsigned)((N + 255) / 256), (unsigned)rows);
            commute_generic_kernel<BITS><<<grid, 256, 0, st>>>(a_xz + (size_t)done * 2 * W, (uint32_t)rows, b_xz, (uint32_t)N,
                                                             W, BITS ? nullptr : out + (size_t)done * N,
                                                             BITS ? out_bits + (size_t)done * stride : nullptr, stride);
            SYM_LAUNCH_OK();
            done += rows;
        }
    }
    return SYM_OK;
}


// qubitwise_commutes_termwise (base.py:985-1009): A[i] and B[j] commute qubit by qubit iff on every qubit
// where both act non-trivially they carry the same Pauli, i.e. no bit survives
//   (xa | za) & (xb | zb) & ((xa ^ xb) | (za ^ zb)).
// Same decomposition as commute_kernel: thread = one B row in registers, CTA = 64 A rows in shared memory.
template <int WT>
__global__ void __launch_bounds__(COM_THREADS) qwc_kernel(const uint64_t *__restrict__ a_xz, uint32_t M,
                                                           const uint64_t *__restrict__ b_xz, uint32_t N, int W,
                                                           uint8_t *__restrict__ out) {
    __shared__ ulonglong2 sa[COM_ICH][WT];   // (x_w, z_w) of the A rows
    const uint32_t i0 = blockIdx.y * COM_ICH;
    const uint32_t ni = min((uint32_t)COM_ICH, M - i0);
    for (int i = threadIdx.x; i < COM_ICH * WT; i += COM_THREADS) {
        int ii = i / WT, w = i % WT;
        ulonglong2 v = make_ulonglong2(0ull, 0ull);
        if ((uint32_t)ii < ni && w < W) {
            const uint64_t *row = a_xz + (size_t)(i0 + ii) * 2 * W;
            v.x = row[w];
            v.y = row[W + w];
        }
        sa[ii][w] = v;
    }
    const uint32_t j = blockIdx.x * COM_THREADS + threadIdx.x;
    const bool active = j < N;
    uint64_t xb[WT], zb[WT];
#pragma unroll
    for (int w = 0; w < WT; ++w) {
        xb[w] = (active && w < W) ? b_xz[(size_t)j * 2 * W + w] : 0ull;
        zb[w] = (active && w < W) ? b_xz[(size_t)j * 2 * W + W + w] : 0ull;
    }
    __syncthreads();
    for (uint32_t ii = 0; ii < ni; ++ii) {
        uint64_t clash = 0;
#pragma unroll
        for (int w = 0; w < WT; ++w) {
            const ulonglong2 a = sa[ii][w];
            clash |= (a.x | a.y) & (xb[w] | zb[w]) & ((a.x ^ xb[w]) | (a.y ^ zb[w]));
        }
        if (active) out[(size_t)(i0 + ii) * N + j] = clash == 0 ? (uint8_t)1 : (uint8_t)0;
    }
}

__global__ void __launch_bounds__(256) qwc_generic_kernel(const uint64_t *__restrict__ a_xz, uint32_t M,
                                                           const uint64_t *__restrict__ b_xz, uint32_t N, int W,
                                                           uint8_t *__restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = blockIdx.y;
    if (j >= N) return;
    const uint64_t *ra = a_xz + (size_t)i * 2 * W, *rb = b_xz + (size_t)j * 2 * W;
    uint64_t clash = 0;
    for (int w = 0; w < W; ++w) {
        const uint64_t xa = ra[w], za = ra[W + w], xb = rb[w], zb = rb[W + w];
        clash |= (xa | za) & (xb | zb) & ((xa ^ xb) | (za ^ zb));
    }
    out[(size_t)i * N + j] = clash == 0 ? (uint8_t)1 : (uint8_t)0;
}

static int qwc_launch(const uint64_t *a_xz, int64_t M, const uint64_t *b_xz, int64_t N, int W, uint8_t *out,
                      cudaStream_t st) {
    if (M == 0 || N == 0) return SYM_OK;
    int64_t done = 0;
    while (done < M) {
        int64_t rows = M - done;
        const uint64_t *a = a_xz + (size_t)done * 2 * W;
        uint8_t *o = out + (size_t)done * N;
        if (W <= 16) {
            const int64_t max_rows = (int64_t)65535 * COM_ICH;   // grid.y limit
            if (rows > max_rows) rows = max_rows;
            dim3 grid((unsigned)((N + COM_THREADS - 1) / COM_THREADS), (unsigned)((rows + COM_ICH - 1) / COM_ICH));
#define QWC_CASE(WT) qwc_kernel<WT><<<grid, COM_THREADS, 0, st>>>(a, (uint32_t)rows, b_xz, (uint32_t)N, W, o)
            if (W <= 1) QWC_CASE(1);
            else if (W <= 2) QWC_CASE(2);
            else if (W <= 4) QWC_CASE(4);
            else if (W <= 8) QWC_CASE(8);
            else QWC_CASE(16);
#undef QWC_CASE
        } else {
            if (rows > 65535) rows = 65535;
            dim3 grid((unsigned)((N + 255) / 256), (unsigned)rows);
            qwc_generic_kernel<<<grid, 256, 0, st>>>(a, (uint32_t)rows, b_xz, (uint32_t)N, W, o);
        }
        SYM_LAUNCH_OK();
        done += rows;
    }
    return SYM_OK;
}

// Symmetric matrices (adjacency of an operator with itself, base.py:1054-1062): only the blocks on and above the
// block diagonal are computed; this kernel mirrors them into the lower triangle. 32 x 32 byte tiles through shared
// memory, so both the reads and the writes are 32-byte row segments. HBM-bound: reads and writes M^2/2 bytes.
__global__ void __launch_bounds__(256) mirror_upper_kernel(uint8_t *__restrict__ m, uint32_t M, size_t pitch, uint32_t blk) {
    __shared__ uint8_t tile[32][33];
    const uint32_t tj = blockIdx.x, ti = blockIdx.y;   // tile (ti, tj) of the upper triangle -> tile (tj, ti)
    // everything right of the block diagonal is mirrored; inside a diagonal block only the strict upper tiles are
    const uint32_t bi = (ti * 32) / blk, bj = (tj * 32) / blk;
    if (bj < bi || (bj == bi && tj <= ti)) return;
    const uint32_t x = threadIdx.x & 31, y0 = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t y = y0 + 8 * r;
        const uint32_t row = ti * 32 + y, col = tj * 32 + x;
        tile[y][x] = (row < M && col < M) ? m[(size_t)row * pitch + col] : (uint8_t)0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t y = y0 + 8 * r;
        const uint32_t row = tj * 32 + y, col = ti * 32 + x;
        if (row < M && col < M) m[(size_t)row * pitch + col] = tile[x][y];
    }
}

}  // namespace symb

using namespace symb;

extern "C" int sym_mirror_upper(uint8_t *matrix, int64_t M, int64_t pitch, int64_t block_rows, void *stream) {
    SYM_REQUIRE(M >= 0 && pitch >= M && block_rows >= 32 && block_rows % 32 == 0, "bad size (block_rows must be a multiple of 32)");
    SYM_REQUIRE(M < ((int64_t)1 << 21), "matrix too large for one mirror launch");
    if (M == 0) return SYM_OK;
    const unsigned nt = (unsigned)((M + 31) / 32);
    mirror_upper_kernel<<<dim3(nt, nt), 256, 0, (cudaStream_t)stream>>>(matrix, (uint32_t)M, (size_t)pitch, (uint32_t)block_rows);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" int sym_commute(const uint64_t *a_xz, int64_t M, const uint64_t *b_xz, int64_t N, int32_t W, uint8_t *out,
                           void *stream) {
    SYM_REQUIRE(M >= 0 && N >= 0 && W >= 1, "bad size");
    SYM_REQUIRE(M < ((int64_t)1 << 31) && N < ((int64_t)1 << 31), "operand too large");
    return commute_launch<false>(a_xz, M, b_xz, N, W, out, nullptr, (cudaStream_t)stream);
}

extern "C" int sym_commute_bits(const uint64_t *a_xz, int64_t M, const uint64_t *b_xz, int64_t N, int32_t W,
                                uint32_t *out_bits, void *stream) {
    SYM_REQUIRE(M >= 0 && N >= 0 && W >= 1, "bad size");
    SYM_REQUIRE(M < ((int64_t)1 << 31) && N < ((int64_t)1 << 31), "operand too large");
    return commute_launch<true>(a_xz, M, b_xz, N, W, nullptr, out_bits, (cudaStream_t)stream);
}

extern "C" int sym_commute_qwc(const uint64_t *a_xz, int64_t M, const uint64_t *b_xz, int64_t N, int32_t W, uint8_t *out,
                               void *stream) {
    SYM_REQUIRE(M >= 0 && N >= 0 && W >= 1, "bad size");
    SYM_REQUIRE(M < ((int64_t)1 << 31) && N < ((int64_t)1 << 31), "operand too large");
    return qwc_launch(a_xz, M, b_xz, N, W, out, (cudaStream_t)stream);
}
