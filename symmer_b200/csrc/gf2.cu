// _rref_binary (symmer/operators/utils.py:292-315) on bit-packed rows: GF(2) row reduction with the
// reference's row-driven pivot rule — for each row i in order, the pivot is the first set column of
// the CURRENT row i, and row i is XORed into every other row with a 1 in that column; rows are
// never swapped. Results are bit-exact with the reference.
//
// Small matrices (the symmetry-generator search: 2n rows x (M+2n) columns) run in ONE CTA with the
// whole matrix in shared memory, one warp-parallel step per pivot row. Large matrices run one
// pivot step per pair of launches over the whole GPU (HBM-bound XOR sweep).
#include "common.cuh"

namespace symb {

constexpr int RREF_THREADS = 1024;

__device__ __forceinline__ int first_set_column(const uint64_t *row, int64_t Cw, int tid, int nthreads, int *s_min) {
    // block-wide: smallest set bit index of `row` (or INT_MAX). s_min must be initialised to INT_MAX.
    int best = 0x7fffffff;
    for (int64_t k = tid; k < Cw; k += nthreads) {
        uint64_t w = row[k];
        if (w) {
            int col = (int)(k * 64 + (__ffsll((long long)w) - 1));
            best = min(best, col);
            break;  // columns grow with k for this thread
        }
    }
    if (best != 0x7fffffff) atomicMin(s_min, best);
    return 0;
}

__global__ void __launch_bounds__(RREF_THREADS) rref_smem_kernel(uint64_t *__restrict__ bits, int R, int64_t Cw,
                                                                  int32_t *__restrict__ pivots) {
    extern __shared__ uint64_t sm[];  // R*Cw words
    __shared__ int s_piv;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t total = (int64_t)R * Cw;
    for (int64_t i = tid; i < total; i += nt) sm[i] = bits[i];
    __syncthreads();
    for (int i = 0; i < R; ++i) {
        if (tid == 0) s_piv = 0x7fffffff;
        __syncthreads();
        first_set_column(sm + (int64_t)i * Cw, Cw, tid, nt, &s_piv);
        __syncthreads();
        const int piv = s_piv;
        if (piv != 0x7fffffff) {
            const int64_t pw = piv >> 6;
            const uint64_t pm = 1ull << (piv & 63);
            // each thread owns whole (row, word) cells; the pivot-column test reads word pw of row r,
            // which is itself updated in this step, so handle it last: first all other words.
            for (int64_t c = tid; c < total; c += nt) {
                const int r = (int)(c / Cw);
                const int64_t k = c - (int64_t)r * Cw;
                if (r == i || k == pw) continue;
                if (sm[(int64_t)r * Cw + pw] & pm) sm[c] ^= sm[(int64_t)i * Cw + k];
            }
            __syncthreads();
            for (int r = tid; r < R; r += nt) {
                if (r == i) continue;
                uint64_t w = sm[(int64_t)r * Cw + pw];
                if (w & pm) sm[(int64_t)r * Cw + pw] = w ^ sm[(int64_t)i * Cw + pw];
            }
        }
        __syncthreads();
    }
    for (int64_t i = tid; i < total; i += nt) bits[i] = sm[i];
    // final pivot column of every row
    for (int r = tid; r < R; r += nt) {
        int p = -1;
        for (int64_t k = 0; k < Cw; ++k) {
            uint64_t w = sm[(int64_t)r * Cw + k];
            if (w) {
                p = (int)(k * 64 + (__ffsll((long long)w) - 1));
                break;
            }
        }
        pivots[r] = p;
    }
}

// ---- large path: one pivot step = (pivot + hit flags) kernel, then a grid-wide XOR sweep
__global__ void __launch_bounds__(RREF_THREADS) rref_pivot_kernel(const uint64_t *__restrict__ bits, int64_t R, int64_t Cw,
                                                                   int64_t i, int *__restrict__ piv_out) {
    __shared__ int s_piv;
    if (threadIdx.x == 0) s_piv = 0x7fffffff;
    __syncthreads();
    first_set_column(bits + i * Cw, Cw, threadIdx.x, blockDim.x, &s_piv);
    __syncthreads();
    if (threadIdx.x == 0) *piv_out = s_piv;
}

__global__ void __launch_bounds__(256) rref_hit_kernel(const uint64_t *__restrict__ bits, int64_t R, int64_t Cw, int64_t i,
                                                        const int *__restrict__ piv_in, uint8_t *__restrict__ hit) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int piv = *piv_in;
    uint8_t h = 0;
    if (piv != 0x7fffffff && r != i) h = (bits[r * Cw + (piv >> 6)] >> (piv & 63)) & 1ull;
    hit[r] = h;
}

__global__ void __launch_bounds__(256) rref_xor_kernel(uint64_t *__restrict__ bits, int64_t R, int64_t Cw, int64_t i,
                                                        const uint8_t *__restrict__ hit) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= R * Cw) return;
    int64_t r = c / Cw, k = c - r * Cw;
    if (hit[r]) bits[c] ^= bits[i * Cw + k];
}

__global__ void __launch_bounds__(256) rref_final_pivots_kernel(const uint64_t *__restrict__ bits, int64_t R, int64_t Cw,
                                                                 int32_t *__restrict__ pivots) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int p = -1;
    for (int64_t k = 0; k < Cw; ++k) {
        uint64_t w = bits[r * Cw + k];
        if (w) {
            p = (int)(k * 64 + (__ffsll((long long)w) - 1));
            break;
        }
    }
    pivots[r] = p;
}

}  // namespace symb

using namespace symb;

extern "C" size_t sym_rref_ws_bytes(int64_t R) { return arena_need((size_t)(R > 0 ? R : 1), 1) + 1024; }

extern "C" int sym_rref(uint64_t *bits, int64_t R, int64_t C, int64_t Cw, int32_t *pivots, void *ws, size_t ws_bytes,
                        void *stream) {
    SYM_REQUIRE(R >= 0 && C >= 0 && Cw * 64 >= C, "bad matrix shape");
    SYM_REQUIRE(C < ((int64_t)1 << 31) && R < ((int64_t)1 << 31), "matrix too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (R == 0) return SYM_OK;
    if (Cw == 0) {
        SYM_CUDA_OK(cudaMemsetAsync(pivots, 0xff, sizeof(int32_t) * (size_t)R, st));
        return SYM_OK;
    }
    const size_t smem = (size_t)R * (size_t)Cw * 8;
    if (smem <= 200 * 1024) {
        static bool attr_set = false;
        if (!attr_set) {
            SYM_CUDA_OK(cudaFuncSetAttribute(rref_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
        rref_smem_kernel<<<1, RREF_THREADS, smem, st>>>(bits, (int)R, Cw, pivots);
        SYM_LAUNCH_OK();
        return SYM_OK;
    }
    if (ws_bytes < sym_rref_ws_bytes(R)) {
        set_error("workspace too small");
        return SYM_E_WORKSPACE;
    }
    Arena ar(ws, ws_bytes);
    uint8_t *hit = ar.take<uint8_t>((size_t)R);
    int *piv = reinterpret_cast<int *>(ar.take<int>(4));
    const int64_t cells = R * Cw;
    for (int64_t i = 0; i < R; ++i) {
        rref_pivot_kernel<<<1, RREF_THREADS, 0, st>>>(bits, R, Cw, i, piv);
        SYM_LAUNCH_OK();
        rref_hit_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(bits, R, Cw, i, piv, hit);
        SYM_LAUNCH_OK();
        rref_xor_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(bits, R, Cw, i, hit);
        SYM_LAUNCH_OK();
    }
    rref_final_pivots_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(bits, R, Cw, pivots);
    SYM_LAUNCH_OK();
    return SYM_OK;
}
