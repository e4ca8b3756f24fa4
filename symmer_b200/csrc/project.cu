// Stabilizer-subspace projection of a rotated operator (symmer/projection/base.py:44-84,
// S3Projection._perform_projection) on packed rows:
//   1. drop every term that anticommutes with one of the single-qubit stabilizers,
//   2. multiply the coefficient by the eigenvalue of every stabilizer whose Pauli appears in the term,
//   3. delete the stabilized qubit positions (bit gather of the free qubits into a narrower row),
// followed by a stream compaction of the surviving rows. The duplicate merge that the reference does
// next (`.cleanup()`) is sym_cleanup. Byte/bit work, HBM-bound: reads 16W+16 B per row, writes
// 16W'+16 B per survivor.
#include "rows.cuh"
#include "sort.cuh"

namespace symb {

// thread = (row, output word). Output word w < Wo is X-block word w, w >= Wo is Z-block word w - Wo.
// The thread of output word 0 also evaluates the commutation test and the eigenvalue product.
__global__ void __launch_bounds__(256) project_rows_kernel(const uint64_t *__restrict__ xz, const double2 *__restrict__ c,
                                                            int64_t M, int W, int n, const int32_t *__restrict__ stab_cols,
                                                            const double *__restrict__ stab_eigs, int S,
                                                            const int32_t *__restrict__ free_q, int n_free, int Wo,
                                                            uint64_t *__restrict__ tmp_xz, double2 *__restrict__ tmp_c,
                                                            uint8_t *__restrict__ keep) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = g / (2 * Wo);
    if (row >= M) return;
    const int w = (int)(g - row * (2 * Wo));
    const uint64_t *r = xz + row * 2 * W;
    const int blk = w / Wo, wo = w - blk * Wo;     // blk 0 = X, 1 = Z
    const uint64_t *src = r + (size_t)blk * W;
    uint64_t out = 0;
    const int k0 = wo * 64;
    const int k1 = min(k0 + 64, n_free);
    for (int k = k0; k < k1; ++k) {
        const int q = free_q[k];
        out |= ((src[q >> 6] >> (q & 63)) & 1ull) << (k - k0);
    }
    tmp_xz[row * 2 * Wo + w] = out;
    if (w == 0) {
        bool commutes = true;
        double f = 1.0;
        for (int j = 0; j < S; ++j) {
            const int col = stab_cols[j];              // column of the stabilizer's single set bit in [X | Z]
            const int q = col < n ? col : col - n;
            const uint64_t xb = (r[q >> 6] >> (q & 63)) & 1ull;
            const uint64_t zb = (r[W + (q >> 6)] >> (q & 63)) & 1ull;
            // X_q anticommutes with a term carrying Z or Y on q; Z_q with one carrying X or Y
            if (col < n ? zb : xb) commutes = false;
            if (col < n ? xb : zb) f *= stab_eigs[j];  // the stabilizer's Pauli appears in the term
        }
        keep[row] = commutes ? 1 : 0;
        const double2 cc = c[row];
        tmp_c[row] = make_double2(cc.x * f, cc.y * f);
    }
}

__global__ void __launch_bounds__(256) project_compact_kernel(const uint64_t *__restrict__ tmp_xz, const double2 *__restrict__ tmp_c,
                                                               const uint8_t *__restrict__ keep, const uint32_t *__restrict__ slot,
                                                               int64_t M, int words, uint64_t *__restrict__ out_xz,
                                                               double2 *__restrict__ out_c) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = g / words;
    if (row >= M || !keep[row]) return;
    const int w = (int)(g - row * words);
    const uint32_t s = slot[row];
    out_xz[(size_t)s * words + w] = tmp_xz[g];
    if (w == 0) out_c[s] = tmp_c[row];
}

__global__ void project_total_kernel(const uint32_t *__restrict__ total, int64_t *__restrict__ n_out) { *n_out = (int64_t)*total; }

}  // namespace symb

using namespace symb;

static inline int words_out(int n_free) { return n_free > 0 ? (n_free + 63) / 64 : 1; }

extern "C" size_t sym_project_ws_bytes(int64_t M, int32_t n_free) {
    if (M < 1) M = 1;
    const size_t Wo = (size_t)words_out(n_free);
    return arena_need((size_t)M * 2 * Wo, 8) + arena_need((size_t)M, 16) + arena_need((size_t)M, 1) + arena_need((size_t)M, 4) +
           arena_need(scan_scratch_elems(M), 4) + arena_need(4, 4) + 2048;
}

extern "C" int sym_project(const uint64_t *xz, const double *c, int64_t M, int32_t W, int32_t n_qubits,
                           const int32_t *stab_cols, const double *stab_eigs, int32_t S, const int32_t *free_qubits,
                           int32_t n_free, uint64_t *out_xz, double *out_c, int64_t *n_out, int64_t *n_out_host, void *ws,
                           size_t ws_bytes, void *stream) {
    SYM_REQUIRE(M >= 0 && W >= 1 && n_qubits >= 0 && n_qubits <= 64 * W, "bad size");
    SYM_REQUIRE(S >= 0 && n_free >= 0 && n_free <= n_qubits, "bad stabilizer / free-qubit counts");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        if (n_out) SYM_CUDA_OK(cudaMemsetAsync(n_out, 0, sizeof(int64_t), st));
        if (n_out_host) *n_out_host = 0;
        return SYM_OK;
    }
    if (ws_bytes < sym_project_ws_bytes(M, n_free)) {
        set_error("workspace too small");
        return SYM_E_WORKSPACE;
    }
    const int Wo = words_out(n_free);
    Arena ar(ws, ws_bytes);
    uint64_t *tmp_xz = ar.take<uint64_t>((size_t)M * 2 * Wo);
    double2 *tmp_c = ar.take<double2>((size_t)M);
    uint8_t *keep = ar.take<uint8_t>((size_t)M);
    uint32_t *slot = ar.take<uint32_t>((size_t)M);
    uint32_t *scratch = ar.take<uint32_t>(scan_scratch_elems(M));
    uint32_t *total = ar.take<uint32_t>(4);
    if (!total) {
        set_error("workspace arena exhausted");
        return SYM_E_WORKSPACE;
    }
    const int64_t threads = M * 2 * Wo;
    project_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        xz, reinterpret_cast<const double2 *>(c), M, W, n_qubits, stab_cols, stab_eigs, S, free_qubits, n_free, Wo, tmp_xz, tmp_c,
        keep);
    SYM_LAUNCH_OK();
    SYM_TRY(scan_exclusive_u8(keep, slot, M, total, scratch, st));
    project_compact_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(tmp_xz, tmp_c, keep, slot, M, 2 * Wo, out_xz,
                                                                             reinterpret_cast<double2 *>(out_c));
    SYM_LAUNCH_OK();
    if (n_out) {
        project_total_kernel<<<1, 1, 0, st>>>(total, n_out);
        SYM_LAUNCH_OK();
    }
    if (n_out_host) {
        uint32_t u = 0;
        SYM_CUDA_OK(cudaMemcpyAsync(&u, total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SYM_CUDA_OK(cudaStreamSynchronize(st));
        *n_out_host = (int64_t)u;
    }
    return SYM_OK;
}
