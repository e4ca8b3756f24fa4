// PauliwordOp._rotate_by_single_Pword (symmer/operators/base.py:1090-1161) as a fused
// commute-test + branch-free product + append of the new terms.
//   R P R^dagger, R = exp(i*angle/2*Q):   [P,Q]=0 -> P ;  {P,Q}=0 -> cos(angle) P + sin(angle) (-i P Q)
// Pass 1 (one warp per row): commutation parity with Q and the phase exponent of P*Q, one byte/row.
// Pass 2 (one thread per row word): copy / XOR rows to their output slots, coalesced.
// HBM-bound: reads M rows once (plus the byte array), writes M (+ M_ac) rows.
#include "common.cuh"
#include "sort.cuh"

namespace symb {

// info byte: bit0 = anticommutes with Q, bits1-2 = phase exponent e of P*Q (coefficient factor i^e)
__global__ void __launch_bounds__(256) rotate_info_kernel(const uint64_t *__restrict__ xz, int64_t M, int W,
                                                           const uint64_t *__restrict__ q_xz, uint8_t *__restrict__ info) {
    const int lane = threadIdx.x & 31;
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    const uint64_t *r = xz + row * 2 * W;
    uint64_t comm = 0, s = 0;
    int ya = 0, yb = 0, yout = 0;
    for (int w = lane; w < W; w += 32) {
        uint64_t xa = r[w], za = r[W + w], xb = q_xz[w], zb = q_xz[W + w];
        comm ^= (xa & zb) ^ (za & xb);
        s ^= xa & zb;
        ya += __popcll(xa & za);
        yb += __popcll(xb & zb);
        yout += __popcll((xa ^ xb) & (za ^ zb));
    }
    int par = __popcll(comm) & 1, sg = __popcll(s) & 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        par ^= __shfl_xor_sync(0xffffffffu, par, o);
        sg ^= __shfl_xor_sync(0xffffffffu, sg, o);
        ya += __shfl_xor_sync(0xffffffffu, ya, o);
        yb += __shfl_xor_sync(0xffffffffu, yb, o);
        yout += __shfl_xor_sync(0xffffffffu, yout, o);
    }
    if (lane == 0) {
        int e = (3 * (ya + yb) + yout + 2 * sg) & 3;
        info[row] = (uint8_t)(par | (e << 1));
    }
}

__global__ void __launch_bounds__(256) rotate_anti_flag_kernel(const uint8_t *__restrict__ info, int64_t M,
                                                                uint8_t *__restrict__ anti) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) anti[i] = info[i] & 1;
}

// mode 0: general; mode 1: Clifford odd; mode 2: Clifford even
__global__ void __launch_bounds__(256) rotate_write_kernel(const uint64_t *__restrict__ xz, const double2 *__restrict__ c,
                                                            int64_t M, int words, const uint64_t *__restrict__ q_xz,
                                                            const uint8_t *__restrict__ info, const uint32_t *__restrict__ rank,
                                                            double cos_a, double sin_a, int mode, double sign,
                                                            uint64_t *__restrict__ out_xz, double2 *__restrict__ out_c) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t row = g / words;
    if (row >= M) return;
    const int k = (int)(g - row * words);
    const uint8_t inf = info[row];
    const bool anti = inf & 1;
    const int e = (inf >> 1) & 3;
    const uint64_t w = xz[g];
    if (mode == 0) {
        out_xz[g] = w;
        if (anti) out_xz[(M + rank[row]) * (int64_t)words + k] = w ^ q_xz[k];
        if (k == 0) {
            double2 cc = c[row];
            if (anti) {
                out_c[row] = make_double2(cc.x * cos_a, cc.y * cos_a);
                // (P*Q coefficient) * (-i sin): i^e then multiply by -i*sin
                double re = cc.x, im = cc.y;
                mul_i_pow(re, im, e);
                out_c[M + rank[row]] = make_double2(im * sin_a, -re * sin_a);
            } else {
                out_c[row] = cc;
            }
        }
    } else if (mode == 1) {
        out_xz[g] = anti ? (w ^ q_xz[k]) : w;
        if (k == 0) {
            double2 cc = c[row];
            if (anti) {
                double re = cc.x, im = cc.y;
                mul_i_pow(re, im, e + 3);  // times i^e, then times -i = i^3
                out_c[row] = make_double2(re * sign, im * sign);
            } else {
                out_c[row] = cc;
            }
        }
    } else {
        out_xz[g] = w;
        if (k == 0) {
            double2 cc = c[row];
            out_c[row] = anti ? make_double2(cc.x * sign, cc.y * sign) : cc;
        }
    }
}

__global__ void rotate_count_kernel(const uint32_t *__restrict__ total, int64_t M, int mode, int64_t *__restrict__ n_out) {
    *n_out = (mode == 0) ? M + (int64_t)*total : M;
}

}  // namespace symb

using namespace symb;

extern "C" size_t sym_rotate_ws_bytes(int64_t M) {
    if (M < 1) M = 1;
    return arena_need((size_t)M, 1) * 2 + arena_need((size_t)M, 4) + arena_need(scan_scratch_elems(M), 4) + 4096;
}

extern "C" int sym_rotate(const uint64_t *xz, const double *c, int64_t M, int32_t W, const uint64_t *q_xz, double cos_a,
                          double sin_a, int32_t mode, double sign, uint64_t *out_xz, double *out_c, int64_t *n_out,
                          void *ws, size_t ws_bytes, void *stream) {
    SYM_REQUIRE(M >= 0 && W >= 1, "bad size");
    SYM_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        SYM_CUDA_OK(cudaMemsetAsync(n_out, 0, sizeof(int64_t), st));
        return SYM_OK;
    }
    if (ws_bytes < sym_rotate_ws_bytes(M)) {
        set_error("workspace too small");
        return SYM_E_WORKSPACE;
    }
    Arena ar(ws, ws_bytes);
    uint8_t *info = ar.take<uint8_t>((size_t)M);
    uint8_t *anti = ar.take<uint8_t>((size_t)M);
    uint32_t *rank = ar.take<uint32_t>((size_t)M);
    uint32_t *scratch = ar.take<uint32_t>(scan_scratch_elems(M));
    uint32_t *total = ar.take<uint32_t>(4);
    rotate_info_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, st>>>(xz, M, W, q_xz, info);
    SYM_LAUNCH_OK();
    if (mode == 0) {
        rotate_anti_flag_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(info, M, anti);
        SYM_LAUNCH_OK();
        SYM_TRY(scan_exclusive_u8(anti, rank, M, total, scratch, st));
    } else {
        SYM_CUDA_OK(cudaMemsetAsync(total, 0, sizeof(uint32_t), st));
    }
    int64_t threads = M * 2 * W;
    rotate_write_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        xz, reinterpret_cast<const double2 *>(c), M, 2 * W, q_xz, info, rank, cos_a, sin_a, mode, sign, out_xz,
        reinterpret_cast<double2 *>(out_c));
    SYM_LAUNCH_OK();
    rotate_count_kernel<<<1, 1, 0, st>>>(total, M, mode, n_out);
    SYM_LAUNCH_OK();
    return SYM_OK;
}
