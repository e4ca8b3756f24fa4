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
                                                           const uint64_t *__restrict__ q_xz, uint8_t *__restrict__ info,
                                                           uint64_t *__restrict__ sk_out, int32_t *__restrict__ y_out,
                                                           uint8_t *__restrict__ anti_out) {
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
        if (anti_out) anti_out[row] = (uint8_t)par;   // 0/1 flags for the rank scan
        if (y_out) y_out[row] = ya;
    }
    if (sk_out) {   // generic widths: a second pass over the row (uniform branch, whole warp)
        const uint64_t h = warp_sketch_row(r, 2 * W, lane);
        if (lane == 0) sk_out[row] = h;
    }
}

// fast path (W even, W <= 16): 8 lanes per row, 16-byte loads, 4 rows per warp
// Optionally also the row sketch (same function as sketch8_kernel) and the Y count of every row, so that
// a product that follows (the fused rotation) need not read the rows again for its tables.
__global__ void __launch_bounds__(256) rotate_info8_kernel(const uint64_t *__restrict__ xz, int64_t M, int W,
                                                            const uint64_t *__restrict__ q_xz, uint8_t *__restrict__ info,
                                                            uint64_t *__restrict__ sk_out, int32_t *__restrict__ y_out,
                                                            uint8_t *__restrict__ anti_out) {
    const int lane = threadIdx.x & 31, g = lane & 7;
    int64_t row = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 2) + (lane >> 3);
    const bool ok = row < M;
    const uint4 *r4 = reinterpret_cast<const uint4 *>(xz + (ok ? row : 0) * 2 * W);
    const uint4 *q4 = reinterpret_cast<const uint4 *>(q_xz);
    const int chunks = W >> 1;   // 16-byte chunks per block
    uint64_t comm = 0, s = 0, h = 0;
    int ya = 0, yb = 0, yout = 0;
#pragma unroll
    for (int rep = 0; rep < 1; ++rep) {
        const int c = g;
        if (c < chunks) {
            const uint4 xv = r4[c], zv = r4[chunks + c], qx = q4[c], qz = q4[chunks + c];
            const uint64_t xa[2] = {((uint64_t)xv.y << 32) | xv.x, ((uint64_t)xv.w << 32) | xv.z};
            const uint64_t za[2] = {((uint64_t)zv.y << 32) | zv.x, ((uint64_t)zv.w << 32) | zv.z};
            const uint64_t xb[2] = {((uint64_t)qx.y << 32) | qx.x, ((uint64_t)qx.w << 32) | qx.z};
            const uint64_t zb[2] = {((uint64_t)qz.y << 32) | qz.x, ((uint64_t)qz.w << 32) | qz.z};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                comm ^= (xa[k] & zb[k]) ^ (za[k] & xb[k]);
                s ^= xa[k] & zb[k];
                ya += __popcll(xa[k] & za[k]);
                yb += __popcll(xb[k] & zb[k]);
                yout += __popcll((xa[k] ^ xb[k]) & (za[k] ^ zb[k]));
                // row words 2c+k (X block) and 2(chunks+c)+k (Z block): the lane index of the sketch is the word index
                h ^= lane_linear(xa[k], 2 * c + k) ^ lane_linear(za[k], 2 * (chunks + c) + k);
            }
        }
    }
    int par = __popcll(comm) & 1, sg = __popcll(s) & 1;
    if (sk_out) {
        h ^= __shfl_xor_sync(0xffffffffu, h, 1);
        h ^= __shfl_xor_sync(0xffffffffu, h, 2);
        h ^= __shfl_xor_sync(0xffffffffu, h, 4);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        par ^= __shfl_xor_sync(0xffffffffu, par, o);
        sg ^= __shfl_xor_sync(0xffffffffu, sg, o);
        ya += __shfl_xor_sync(0xffffffffu, ya, o);
        yb += __shfl_xor_sync(0xffffffffu, yb, o);
        yout += __shfl_xor_sync(0xffffffffu, yout, o);
    }
    if (ok && g == 0) {
        int e = (3 * (ya + yb) + yout + 2 * sg) & 3;
        info[row] = (uint8_t)(par | (e << 1));
        if (anti_out) anti_out[row] = (uint8_t)par;   // 0/1 flags for the rank scan
        if (sk_out) sk_out[row] = h;
        if (y_out) y_out[row] = ya;
    }
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

// 16-byte form of rotate_write_kernel (words even): LC = log2(chunks per row) or -1 for generic
template <int LC>
__global__ void __launch_bounds__(256) rotate_write4_kernel(const uint4 *__restrict__ xz, const double2 *__restrict__ c,
                                                             int64_t M, uint32_t chunks, const uint4 *__restrict__ q_xz,
                                                             const uint8_t *__restrict__ info, const uint32_t *__restrict__ rank,
                                                             double cos_a, double sin_a, int mode, double sign,
                                                             uint4 *__restrict__ out_xz, double2 *__restrict__ out_c) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t row;
    uint32_t k;
    if (LC >= 0) {
        row = g >> (LC >= 0 ? LC : 0);
        k = (uint32_t)g & ((1u << (LC >= 0 ? LC : 0)) - 1u);
    } else {
        row = g / chunks;
        k = (uint32_t)(g - row * chunks);
    }
    if (row >= (size_t)M) return;
    const uint8_t inf = info[row];
    const bool anti = inf & 1;
    const int e = (inf >> 1) & 3;
    const uint4 w = xz[g];
    const uint4 q = q_xz[k];
    const uint4 wq = make_uint4(w.x ^ q.x, w.y ^ q.y, w.z ^ q.z, w.w ^ q.w);
    if (mode == 0) {
        out_xz[g] = w;
        if (anti) out_xz[((size_t)M + rank[row]) * chunks + k] = wq;
        if (k == 0) {
            double2 cc = c[row];
            if (anti) {
                out_c[row] = make_double2(cc.x * cos_a, cc.y * cos_a);
                double re = cc.x, im = cc.y;
                mul_i_pow(re, im, e);
                out_c[M + rank[row]] = make_double2(im * sin_a, -re * sin_a);
            } else {
                out_c[row] = cc;
            }
        }
    } else if (mode == 1) {
        out_xz[g] = anti ? wq : w;
        if (k == 0) {
            double2 cc = c[row];
            if (anti) {
                double re = cc.x, im = cc.y;
                mul_i_pow(re, im, e + 3);
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

// One-pass rotation for the 8-lane row form (W even, W <= 16): the commutation test, the phase exponent and the
// write of the rotated rows in ONE kernel, no flag scan, no count to read back.
//   MODE 1 / 2: the Clifford relabellings of rotate_write_kernel.
//   MODE 4: general angle, PADDED: out has exactly 2M rows. Row p keeps P_p (cos*c if it anticommutes with Q);
//           row M + p holds the second term of an anticommuting row (P_p Q, -i sin * c * i^e) and, for a commuting
//           row, a copy of P_p with coefficient 0. The cleanup that always follows a general rotation merges the
//           copy into its twin (c + 0 = c; the twin comes first, so first-occurrence order is that of the compact
//           form) -- for operators of config size the extra rows cost less than the scan of the anticommuting
//           flags plus the host round trip for the row count that the compact form needs.
template <int MODE>
__global__ void __launch_bounds__(256) rotate_fused8_kernel(const uint64_t *__restrict__ xz, const double2 *__restrict__ c, int64_t M,
                                                            int W, const uint64_t *__restrict__ q_xz, double cos_a, double sin_a,
                                                            double sign, uint4 *__restrict__ out_xz, double2 *__restrict__ out_c,
                                                            int64_t *__restrict__ n_out) {
    const int lane = threadIdx.x & 31, g = lane & 7;
    const int64_t row = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 2) + (lane >> 3);
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = MODE == 4 ? 2 * M : M;
    const bool ok = row < M;
    const uint4 *r4 = reinterpret_cast<const uint4 *>(xz + (ok ? row : 0) * 2 * W);
    const uint4 *q4 = reinterpret_cast<const uint4 *>(q_xz);
    const int chunks = W >> 1;   // 16-byte chunks per block (X or Z)
    const bool mine = g < chunks;
    uint4 xv = make_uint4(0u, 0u, 0u, 0u), zv = xv, qx = xv, qz = xv;
    if (mine) {
        xv = r4[g];
        zv = r4[chunks + g];
        qx = q4[g];
        qz = q4[chunks + g];
    }
    const uint64_t xa[2] = {((uint64_t)xv.y << 32) | xv.x, ((uint64_t)xv.w << 32) | xv.z};
    const uint64_t za[2] = {((uint64_t)zv.y << 32) | zv.x, ((uint64_t)zv.w << 32) | zv.z};
    const uint64_t xb[2] = {((uint64_t)qx.y << 32) | qx.x, ((uint64_t)qx.w << 32) | qx.z};
    const uint64_t zb[2] = {((uint64_t)qz.y << 32) | qz.x, ((uint64_t)qz.w << 32) | qz.z};
    uint64_t comm = 0, s = 0;
    int ya = 0, yb = 0, yout = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        comm ^= (xa[k] & zb[k]) ^ (za[k] & xb[k]);
        s ^= xa[k] & zb[k];
        ya += __popcll(xa[k] & za[k]);
        yb += __popcll(xb[k] & zb[k]);
        yout += __popcll((xa[k] ^ xb[k]) & (za[k] ^ zb[k]));
    }
    int par = __popcll(comm) & 1, sg = __popcll(s) & 1;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        par ^= __shfl_xor_sync(0xffffffffu, par, o);
        sg ^= __shfl_xor_sync(0xffffffffu, sg, o);
        ya += __shfl_xor_sync(0xffffffffu, ya, o);
        yb += __shfl_xor_sync(0xffffffffu, yb, o);
        yout += __shfl_xor_sync(0xffffffffu, yout, o);
    }
    if (!ok) return;
    const bool anti = par != 0;
    const int e = (3 * (ya + yb) + yout + 2 * sg) & 3;
    const uint4 xq = make_uint4(xv.x ^ qx.x, xv.y ^ qx.y, xv.z ^ qx.z, xv.w ^ qx.w);
    const uint4 zq = make_uint4(zv.x ^ qz.x, zv.y ^ qz.y, zv.z ^ qz.z, zv.w ^ qz.w);
    uint4 *o1 = out_xz + (size_t)row * W;
    if (MODE == 4) {
        uint4 *o2 = out_xz + (size_t)(M + row) * W;
        if (mine) {
            o1[g] = xv;
            o1[chunks + g] = zv;
            o2[g] = anti ? xq : xv;
            o2[chunks + g] = anti ? zq : zv;
        }
        if (g == 0) {
            const double2 cc = c[row];
            if (anti) {
                out_c[row] = make_double2(cc.x * cos_a, cc.y * cos_a);
                double re = cc.x, im = cc.y;
                mul_i_pow(re, im, e);   // (P*Q coefficient) * (-i sin): i^e, then times -i*sin
                out_c[M + row] = make_double2(im * sin_a, -re * sin_a);
            } else {
                out_c[row] = cc;
                out_c[M + row] = make_double2(0.0, 0.0);
            }
        }
    } else if (MODE == 1) {
        if (mine) {
            o1[g] = anti ? xq : xv;
            o1[chunks + g] = anti ? zq : zv;
        }
        if (g == 0) {
            const double2 cc = c[row];
            if (anti) {
                double re = cc.x, im = cc.y;
                mul_i_pow(re, im, e + 3);   // times i^e, then times -i = i^3
                out_c[row] = make_double2(re * sign, im * sign);
            } else {
                out_c[row] = cc;
            }
        }
    } else {
        if (mine) {
            o1[g] = xv;
            o1[chunks + g] = zv;
        }
        if (g == 0) {
            const double2 cc = c[row];
            out_c[row] = anti ? make_double2(cc.x * sign, cc.y * sign) : cc;
        }
    }
}

// The same one-pass rotation for any word count, one thread per row (operators of <= 64 qubits have W = 1: a row is
// 16 bytes, and the stabilizer rotations of a tapering are dozens of such calls on a handful of rows each).
template <int MODE>
__global__ void __launch_bounds__(256) rotate_fused_row_kernel(const uint64_t *__restrict__ xz, const double2 *__restrict__ c, int64_t M,
                                                               int W, const uint64_t *__restrict__ q_xz, double cos_a, double sin_a,
                                                               double sign, uint64_t *__restrict__ out_xz, double2 *__restrict__ out_c,
                                                               int64_t *__restrict__ n_out) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row == 0) *n_out = MODE == 4 ? 2 * M : M;
    if (row >= M) return;
    const uint64_t *r = xz + row * 2 * W;
    uint64_t comm = 0, s = 0;
    int ya = 0, yb = 0, yout = 0;
    for (int w = 0; w < W; ++w) {
        const uint64_t xa = r[w], za = r[W + w], xb = q_xz[w], zb = q_xz[W + w];
        comm ^= (xa & zb) ^ (za & xb);
        s ^= xa & zb;
        ya += __popcll(xa & za);
        yb += __popcll(xb & zb);
        yout += __popcll((xa ^ xb) & (za ^ zb));
    }
    const bool anti = (__popcll(comm) & 1) != 0;
    const int e = (3 * (ya + yb) + yout + 2 * (__popcll(s) & 1)) & 3;
    uint64_t *o1 = out_xz + row * 2 * W;
    const double2 cc = c[row];
    if (MODE == 4) {
        uint64_t *o2 = out_xz + (M + row) * 2 * W;
        for (int w = 0; w < 2 * W; ++w) {
            const uint64_t v = r[w];
            o1[w] = v;
            o2[w] = anti ? v ^ q_xz[w] : v;
        }
        if (anti) {
            out_c[row] = make_double2(cc.x * cos_a, cc.y * cos_a);
            double re = cc.x, im = cc.y;
            mul_i_pow(re, im, e);
            out_c[M + row] = make_double2(im * sin_a, -re * sin_a);
        } else {
            out_c[row] = cc;
            out_c[M + row] = make_double2(0.0, 0.0);
        }
    } else if (MODE == 1) {
        for (int w = 0; w < 2 * W; ++w) o1[w] = anti ? r[w] ^ q_xz[w] : r[w];
        if (anti) {
            double re = cc.x, im = cc.y;
            mul_i_pow(re, im, e + 3);
            out_c[row] = make_double2(re * sign, im * sign);
        } else {
            out_c[row] = cc;
        }
    } else {
        for (int w = 0; w < 2 * W; ++w) o1[w] = r[w];
        out_c[row] = anti ? make_double2(cc.x * sign, cc.y * sign) : cc;
    }
}

// sym_rotate_split: stable split of the rows into (commuting with Q | anticommuting with Q), coefficients untouched.
// The general rotation is then ONE block-list product (sym_mul_blocks_*) of the split operator with the three-row
// operator [I, cos I, -i sin Q]: commuting rows x [I], anticommuting rows x [cos I, -i sin Q] — the rotated
// rows are never materialised before the dedup (base.py:1090-1161 as a product, base.py:764-794).
__global__ void __launch_bounds__(256) rotate_split_kernel(const uint4 *__restrict__ xz, const double2 *__restrict__ c, int64_t M,
                                                            uint32_t chunks, const uint8_t *__restrict__ info,
                                                            const uint32_t *__restrict__ rank, const uint32_t *__restrict__ total,
                                                            const uint64_t *__restrict__ sk, const int32_t *__restrict__ yc,
                                                            uint4 *__restrict__ out_xz, double2 *__restrict__ out_c,
                                                            uint64_t *__restrict__ out_sk, int32_t *__restrict__ out_y) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t row = g / chunks;
    if (row >= (size_t)M) return;
    const uint32_t k = (uint32_t)(g - row * chunks);
    const bool anti = info[row] & 1;
    const size_t n_comm = (size_t)M - *total;
    const size_t dst = anti ? n_comm + rank[row] : row - rank[row];
    out_xz[dst * chunks + k] = xz[g];
    if (k == 0) {
        out_c[dst] = c[row];
        out_sk[dst] = sk[row];
        out_y[dst] = yc[row];
    }
}

__global__ void rotate_count_kernel(const uint32_t *__restrict__ total, int64_t M, int mode, int64_t *__restrict__ n_out) {
    *n_out = (mode == 0) ? M + (int64_t)*total : (mode == 3 ? M - (int64_t)*total : M);   // 3: sym_rotate_split
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
    SYM_REQUIRE((mode >= 0 && mode <= 2) || mode == 4, "mode must be 0, 1, 2 or 4");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        SYM_CUDA_OK(cudaMemsetAsync(n_out, 0, sizeof(int64_t), st));
        return SYM_OK;
    }
    if (mode != 0 && group8_ok(W)) {   // one pass: no flags, no scan, no workspace
        const unsigned nbf = (unsigned)((((M + 3) / 4) * 32 + 255) / 256);
        const double2 *c2 = reinterpret_cast<const double2 *>(c);
        uint4 *o4 = reinterpret_cast<uint4 *>(out_xz);
        double2 *oc2 = reinterpret_cast<double2 *>(out_c);
        if (mode == 4) rotate_fused8_kernel<4><<<nbf, 256, 0, st>>>(xz, c2, M, W, q_xz, cos_a, sin_a, sign, o4, oc2, n_out);
        else if (mode == 1) rotate_fused8_kernel<1><<<nbf, 256, 0, st>>>(xz, c2, M, W, q_xz, cos_a, sin_a, sign, o4, oc2, n_out);
        else rotate_fused8_kernel<2><<<nbf, 256, 0, st>>>(xz, c2, M, W, q_xz, cos_a, sin_a, sign, o4, oc2, n_out);
        SYM_LAUNCH_OK();
        return SYM_OK;
    }
    if (mode != 0 && (W <= 2 || mode == 4)) {   // narrow rows (or any width for the padded form): one thread per row
        const unsigned nbr = (unsigned)((M + 255) / 256);
        const double2 *c2 = reinterpret_cast<const double2 *>(c);
        double2 *oc2 = reinterpret_cast<double2 *>(out_c);
        if (mode == 4) rotate_fused_row_kernel<4><<<nbr, 256, 0, st>>>(xz, c2, M, W, q_xz, cos_a, sin_a, sign, out_xz, oc2, n_out);
        else if (mode == 1) rotate_fused_row_kernel<1><<<nbr, 256, 0, st>>>(xz, c2, M, W, q_xz, cos_a, sin_a, sign, out_xz, oc2, n_out);
        else rotate_fused_row_kernel<2><<<nbr, 256, 0, st>>>(xz, c2, M, W, q_xz, cos_a, sin_a, sign, out_xz, oc2, n_out);
        SYM_LAUNCH_OK();
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
    if (group8_ok(W))
        rotate_info8_kernel<<<(unsigned)((((M + 3) / 4) * 32 + 255) / 256), 256, 0, st>>>(xz, M, W, q_xz, info, nullptr, nullptr, mode == 0 ? anti : nullptr);
    else
        rotate_info_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, st>>>(xz, M, W, q_xz, info, nullptr, nullptr, mode == 0 ? anti : nullptr);
    SYM_LAUNCH_OK();
    if (mode == 0) {
        SYM_TRY(scan_exclusive_u8(anti, rank, M, total, scratch, st));
    } else {
        SYM_CUDA_OK(cudaMemsetAsync(total, 0, sizeof(uint32_t), st));
    }
    {
        const uint32_t chunks = (uint32_t)W;   // 2W words = W 16-byte chunks
        const size_t threads = (size_t)M * chunks;
        const unsigned nbw = (unsigned)((threads + 255) / 256);
        const uint4 *x4 = reinterpret_cast<const uint4 *>(xz);
        const uint4 *q4 = reinterpret_cast<const uint4 *>(q_xz);
        const double2 *c2 = reinterpret_cast<const double2 *>(c);
        uint4 *o4 = reinterpret_cast<uint4 *>(out_xz);
        double2 *oc2 = reinterpret_cast<double2 *>(out_c);
#define RW_CASE(LC) rotate_write4_kernel<LC><<<nbw, 256, 0, st>>>(x4, c2, M, chunks, q4, info, rank, cos_a, sin_a, mode, sign, o4, oc2)
        switch (chunks) {
            case 1: RW_CASE(0); break;
            case 2: RW_CASE(1); break;
            case 4: RW_CASE(2); break;
            case 8: RW_CASE(3); break;
            case 16: RW_CASE(4); break;
            default: RW_CASE(-1); break;
        }
#undef RW_CASE
        SYM_LAUNCH_OK();
    }
    rotate_count_kernel<<<1, 1, 0, st>>>(total, M, mode, n_out);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

extern "C" size_t sym_rotate_split_ws_bytes(int64_t M) {
    if (M < 1) M = 1;
    return sym_rotate_ws_bytes(M) + arena_need((size_t)M, 8) + arena_need((size_t)M, 4) + 1024;
}

extern "C" int sym_rotate_split(const uint64_t *xz, const double *c, int64_t M, int32_t W, const uint64_t *q_xz,
                                uint64_t *out_xz, double *out_c, uint64_t *out_sketch, int32_t *out_ycount,
                                int64_t *n_commuting, void *ws, size_t ws_bytes, void *stream) {
    SYM_REQUIRE(M >= 0 && W >= 1, "bad size");
    SYM_REQUIRE(M < ((int64_t)1 << 32), "too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        SYM_CUDA_OK(cudaMemsetAsync(n_commuting, 0, sizeof(int64_t), st));
        return SYM_OK;
    }
    if (ws_bytes < sym_rotate_split_ws_bytes(M)) {
        set_error("workspace too small");
        return SYM_E_WORKSPACE;
    }
    Arena ar(ws, ws_bytes);
    uint8_t *info = ar.take<uint8_t>((size_t)M);
    uint8_t *anti = ar.take<uint8_t>((size_t)M);
    uint32_t *rank = ar.take<uint32_t>((size_t)M);
    uint32_t *scratch = ar.take<uint32_t>(scan_scratch_elems(M));
    uint32_t *total = ar.take<uint32_t>(4);
    uint64_t *sk = ar.take<uint64_t>((size_t)M);
    int32_t *yc = ar.take<int32_t>((size_t)M);
    if (group8_ok(W))
        rotate_info8_kernel<<<(unsigned)((((M + 3) / 4) * 32 + 255) / 256), 256, 0, st>>>(xz, M, W, q_xz, info, sk, yc, anti);
    else
        rotate_info_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, st>>>(xz, M, W, q_xz, info, sk, yc, anti);
    SYM_LAUNCH_OK();
    SYM_TRY(scan_exclusive_u8(anti, rank, M, total, scratch, st));
    const uint32_t chunks = (uint32_t)W;
    const size_t threads = (size_t)M * chunks;
    rotate_split_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const uint4 *>(xz), reinterpret_cast<const double2 *>(c), M, chunks, info, rank, total, sk, yc,
        reinterpret_cast<uint4 *>(out_xz), reinterpret_cast<double2 *>(out_c), out_sketch, out_ycount);
    SYM_LAUNCH_OK();
    rotate_count_kernel<<<1, 1, 0, st>>>(total, M, 3, n_commuting);
    SYM_LAUNCH_OK();
    return SYM_OK;
}
