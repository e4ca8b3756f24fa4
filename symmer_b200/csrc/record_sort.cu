// Stable LSD radix sort of 64-bit records (hash | t | phase) — the dedup hot path.
// Per pass: (1) per-tile digit histogram, (2) exclusive scan of the digit-major histogram,
// (3) scatter: rank inside the tile with warp match.any, stage the tile in shared memory in
// digit order, then store — consecutive threads write consecutive addresses within a digit run.
// HBM-bound: 8 B read (hist) + 8 B read + 8 B write (scatter) per record per pass.
#include "sort.cuh"

namespace symb {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 records
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t *__restrict__ in, int64_t T, int shift,
                                                              uint32_t mask, uint32_t *__restrict__ hist, int64_t ntiles) {
    __shared__ uint32_t h[RS_RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int64_t idx = base + j * RS_THREADS + threadIdx.x;
        if (idx < T) atomicAdd(&h[(uint32_t)(in[idx] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS, 3) rs_scatter_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out,
                                                                 int64_t T, int shift, uint32_t mask,
                                                                 const uint32_t *__restrict__ hist, int64_t ntiles) {
    __shared__ uint64_t srec[RS_TILE];
    __shared__ uint32_t wh[RS_WARPS][RS_RADIX];
    __shared__ uint32_t dbase[RS_RADIX];
    __shared__ uint32_t delta[RS_RADIX];
    __shared__ uint32_t wtot[RS_WARPS];
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t tile_base = (int64_t)blockIdx.x * RS_TILE;
    const int64_t wbase = tile_base + (int64_t)wid * (32 * RS_ITEMS);
    uint64_t k[RS_ITEMS];
    uint32_t r[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int64_t idx = wbase + j * 32 + lane;
        k[j] = (idx < T) ? in[idx] : 0ull;
    }
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int64_t idx = wbase + j * 32 + lane;
        bool valid = idx < T;
        uint32_t d = valid ? ((uint32_t)(k[j] >> shift) & mask) : 0xffffffffu;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t before = valid ? wh[wid][d] : 0u;
        r[j] = before + __popc(peers & lt);
        __syncwarp();
        if (valid && lane == (__ffs(peers) - 1)) wh[wid][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive offsets of the warps, tile total, then a block scan over the 256 totals
    uint32_t cnt = 0;
    {
        const int d = threadIdx.x;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = wh[w][d];
            wh[w][d] = cnt;
            cnt += c;
        }
    }
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wtot[wid] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w)
        if (w < wid) woff += wtot[w];
    const uint32_t excl = woff + inc - cnt;
    dbase[threadIdx.x] = excl;
    delta[threadIdx.x] = hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] - excl;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int64_t idx = wbase + j * 32 + lane;
        if (idx < T) {
            uint32_t d = (uint32_t)(k[j] >> shift) & mask;
            srec[dbase[d] + wh[wid][d] + r[j]] = k[j];
        }
    }
    __syncthreads();
    const int64_t remain = T - tile_base;
    const int count = remain < RS_TILE ? (int)remain : RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int pos = j * RS_THREADS + threadIdx.x;
        if (pos < count) {
            uint64_t rec = srec[pos];
            uint32_t d = (uint32_t)(rec >> shift) & mask;
            out[(size_t)(delta[d] + (uint32_t)pos)] = rec;
        }
    }
}

__global__ void rs_bucket_counts_kernel(const uint32_t *__restrict__ scanned, int64_t ntiles, int nb, int64_t T,
                                        int64_t *__restrict__ counts) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < nb) {
        int64_t lo = scanned[(int64_t)d * ntiles];
        int64_t hi = (d + 1 < RS_RADIX) ? (int64_t)scanned[(int64_t)(d + 1) * ntiles] : T;
        counts[d] = hi - lo;
    }
}

size_t record_hist_elems(int64_t T) {
    int64_t ntiles = (T + RS_TILE - 1) / RS_TILE;
    if (ntiles < 1) ntiles = 1;
    size_t h = (size_t)RS_RADIX * (size_t)ntiles;
    return h + scan_scratch_elems((int64_t)h) + 64;
}

static int rs_pass(const uint64_t *in, uint64_t *out, int64_t T, int shift, uint32_t mask, uint32_t *hist, cudaStream_t st) {
    int64_t ntiles = (T + RS_TILE - 1) / RS_TILE;
    int64_t hn = (int64_t)RS_RADIX * ntiles;
    uint32_t *scratch = hist + hn;
    rs_hist_kernel<<<(unsigned)ntiles, RS_THREADS, 0, st>>>(in, T, shift, mask, hist, ntiles);
    SYM_LAUNCH_OK();
    SYM_TRY(scan_exclusive_u32(hist, hist, hn, nullptr, scratch, st));
    rs_scatter_kernel<<<(unsigned)ntiles, RS_THREADS, 0, st>>>(in, out, T, shift, mask, hist, ntiles);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

int radix_sort_records(uint64_t *keys, uint64_t *alt, int64_t T, int begin_bit, uint32_t *hist, uint64_t **result,
                       cudaStream_t st) {
    *result = keys;
    if (T <= 1) return SYM_OK;
    if (begin_bit < 0) begin_bit = 0;
    if (begin_bit > 63) begin_bit = 63;
    uint64_t *a = keys, *b = alt;
    int bit = begin_bit;
    while (bit < 64) {
        int width = (64 - bit) < 8 ? (64 - bit) : 8;
        SYM_TRY(rs_pass(a, b, T, bit, (1u << width) - 1u, hist, st));
        uint64_t *t = a; a = b; b = t;
        bit += width;
    }
    *result = a;
    return SYM_OK;
}

int radix_partition_records(const uint64_t *keys, uint64_t *out, int64_t T, int bits, int64_t *counts, uint32_t *hist,
                            cudaStream_t st) {
    const int nb = 1 << bits;
    if (T <= 0) {
        SYM_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * nb, st));
        return SYM_OK;
    }
    if (bits == 0) {
        SYM_CUDA_OK(cudaMemcpyAsync(out, keys, sizeof(uint64_t) * (size_t)T, cudaMemcpyDeviceToDevice, st));
        int64_t t = T;
        SYM_CUDA_OK(cudaMemcpyAsync(counts, &t, sizeof(int64_t), cudaMemcpyHostToDevice, st));
        SYM_CUDA_OK(cudaStreamSynchronize(st));
        return SYM_OK;
    }
    SYM_TRY(rs_pass(keys, out, T, 64 - bits, (1u << bits) - 1u, hist, st));
    int64_t ntiles = (T + RS_TILE - 1) / RS_TILE;
    rs_bucket_counts_kernel<<<1, 256, 0, st>>>(hist, ntiles, nb, T, counts);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

}  // namespace symb
