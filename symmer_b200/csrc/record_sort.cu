// Stable LSD radix sort of 64-bit records (hash | t | phase) — the dedup hot path.
// Per pass: (1) per-tile digit histogram, (2) exclusive scan of the digit-major histogram,
// (3) scatter: rank inside the tile with warp match.any, stage the tile in shared memory in
// digit order, then store — consecutive threads write consecutive addresses within a digit run.
// HBM-bound: 8 B read (hist) + 8 B read + 8 B write (scatter) per record per pass.
#include "sort.cuh"

namespace symb {

constexpr int RS_THREADS = 256;
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

// THREADS x ITEMS = RS_TILE. Ranks are < 4096 and are kept as 16-bit halves to save registers.
template <int THREADS, int ITEMS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) rs_scatter_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out,
                                                                    int64_t T, int shift, uint32_t mask,
                                                                    const uint32_t *__restrict__ hist, int64_t ntiles) {
    static_assert(THREADS * ITEMS == RS_TILE, "tile size is fixed");
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t *srec = reinterpret_cast<uint64_t *>(rs_smem);
    uint32_t(*wh)[RS_RADIX] = reinterpret_cast<uint32_t(*)[RS_RADIX]>(rs_smem + RS_TILE * 8);
    uint32_t *dbase = reinterpret_cast<uint32_t *>(rs_smem + RS_TILE * 8 + WARPS * RS_RADIX * 4);
    uint32_t *delta = dbase + RS_RADIX;
    uint32_t *wtot = delta + RS_RADIX;
    for (int i = threadIdx.x; i < WARPS * RS_RADIX; i += THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t tile_base = (int64_t)blockIdx.x * RS_TILE;
    const int64_t wbase = tile_base + (int64_t)wid * (32 * ITEMS);
    uint64_t k[ITEMS];
    uint16_t r[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        int64_t idx = wbase + j * 32 + lane;
        k[j] = (idx < T) ? in[idx] : 0ull;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        int64_t idx = wbase + j * 32 + lane;
        bool valid = idx < T;
        uint32_t d = valid ? ((uint32_t)(k[j] >> shift) & mask) : 0xffffffffu;
        // lanes holding the same digit, built from 8 ballots (match.any is far slower on sm_100)
        uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
        uint32_t before = valid ? wh[wid][d] : 0u;
        r[j] = (uint16_t)(before + __popc(peers & lt));
        __syncwarp();
        if (valid && lane == (__ffs(peers) - 1)) wh[wid][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive offsets of the warps, tile total, then a block scan over the 256 totals
    if (threadIdx.x < RS_RADIX) {
        uint32_t cnt = 0;
        const int d = threadIdx.x;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            uint32_t c = wh[w][d];
            wh[w][d] = cnt;
            cnt += c;
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) wtot[wid] = inc;
        dbase[d] = inc - cnt;   // exclusive inside the warp; warp offset added below
    }
    __syncthreads();
    if (threadIdx.x < RS_RADIX) {
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_RADIX / 32; ++w)
            if (w < wid) woff += wtot[w];
        const uint32_t excl = dbase[threadIdx.x] + woff;
        dbase[threadIdx.x] = excl;
        delta[threadIdx.x] = hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] - excl;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
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
    for (int j = 0; j < ITEMS; ++j) {
        int pos = j * THREADS + threadIdx.x;
        if (pos < count) {
            uint64_t rec = srec[pos];
            uint32_t d = (uint32_t)(rec >> shift) & mask;
            out[(size_t)(delta[d] + (uint32_t)pos)] = rec;
        }
    }
}

int g_scatter_variant = 3;  // tuning knob 2: 256 threads x 16 records, 4 CTAs/SM (measured best)

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
#define RS_LAUNCH(TH, IT, MB)                                                                                     \
    {                                                                                                             \
        constexpr size_t smem = RS_TILE * 8 + (TH / 32) * RS_RADIX * 4 + 2 * RS_RADIX * 4 + 64;                   \
        static bool attr_done = false;                                                                            \
        if (!attr_done) {                                                                                         \
            SYM_CUDA_OK(cudaFuncSetAttribute(rs_scatter_kernel<TH, IT, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem));                                                         \
            attr_done = true;                                                                                     \
        }                                                                                                         \
        rs_scatter_kernel<TH, IT, MB><<<(unsigned)ntiles, TH, smem, st>>>(in, out, T, shift, mask, hist, ntiles); \
    }
    switch (g_scatter_variant) {
        case 1: RS_LAUNCH(512, 8, 2); break;
        case 2: RS_LAUNCH(512, 8, 3); break;
        case 3: RS_LAUNCH(256, 16, 4); break;
        case 4: RS_LAUNCH(1024, 4, 1); break;
        case 5: RS_LAUNCH(1024, 4, 2); break;
        default: RS_LAUNCH(256, 16, 3); break;
    }
#undef RS_LAUNCH
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
