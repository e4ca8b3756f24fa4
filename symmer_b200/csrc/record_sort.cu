// Stable LSD radix sort of 64-bit records (hash | t | phase) — the dedup hot path.
// Per pass: (1) per-tile digit histogram, (2) exclusive scan of the digit-major histogram,
// (3) scatter: rank inside the tile with warp match.any, stage the tile in shared memory in
// digit order, then store — consecutive threads write consecutive addresses within a digit run.
// HBM-bound: 8 B read (hist) + 8 B read + 8 B write (scatter) per record per pass.
#include <algorithm>

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

// ---------------------------------------------------------------------------------------------
// One-sweep form (tuning knob 8, default): ONE kernel builds the digit histograms of every pass
// (the keys never change, only their order), then each pass is a single kernel — a tile learns its
// global offsets from the tiles before it by decoupled look-back instead of from a per-pass
// histogram + scan. Per pass 16 B/record of HBM traffic instead of 24 B and two launches fewer.
// Tiles take their index from a ticket counter, so every tile a CTA waits for is already running.
// ---------------------------------------------------------------------------------------------
constexpr int OS_MAX_PASSES = 8;
constexpr int OS_EXTRA_ELEMS = OS_MAX_PASSES * RS_RADIX * 2 + 64;   // global counts + bases + tickets

struct PlainKeySrc {
    const uint64_t *in;
    // k[j] = record at first + j*32 (0 beyond T)
    template <int ITEMS>
    __device__ __forceinline__ void load(int64_t first, int64_t T, uint64_t (&k)[ITEMS]) const {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int64_t idx = first + j * 32;
            k[j] = idx < T ? in[idx] : 0ull;
        }
    }
};

struct ProductKeyGen {
    ProductKeySrc s;
    __device__ __forceinline__ void locate(uint32_t idx, int &b, uint32_t &p, uint32_t &q) const {
        b = 0;
        while (b + 1 < s.nblk && idx >= s.rec_off[b + 1]) ++b;
        const uint32_t local = idx - s.rec_off[b];
        q = local / s.m_blk[b];
        p = local - q * s.m_blk[b];
    }
    template <int ITEMS>
    __device__ __forceinline__ void load(int64_t first, int64_t T, uint64_t (&k)[ITEMS]) const {
        int b = 0;
        uint32_t p = 0, q = 0;
        if (first < T) locate((uint32_t)first, b, p, q);
        const int sh = s.tb + 2;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int64_t idx = first + j * 32;
            uint64_t rec = 0ull;
            if (idx < T) {
                const uint32_t pg = s.p0[b] + p, qg = s.q0[b] + q;
                const uint64_t h = mix64(s.a_sk[pg] ^ s.b_sk[qg]) & s.key_mask;
                rec = ((h >> sh) << sh) | (((uint64_t)qg * s.M_total + pg) << 2);
                if (idx + 32 < T) {
                    if ((uint32_t)(idx + 32) >= s.rec_off[b + 1]) {
                        locate((uint32_t)(idx + 32), b, p, q);   // next item is in another block
                    } else {
                        p += 32;
                        while (p >= s.m_blk[b]) {
                            p -= s.m_blk[b];
                            ++q;
                        }
                    }
                }
            }
            k[j] = rec;
        }
    }
};

// histograms of all passes in one read (or one generation) of the keys; grid-stride over tiles
template <class Src>
__global__ void __launch_bounds__(RS_THREADS) os_hist_kernel(Src src, int64_t T, int begin_bit, int passes, int64_t ntiles,
                                                              uint32_t *__restrict__ counts) {
    __shared__ uint32_t h[OS_MAX_PASSES][RS_RADIX];
    for (int i = threadIdx.x; i < OS_MAX_PASSES * RS_RADIX; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint64_t k[RS_ITEMS];
        const int64_t first = tile * RS_TILE + (int64_t)wid * (32 * RS_ITEMS) + lane;
        src.template load<RS_ITEMS>(first, T, k);
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            if (first + j * 32 < T) {
                for (int ps = 0; ps < passes; ++ps) {
                    const int bit = begin_bit + 8 * ps;
                    atomicAdd(&h[ps][(uint32_t)(k[j] >> bit) & 0xffu], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_THREADS) {
        const uint32_t v = (&h[0][0])[i];
        if (v) atomicAdd(counts + i, v);
    }
}

// per pass: exclusive scan of the 256 digit counts -> bases[pass][digit]; one block of 256 threads per pass
__global__ void __launch_bounds__(RS_RADIX) os_bases_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ bases) {
    __shared__ uint32_t wt[RS_RADIX / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t c = counts[blockIdx.x * RS_RADIX + threadIdx.x];
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wt[wid] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < wid; ++w) woff += wt[w];
    bases[blockIdx.x * RS_RADIX + threadIdx.x] = woff + inc - c;
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One radix pass. state[tile][digit]: bits 31:30 = 0 not published / 1 tile count / 2 inclusive
// prefix over tiles 0..tile; bits 29:0 = value (T < 2^30 on this path).
template <class Src, int THREADS, int ITEMS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) os_pass_kernel(Src src, uint64_t *__restrict__ out, int64_t T, int shift,
                                                                 uint32_t mask, const uint32_t *__restrict__ bases,
                                                                 uint32_t *__restrict__ state, uint32_t *__restrict__ ticket) {
    static_assert(THREADS * ITEMS == RS_TILE, "tile size is fixed");
    static_assert(THREADS >= RS_RADIX, "one look-back thread per digit");
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t *srec = reinterpret_cast<uint64_t *>(rs_smem);
    uint32_t(*wh)[RS_RADIX] = reinterpret_cast<uint32_t(*)[RS_RADIX]>(rs_smem + RS_TILE * 8);
    uint32_t *dbase = reinterpret_cast<uint32_t *>(rs_smem + RS_TILE * 8 + WARPS * RS_RADIX * 4);
    uint32_t *delta = dbase + RS_RADIX;
    uint32_t *wtot = delta + RS_RADIX;
    uint32_t *s_tile = wtot + 32;
    for (int i = threadIdx.x; i < WARPS * RS_RADIX; i += THREADS) (&wh[0][0])[i] = 0;
    if (threadIdx.x == 0) *s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = *s_tile;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t tile_base = (int64_t)tile * RS_TILE;
    const int64_t wbase = tile_base + (int64_t)wid * (32 * ITEMS);
    uint64_t k[ITEMS];
    uint16_t r[ITEMS];
    src.template load<ITEMS>(wbase + lane, T, k);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        int64_t idx = wbase + j * 32 + lane;
        bool valid = idx < T;
        uint32_t d = valid ? ((uint32_t)(k[j] >> shift) & mask) : 0xffffffffu;
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
    uint32_t cnt = 0, excl_d = 0;
    if (threadIdx.x < RS_RADIX) {
        const int d = threadIdx.x;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            uint32_t c = wh[w][d];
            wh[w][d] = cnt;
            cnt += c;
        }
        // publish the tile count as early as possible: the tiles after this one are waiting for it
        st_volatile_u32(state + (size_t)tile * RS_RADIX + d, (tile == 0 ? 0x80000000u : 0x40000000u) | cnt);
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
        const int d = threadIdx.x;
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_RADIX / 32; ++w)
            if (w < wid) woff += wtot[w];
        excl_d = dbase[d] + woff;
        dbase[d] = excl_d;
    }
    __syncthreads();
    // stage the tile in digit order first: it needs no global offsets, and the tiles this one is about to
    // wait for get that much closer to publishing theirs
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        int64_t idx = wbase + j * 32 + lane;
        if (idx < T) {
            uint32_t d = (uint32_t)(k[j] >> shift) & mask;
            srec[dbase[d] + wh[wid][d] + r[j]] = k[j];
        }
    }
    if (threadIdx.x < RS_RADIX) {
        const int d = threadIdx.x;
        // decoupled look-back over the tiles before this one, four state words in flight per step
        uint32_t prev = 0;
        if (tile > 0) {
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = (t - u >= 0) ? ld_volatile_u32(state + (size_t)(t - u) * RS_RADIX + d) : 0x80000000u;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (done) break;
                    uint32_t x = v[u], spins = 0;
                    while ((x >> 30) == 0u) {
                        if (++spins > (1u << 22)) __trap();   // a predecessor never published: fail loudly, never hang
                        x = ld_volatile_u32(state + (size_t)(t - u) * RS_RADIX + d);
                    }
                    prev += x & 0x3fffffffu;
                    done = (x >> 30) == 2u;
                }
                t -= 4;
            }
            st_volatile_u32(state + (size_t)tile * RS_RADIX + d, 0x80000000u | (prev + cnt));
        }
        delta[d] = bases[d] + prev - excl_d;
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

int g_onesweep = 1;   // tuning knob 8

// Short sorts keep one look-back state array PER PASS, so that a single memset clears them all (and the pass
// counters behind them) instead of one memset in front of every pass: on a 250 000-record sort the passes take
// 10 us each and every extra stream operation costs 3-4 us of launch gap.
constexpr int64_t OS_MULTI_STATE_TILES = 1024;
static inline int os_state_copies(int64_t ntiles) { return ntiles <= OS_MULTI_STATE_TILES ? OS_MAX_PASSES : 1; }

size_t record_hist_elems(int64_t T) {
    int64_t ntiles = (T + RS_TILE - 1) / RS_TILE;
    if (ntiles < 1) ntiles = 1;
    size_t h = (size_t)RS_RADIX * (size_t)ntiles;
    return h * (size_t)os_state_copies(ntiles) + scan_scratch_elems((int64_t)h) + 64 + OS_EXTRA_ELEMS;
}

template <class Src, int TH, int IT, int MB>
static int os_launch_shape(const Src &src, uint64_t *out, int64_t T, int shift, uint32_t mask, const uint32_t *bases,
                           uint32_t *state, uint32_t *ticket, int64_t ntiles, cudaStream_t st, bool clear_state) {
    constexpr size_t smem = RS_TILE * 8 + (TH / 32) * RS_RADIX * 4 + 2 * RS_RADIX * 4 + 32 * 4 + 64;
    static bool attr_done[64] = {};   // per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        SYM_CUDA_OK(cudaFuncSetAttribute(os_pass_kernel<Src, TH, IT, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[dev] = true;
    }
    if (clear_state) SYM_CUDA_OK(cudaMemsetAsync(state, 0, sizeof(uint32_t) * (size_t)RS_RADIX * (size_t)ntiles, st));
    os_pass_kernel<Src, TH, IT, MB><<<(unsigned)ntiles, TH, smem, st>>>(src, out, T, shift, mask, bases, state, ticket);
    SYM_LAUNCH_OK();
    return SYM_OK;
}

template <class Src>
static int os_launch_pass(const Src &src, uint64_t *out, int64_t T, int shift, uint32_t mask, const uint32_t *bases,
                          uint32_t *state, uint32_t *ticket, int64_t ntiles, cudaStream_t st, bool clear_state) {
    switch (g_scatter_variant) {   // tuning knob 2 (CTA shape); 3 = 256 threads x 16 records, 4 CTAs/SM is the default
        case 1: return os_launch_shape<Src, 512, 8, 2>(src, out, T, shift, mask, bases, state, ticket, ntiles, st, clear_state);
        case 2: return os_launch_shape<Src, 512, 8, 3>(src, out, T, shift, mask, bases, state, ticket, ntiles, st, clear_state);
        case 5: return os_launch_shape<Src, 256, 16, 5>(src, out, T, shift, mask, bases, state, ticket, ntiles, st, clear_state);
        default: return os_launch_shape<Src, 256, 16, 4>(src, out, T, shift, mask, bases, state, ticket, ntiles, st, clear_state);
    }
}

// keys of pass 0 come from `first` (a buffer or the product key generator), later passes ping-pong a <-> b
template <class Src>
static int onesweep_sort(const Src &first, uint64_t *a, uint64_t *b, int64_t T, int begin_bit, uint32_t *hist,
                         uint64_t **result, cudaStream_t st) {
    const int passes = (64 - begin_bit + 7) / 8;
    const int64_t ntiles = (T + RS_TILE - 1) / RS_TILE;
    uint32_t *state = hist;
    const size_t h = (size_t)RS_RADIX * (size_t)ntiles;
    const int copies = os_state_copies(ntiles);
    uint32_t *extra = hist + h * (size_t)copies + scan_scratch_elems((int64_t)h) + 64;
    uint32_t *counts = extra, *bases = extra + OS_MAX_PASSES * RS_RADIX, *tickets = bases + OS_MAX_PASSES * RS_RADIX;
    if (copies > 1)   // every pass's state, the (unused) scan scratch between, and the counters: one memset
        SYM_CUDA_OK(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * ((size_t)(extra - hist) + OS_EXTRA_ELEMS), st));
    else
        SYM_CUDA_OK(cudaMemsetAsync(extra, 0, sizeof(uint32_t) * OS_EXTRA_ELEMS, st));
    const unsigned hgrid = (unsigned)std::min<int64_t>(ntiles, (int64_t)num_sms() * 8);
    os_hist_kernel<Src><<<hgrid, RS_THREADS, 0, st>>>(first, T, begin_bit, passes, ntiles, counts);
    SYM_LAUNCH_OK();
    os_bases_kernel<<<passes, RS_RADIX, 0, st>>>(counts, bases);
    SYM_LAUNCH_OK();
    int bit = begin_bit;
    for (int ps = 0; ps < passes; ++ps) {
        const int width = (64 - bit) < 8 ? (64 - bit) : 8;
        const uint32_t mask = (1u << width) - 1u;
        uint32_t *state_ps = copies > 1 ? state + (size_t)ps * h : state;
        if (ps == 0) {
            SYM_TRY(os_launch_pass(first, b, T, bit, mask, bases + ps * RS_RADIX, state_ps, tickets + ps, ntiles, st, copies == 1));
        } else {
            PlainKeySrc src{a};
            SYM_TRY(os_launch_pass(src, b, T, bit, mask, bases + ps * RS_RADIX, state_ps, tickets + ps, ntiles, st, copies == 1));
        }
        uint64_t *t = a; a = b; b = t;
        bit += width;
    }
    *result = a;
    return SYM_OK;
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
        static bool attr_done[64] = {};                                                                           \
        int dev_ = 0;                                                                                             \
        cudaGetDevice(&dev_);                                                                                     \
        if (dev_ >= 0 && dev_ < 64 && !attr_done[dev_]) {                                                         \
            SYM_CUDA_OK(cudaFuncSetAttribute(rs_scatter_kernel<TH, IT, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem));                                                         \
            attr_done[dev_] = true;                                                                               \
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

// materialise the generated records (fallback when the one-sweep path is off or T >= 2^30)
__global__ void __launch_bounds__(RS_THREADS) os_materialise_kernel(ProductKeyGen gen, int64_t T, uint64_t *__restrict__ out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t first = (int64_t)blockIdx.x * RS_TILE + (int64_t)wid * (32 * RS_ITEMS) + lane;
    uint64_t k[RS_ITEMS];
    gen.template load<RS_ITEMS>(first, T, k);
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j)
        if (first + j * 32 < T) out[first + j * 32] = k[j];
}

int radix_sort_product_keys(const ProductKeySrc &src, uint64_t *keys, uint64_t *alt, int64_t T, int begin_bit,
                            uint32_t *hist, uint64_t **result, cudaStream_t st) {
    *result = keys;
    if (T <= 0) return SYM_OK;
    if (begin_bit < 0) begin_bit = 0;
    if (begin_bit > 63) begin_bit = 63;
    ProductKeyGen gen{src};
    if (T > 1 && g_onesweep && T < ((int64_t)1 << 30)) return onesweep_sort(gen, keys, alt, T, begin_bit, hist, result, st);
    const int64_t ntiles = (T + RS_TILE - 1) / RS_TILE;
    os_materialise_kernel<<<(unsigned)ntiles, RS_THREADS, 0, st>>>(gen, T, keys);
    SYM_LAUNCH_OK();
    return radix_sort_records(keys, alt, T, begin_bit, hist, result, st);
}

int radix_sort_records(uint64_t *keys, uint64_t *alt, int64_t T, int begin_bit, uint32_t *hist, uint64_t **result,
                       cudaStream_t st) {
    *result = keys;
    if (T <= 1) return SYM_OK;
    if (begin_bit < 0) begin_bit = 0;
    if (begin_bit > 63) begin_bit = 63;
    if (g_onesweep && T < ((int64_t)1 << 30)) {
        PlainKeySrc src{keys};
        return onesweep_sort(src, keys, alt, T, begin_bit, hist, result, st);
    }
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
