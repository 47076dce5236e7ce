// scan_sort.cu — device-wide exclusive scan and stable LSD radix sort (u64 key, u32 value).
//
// The sort is the "LSD radix sort of (minimizer, read, pos) tuples staged through shared memory" of
// the north star: 8-bit digits; per pass (1) per-tile digit histograms, (2) one exclusive scan over the
// digit-major histogram matrix, (3) a scatter kernel that ranks the tile's keys stably with warp
// match/ballot, regroups them by digit in shared memory and writes each digit run coalesced.
// HBM-bound: algorithmic bytes per key per pass = 2 * 12 B (read + write) + 12 B re-read for the
// histogram.
#include <algorithm>

#include "kernels.cuh"

namespace hlo {

// ------------------------------------------------------------------------------------------------
// exclusive scan
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across a 256-thread block; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total, uint32_t *smem /* >= 9 words */) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v);
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < (SCAN_THREADS / 32) ? smem[lane] : 0u;
        uint32_t si = warp_incl_scan(s);
        if (lane < (SCAN_THREADS / 32)) smem[lane] = si - s;
        if (lane == (SCAN_THREADS / 32) - 1) smem[SCAN_THREADS / 32] = si;
    }
    __syncthreads();
    const uint32_t r = smem[wid] + inc - v;
    *total = smem[SCAN_THREADS / 32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                   uint32_t *__restrict__ sums) {
    __shared__ uint32_t sm[16];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const uint64_t i = base + (uint64_t)j * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t tot;
    block_excl_scan(s, &tot, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// scans one tile; adds offs[blockIdx.x] (or 0).  If total != nullptr and this is the last block, writes the grand total.
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const uint32_t *in, uint32_t *out,
                                                                 uint64_t n, const uint32_t *__restrict__ offs,
                                                                 uint32_t *__restrict__ total) {
    __shared__ uint32_t sm[16];
    __shared__ uint32_t stage[SCAN_TILE];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    // coalesced load into shared memory, then a blocked arrangement per thread
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const uint64_t i = base + (uint64_t)j * SCAN_THREADS + threadIdx.x;
        stage[j * SCAN_THREADS + threadIdx.x] = i < n ? in[i] : 0u;
    }
    __syncthreads();
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        v[j] = stage[threadIdx.x * SCAN_ITEMS + j];
        s += v[j];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan(s, &tot, sm);
    const uint32_t off = offs ? offs[blockIdx.x] : 0u;
    ex += off;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        stage[threadIdx.x * SCAN_ITEMS + j] = ex;
        ex += v[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const uint64_t i = base + (uint64_t)j * SCAN_THREADS + threadIdx.x;
        if (i < n) out[i] = stage[j * SCAN_THREADS + threadIdx.x];
    }
    if (total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total = off + tot;
}

static inline uint64_t ceil_div_u64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

size_t scan_tmp_words(uint64_t n) {
    size_t w = 0;
    while (n > (uint64_t)SCAN_TILE) {
        n = ceil_div_u64(n, SCAN_TILE);
        w += (size_t)((n + 63) & ~63ull);
    }
    return w + 64;
}

void launch_exclusive_scan(const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *tmp, uint32_t *total, cudaStream_t s,
                           LaunchCounter *lc) {
    if (n == 0) {
        if (total) cudaMemsetAsync(total, 0, sizeof(uint32_t), s);
        return;
    }
    const uint64_t nb = ceil_div_u64(n, SCAN_TILE);
    if (nb == 1) {
        scan_tile_kernel<<<1, SCAN_THREADS, 0, s>>>(in, out, n, nullptr, total);
        if (lc) lc->total++;
        return;
    }
    uint32_t *sums = tmp;
    uint32_t *next_tmp = tmp + ((nb + 63) & ~63ull);
    scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, n, sums);
    if (lc) lc->total++;
    launch_exclusive_scan(sums, sums, nb, next_tmp, nullptr, s, lc);
    scan_tile_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, out, n, sums, total);
    if (lc) lc->total++;
}

// ------------------------------------------------------------------------------------------------
// radix sort
constexpr int RS_BITS = 8;
constexpr int RS_BINS = 1 << RS_BITS;
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;

// Segments: the array is a sequence of independent segments (the clusters: their minimizers are contiguous), each
// sorted on its own.  Tiles never straddle a segment and the histogram is laid out (segment, digit, tile), so the one
// exclusive scan still yields final positions — the cluster index costs no key bits and no pass.
// seg_start[c] = first element of segment c, seg_tile[c] = its first tile; seg_start == nullptr: one segment.
struct RsSeg {
    const uint32_t *seg_start = nullptr;
    const uint32_t *seg_tile = nullptr;
    uint32_t n_seg = 0;
};
struct RsTile {
    uint64_t base;      // first element
    uint32_t n;         // elements (0: no such tile)
    uint64_t hist0;     // index of (digit 0, this tile) in the histogram
    uint32_t hstride;   // distance between digits there
};
__device__ __forceinline__ RsTile rs_tile(const RsSeg &g, uint64_t n, uint32_t n_tiles) {
    RsTile t;
    const uint32_t b = blockIdx.x;
    if (!g.seg_start) {
        t.base = (uint64_t)b * RS_TILE;
        t.n = (uint32_t)((n - t.base) < (uint64_t)RS_TILE ? (n - t.base) : (uint64_t)RS_TILE);
        t.hist0 = b, t.hstride = n_tiles;
        return t;
    }
    t.base = 0, t.n = 0, t.hist0 = 0, t.hstride = 0;
    if (b >= g.seg_tile[g.n_seg]) return t;
    uint32_t lo = 0, hi = g.n_seg;  // the segment with seg_tile[lo] <= b < seg_tile[lo + 1]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (g.seg_tile[mid] <= b) lo = mid; else hi = mid;
    }
    const uint32_t t0 = g.seg_tile[lo], nt = g.seg_tile[lo + 1] - t0, lt = b - t0;
    t.base = (uint64_t)g.seg_start[lo] + (uint64_t)lt * RS_TILE;
    const uint64_t left = (uint64_t)g.seg_start[lo + 1] - t.base;
    t.n = (uint32_t)(left < (uint64_t)RS_TILE ? left : (uint64_t)RS_TILE);
    t.hist0 = (uint64_t)RS_BINS * t0 + lt, t.hstride = nt;
    return t;
}

__global__ void rs_segments_kernel(const uint32_t *__restrict__ elem_off, const uint32_t *__restrict__ seg_off, uint32_t n_seg,
                                   uint32_t *__restrict__ seg_start, uint32_t *__restrict__ seg_tile) {
    for (uint32_t c = threadIdx.x; c <= n_seg; c += blockDim.x) seg_start[c] = elem_off[seg_off[c]];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (uint32_t c = 0; c < n_seg; ++c) {
            seg_tile[c] = t;
            t += (seg_start[c + 1] - seg_start[c] + RS_TILE - 1) / RS_TILE;
        }
        seg_tile[n_seg] = t;
    }
}
// the same for many segments: starts and tile counts in parallel (seg_tile[n_seg] = 0), then an exclusive scan
__global__ void rs_segments_wide_kernel(const uint32_t *__restrict__ elem_off, const uint32_t *__restrict__ seg_off, uint32_t n_seg,
                                        uint32_t *__restrict__ seg_start, uint32_t *__restrict__ seg_tile) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_seg) return;
    const uint32_t a = elem_off[seg_off[c]];
    seg_start[c] = a;
    seg_tile[c] = c < n_seg ? (elem_off[seg_off[c + 1]] - a + RS_TILE - 1) / RS_TILE : 0u;
}

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, int shift,
                                                             uint32_t n_tiles, uint32_t *__restrict__ hist, RsSeg seg) {
    __shared__ uint32_t h[RS_BINS];
    const RsTile t = rs_tile(seg, n, n_tiles);
    if (t.n == 0) return;
    h[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint32_t e = (uint32_t)j * RS_THREADS + threadIdx.x;
        if (e < t.n) atomicAdd(&h[(uint32_t)(keys[t.base + e] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    hist[t.hist0 + (uint64_t)threadIdx.x * t.hstride] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS, 4) rs_scatter_kernel(const uint64_t *__restrict__ keys_in,
                                                                const uint32_t *__restrict__ vals_in,
                                                                uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                                uint64_t n, int shift, uint32_t n_tiles,
                                                                const uint32_t *__restrict__ hist_scanned, RsSeg seg) {
    __shared__ uint32_t wcnt[RS_WARPS][RS_BINS];
    __shared__ uint32_t tile_start[RS_BINS];
    __shared__ uint32_t gbase[RS_BINS];
    __shared__ uint32_t sm[16];
    __shared__ uint64_t stage_k[RS_TILE];
    __shared__ uint32_t stage_v[RS_TILE];
    const unsigned lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const RsTile tl = rs_tile(seg, n, n_tiles);
    if (tl.n == 0) return;
    const uint64_t base = tl.base;
    const uint32_t tile_n = tl.n;
#pragma unroll
    for (int w2 = 0; w2 < RS_WARPS; ++w2) wcnt[w2][threadIdx.x] = 0;
    __syncthreads();
    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS], rank[RS_ITEMS];
    const unsigned lt_mask = (1u << lane) - 1u;
    // warp-striped: warp wp owns tile elements [wp*256, wp*256+256), item j of lane l is element wp*256 + j*32 + l
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint32_t e = wp * (RS_ITEMS * 32) + j * 32 + lane;
        const bool ok = e < tile_n;
        key[j] = ok ? keys_in[base + e] : ~0ull;
        val[j] = ok ? vals_in[base + e] : 0u;
        const uint32_t d = ok ? ((uint32_t)(key[j] >> shift) & (RS_BINS - 1)) : (uint32_t)RS_BINS;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t prev = 0;
        if (ok) prev = wcnt[wp][d];
        __syncwarp();
        rank[j] = prev + __popc(peers & lt_mask);
        if (ok && (peers & lt_mask) == 0) wcnt[wp][d] = prev + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {  // thread d: exclusive scan of digit d over the warps, then over the digits
        const uint32_t d = threadIdx.x;
        uint32_t off = 0;
#pragma unroll
        for (int w2 = 0; w2 < RS_WARPS; ++w2) {
            const uint32_t t = wcnt[w2][d];
            wcnt[w2][d] = off;
            off += t;
        }
        uint32_t tot;
        const uint32_t ex = block_excl_scan(off, &tot, sm);
        tile_start[d] = ex;
        gbase[d] = hist_scanned[tl.hist0 + (uint64_t)d * tl.hstride];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint32_t e = wp * (RS_ITEMS * 32) + j * 32 + lane;
        if (e < tile_n) {
            const uint32_t d = (uint32_t)(key[j] >> shift) & (RS_BINS - 1);
            const uint32_t p = tile_start[d] + wcnt[wp][d] + rank[j];
            stage_k[p] = key[j];
            stage_v[p] = val[j];
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tile_n; i += RS_THREADS) {
        const uint64_t kx = stage_k[i];
        const uint32_t d = (uint32_t)(kx >> shift) & (RS_BINS - 1);
        const uint64_t dst = (uint64_t)gbase[d] + (i - tile_start[d]);
        keys_out[dst] = kx;
        vals_out[dst] = stage_v[i];
    }
}

// ------------------------------------------------------------------------------------------------
// One-sweep variant of the same sort (option "sort_onesweep"; off by default, see g_onesweep): the digit counts of ALL passes are taken in one read of
// the keys (they do not depend on the order), and each pass is ONE kernel that reads and writes every
// key once: a block takes the next tile in ticket order, publishes its per-digit counts, and gets its
// global offsets by looking back over the tiles before it (decoupled look-back: a tile publishes first
// its own counts, then the inclusive prefix; a successor adds counts until it meets a prefix).
// Per key and pass: 12 B read + 12 B written (+ 8 B once) instead of 8 + 12 + 12 and five launches.
constexpr int OS_MAX_PASSES = 8;
constexpr uint32_t OS_FLAG_AGG = 1u << 30, OS_FLAG_PREFIX = 2u << 30, OS_VAL_MASK = (1u << 30) - 1u;

__global__ void __launch_bounds__(256) os_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, int begin_bit, int n_pass,
                                                      uint32_t *__restrict__ ghist) {
    __shared__ uint32_t h[OS_MAX_PASSES * RS_BINS];
    for (int i = threadIdx.x; i < n_pass * RS_BINS; i += 256) h[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const uint64_t k = keys[i] >> begin_bit;
        for (int p = 0; p < n_pass; ++p) atomicAdd(&h[p * RS_BINS + ((uint32_t)(k >> (p * RS_BITS)) & (RS_BINS - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pass * RS_BINS; i += 256)
        if (h[i]) atomicAdd(&ghist[i], h[i]);
}

// one block per pass: exclusive scan of the 256 digit totals, in place
__global__ void __launch_bounds__(256) os_scan_kernel(uint32_t *__restrict__ ghist) {
    __shared__ uint32_t sm[16];
    uint32_t *g = ghist + blockIdx.x * RS_BINS;
    uint32_t tot;
    const uint32_t ex = block_excl_scan(g[threadIdx.x], &tot, sm);
    g[threadIdx.x] = ex;
}

__global__ void __launch_bounds__(RS_THREADS) os_pass_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                             uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                             uint64_t n, int shift, const uint32_t *__restrict__ ghist_excl,
                                                             uint32_t *status, uint32_t *ticket, uint32_t *err) {
    __shared__ uint32_t wcnt[RS_WARPS][RS_BINS];
    __shared__ uint32_t tile_start[RS_BINS];
    __shared__ uint32_t gbase[RS_BINS];
    __shared__ uint32_t sm[16];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t stage_k[RS_TILE];
    __shared__ uint32_t stage_v[RS_TILE];
    const unsigned lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);  // tiles are taken in launch order: every earlier tile is running or done
#pragma unroll
    for (int w2 = 0; w2 < RS_WARPS; ++w2) wcnt[w2][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * RS_TILE;
    const uint32_t tile_n = (uint32_t)((n - base) < (uint64_t)RS_TILE ? (n - base) : (uint64_t)RS_TILE);
    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS], rank[RS_ITEMS];
    const unsigned lt_mask = (1u << lane) - 1u;
    // warp-striped: warp wp owns tile elements [wp*256, wp*256+256), item j of lane l is element wp*256 + j*32 + l
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint32_t e = wp * (RS_ITEMS * 32) + j * 32 + lane;
        const bool ok = e < tile_n;
        key[j] = ok ? keys_in[base + e] : ~0ull;
        val[j] = ok ? vals_in[base + e] : 0u;
    }
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint32_t e = wp * (RS_ITEMS * 32) + j * 32 + lane;
        const bool ok = e < tile_n;
        const uint32_t d = ok ? ((uint32_t)(key[j] >> shift) & (RS_BINS - 1)) : (uint32_t)RS_BINS;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t prev = 0;
        if (ok) prev = wcnt[wp][d];
        __syncwarp();
        rank[j] = prev + __popc(peers & lt_mask);
        if (ok && (peers & lt_mask) == 0) wcnt[wp][d] = prev + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {  // thread d: digit d of this tile
        const uint32_t d = threadIdx.x;
        uint32_t cnt = 0;
#pragma unroll
        for (int w2 = 0; w2 < RS_WARPS; ++w2) {
            const uint32_t t = wcnt[w2][d];
            wcnt[w2][d] = cnt;  // exclusive offset of the warp inside the digit
            cnt += t;
        }
        uint32_t *st = status + (size_t)tile * RS_BINS + d;
        atomicExch(st, cnt | (tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG));
        // decoupled look-back over the tiles before this one
        uint32_t excl = 0;
        if (tile > 0) {
            uint32_t t = tile - 1;
            for (;;) {
                const volatile uint32_t *ps = status + (size_t)t * RS_BINS + d;
                uint32_t v = *ps;
                uint32_t spins = 0;
                while ((v >> 30) == 0u) {
                    if (++spins > (1u << 26)) {  // never expected: report instead of hanging the GPU
                        atomicExch(err, 1u);
                        break;
                    }
                    __nanosleep(20);
                    v = *ps;
                }
                excl += v & OS_VAL_MASK;
                if ((v & OS_FLAG_PREFIX) || (v >> 30) == 0u || t == 0) break;
                --t;
            }
            atomicExch(st, (excl + cnt) | OS_FLAG_PREFIX);
        }
        gbase[d] = ghist_excl[d] + excl;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(cnt, &tot, sm);
        tile_start[d] = ex;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint32_t e = wp * (RS_ITEMS * 32) + j * 32 + lane;
        if (e < tile_n) {
            const uint32_t d = (uint32_t)(key[j] >> shift) & (RS_BINS - 1);
            const uint32_t p = tile_start[d] + wcnt[wp][d] + rank[j];
            stage_k[p] = key[j];
            stage_v[p] = val[j];
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tile_n; i += RS_THREADS) {
        const uint64_t kx = stage_k[i];
        const uint32_t d = (uint32_t)(kx >> shift) & (RS_BINS - 1);
        const uint64_t dst = (uint64_t)gbase[d] + (i - tile_start[d]);
        keys_out[dst] = kx;
        vals_out[dst] = stage_v[i];
    }
}

static bool g_onesweep = false;  // measured on B200: 264 us per pass of 20 M keys vs 223 us for the classic three-kernel pass (serial look-back chains)
void radix_sort_use_onesweep(bool on) { g_onesweep = on; }

size_t sort_tmp_words(uint64_t n) {
    const uint64_t n_tiles = ceil_div_u64(n ? n : 1, RS_TILE);
    const uint64_t hist = (uint64_t)RS_BINS * n_tiles;
    const size_t classic = (size_t)(((hist + 63) & ~63ull) + scan_tmp_words(hist));
    const size_t onesweep = (size_t)(OS_MAX_PASSES * RS_BINS + 64 + (uint64_t)OS_MAX_PASSES * n_tiles * RS_BINS);
    return classic > onesweep ? classic : onesweep;
}

int launch_radix_sort(uint64_t *keys[2], uint32_t *vals[2], uint64_t n, int begin_bit, int end_bit, uint32_t *tmp,
                      cudaStream_t s, LaunchCounter *lc) {
    int cur = 0;
    if (n == 0) return cur;
    const uint32_t n_tiles = (uint32_t)ceil_div_u64(n, RS_TILE);
    const int n_pass = (end_bit - begin_bit + RS_BITS - 1) / RS_BITS;
    if (g_onesweep && n_pass <= OS_MAX_PASSES && n < (1ull << 30)) {
        // tmp layout: ghist [n_pass][256] | ticket [8] + err | status [n_pass][n_tiles][256]
        uint32_t *ghist = tmp, *ticket = tmp + OS_MAX_PASSES * RS_BINS, *err = ticket + 32;
        uint32_t *status = tmp + OS_MAX_PASSES * RS_BINS + 64;
        cudaMemsetAsync(tmp, 0, ((size_t)OS_MAX_PASSES * RS_BINS + 64 + (size_t)n_pass * n_tiles * RS_BINS) * 4, s);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const uint64_t hb = std::min<uint64_t>((uint64_t)sms * 8, (n + 255) / 256);
        os_hist_kernel<<<(unsigned)hb, 256, 0, s>>>(keys[cur], n, begin_bit, n_pass, ghist);
        os_scan_kernel<<<n_pass, 256, 0, s>>>(ghist);
        if (lc) lc->total += 2;
        for (int p = 0; p < n_pass; ++p) {
            os_pass_kernel<<<n_tiles, RS_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, begin_bit + p * RS_BITS,
                                                          ghist + p * RS_BINS, status + (size_t)p * n_tiles * RS_BINS, ticket + p, err);
            if (lc) lc->total++;
            cur ^= 1;
        }
        return cur;
    }
    const uint64_t hist_n = (uint64_t)RS_BINS * n_tiles;
    uint32_t *hist = tmp;
    uint32_t *scan_tmp = tmp + ((hist_n + 63) & ~63ull);
    for (int shift = begin_bit; shift < end_bit; shift += RS_BITS) {
        rs_hist_kernel<<<n_tiles, RS_THREADS, 0, s>>>(keys[cur], n, shift, n_tiles, hist, RsSeg());
        if (lc) lc->total++;
        launch_exclusive_scan(hist, hist, hist_n, scan_tmp, nullptr, s, lc);
        rs_scatter_kernel<<<n_tiles, RS_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift, n_tiles,
                                                         hist, RsSeg());
        if (lc) lc->total++;
        cur ^= 1;
    }
    return cur;
}

static uint64_t seg_tile_bound(uint64_t n, uint32_t n_seg) { return ceil_div_u64(n, RS_TILE) + n_seg; }

size_t sort_segmented_tmp_words(uint64_t n, uint32_t n_seg) {
    const uint64_t hist_n = (uint64_t)RS_BINS * seg_tile_bound(n, n_seg);
    return (size_t)(((hist_n + 63) & ~63ull) + scan_tmp_words(hist_n) + 2 * (((uint64_t)n_seg + 64) & ~63ull));
}

int launch_radix_sort_segmented(uint64_t *keys[2], uint32_t *vals[2], uint64_t n, int begin_bit, int end_bit, const uint32_t *elem_off,
                                const uint32_t *seg_off, uint32_t n_seg, uint32_t *tmp, cudaStream_t s, LaunchCounter *lc) {
    int cur = 0;
    if (n == 0 || begin_bit >= end_bit) return cur;
    const uint32_t t_bound = (uint32_t)seg_tile_bound(n, n_seg);  // the tile count lives on the device: launch the bound
    const uint64_t hist_n = (uint64_t)RS_BINS * t_bound;
    uint32_t *hist = tmp;
    uint32_t *scan_tmp = tmp + ((hist_n + 63) & ~63ull);
    uint32_t *seg_start = scan_tmp + scan_tmp_words(hist_n);
    uint32_t *seg_tile = seg_start + (((uint64_t)n_seg + 64) & ~63ull);
    if (n_seg <= 2048) {
        rs_segments_kernel<<<1, 256, 0, s>>>(elem_off, seg_off, n_seg, seg_start, seg_tile);
    } else {  // (the histogram area is free until the first pass and large enough for the scan's partial sums)
        rs_segments_wide_kernel<<<(n_seg + 1 + 255) / 256, 256, 0, s>>>(elem_off, seg_off, n_seg, seg_start, seg_tile);
        launch_exclusive_scan(seg_tile, seg_tile, (uint64_t)n_seg + 1, hist, nullptr, s, lc);
    }
    if (lc) lc->total++;
    RsSeg g;
    g.seg_start = seg_start, g.seg_tile = seg_tile, g.n_seg = n_seg;
    for (int shift = begin_bit; shift < end_bit; shift += RS_BITS) {
        rs_hist_kernel<<<t_bound, RS_THREADS, 0, s>>>(keys[cur], n, shift, t_bound, hist, g);
        if (lc) lc->total++;
        launch_exclusive_scan(hist, hist, hist_n, scan_tmp, nullptr, s, lc);
        rs_scatter_kernel<<<t_bound, RS_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift, t_bound, hist, g);
        if (lc) lc->total++;
        cur ^= 1;
    }
    return cur;
}

// ------------------------------------------------------------------------------------------------
__global__ void copy_words_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, uint32_t n) {
    if (threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}
void launch_copy_words(const uint32_t *src, uint32_t *dst, uint32_t n, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    copy_words_kernel<<<1, 64, 0, s>>>(src, dst, n);
    if (lc) lc->total++;
}

// ---- grouping by a PREFIX of the key -------------------------------------------------------------------------
// The index only needs equal keys adjacent, in stable order — not the order between different keys.  Sorting on
// the top `sort_bits` of the key therefore suffices wherever the keys of a group (a run of equal prefixes) are all
// equal; the few groups that mix two or more keys (a prefix collision) are listed by collision_find_kernel and put
// in full-key order, stably, by collision_fix_kernel.  state[0] = groups listed, state[1] = 1 if a mixed group was
// too large for the one-warp fix (the caller then redoes the sort on all bits).
constexpr uint32_t CF_MAX_GROUP = 512;
constexpr int CF_WARPS = 4;

__global__ void collision_find_kernel(const uint64_t *__restrict__ keys, uint64_t n, int shift, uint32_t *__restrict__ list,
                                      uint32_t *__restrict__ state) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 || i >= n) return;
    const uint64_t b = keys[i];
    uint64_t a = keys[i - 1];
    if (a == b || (a >> shift) != (b >> shift)) return;  // not a witness of a mixed group
    // first witness of its group?  walk back to the group start; an earlier key change inside the group = no
    uint64_t j = i - 1;
    while (j > 0) {
        const uint64_t c = keys[j - 1];
        if ((c >> shift) != (b >> shift)) break;
        if (c != a) return;
        --j;
    }
    list[atomicAdd(&state[0], 1u)] = (uint32_t)j;
}

// one warp per listed group: the group is staged in shared memory, every element counts the elements that must
// precede it (smaller key, or equal key and earlier position) and is written to that rank
__global__ void __launch_bounds__(CF_WARPS * 32) collision_fix_kernel(uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, uint64_t n,
                                                                      int shift, const uint32_t *__restrict__ list,
                                                                      uint32_t *__restrict__ state) {
    __shared__ uint64_t sk[CF_WARPS][CF_MAX_GROUP];
    __shared__ uint32_t sv[CF_WARPS][CF_MAX_GROUP];
    const uint32_t n_groups = state[0];
    const unsigned lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (uint32_t g = blockIdx.x * CF_WARPS + wp; g < n_groups; g += gridDim.x * CF_WARPS) {
        const uint64_t b0 = list[g];
        const uint64_t pre = keys[b0] >> shift;
        uint32_t m = 0;
        for (;;) {
            const uint64_t idx = b0 + m + lane;
            const uint64_t kx = idx < n ? keys[idx] : 0ull;
            const bool ok = idx < n && (kx >> shift) == pre;
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            const uint32_t cnt = bal == 0xffffffffu ? 32u : (uint32_t)__ffs(~bal) - 1u;  // leading lanes still inside the group
            if (lane < cnt && m + lane < CF_MAX_GROUP) sk[wp][m + lane] = kx, sv[wp][m + lane] = vals[idx];
            m += cnt;
            if (cnt < 32u || m > CF_MAX_GROUP) break;
        }
        if (m > CF_MAX_GROUP) {
            if (lane == 0) state[1] = 1u;
            continue;
        }
        __syncwarp();
        for (uint32_t e = lane; e < m; e += 32) {
            const uint64_t ke = sk[wp][e];
            uint32_t r = 0;
            for (uint32_t f = 0; f < m; ++f) {
                const uint64_t kf = sk[wp][f];
                r += (kf < ke || (kf == ke && f < e)) ? 1u : 0u;
            }
            keys[b0 + r] = ke, vals[b0 + r] = sv[wp][e];
        }
        __syncwarp();
    }
}

void launch_collision_fix(uint64_t *keys, uint32_t *vals, uint64_t n, int shift, uint32_t *list, uint32_t *state, cudaStream_t s,
                          LaunchCounter *lc) {
    if (n < 2) return;
    collision_find_kernel<<<(unsigned)ceil_div_u64(n, 256), 256, 0, s>>>(keys, n, shift, list, state);
    collision_fix_kernel<<<148 * 8, CF_WARPS * 32, 0, s>>>(keys, vals, n, shift, list, state);
    if (lc) lc->total += 2;
}


__global__ void unpermute_pairs_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
                                       uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = vals[i];
    keys_out[v] = keys[i], vals_out[v] = v;
}
void launch_unpermute_pairs(const uint64_t *keys, const uint32_t *vals, uint64_t n, uint64_t *keys_out, uint32_t *vals_out, cudaStream_t s,
                            LaunchCounter *lc) {
    if (n == 0) return;
    unpermute_pairs_kernel<<<(unsigned)ceil_div_u64(n, 256), 256, 0, s>>>(keys, vals, n, keys_out, vals_out);
    if (lc) lc->total++;
}

}  // namespace hlo
