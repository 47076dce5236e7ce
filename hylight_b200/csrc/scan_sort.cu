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

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, int shift,
                                                             uint32_t n_tiles, uint32_t *__restrict__ hist) {
    __shared__ uint32_t h[RS_BINS];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint64_t i = base + (uint64_t)j * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS, 4) rs_scatter_kernel(const uint64_t *__restrict__ keys_in,
                                                                const uint32_t *__restrict__ vals_in,
                                                                uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                                uint64_t n, int shift, uint32_t n_tiles,
                                                                const uint32_t *__restrict__ hist_scanned) {
    __shared__ uint32_t wcnt[RS_WARPS][RS_BINS];
    __shared__ uint32_t tile_start[RS_BINS];
    __shared__ uint32_t gbase[RS_BINS];
    __shared__ uint32_t sm[16];
    __shared__ uint64_t stage_k[RS_TILE];
    __shared__ uint32_t stage_v[RS_TILE];
    const unsigned lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
    const uint32_t tile_n = (uint32_t)((n - base) < (uint64_t)RS_TILE ? (n - base) : (uint64_t)RS_TILE);
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
        gbase[d] = hist_scanned[(uint64_t)d * n_tiles + blockIdx.x];
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
        rs_hist_kernel<<<n_tiles, RS_THREADS, 0, s>>>(keys[cur], n, shift, n_tiles, hist);
        if (lc) lc->total++;
        launch_exclusive_scan(hist, hist, hist_n, scan_tmp, nullptr, s, lc);
        rs_scatter_kernel<<<n_tiles, RS_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift, n_tiles,
                                                         hist);
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
__global__ void fill_kernel(uint32_t *v, uint32_t value, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = value;
}
void launch_fill_u32(uint32_t *v, uint32_t value, uint64_t n, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(v, value, n);
    if (lc) lc->total++;
}

}  // namespace hlo
