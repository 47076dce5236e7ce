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
// seg_start[c] = first element of segment c, seg_tile[c] = its first tile.  Both live on the device, computed from the
// per-read minimizer offsets: the host launches an upper bound of tiles and never learns the element count.
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
__device__ __forceinline__ RsTile rs_tile(const RsSeg &g) {
    RsTile t;
    const uint32_t b = blockIdx.x;
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

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t *__restrict__ keys, int shift,
                                                             uint32_t *__restrict__ hist, RsSeg seg) {
    __shared__ uint32_t h[RS_BINS];
    const RsTile t = rs_tile(seg);
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
                                                                int shift, const uint32_t *__restrict__ hist_scanned, RsSeg seg) {
    __shared__ uint32_t wcnt[RS_WARPS][RS_BINS];
    __shared__ uint32_t tile_start[RS_BINS];
    __shared__ uint32_t gbase[RS_BINS];
    __shared__ uint32_t sm[16];
    __shared__ uint64_t stage_k[RS_TILE];
    __shared__ uint32_t stage_v[RS_TILE];
    const unsigned lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const RsTile tl = rs_tile(seg);
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
            HLO_DBG_ASSERT(p < tile_n);
            stage_k[p] = key[j];
            stage_v[p] = val[j];
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tile_n; i += RS_THREADS) {
        const uint64_t kx = stage_k[i];
        const uint32_t d = (uint32_t)(kx >> shift) & (RS_BINS - 1);
        const uint64_t dst = (uint64_t)gbase[d] + (i - tile_start[d]);
        HLO_DBG_ASSERT(dst >= seg.seg_start[0] && dst < seg.seg_start[seg.n_seg]);
        keys_out[dst] = kx;
        vals_out[dst] = stage_v[i];
    }
}

static uint64_t seg_tile_bound(uint64_t n, uint32_t n_seg) { return ceil_div_u64(n, RS_TILE) + n_seg; }

size_t sort_segmented_tmp_words(uint64_t n, uint32_t n_seg) {
    const uint64_t hist_n = (uint64_t)RS_BINS * seg_tile_bound(n, n_seg);
    return (size_t)(((hist_n + 63) & ~63ull) + scan_tmp_words(hist_n) + 2 * (((uint64_t)n_seg + 64) & ~63ull));
}

int launch_radix_sort_segmented(uint64_t *keys[2], uint32_t *vals[2], uint64_t n, int begin_bit, int end_bit, const uint32_t *elem_off,
                                const uint32_t *seg_off, uint32_t n_seg, uint32_t *tmp, cudaStream_t s, LaunchCounter *lc) {
    int cur = 0;
    if (n == 0 || begin_bit >= end_bit) return cur;  // n: an upper bound of the element count (the real one is on the device)
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
        rs_hist_kernel<<<t_bound, RS_THREADS, 0, s>>>(keys[cur], shift, hist, g);
        if (lc) lc->total++;
        launch_exclusive_scan(hist, hist, hist_n, scan_tmp, nullptr, s, lc);
        rs_scatter_kernel<<<t_bound, RS_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], shift, hist, g);
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

__global__ void collision_find_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ n_ptr, int shift,
                                      uint32_t *__restrict__ list, uint32_t *__restrict__ state) {
    const uint64_t n = *n_ptr;
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
__global__ void __launch_bounds__(CF_WARPS * 32) collision_fix_kernel(uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                                                      const uint32_t *__restrict__ n_ptr, int shift,
                                                                      const uint32_t *__restrict__ list, uint32_t *__restrict__ state) {
    __shared__ uint64_t sk[CF_WARPS][CF_MAX_GROUP];
    __shared__ uint32_t sv[CF_WARPS][CF_MAX_GROUP];
    const uint64_t n = *n_ptr;
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
            HLO_DBG_ASSERT(r < m && b0 + r < n);
            keys[b0 + r] = ke, vals[b0 + r] = sv[wp][e];
        }
        __syncwarp();
    }
}

void launch_collision_fix(uint64_t *keys, uint32_t *vals, uint64_t n_cap, const uint32_t *n_ptr, int shift, uint32_t *list, uint32_t *state,
                          cudaStream_t s, LaunchCounter *lc) {
    if (n_cap < 2) return;
    collision_find_kernel<<<(unsigned)ceil_div_u64(n_cap, 256), 256, 0, s>>>(keys, n_ptr, shift, list, state);
    collision_fix_kernel<<<148 * 8, CF_WARPS * 32, 0, s>>>(keys, vals, n_ptr, shift, list, state);
    if (lc) lc->total += 2;
}


__global__ void unpermute_pairs_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ n_ptr,
                                       uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_ptr) return;
    const uint32_t v = vals[i];
    keys_out[v] = keys[i], vals_out[v] = v;
}
void launch_unpermute_pairs(const uint64_t *keys, const uint32_t *vals, uint64_t n_cap, const uint32_t *n_ptr, uint64_t *keys_out,
                            uint32_t *vals_out, cudaStream_t s, LaunchCounter *lc) {
    if (n_cap == 0) return;
    unpermute_pairs_kernel<<<(unsigned)ceil_div_u64(n_cap, 256), 256, 0, s>>>(keys, vals, n_ptr, keys_out, vals_out);
    if (lc) lc->total++;
}

}  // namespace hlo
