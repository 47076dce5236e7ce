// pairgen.cu — seed-bucket pair-and-diagonal generator = minimap2's seeding + "-r 0" chaining.
//
// Restates (oracle/hlo_oracle.c:orc_candidates; minimap2 is absent from /root/reference, call site
// polyte.tune_params.py:510 "--sr -X ... -m 30 -n 2 -r 0"):
//   * index: minimizers bucketed by (cluster, hash); a bucket with more than mid_occ entries seeds nothing;
//   * -X: no self hits, a seed (q -> t) is skipped when strcmp(qname, tname) > 0;
//   * anchors of one (t, strand, diagonal) form one chain (bw = 0); chain score = k + sum(min(k, gap));
//     kept iff #anchors >= -n and score >= -m, which for -m <= 2k and -n <= 2 is
//     (max anchor pos - min anchor pos) >= max(m - k, n - 1, 0)  (DESIGN.md §pairgen).
//
// The minimizer arrays are in name order (sketch.cu), and the hash sort is stable, so every bucket lists
// its reads in strcmp order of their names: the seeds minimap2 -X keeps for a query minimizer are exactly
// the bucket entries AFTER its own entry (minus further entries of the same read).  No name comparison
// and no wasted half of the bucket.
//
// pairgen_group_kernel: 16 lanes per query read (2 reads per warp, 8 per block) fold the read's seed hits into a
// shared-memory hash table keyed by a 32-bit (t - cluster base, strand, diagonal) that keeps min/max anchor position —
// the warp-level dedup of the north star; the details stand above the kernel.
// Reads whose table overflows run again with tables of twice the size (ctx.cu:enqueue_pairgen).
// A read that overflows even 2048 slots, or a batch whose keys do not fit 32 bits, is done exactly by
// pairgen_generic_kernel with a 64-bit-keyed table in HBM.
// Algorithmic bytes: 8 B per seed hit in + 16 B per candidate out (the kernel itself reads 4 B per hit).
#include <type_traits>

#include "kernels.cuh"

namespace hlo {

#define HLO_EMPTY_KEY 0xffffffffffffffffULL
#define HLO_EMPTY32 0xffffffffu

__global__ void hash_keys_kernel(const uint64_t *__restrict__ x, const uint2 *__restrict__ mz_rr,
                                 const uint32_t *__restrict__ cidx, int hash_bits, const uint32_t *__restrict__ n_ptr,
                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *n_ptr) keys[i] = (x[i] >> 8) | (uint64_t)cidx[mz_rr[i].x] << hash_bits, vals[i] = (uint32_t)i;
}
void launch_hash_keys(const uint64_t *mz_x, const uint2 *mz_rr, const uint32_t *cidx, int hash_bits, uint64_t n_cap, const uint32_t *n_ptr,
                      uint64_t *keys, uint32_t *vals, cudaStream_t s, LaunchCounter *lc) {
    if (n_cap == 0) return;
    hash_keys_kernel<<<(unsigned)((n_cap + 255) / 256), 256, 0, s>>>(mz_x, mz_rr, cidx, hash_bits, n_ptr, keys, vals);
    if (lc) lc->total++;
}

// The bucket index over the (cluster, hash)-grouped minimizers, in two modes of one tile kernel:
//   EMIT = false: occurrence histogram.  The first entry of every bucket adds one to hist[cluster][min(size, bins - 1)]:
//              the input of minimap2's data-derived occurrence cut-off (mm_idx_cal_max_occ: the (1 - 2e-4) quantile of the
//              occurrences of the distinct minimizers, + 1, clamped to [10, 10^6]; midocc_kernel below).  Skipped when the
//              caller fixes mid_occ.
//   EMIT = true:  hits[i]  = the i-th entry (one 8-byte gather from the read-order array): (read, pos<<1|strand), or, with
//              dg_bits > 0, the 32-bit word (read - cluster base) << (dg_bits + 1) | strand << dg_bits | pos that
//              pairgen_group_kernel turns into its table key with one XOR and one add;
//              range[m] = the hits minimizer m seeds: the part of its bucket behind its own entry (i + 1 .. bucket end),
//              or nothing when the bucket holds more than the cluster's cut-off (scatter to read order).
// A tile of 1024 entries finds its bucket boundaries from warp ballots of "key differs from the previous key";
// only the two buckets that stick out of the tile need a look at the neighbouring tiles' keys, and that search
// stops after `reach` entries (EMIT: the cut-off, a longer bucket is dropped anyway; histogram: the clamp bin).
// The keys are cluster-major, so the gather and the scatter stay inside one cluster's few MB.
constexpr int IX_THREADS = 256;
constexpr int IX_ITEMS = 4;
constexpr int IX_TILE = IX_THREADS * IX_ITEMS;
constexpr int IX_RUNS = IX_TILE / 32;
constexpr uint32_t IX_LO = 1024, IX_REP = 8;  // occurrence histogram: [cluster][IX_REP][IX_LO] below IX_LO, [cluster][bins] above

template <bool EMIT>
__global__ void __launch_bounds__(IX_THREADS) bucket_index_kernel(const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ svals,
                                                                  const uint2 *__restrict__ mz_rr, const uint32_t *__restrict__ n_ptr,
                                                                  int hash_bits, const uint32_t *__restrict__ midocc, uint32_t *__restrict__ hist,
                                                                  uint32_t *__restrict__ hist_hi, uint32_t *hist_max, uint32_t bins,
                                                                  const uint32_t *__restrict__ cluster_off, int dg_bits, void *__restrict__ hits_out,
                                                                  uint2 *__restrict__ range) {
    __shared__ int run_first[IX_RUNS], run_last[IX_RUNS];     // first / last bucket start inside a run of 32 entries (tile-relative)
    __shared__ int next_after[IX_RUNS], last_before[IX_RUNS]; // first start in a later run / last start in an earlier run
    __shared__ uint32_t tile_prev, tile_next;                 // start of the tile's first bucket, end of its last (global)
    const uint64_t n = *n_ptr;
    const unsigned lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const uint64_t base = (uint64_t)blockIdx.x * IX_TILE;
    if (base >= n) return;
    const uint32_t tn = (uint32_t)((n - base) < (uint64_t)IX_TILE ? (n - base) : (uint64_t)IX_TILE);
    unsigned bal[IX_ITEMS];
    uint32_t idx[IX_ITEMS];
    uint32_t clu[IX_ITEMS];
#pragma unroll
    for (int j = 0; j < IX_ITEMS; ++j) {
        const uint32_t r = (uint32_t)j * (IX_THREADS / 32) + wp, e = r * 32 + lane;
        const uint64_t i = base + e;
        const bool ok = e < tn;
        bool flag = false;
        idx[j] = 0, clu[j] = 0;
        if (ok) {
            const uint64_t kx = skeys[i];
            flag = i == 0 || kx != skeys[i - 1];
            clu[j] = (uint32_t)(kx >> hash_bits);
            if (EMIT) {
                idx[j] = svals[i];
                const uint2 rr = mz_rr[idx[j]];
                if (dg_bits) reinterpret_cast<uint32_t *>(hits_out)[i] = (rr.x - cluster_off[clu[j]]) << (dg_bits + 1) | (rr.y & 1u) << dg_bits | rr.y >> 1;
                else reinterpret_cast<uint2 *>(hits_out)[i] = rr;
            }
        }
        bal[j] = __ballot_sync(FULL, flag);
        if (lane == 0) {
            run_first[r] = bal[j] ? (int)(r * 32) + __ffs(bal[j]) - 1 : INT_MAX;
            run_last[r] = bal[j] ? (int)(r * 32) + 31 - __clz(bal[j]) : -1;
        }
    }
    __syncthreads();
    if (wp == 0) {  // IX_RUNS == 32: one lane per run
        int nf = run_first[lane], lb = run_last[lane];
        // inclusive suffix minimum / prefix maximum, then shift by one run
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int a = __shfl_down_sync(FULL, nf, d), c = __shfl_up_sync(FULL, lb, d);
            if (lane + d < 32) nf = min(nf, a);
            if ((int)lane >= d) lb = max(lb, c);
        }
        const int nf_x = __shfl_down_sync(FULL, nf, 1), lb_x = __shfl_up_sync(FULL, lb, 1);
        next_after[lane] = lane == 31 ? INT_MAX : nf_x;
        last_before[lane] = lane == 0 ? -1 : lb_x;
    } else if (wp == 1) {  // end of the bucket the tile's last entry belongs to
        uint64_t end = base + tn;
        if (end < n) {
            const uint64_t key_last = skeys[end - 1];
            const uint64_t reach = EMIT ? (uint64_t)midocc[(uint32_t)(key_last >> hash_bits)] : (uint64_t)bins;
            for (;;) {
                const uint64_t i = end + lane;
                const bool differs = i >= n || skeys[i] != key_last;
                const unsigned b = __ballot_sync(FULL, differs);
                if (b) {
                    end += (uint64_t)(__ffs(b) - 1);
                    break;
                }
                end += 32;
                if (end - (base + tn) > reach) break;  // already larger than the cut-off (or the clamp bin): no need to know more
            }
        }
        if (lane == 0) tile_next = (uint32_t)end;
    } else if (wp == 2 && EMIT) {  // start of the bucket the tile's first entry belongs to
        uint64_t start = base;
        if (base > 0 && tn > 0) {
            const uint64_t key_first = skeys[base];
            const uint64_t reach = (uint64_t)midocc[(uint32_t)(key_first >> hash_bits)];
            for (;;) {
                // lane l looks at entry start - 1 - l
                const bool differs = start < (uint64_t)lane + 1 || skeys[start - 1 - lane] != key_first;
                const unsigned b = __ballot_sync(FULL, differs);
                if (b) {
                    start -= (uint64_t)(__ffs(b) - 1);
                    break;
                }
                start -= 32;
                if (base - start > reach) break;
            }
        }
        if (lane == 0) tile_prev = (uint32_t)start;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IX_ITEMS; ++j) {
        const uint32_t r = (uint32_t)j * (IX_THREADS / 32) + wp, e = r * 32 + lane;
        if (e >= tn) continue;
        const unsigned above = lane == 31 ? 0u : bal[j] >> (lane + 1);
        const int nb = above ? (int)e + __ffs(above) : next_after[r];
        const uint32_t en = nb == INT_MAX ? tile_next : (uint32_t)(base + (uint32_t)nb);
        if (EMIT) {
            const unsigned upto = bal[j] & (0xffffffffu >> (31 - lane));
            const int pb = upto ? (int)(r * 32) + 31 - __clz(upto) : last_before[r];
            const uint32_t st = pb < 0 ? tile_prev : (uint32_t)(base + (uint32_t)pb);
            HLO_DBG_ASSERT(idx[j] < n && en <= n && st <= base + e && en > base + e);
            range[idx[j]] = (en - st) > midocc[clu[j]] ? make_uint2(0u, 0u) : make_uint2((uint32_t)(base + e) + 1u, en);
        } else if ((bal[j] >> lane) & 1u) {  // first entry of its bucket: one count for the bucket's size
            // sizes below IX_LO (nearly all: a bucket holds about one entry per read covering the minimizer) go to one of
            // IX_REP replicas of the low part of the histogram — a cluster's buckets crowd on a few dozen sizes, and the
            // replicas spread what would be same-address atomics
            uint32_t sz = en - (uint32_t)(base + e);
            sz = sz < bins - 1u ? sz : bins - 1u;
            const uint32_t lo = bins < IX_LO ? bins : IX_LO;
            if (sz < lo) atomicAdd(&hist[((size_t)clu[j] * IX_REP + (blockIdx.x & (IX_REP - 1))) * lo + sz], 1u);
            else atomicAdd(&hist_hi[(size_t)clu[j] * bins + sz], 1u);
            if (sz > hist_max[clu[j]]) atomicMax(&hist_max[clu[j]], sz);  // (the plain read first: nearly every bucket is below the maximum already)
        }
    }
}

// one block per cluster: mid_occ = clamp(k-th smallest occurrence + 1, 10, 10^6), k = (uint32_t)((1 - 2e-4f) * distinct)
// (minimap2 mm_idx_cal_max_occ / mm_mapopt_update; the float constant is promoted to double there as here)
__global__ void __launch_bounds__(256) midocc_kernel(const uint32_t *__restrict__ hist, const uint32_t *__restrict__ hist_hi,
                                                     const uint32_t *__restrict__ hist_max, uint32_t bins_all, uint32_t n_clusters, int fixed,
                                                     uint32_t *__restrict__ midocc, unsigned long long *__restrict__ counters) {
    __shared__ uint32_t part[256];
    __shared__ uint32_t s_thr;
    const uint32_t c = blockIdx.x;
    if (c >= n_clusters) return;
    uint32_t thr;
    if (fixed > 0) {
        thr = (uint32_t)fixed;
    } else {
        // distinct minimizers of cluster c that occur v times
        const uint32_t lo = bins_all < IX_LO ? bins_all : IX_LO;
        // only sizes up to the cluster's largest bucket are occupied: the walk below covers those
        const uint32_t bins = min(bins_all, hist_max[c] + 1u);
        auto h = [&](uint32_t v) {
            if (v >= lo) return hist_hi[(size_t)c * bins_all + v];
            uint32_t t = 0;
            for (uint32_t r = 0; r < IX_REP; ++r) t += hist[((size_t)c * IX_REP + r) * lo + v];
            return t;
        };
        const uint32_t per = (bins + 255u) / 256u, b0 = threadIdx.x * per, b1 = min(bins, b0 + per);
        uint32_t sum = 0;
        for (uint32_t b = b0; b < b1; ++b) sum += h(b);
        part[threadIdx.x] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t n = 0;
            for (int t = 0; t < 256; ++t) n += part[t];
            uint32_t v = 0;
            if (n) {
                const uint32_t k = (uint32_t)((1. - (double)2e-4f) * (double)n);
                uint32_t cum = 0;
                int t = 0;
                while (t < 255 && cum + part[t] <= k) cum += part[t], ++t;
                uint32_t b = (uint32_t)t * per;
                const uint32_t be = min(bins, b + per);
                while (b + 1 < be && cum + h(b) <= k) cum += h(b), ++b;
                v = b;
                if (v == bins_all - 1u) atomicOr(&counters[C_FLAGS], HLO_FLAG_MIDOCC_AMBIGUOUS);  // the true occurrence is not known
            }
            long long m = (long long)v + 1;
            if (m < 10) m = 10;
            if (m > 1000000) m = 1000000;
            s_thr = (uint32_t)m;
        }
        __syncthreads();
        thr = s_thr;
    }
    if (threadIdx.x == 0) {
        midocc[c] = thr;
        atomicMin(&counters[C_MIDOCC_LO], (unsigned long long)thr);
        atomicMax(&counters[C_MIDOCC_HI], (unsigned long long)thr);
    }
}

size_t bucket_hist_words(uint32_t n_clusters, uint32_t bins) {  // low replicas | high part | largest size per cluster
    return (size_t)n_clusters * ((size_t)IX_REP * (bins < IX_LO ? bins : IX_LO) + bins + 1);
}

void launch_bucket_index(const uint64_t *skeys, const uint32_t *svals, const uint2 *mz_rr, uint64_t n_cap, const uint32_t *n_ptr, int hash_bits,
                         int fixed_mid_occ, uint32_t n_clusters, uint32_t *midocc, uint32_t *hist, uint32_t bins, unsigned long long *counters,
                         const uint32_t *cluster_off, int dg_bits, void *hits, uint2 *range, cudaStream_t s, LaunchCounter *lc) {
    if (n_cap == 0 || n_clusters == 0) return;
    const unsigned grid = (unsigned)((n_cap + IX_TILE - 1) / IX_TILE);
    uint32_t *hist_hi = hist ? hist + (size_t)n_clusters * IX_REP * (bins < IX_LO ? bins : IX_LO) : nullptr;
    uint32_t *hist_max = hist ? hist_hi + (size_t)n_clusters * bins : nullptr;
    if (fixed_mid_occ <= 0) {
        cudaMemsetAsync(hist, 0, bucket_hist_words(n_clusters, bins) * sizeof(uint32_t), s);
        bucket_index_kernel<false><<<grid, IX_THREADS, 0, s>>>(skeys, svals, mz_rr, n_ptr, hash_bits, nullptr, hist, hist_hi, hist_max, bins,
                                                               cluster_off, dg_bits, hits, range);
        if (lc) lc->total++;
    }
    midocc_kernel<<<n_clusters, 256, 0, s>>>(hist, hist_hi, hist_max, bins, n_clusters, fixed_mid_occ, midocc, counters);
    bucket_index_kernel<true><<<grid, IX_THREADS, 0, s>>>(skeys, svals, mz_rr, n_ptr, hash_bits, midocc, hist, hist_hi, hist_max, bins, cluster_off,
                                                          dg_bits, hits, range);
    if (lc) lc->total += 2;
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_cand(const PairgenArgs &a, unsigned long long p, uint32_t q, uint32_t t, uint32_t strand,
                                          int diag, uint32_t mn, uint32_t mx, uint32_t ql, uint32_t tl, uint32_t amb) {
    if (p >= a.cand_cap) return;  // only counted; the host grows the buffer and reruns
    const int a_start = (int)mn - a.k + 1, a_end = (int)mx + 1;
    uint4 o;
    if (a.external) {
        o = make_uint4(q, t, (uint32_t)diag, HLO_CAND_PACK(strand, a_start, a_end));
    } else {
        uint32_t far = 0;
        const CandI c = make_candi(q, t, diag, strand, ql, tl, a_start, a_end, amb, &far);
        o = make_uint4(c.q, c.t, c.z, c.w);
        if (c.w & HLO_CAND_FAR) a.cand_far[p] = far;  // (rare: an anchor more than 127 bases from its overlap end)
    }
    reinterpret_cast<uint4 *>(a.cands)[p] = o;
}

// ------------------------------------------------------------------------------------------------
// fast path: PG_G lanes per query read, PG_RPW reads per warp in LOCKSTEP (every loop bound is the maximum over
// the warp's reads, so the warp never diverges into per-read paths), table in shared memory.
//
// What one seed hit costs is the whole story of this kernel (round 1: issue-bound at 8.9 warp instructions per hit; now
// 5.4, and the shared-memory data pipe, 81 % of its wavefront rate, is the first limit with issue at 76 % right behind;
// HBM 7 %), so the inner loop is built to be short:
//   * a hit is ONE 32-bit word W = (t - cluster base) << (dg_bits + 1) | strand(t) << dg_bits | pos(t), laid down by
//     the bucket index in the field layout of the table key: key = (W ^ strand(q) << dg_bits) + max_len - qp, where
//     qp (the query anchor position on the hit's strand) is one of two per-minimizer values — five instructions
//     instead of the unpack / select / shift / or sequence on (read, pos<<1|strand) pairs, and half the bytes;
//   * the read's bucket suffixes are expanded into a list of CHUNKS (first hit, #hits <= PG_G, strand and position of
//     the query minimizer) in shared memory, PG_RPL * PG_G suffixes at a time; the loop over chunks is a broadcast
//     8-byte load, no "advance to the next range" logic, unrolled by two with the hits loaded one chunk ahead;
//   * a slot's value is (min anchor position, 0xffff - max anchor position) in two 16-bit halves: one VIMNMX.U16x2
//     between a load and a store updates both;
//   * nothing is recorded while inserting: after the last chunk the lanes scan the value array once (a fresh slot's
//     value fails the span test by itself), list the qualifying slots through a shared counter (positions from ballots
//     were measured slower: more instructions for fewer wavefronts), and emit from the list.
// Within one chunk all keys are distinct (one minimizer, one bucket), so only the key claim is an atomic (ATOMS.CAS);
// the value update is a plain read-modify-write ordered by a __syncwarp per chunk.
constexpr int PG_THREADS = 128;
#ifndef HLO_PG_G
#define HLO_PG_G 16
#endif
#ifndef HLO_PG_MINB
#define HLO_PG_MINB 10  // 48 registers: 10 blocks per SM (measured: 9 blocks 1.59 ms, 10 blocks 1.54 ms, 40 registers 1.60 ms; left to
                        // itself ptxas takes 86 registers and the kernel 2.34 ms)
#endif
constexpr int PG_G = HLO_PG_G;                 // lanes per read
constexpr unsigned PG_GMASK = PG_G == 32 ? 0xffffffffu : ((1u << PG_G) - 1u);
constexpr int PG_RPW = 32 / PG_G;              // reads per warp
constexpr int PG_GROUPS = PG_THREADS / PG_G;   // reads in flight per block
constexpr int PG_PAD = 8;                      // words between the groups' areas (bank spread; word 0 = the list counter)
constexpr int PG_RPL = 2;                      // bucket suffixes per lane and stage
constexpr int PG_CAP = 64;                     // chunk entries per read and stage
// words per read: keys | values | chunk entries (8 B), later the list of qualifying slots (2 B) | pad
__host__ __device__ inline uint32_t pairgen_aux_words(int slots_log2) {
    const uint32_t half = (1u << slots_log2) / 2u, chunks = 2u * (PG_CAP + 2) + 4u;  // (+2 entries: the loop looks two chunks ahead)
    return half > chunks ? half : chunks;
}
__host__ __device__ inline uint32_t pairgen_group_words(int slots_log2) {
    return (2u << slots_log2) + pairgen_aux_words(slots_log2) + PG_PAD;
}

__global__ void __launch_bounds__(PG_THREADS, HLO_PG_MINB) pairgen_group_kernel(PairgenArgs a, int slots_log2) {
    extern __shared__ __align__(16) uint32_t pg_smem[];
    const uint32_t S = 1u << slots_log2, smask = S - 1u;
    const uint32_t g = threadIdx.x / PG_G, l = threadIdx.x % PG_G;
    uint32_t *keys = pg_smem + (size_t)g * pairgen_group_words(slots_log2);  // (t - cluster base, strand, diagonal)
    uint32_t *vals = keys + S;                                               // min | (0xffff - max) << 16 anchor position
    uint2 *stg = reinterpret_cast<uint2 *>(keys + 2 * S);                    // chunks: (first hit, #hits | strand << 5 | pos << 6)
    uint16_t *list = reinterpret_cast<uint16_t *>(keys + 2 * S);             // qualifying slots (after the last chunk)
    uint32_t *cnt = keys + 2 * S + pairgen_aux_words(slots_log2);            // their number
    const unsigned lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const unsigned gshift = lane & ~(unsigned)(PG_G - 1);
    const int kshift = a.dg_bits + 1;
    const uint32_t sbit = 1u << a.dg_bits;
    const uint32_t max_len = a.reads.max_len;
    const uint32_t span_max = 0xffffu - (uint32_t)a.gap_min;  // a slot qualifies iff min + (0xffff - max) <= span_max
    // the listed reads (x indices; their number is a device counter left by the previous pass), or all of them
    uint32_t n_work = a.reads.n_reads;
    if (a.work) {
        const unsigned long long nw = *a.n_work_ptr;
        n_work = nw < (unsigned long long)a.overflow_cap ? (uint32_t)nw : a.overflow_cap;
    }
    unsigned long long nh_total = 0;
    for (;;) {
        unsigned long long qq = 0;
        if (lane == 0) qq = atomicAdd(a.work_counter, (unsigned long long)PG_RPW);
        qq = __shfl_sync(FULL, qq, 0);
        if (qq >= n_work) break;
        const uint32_t wi = (uint32_t)qq + lane / PG_G;
        const bool have_read = wi < n_work;
        const uint32_t xq = !have_read ? 0u : a.work ? __ldg(a.work + wi) : wi;
        uint32_t m0 = 0, m1 = 0, q = 0, lq = 0, cb = 0;
        if (have_read) {
            m0 = __ldg(a.mz_off + xq), m1 = __ldg(a.mz_off + xq + 1);
            q = __ldg(a.reads.perm + xq);
            lq = __ldg(a.reads.len + q);
            cb = __ldg(a.reads.cbase + q);
        }
        const uint32_t ql = lq & HLO_LEN_MASK;
        const uint32_t nm = m1 - m0;
        if (__reduce_max_sync(FULL, nm) == 0) continue;
        {  // keys = empty, values = (min 0xffff, max 0): the same all-ones pattern
            uint4 *t4 = reinterpret_cast<uint4 *>(keys);
            const uint4 e4 = make_uint4(HLO_EMPTY32, HLO_EMPTY32, HLO_EMPTY32, HLO_EMPTY32);
            for (uint32_t i = l; i < S / 2; i += PG_G) t4[i] = e4;
            if (l == 0) *cnt = 0u;
        }
        __syncwarp();
        const uint32_t qrel = q - cb;
        const uint32_t w_self = qrel << kshift;       // a hit word that names the query itself: skipped (-X), also stands for "no hit"
        const uint32_t KQ = ql + (uint32_t)a.k - 2u;  // reverse-strand anchor position = KQ - pos
        uint32_t nh = 0;
        uint32_t ovf = 0;
        uint32_t mb = 0, skip = 0;  // minimizers staged so far; hits already taken from the first unstaged suffix
        while (__any_sync(FULL, mb < nm)) {
            // ---- stage: PG_RPL consecutive suffixes per lane -> chunk entries, as many suffixes as fit PG_CAP chunks
            uint2 rg[PG_RPL];
            uint32_t yy[PG_RPL], cc[PG_RPL];
            uint32_t incl = 0;
#pragma unroll
            for (int rr = 0; rr < PG_RPL; ++rr) {
                const uint32_t idx = mb + l * PG_RPL + rr;
                rg[rr] = make_uint2(0u, 0u), yy[rr] = 0u;
                if (idx < nm) rg[rr] = __ldg(a.range + m0 + idx), yy[rr] = __ldg(&a.mz_rr[m0 + idx].y);
                if (rr == 0 && l == 0) rg[rr].x += skip;
                cc[rr] = (rg[rr].y - rg[rr].x + PG_G - 1) / PG_G;
                incl += cc[rr];
            }
            const uint32_t mine = incl;
#pragma unroll
            for (int d = 1; d < PG_G; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, incl, d, PG_G);
                if ((int)l >= d) incl += o;
            }
            // the chunk counts are a running sum in suffix order, so "fits" is a prefix property
            uint32_t pos[PG_RPL];
            bool fit[PG_RPL];
            uint32_t n_fit = 0, upto = 0, run = incl - mine;
#pragma unroll
            for (int rr = 0; rr < PG_RPL; ++rr) {
                pos[rr] = run;
                run += cc[rr];
                fit[rr] = run <= (uint32_t)PG_CAP;
                n_fit += __popc((__ballot_sync(FULL, fit[rr]) >> gshift) & PG_GMASK);
                if (fit[rr]) upto = run;
            }
#pragma unroll
            for (int d = PG_G / 2; d > 0; d >>= 1) upto = max(upto, __shfl_xor_sync(FULL, upto, d));
            const uint32_t fx = __shfl_sync(FULL, rg[0].x, 0, PG_G), fy = __shfl_sync(FULL, yy[0], 0, PG_G);
            uint32_t n_chunks = upto;
            if (n_fit == 0) {  // the first suffix alone holds more than PG_CAP chunks: PG_CAP full chunks of it now, the rest next time
                const uint32_t meta = (uint32_t)PG_G | (fy & 1u) << 5 | (fy >> 1) << 6;
                for (uint32_t j = l; j < (uint32_t)PG_CAP; j += PG_G) stg[j] = make_uint2(fx + j * PG_G, meta);
                skip += (uint32_t)PG_CAP * PG_G;
                n_chunks = PG_CAP;
            } else {
#pragma unroll
                for (int rr = 0; rr < PG_RPL; ++rr) {
                    if (!fit[rr]) continue;
                    const uint32_t meta = (yy[rr] & 1u) << 5 | (yy[rr] >> 1) << 6;
                    for (uint32_t j = 0, x = rg[rr].x; j < cc[rr]; ++j, x += PG_G) {
                        HLO_DBG_ASSERT(pos[rr] + j < (uint32_t)PG_CAP);
                        stg[pos[rr] + j] = make_uint2(x, min((uint32_t)PG_G, rg[rr].y - x) | meta);
                    }
                }
                skip = 0;
                mb += n_fit;
            }
            // empty entries up to the end of the area: the chunk loop runs to the longest list of the warp and looks two ahead
            for (uint32_t j = n_chunks + l; j < (uint32_t)PG_CAP + 2u; j += PG_G) stg[j] = make_uint2(0u, 0u);
            __syncwarp();
            const uint32_t trips = __reduce_max_sync(FULL, n_chunks);
            // ---- chunks: PG_G hits of one minimizer each, loaded one chunk ahead
            auto fetch = [&](uint32_t c, uint32_t &meta, uint32_t &W) {
                const unsigned long long e = reinterpret_cast<const unsigned long long *>(stg)[c];  // (one 8-byte broadcast load)
                meta = (uint32_t)(e >> 32);
                W = w_self;
                if (l < (meta & 31u)) W = __ldg(a.hits32 + ((uint32_t)e + l));
            };
            auto fold = [&](uint32_t meta, uint32_t W) {
                if ((W >> kshift) != qrel) {  // == : another minimizer of the query itself (-X: no self hits), or no hit in this lane
                    ++nh;
                    const uint32_t qpos = meta >> 6;
                    const uint32_t Wx = W ^ ((meta & 32u) << (a.dg_bits - 5));  // strand of the pair in the key's strand bit
                    const uint32_t qp = (Wx & sbit) ? KQ - qpos : qpos;          // ql - (qpos + 1 - k) - 1 on the reverse strand
                    const uint32_t key = Wx + max_len - qp;                      // the position field becomes diagonal + max_len
                    const uint32_t upd = 0xffff0000u - qp * 0xffffu;             // qp | (0xffff - qp) << 16
                    uint32_t slot = (key * 0x9E3779B1u) >> (32 - slots_log2);
                    uint32_t left = S, old;  // claim-or-find, at most S probes (a full table: the read is redone with a larger one)
                    while ((old = atomicCAS(&keys[slot], HLO_EMPTY32, key)) != HLO_EMPTY32 && old != key && --left) slot = (slot + 1u) & smask;
                    HLO_DBG_ASSERT(slot < S && (W >> kshift) + cb < a.reads.n_reads && (key & (sbit - 1u)) < 2u * max_len);
                    if (left) vals[slot] = __vminu2(vals[slot], upd);  // no other lane holds this key in this chunk
                    else ovf = 1u;
                }
                __syncwarp();  // orders the value updates of consecutive chunks (they may belong to different minimizers)
            };
            uint32_t mA, wA, mB, wB;
            fetch(0, mA, wA);
            for (uint32_t c = 0; c < trips; c += 2) {
                fetch(c + 1, mB, wB);
                fold(mA, wA);
                fetch(c + 2, mA, wA);
                fold(mB, wB);
            }
            __syncwarp();  // the chunk area is rewritten by the next stage
        }
        // ---- the read's chains: slots whose anchors span at least gap_min
        const bool ovf_any = ((__ballot_sync(FULL, ovf != 0u) >> gshift) & PG_GMASK) != 0u;
        if (ovf_any) {  // redo this read with a larger table, or exactly in the generic kernel (its hits are counted there)
            if (l == 0 && have_read) {
                const unsigned long long p = atomicAdd(a.ovf_counter, 1ULL);
                if (p < a.overflow_cap) a.overflow_reads[p] = xq;
            }
        } else {
            nh_total += nh;
            const uint4 *v4 = reinterpret_cast<const uint4 *>(vals);
            for (uint32_t i = l; i < S / 4; i += PG_G) {
                const uint4 v = v4[i];
                const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((vv[j] >> 16) + (vv[j] & 0xffffu) <= span_max) {
                        const uint32_t p = atomicAdd(cnt, 1u);
                        HLO_DBG_ASSERT(p < S);
                        list[p] = (uint16_t)(i * 4 + j);
                    }
            }
        }
        __syncwarp();
        const uint32_t nq = ovf_any ? 0u : *cnt;
        // reserve space for the warp's reads with one atomic: inclusive scan of the per-read counts
        uint32_t inc = l == 0 ? nq : 0u;
#pragma unroll
        for (int d = PG_G; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, d);
            if ((int)lane >= d) inc += o;
        }
        const uint32_t gsum = __shfl_sync(FULL, inc, gshift);  // counts of the groups up to and including mine
        const uint32_t total = __shfl_sync(FULL, inc, 32 - PG_G);
        unsigned long long base = 0;
        if (lane == 0 && total) base = atomicAdd(a.cand_counter, (unsigned long long)total);
        base = __shfl_sync(FULL, base, 0) + (gsum - nq);
        for (uint32_t j = l; j < nq; j += PG_G) {
            const uint32_t e = list[j];
            const uint32_t key = keys[e], v = vals[e];
            const uint32_t t = cb + (key >> kshift);
            const uint32_t lt = __ldg(a.reads.len + t);
            emit_cand(a, base + j, q, t, (key >> a.dg_bits) & 1u, (int)(key & (sbit - 1u)) - (int)max_len, v & 0xffffu, 0xffffu - (v >> 16), ql,
                      lt & 0x7fffu, (lt | lq) >> 31);
        }
        __syncwarp();
    }
    // seed-hit statistic: one atomic per warp
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nh_total += __shfl_xor_sync(FULL, nh_total, d);
    if (lane == 0 && nh_total) atomicAdd(a.hit_counter, nh_total);
}

// ------------------------------------------------------------------------------------------------
size_t pairgen_smem_bytes(int slots_log2) { return (size_t)PG_GROUPS * pairgen_group_words(slots_log2) * 4; }

bool pairgen_fast_supported(const DevReads &r, int *dg_bits_out) {
    int dg = 5;  // (at least 5: the chunk entries keep the query strand in bit 5 and shift it up to the key's strand bit)
    while ((1u << dg) < 2u * r.max_len + 2u) ++dg;
    int tb = 1;
    while ((1ull << tb) < (unsigned long long)r.max_cluster + 1ull) ++tb;
    if (dg_bits_out) *dg_bits_out = dg;
    return r.max_len <= 32767u && tb + dg + 1 <= 31;  // the top bit stays clear, so no key equals the empty marker
}

void launch_pairgen(const PairgenArgs &a, int slots_log2, cudaStream_t s, LaunchCounter *lc) {
    const uint32_t n_work = a.reads.n_reads;  // (a listed pass usually finds its device-side count at zero and leaves at once)
    if (n_work == 0) return;
    const size_t smem = pairgen_smem_bytes(slots_log2);
    void (*kern)(PairgenArgs, int) = pairgen_group_kernel;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PG_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    uint64_t want = ((uint64_t)n_work + PG_GROUPS - 1) / PG_GROUPS;
    uint64_t grid = (uint64_t)sms * per_sm;
    if (grid > want) grid = want;
    kern<<<(unsigned)grid, PG_THREADS, smem, s>>>(a, slots_log2);
    if (lc) lc->total++;
}

// ------------------------------------------------------------------------------------------------
// generic path: one block per listed read (x index; list == nullptr -> every read), table in HBM with
// 64-bit keys (t, strand, diagonal), sized from the read's bucket sizes

__global__ void generic_hit_counts_kernel(PairgenArgs a, const uint32_t *__restrict__ list, uint32_t n_list,
                                          uint32_t *__restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_list) return;
    const uint32_t xq = list ? list[i] : i;
    uint32_t c = 0;
    for (uint32_t m = a.mz_off[xq]; m < a.mz_off[xq + 1]; ++m) {
        const uint2 rg = a.range[m];
        c += rg.y - rg.x;
    }
    counts[i] = c;
}
void launch_generic_hit_counts(const PairgenArgs &a, const uint32_t *list, uint32_t n_list, uint32_t *counts, cudaStream_t s,
                               LaunchCounter *lc) {
    if (n_list == 0) return;
    generic_hit_counts_kernel<<<(n_list + 127) / 128, 128, 0, s>>>(a, list, n_list, counts);
    if (lc) lc->total++;
}

__device__ __forceinline__ uint32_t slot_hash64(uint64_t key, uint32_t mask) {
    return (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> 40) & mask;
}

__global__ void __launch_bounds__(256) pairgen_generic_kernel(PairgenArgs a, const uint32_t *__restrict__ list, uint32_t n_list,
                                                              unsigned long long *gkeys, uint32_t *gmn, uint32_t *gmx,
                                                              const uint64_t *__restrict__ goff) {
    __shared__ unsigned long long nh_tot;
    const uint32_t bi = blockIdx.x;
    if (bi >= n_list) return;
    const uint32_t xq = list ? list[bi] : bi;
    const uint32_t q = a.reads.perm[xq];
    const uint32_t lq = a.reads.len[q], ql = lq & HLO_LEN_MASK, cb = a.reads.cbase[q];
    const uint64_t o0 = goff[bi];
    const uint32_t S = (uint32_t)(goff[bi + 1] - o0), mask = S - 1u;  // S is a power of two
    unsigned long long *keys = gkeys + o0;
    uint32_t *mn = gmn + o0, *mx = gmx + o0;
    if (threadIdx.x == 0) nh_tot = 0ULL;
    __syncthreads();
    uint32_t nh = 0;
    for (uint32_t m = a.mz_off[xq]; m < a.mz_off[xq + 1]; ++m) {
        const uint2 rg = a.range[m];
        const uint32_t qrp = a.mz_rr[m].y;
        const int qpos = (int)(qrp >> 1);
        for (uint32_t i = rg.x + threadIdx.x; i < rg.y; i += blockDim.x) {
            uint2 h;
            if (a.hits32) {  // (t - cluster base) << (dg_bits + 1) | strand << dg_bits | pos: the targets are in the query's cluster
                const uint32_t w = a.hits32[i];
                h = make_uint2(cb + (w >> (a.dg_bits + 1)), (w & ((1u << a.dg_bits) - 1u)) << 1 | ((w >> a.dg_bits) & 1u));
            } else {
                h = a.hits[i];
            }
            const uint32_t t = h.x;
            if (t == q) continue;
            ++nh;
            const uint32_t strand = (h.y ^ qrp) & 1u;
            const int tp = (int)((h.y >> 1) & 0x7fffu);
            const int qp = strand ? (int)ql - qpos + a.k - 2 : qpos;
            const int diag = tp - qp;
            const uint64_t key = (uint64_t)t << 32 | (uint64_t)(strand << 31) | (uint32_t)(diag + HLO_DIAG_BIAS);
            uint32_t slot = slot_hash64(key, mask);
            for (uint32_t probe = 0; probe <= mask; ++probe) {  // the table is twice the hit count: never full
                const unsigned long long prev = atomicCAS(&keys[slot], HLO_EMPTY_KEY, (unsigned long long)key);
                if (prev == HLO_EMPTY_KEY || prev == key) {
                    atomicMin(&mn[slot], (uint32_t)qp);
                    atomicMax(&mx[slot], (uint32_t)qp);
                    break;
                }
                slot = (slot + 1) & mask;
            }
        }
    }
    if (nh) atomicAdd(&nh_tot, (unsigned long long)nh);
    __syncthreads();
    if (threadIdx.x == 0 && nh_tot) atomicAdd(a.hit_counter, nh_tot);
    for (uint32_t j = threadIdx.x; j < S; j += blockDim.x) {
        const unsigned long long key = keys[j];
        const uint32_t lo = mn[j], hi = mx[j];
        if (key != HLO_EMPTY_KEY && (int)(hi - lo) >= a.gap_min) {
            const unsigned long long p = atomicAdd(a.cand_counter, 1ULL);
            const uint32_t t = (uint32_t)(key >> 32), low = (uint32_t)key;
            const uint32_t lt = a.reads.len[t];
            emit_cand(a, p, q, t, low >> 31, (int)(low & 0x7fffffffu) - HLO_DIAG_BIAS, lo, hi, ql, lt & 0x7fffu, (lt | lq) >> 31);
        }
    }
}

void launch_pairgen_generic(const PairgenArgs &a, const uint32_t *list, uint32_t n_list, unsigned long long *gtable_keys,
                            uint32_t *gtable_mn, uint32_t *gtable_mx, const uint64_t *gtable_off, cudaStream_t s,
                            LaunchCounter *lc) {
    if (n_list == 0) return;
    pairgen_generic_kernel<<<n_list, 256, 0, s>>>(a, list, n_list, gtable_keys, gtable_mn, gtable_mx, gtable_off);
    if (lc) lc->total++;
}

}  // namespace hlo
