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
// pairgen_group_kernel: 16 lanes per query read (2 reads per warp, 8 per block).  The non-empty bucket suffixes
// of the read are staged in shared memory and walked as one flat sequence of chunks (16 hits of one minimizer):
// the lanes stream the hits from the hash-ordered hit array (coalesced 8-byte entries, loaded one chunk ahead) and
// fold each into the group's shared-memory hash table keyed by a 32-bit (t - cluster base, strand, diagonal) that
// keeps min/max anchor position — the warp-level dedup of the north star.  Within one chunk all keys are
// distinct, so only the key claim is an atomic (ATOMS.CAS; a claim by plain store + read-back after __syncwarp,
// and a plain load before the CAS, were measured slower); min/max are plain read-modify-writes ordered by a
// __syncwarp per chunk.  Claimed slots go to a per-read list and the number of qualifying chains is counted
// while inserting, so the emission is one pass over ~100 list entries instead of two over the table.  Lane use is
// set by the chunk granularity (a bucket suffix holds ~12 hits): ~53 % (profiles/r1k_final.md).
// Reads whose table overflows run again with tables of twice the size (ctx.cu:run_pairgen).
// A read that overflows even 2048 slots, or a batch whose keys do not fit 32 bits, is done exactly by
// pairgen_generic_kernel with a 64-bit-keyed table in HBM.
// Algorithmic bytes: 8 B per seed hit in + 16 B per candidate out.
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
//   EMIT = true:  hits[i]  = (read, pos<<1|strand) of the i-th entry (one 8-byte gather from the read-order array);
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
                                                                  uint2 *__restrict__ hits, uint2 *__restrict__ range) {
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
                hits[i] = mz_rr[idx[j]];
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
                         uint2 *hits, uint2 *range, cudaStream_t s, LaunchCounter *lc) {
    if (n_cap == 0 || n_clusters == 0) return;
    const unsigned grid = (unsigned)((n_cap + IX_TILE - 1) / IX_TILE);
    uint32_t *hist_hi = hist ? hist + (size_t)n_clusters * IX_REP * (bins < IX_LO ? bins : IX_LO) : nullptr;
    uint32_t *hist_max = hist ? hist_hi + (size_t)n_clusters * bins : nullptr;
    if (fixed_mid_occ <= 0) {
        cudaMemsetAsync(hist, 0, bucket_hist_words(n_clusters, bins) * sizeof(uint32_t), s);
        bucket_index_kernel<false><<<grid, IX_THREADS, 0, s>>>(skeys, svals, mz_rr, n_ptr, hash_bits, nullptr, hist, hist_hi, hist_max, bins, hits,
                                                               range);
        if (lc) lc->total++;
    }
    midocc_kernel<<<n_clusters, 256, 0, s>>>(hist, hist_hi, hist_max, bins, n_clusters, fixed_mid_occ, midocc, counters);
    bucket_index_kernel<true><<<grid, IX_THREADS, 0, s>>>(skeys, svals, mz_rr, n_ptr, hash_bits, midocc, hist, hist_hi, hist_max, bins, hits, range);
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
        const CandI c = make_candi(q, t, diag, strand, ql, tl, a_start, a_end, amb);
        o = make_uint4(c.q, c.t, c.z, c.w);
    }
    reinterpret_cast<uint4 *>(a.cands)[p] = o;
}

// ------------------------------------------------------------------------------------------------
// fast path: PG_G lanes per query read, PG_RPW reads per warp in LOCKSTEP (every loop bound is the maximum over
// the warp's reads, so the warp never diverges into per-read paths), table in shared memory
constexpr int PG_THREADS = 128;
constexpr int PG_G = 16;                       // lanes per read (8 and 16 measured; 16: fewer, fuller chunks and less smem per block)
constexpr unsigned PG_GMASK = PG_G == 32 ? 0xffffffffu : ((1u << PG_G) - 1u);
constexpr int PG_RPW = 32 / PG_G;              // reads per warp
constexpr int PG_GROUPS = PG_THREADS / PG_G;   // reads in flight per block
constexpr int PG_PAD = 8;                      // words of padding between the groups' tables (bank spread)
constexpr int PG_STAGE = 32;                   // bucket ranges of a read staged at a time (a multiple of PG_G)
__host__ __device__ inline size_t pairgen_table_words(int slots_log2, bool narrow) {
    const size_t ST = ((size_t)1 << slots_log2) + PG_PAD;
    return ((size_t)PG_GROUPS * ST * (narrow ? 7 : 10) + 15) / 16 * 4;  // keys u32 + vals u16|u32 + list u8|u16, rounded to 16 B
}

// NARROW: reads of at most 255 bases and tables of at most 256 slots: min/max anchor positions are bytes and
// the slot list is bytes, 7 instead of 10 bytes of shared memory per slot (more resident warps)
template <bool NARROW>
__global__ void __launch_bounds__(PG_THREADS) pairgen_group_kernel(PairgenArgs a, int slots_log2) {
    extern __shared__ uint32_t pg_smem[];
    typedef typename std::conditional<NARROW, uint16_t, uint32_t>::type val_t;
    typedef typename std::conditional<NARROW, uint8_t, uint16_t>::type list_t;
    constexpr int VSH = NARROW ? 8 : 16;
    constexpr uint32_t VMASK = NARROW ? 0xffu : 0xffffu;
    const uint32_t S = 1u << slots_log2, smask = S - 1u, ST = S + PG_PAD;
    const uint32_t g = threadIdx.x / PG_G, l = threadIdx.x % PG_G;
    uint32_t *keys = pg_smem + (size_t)g * ST;                                // (t - cluster base, strand, diagonal)
    val_t *vals = reinterpret_cast<val_t *>(pg_smem + (size_t)PG_GROUPS * ST) + (size_t)g * ST;  // min | max << VSH anchor position
    list_t *list = reinterpret_cast<list_t *>(reinterpret_cast<val_t *>(pg_smem + (size_t)PG_GROUPS * ST) + (size_t)PG_GROUPS * ST) +
                   (size_t)g * ST;                                            // claimed slots
    // the read's non-empty bucket ranges, long ones first: (first hit, end, query pos<<1|strand, -)
    uint4 *stg = reinterpret_cast<uint4 *>(pg_smem + pairgen_table_words(slots_log2, NARROW)) + (size_t)g * PG_STAGE;
    const unsigned lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const unsigned gshift = lane & ~(unsigned)(PG_G - 1);
    const unsigned lt8 = (1u << l) - 1u;
    const int kshift = a.dg_bits + 1;
    const uint32_t max_len = a.reads.max_len;
    // the listed reads (x indices; their number is a device counter left by the previous pass), or all of them
    uint32_t n_work = a.reads.n_reads;
    if (a.work) {
        const unsigned long long nw = *a.n_work_ptr;
        n_work = nw < (unsigned long long)a.overflow_cap ? (uint32_t)nw : a.overflow_cap;
    }
    const int gap_min = a.gap_min;
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
        const uint32_t nm_max = __reduce_max_sync(FULL, nm);
        if (nm_max == 0) continue;
        {
            uint4 *k4 = reinterpret_cast<uint4 *>(keys);
            const uint4 e4 = make_uint4(HLO_EMPTY32, HLO_EMPTY32, HLO_EMPTY32, HLO_EMPTY32);
            for (uint32_t i = l; i < S / 4; i += PG_G) k4[i] = e4;
            const uint32_t f = NARROW ? VMASK * 0x10001u : VMASK;  // every value: min = VMASK, max = 0
            uint4 *v4 = reinterpret_cast<uint4 *>(vals);
            for (uint32_t i = l; i < S * (uint32_t)sizeof(val_t) / 16; i += PG_G) v4[i] = make_uint4(f, f, f, f);
        }
        __syncwarp();
        uint32_t nh = 0, nq = 0, list_n = 0;
        bool ovf = false;
        // The minimizers are taken PG_STAGE at a time: their non-empty bucket ranges are staged in shared memory, and the
        // read then walks ONE flat sequence of chunks (PG_G hits of one minimizer each).  The two reads of the warp run
        // in lockstep chunk by chunk, not minimizer by minimizer: a short range of one read no longer idles while the
        // other read finishes a long one, steps without hits vanish, and there is no per-minimizer prologue — moving on
        // to the next range is a few predicated instructions inside the chunk loop.
        for (uint32_t mb = 0; mb < nm_max; mb += PG_STAGE) {
            uint32_t n_eff = 0, n_chunks = 0;
            {
                uint2 rg_r[PG_STAGE / PG_G];
                uint32_t qrp_r[PG_STAGE / PG_G];
                unsigned has_b[PG_STAGE / PG_G];
                uint32_t my_chunks = 0;
#pragma unroll
                for (int rr = 0; rr < PG_STAGE / PG_G; ++rr) {
                    const uint32_t idx = mb + rr * PG_G + l;
                    rg_r[rr] = make_uint2(0u, 0u), qrp_r[rr] = 0u;
                    if (idx < nm) rg_r[rr] = __ldg(a.range + m0 + idx), qrp_r[rr] = __ldg(&a.mz_rr[m0 + idx].y);
                    const uint32_t sz = rg_r[rr].y - rg_r[rr].x;
                    has_b[rr] = (__ballot_sync(FULL, sz != 0u) >> gshift) & PG_GMASK;
                    my_chunks += (sz + PG_G - 1) / PG_G;
                }
                uint32_t pos = 0;
#pragma unroll
                for (int rr = 0; rr < PG_STAGE / PG_G; ++rr) {
                    if (rg_r[rr].y != rg_r[rr].x) stg[pos + __popc(has_b[rr] & lt8)] = make_uint4(rg_r[rr].x, rg_r[rr].y, qrp_r[rr], 0u);
                    pos += __popc(has_b[rr]);
                }
                n_eff = pos;
                // chunks of this read = sum over its lanes
#pragma unroll
                for (int dd = PG_G / 2; dd > 0; dd >>= 1) my_chunks += __shfl_xor_sync(FULL, my_chunks, dd);
                n_chunks = my_chunks;
            }
            __syncwarp();
            const uint32_t trips = __reduce_max_sync(FULL, n_chunks);
            // current chunk = hits [i0, min(i0 + PG_G, end)) of the range with query position qrp; next staged entry in e_n
            uint32_t i0 = 0, end = 0, qrp = 0, sidx = 0;
            uint4 e_n = make_uint4(0u, 0u, 0u, 0u);
            if (n_eff) {
                const uint4 e0 = stg[0];
                i0 = e0.x, end = e0.y, qrp = e0.z;
                if (n_eff > 1) e_n = stg[1];
            }
            uint2 h_n = make_uint2(q, 0u);
            if (i0 + l < end) h_n = __ldg(a.hits + i0 + l);
            for (uint32_t c = 0; c < trips; ++c) {
                const uint2 h = h_n;
                const bool in_range = i0 + l < end;
                const uint32_t qrp_c = qrp;
                // the chunk after this one: the rest of the range, or the start of the next staged range
                if (i0 + PG_G < end) {
                    i0 += PG_G;
                } else {
                    i0 = e_n.x, end = e_n.y, qrp = e_n.z;
                    ++sidx;
                    e_n = make_uint4(0u, 0u, 0u, 0u);
                    if (sidx + 1 < n_eff) e_n = stg[sidx + 1];
                }
                h_n = make_uint2(q, 0u);
                HLO_DBG_ASSERT(end <= a.mz_off[a.reads.n_reads]);
                if (i0 + l < end) h_n = __ldg(a.hits + i0 + l);
                int newslot = -1;
                if (in_range && h.x != q) {  // h.x == q: another minimizer of the query itself (-X: no self hits)
                    const uint32_t t = h.x;
                    ++nh;
                    const int qpos = (int)(qrp_c >> 1);
                    const uint32_t strand = (h.y ^ qrp_c) & 1u;
                    const int tp = (int)((h.y >> 1) & 0x7fffu);
                    const int qpi = strand ? (int)ql - qpos + a.k - 2 : qpos;  // ql - (qpos + 1 - k) - 1
                    const int diag = tp - qpi;
                    const uint32_t qp = (uint32_t)qpi;
                    const uint32_t key = (t - cb) << kshift | strand << a.dg_bits | (uint32_t)(diag + (int)max_len);
                    uint32_t slot = (key * 0x9E3779B1u) >> (32 - slots_log2);
                    // claim-or-find, then ONE update path for both outcomes: a fresh slot holds (min = all ones, max = 0),
                    // so "min/max with qp" initialises it and its old span (negative) is below any threshold
                    bool done = false;
                    uint32_t old = 0;
                    for (uint32_t probe = 0; probe < S; ++probe) {
                        old = atomicCAS(&keys[slot], HLO_EMPTY32, key);
                        if (old == HLO_EMPTY32 || old == key) {
                            done = true;
                            break;
                        }
                        slot = (slot + 1u) & smask;
                    }
                    if (done) {  // no other lane of the group holds this key in this chunk: plain read-modify-write
                        HLO_DBG_ASSERT(slot < S && t < a.reads.n_reads && (uint32_t)(diag + (int)max_len) < (1u << a.dg_bits));
                        const uint32_t v = vals[slot];
                        const uint32_t mn = min(v & VMASK, qp), mx = max(v >> VSH, qp);
                        vals[slot] = (val_t)(mn | mx << VSH);
                        nq += ((int)((v >> VSH) - (v & VMASK)) < gap_min) && ((int)(mx - mn) >= gap_min);
                        if (old == HLO_EMPTY32) newslot = (int)slot;
                    }
                    if (!done) ovf = true;
                }
                // remember the claimed slots (per-group list, positions from a ballot: no atomics)
                const unsigned gb = (__ballot_sync(FULL, newslot >= 0) >> gshift) & PG_GMASK;
                HLO_DBG_ASSERT(list_n + __popc(gb) <= S);
                if (newslot >= 0) list[list_n + __popc(gb & lt8)] = (list_t)newslot;
                list_n += __popc(gb);
                __syncwarp();  // orders the plain min/max updates of consecutive chunks (they may belong to different minimizers)
            }
            __syncwarp();  // the staging area is rewritten by the next batch
        }
        // per-read reductions inside the lane groups (uniform across the warp)
        const bool ovf_any = ((__ballot_sync(FULL, ovf) >> gshift) & PG_GMASK) != 0u;
        if (ovf_any) {  // redo this read exactly in the generic kernel (its hits are counted there)
            if (l == 0 && have_read) {
                const unsigned long long p = atomicAdd(a.ovf_counter, 1ULL);
                if (p < a.overflow_cap) a.overflow_reads[p] = xq;
            }
            nq = 0, list_n = 0;
        } else {
            nh_total += nh;
        }
        // candidates of the read = group sum of nq; reserve space for the warp's reads with one atomic
        uint32_t inc = nq;  // inclusive scan over the whole warp: groups are laid out consecutively
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, d);
            if ((int)lane >= d) inc += o;
        }
        const uint32_t total = __shfl_sync(FULL, inc, 31);
        unsigned long long base = 0;
        if (lane == 0 && total) base = atomicAdd(a.cand_counter, (unsigned long long)total);
        base = __shfl_sync(FULL, base, 0);
        // exclusive prefix of my group = inclusive prefix of the last lane of the previous group
        const uint32_t gprev = __shfl_sync(FULL, inc, (gshift + 31u) & 31u);
        unsigned long long p = base + (gshift ? gprev : 0u);
        const uint32_t list_max = __reduce_max_sync(FULL, list_n);
        for (uint32_t j0 = 0; j0 < list_max; j0 += PG_G) {
            const uint32_t j = j0 + l;
            bool okc = false;
            uint32_t key = 0, v = 0;
            if (j < list_n) {
                const uint32_t e = list[j];
                key = keys[e], v = vals[e];
                okc = (int)((v >> VSH) - (v & VMASK)) >= gap_min;
            }
            const unsigned ob = (__ballot_sync(FULL, okc) >> gshift) & PG_GMASK;
            if (okc) {
                const uint32_t t = cb + (key >> kshift);
                const uint32_t lt = __ldg(a.reads.len + t);
                const uint32_t strand = (key >> a.dg_bits) & 1u;
                const int diag = (int)(key & ((1u << a.dg_bits) - 1u)) - (int)max_len;
                emit_cand(a, p + __popc(ob & lt8), q, t, strand, diag, v & VMASK, v >> VSH, ql, lt & 0x7fffu, (lt | lq) >> 31);
            }
            p += __popc(ob);
        }
        __syncwarp();
    }
    // seed-hit statistic: one atomic per warp
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nh_total += __shfl_xor_sync(FULL, nh_total, d);
    if (lane == 0 && nh_total) atomicAdd(a.hit_counter, nh_total);
}

// ------------------------------------------------------------------------------------------------
// pairgen_flat_kernel: the same tables, the same lockstep of two reads per warp, but the lanes of a group walk the
// read's seed hits as ONE FLAT SEQUENCE: lane l of a chunk takes hit number f0 + l of the concatenation of the staged
// bucket suffixes (an exclusive prefix of their lengths maps a hit number to its suffix; a chunk spans at most
// PF_SPAN suffixes).  pairgen_group_kernel gives every chunk to ONE minimizer, and a suffix holds ~12 hits for 16
// lanes: 53 % of the lane slots carry a hit; here ~90 % do.  The price: hits of DIFFERENT minimizers of the query meet
// in one chunk, and overlapping reads share most of their minimizers, so one chunk often holds the same (target,
// strand, diagonal) two or three times.  MATCH.ANY on the key finds those lanes; the lowest lane of each peer set
// takes the minimum and maximum anchor position of the set (REDUX over the peer mask) and is the only one to touch the
// table — within a chunk every table key is still touched by exactly one lane, so min/max stay plain read-modify-writes.
// MEASURED (B200, 32 config-2 clusters): 7.03 ms against 2.34 ms for pairgen_group_kernel — bit-identical candidates,
// three times the time: MATCH.ANY over 32 mostly distinct values and the two REDUX over per-lane peer masks cost far more
// than the idle lanes they remove.  Not the default (option "pairgen_flat" = 1 selects it; one parity test keeps it honest);
// kept as the record of the experiment.
constexpr int PF_SPAN = 5;  // bucket suffixes one chunk may draw from

template <bool NARROW>
__global__ void __launch_bounds__(PG_THREADS) pairgen_flat_kernel(PairgenArgs a, int slots_log2) {
    extern __shared__ uint32_t pg_smem[];
    typedef typename std::conditional<NARROW, uint16_t, uint32_t>::type val_t;
    typedef typename std::conditional<NARROW, uint8_t, uint16_t>::type list_t;
    constexpr int VSH = NARROW ? 8 : 16;
    constexpr uint32_t VMASK = NARROW ? 0xffu : 0xffffu;
    const uint32_t S = 1u << slots_log2, smask = S - 1u, ST = S + PG_PAD;
    const uint32_t g = threadIdx.x / PG_G, l = threadIdx.x % PG_G;
    uint32_t *keys = pg_smem + (size_t)g * ST;
    val_t *vals = reinterpret_cast<val_t *>(pg_smem + (size_t)PG_GROUPS * ST) + (size_t)g * ST;
    list_t *list = reinterpret_cast<list_t *>(reinterpret_cast<val_t *>(pg_smem + (size_t)PG_GROUPS * ST) + (size_t)PG_GROUPS * ST) +
                   (size_t)g * ST;
    // staging area of the group (PG_STAGE * 4 words): prefix of the suffix lengths | first hit | query pos<<1|strand
    uint32_t *sP = pg_smem + pairgen_table_words(slots_log2, NARROW) + (size_t)g * PG_STAGE * 4;  // [PG_STAGE + PF_SPAN + 1]
    uint32_t *sS = sP + PG_STAGE + PF_SPAN + 3;                                                    // [PG_STAGE]
    uint32_t *sQ = sS + PG_STAGE;                                                                  // [PG_STAGE]
    const unsigned lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const unsigned gshift = lane & ~(unsigned)(PG_G - 1);
    const unsigned lt8 = (1u << l) - 1u;
    const unsigned lt_lane = (1u << lane) - 1u;
    const int kshift = a.dg_bits + 1;
    const uint32_t max_len = a.reads.max_len;
    uint32_t n_work = a.reads.n_reads;
    if (a.work) {
        const unsigned long long nw = *a.n_work_ptr;
        n_work = nw < (unsigned long long)a.overflow_cap ? (uint32_t)nw : a.overflow_cap;
    }
    const int gap_min = a.gap_min;
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
        const uint32_t nm_max = __reduce_max_sync(FULL, nm);
        if (nm_max == 0) continue;
        {
            uint4 *k4 = reinterpret_cast<uint4 *>(keys);
            const uint4 e4 = make_uint4(HLO_EMPTY32, HLO_EMPTY32, HLO_EMPTY32, HLO_EMPTY32);
            for (uint32_t i = l; i < S / 4; i += PG_G) k4[i] = e4;
            const uint32_t f = NARROW ? VMASK * 0x10001u : VMASK;  // every value: min = VMASK, max = 0
            uint4 *v4 = reinterpret_cast<uint4 *>(vals);
            for (uint32_t i = l; i < S * (uint32_t)sizeof(val_t) / 16; i += PG_G) v4[i] = make_uint4(f, f, f, f);
        }
        __syncwarp();
        uint32_t nh = 0, nq = 0, list_n = 0;
        bool ovf = false;
        for (uint32_t mb = 0; mb < nm_max; mb += PG_STAGE) {
            // stage the non-empty bucket suffixes of minimizers mb .. mb + PG_STAGE - 1, with the prefix of their lengths
            uint32_t T = 0;  // hits of this stage
            {
                static_assert(PG_STAGE == 2 * PG_G, "two staged entries per lane");
                uint2 rg0 = make_uint2(0u, 0u), rg1 = make_uint2(0u, 0u);
                uint32_t q0 = 0, q1 = 0;
                const uint32_t i0 = mb + l, i1 = mb + PG_G + l;
                if (i0 < nm) rg0 = __ldg(a.range + m0 + i0), q0 = __ldg(&a.mz_rr[m0 + i0].y);
                if (i1 < nm) rg1 = __ldg(a.range + m0 + i1), q1 = __ldg(&a.mz_rr[m0 + i1].y);
                const uint32_t z0 = rg0.y - rg0.x, z1 = rg1.y - rg1.x;
                const unsigned b0 = (__ballot_sync(FULL, z0 != 0u) >> gshift) & PG_GMASK;
                const unsigned b1 = (__ballot_sync(FULL, z1 != 0u) >> gshift) & PG_GMASK;
                // exclusive prefix of the lengths in staged order: entries of the first half, then of the second
                uint32_t inc0 = z0, inc1 = z1;
#pragma unroll
                for (int d = 1; d < PG_G; d <<= 1) {
                    const uint32_t o0 = __shfl_up_sync(FULL, inc0, d), o1 = __shfl_up_sync(FULL, inc1, d);
                    if ((int)l >= d) inc0 += o0, inc1 += o1;
                }
                const uint32_t tot0 = __shfl_sync(FULL, inc0, (int)(gshift + PG_G - 1));
                const uint32_t tot1 = __shfl_sync(FULL, inc1, (int)(gshift + PG_G - 1));
                const uint32_t n0 = __popc(b0), n_eff = n0 + __popc(b1);
                T = tot0 + tot1;
                if (z0) {
                    const uint32_t p = __popc(b0 & lt8);
                    sP[p] = inc0 - z0, sS[p] = rg0.x, sQ[p] = q0;
                }
                if (z1) {
                    const uint32_t p = n0 + __popc(b1 & lt8);
                    sP[p] = tot0 + inc1 - z1, sS[p] = rg1.x, sQ[p] = q1;
                }
                if (l < PF_SPAN + 1) sP[n_eff + l] = T;  // the walk below may look PF_SPAN entries past the last suffix
            }
            __syncwarp();
            // chunk = hits [f0, f0 + PG_G) of the flat sequence, cut at the end of suffix r0 + PF_SPAN - 1
            uint32_t f0 = 0, r0 = 0;
            // this chunk's hit of the lane (loaded one chunk ahead)
            uint2 h = make_uint2(0u, 0u);
            uint32_t qrp = 0;
            bool ok = false;
            auto fetch = [&]() {  // maps f0 + l to its suffix, loads the hit, and moves (f0, r0) to the next chunk
                const uint32_t f = f0 + l;
                const uint32_t p1 = sP[r0 + 1], p2 = sP[r0 + 2], p3 = sP[r0 + 3], p4 = sP[r0 + 4], p5 = sP[r0 + 5];
                const uint32_t r = r0 + (f >= p1) + (f >= p2) + (f >= p3) + (f >= p4);
                ok = f < p5;
                h = make_uint2(0u, 0u), qrp = 0u;
                if (ok) {
                    const uint32_t hi = sS[r] + (f - sP[r]);
                    HLO_DBG_ASSERT(hi < a.mz_off[a.reads.n_reads]);
                    h = __ldg(a.hits + hi), qrp = sQ[r];
                }
                const uint32_t nf = min(f0 + (uint32_t)PG_G, p5);
                r0 += (nf >= p1) + (nf >= p2) + (nf >= p3) + (nf >= p4) + (nf >= p5);
                f0 = nf;
            };
            static_assert(PF_SPAN == 5, "fetch() is written for five suffixes per chunk");
            bool more = f0 < T;  // (T == 0: nothing staged for this read)
            if (more) fetch();
            else ok = false;
            while (__any_sync(FULL, more)) {
                const uint2 hc = h;
                const uint32_t qrp_c = qrp;
                const bool ok_c = more && ok;
                more = more && f0 < T;
                if (more) fetch();  // next chunk's hit, in flight while this one is folded in
                // key and anchor position of this lane's hit
                const bool live = ok_c && hc.x != q;  // hc.x == q: another minimizer of the query itself (-X: no self hits)
                uint32_t key = 0xffffffffu - lane;     // (dead lanes: distinct values that equal no key — keys keep bit 31 clear... see below)
                uint32_t qp = 0;
                if (live) {
                    const uint32_t t = hc.x;
                    ++nh;
                    const int qpos = (int)(qrp_c >> 1);
                    const uint32_t strand = (hc.y ^ qrp_c) & 1u;
                    const int tp = (int)((hc.y >> 1) & 0x7fffu);
                    const int qpi = strand ? (int)ql - qpos + a.k - 2 : qpos;  // ql - (qpos + 1 - k) - 1
                    const int diag = tp - qpi;
                    qp = (uint32_t)qpi;
                    key = (t - cb) << kshift | strand << a.dg_bits | (uint32_t)(diag + (int)max_len);
                }
                // lanes of the group that hold the same key: fold their anchor positions into the lowest one.  (The two
                // groups of the warp work on different reads: bit 31, clear in every key, tells them apart.)
                const unsigned peers = __match_any_sync(FULL, live ? (key | (gshift ? 0x80000000u : 0u)) : key);
                const uint32_t mn_p = __reduce_min_sync(peers, qp), mx_p = __reduce_max_sync(peers, qp);
                const bool leader = live && (peers & lt_lane) == 0u;
                int newslot = -1;
                if (leader) {
                    uint32_t slot = (key * 0x9E3779B1u) >> (32 - slots_log2);
                    bool done = false;
                    uint32_t old = 0;
                    for (uint32_t probe = 0; probe < S; ++probe) {
                        old = atomicCAS(&keys[slot], HLO_EMPTY32, key);
                        if (old == HLO_EMPTY32 || old == key) {
                            done = true;
                            break;
                        }
                        slot = (slot + 1u) & smask;
                    }
                    if (done) {  // no other lane touches this key in this chunk: plain read-modify-write
                        HLO_DBG_ASSERT(slot < S);
                        const uint32_t v = vals[slot];
                        const uint32_t mn = min(v & VMASK, mn_p), mx = max(v >> VSH, mx_p);
                        vals[slot] = (val_t)(mn | mx << VSH);
                        nq += ((int)((v >> VSH) - (v & VMASK)) < gap_min) && ((int)(mx - mn) >= gap_min);
                        if (old == HLO_EMPTY32) newslot = (int)slot;
                    }
                    if (!done) ovf = true;
                }
                const unsigned gb = (__ballot_sync(FULL, newslot >= 0) >> gshift) & PG_GMASK;
                HLO_DBG_ASSERT(list_n + __popc(gb) <= S);
                if (newslot >= 0) list[list_n + __popc(gb & lt8)] = (list_t)newslot;
                list_n += __popc(gb);
                __syncwarp();  // orders the plain min/max updates of consecutive chunks
            }
            __syncwarp();  // the staging area is rewritten by the next stage
        }
        // per-read reductions inside the lane groups (uniform across the warp)
        const bool ovf_any = ((__ballot_sync(FULL, ovf) >> gshift) & PG_GMASK) != 0u;
        // seed hits of the read = group sum
        uint32_t nh_g = nh;
#pragma unroll
        for (int dd = PG_G / 2; dd > 0; dd >>= 1) nh_g += __shfl_xor_sync(FULL, nh_g, dd);
        if (ovf_any) {  // redo this read with a larger table (its hits are counted there)
            if (l == 0 && have_read) {
                const unsigned long long p = atomicAdd(a.ovf_counter, 1ULL);
                if (p < a.overflow_cap) a.overflow_reads[p] = xq;
            }
            nq = 0, list_n = 0;
        } else if (l == 0) {
            nh_total += nh_g;
        }
        uint32_t inc = nq;  // inclusive scan over the whole warp: groups are laid out consecutively
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, d);
            if ((int)lane >= d) inc += o;
        }
        const uint32_t total = __shfl_sync(FULL, inc, 31);
        unsigned long long base = 0;
        if (lane == 0 && total) base = atomicAdd(a.cand_counter, (unsigned long long)total);
        base = __shfl_sync(FULL, base, 0);
        const uint32_t gprev = __shfl_sync(FULL, inc, (gshift + 31u) & 31u);
        unsigned long long p = base + (gshift ? gprev : 0u);
        const uint32_t list_max = __reduce_max_sync(FULL, list_n);
        for (uint32_t j0 = 0; j0 < list_max; j0 += PG_G) {
            const uint32_t j = j0 + l;
            bool okc = false;
            uint32_t key = 0, v = 0;
            if (j < list_n) {
                const uint32_t e = list[j];
                key = keys[e], v = vals[e];
                okc = (int)((v >> VSH) - (v & VMASK)) >= gap_min;
            }
            const unsigned ob = (__ballot_sync(FULL, okc) >> gshift) & PG_GMASK;
            if (okc) {
                const uint32_t t = cb + (key >> kshift);
                const uint32_t lt = __ldg(a.reads.len + t);
                const uint32_t strand = (key >> a.dg_bits) & 1u;
                const int diag = (int)(key & ((1u << a.dg_bits) - 1u)) - (int)max_len;
                emit_cand(a, p + __popc(ob & lt8), q, t, strand, diag, v & VMASK, v >> VSH, ql, lt & 0x7fffu, (lt | lq) >> 31);
            }
            p += __popc(ob);
        }
        __syncwarp();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nh_total += __shfl_xor_sync(FULL, nh_total, d);
    if (lane == 0 && nh_total) atomicAdd(a.hit_counter, nh_total);
}

static bool pairgen_narrow(const DevReads &r, int slots_log2) { return r.max_len <= 255 && slots_log2 <= 8; }
size_t pairgen_smem_bytes(int slots_log2, bool narrow) {
    return pairgen_table_words(slots_log2, narrow) * 4 + (size_t)PG_GROUPS * PG_STAGE * sizeof(uint4);
}

bool pairgen_fast_supported(const DevReads &r, int *dg_bits_out) {
    int dg = 1;
    while ((1u << dg) < 2u * r.max_len + 2u) ++dg;
    int tb = 1;
    while ((1ull << tb) < (unsigned long long)r.max_cluster + 1ull) ++tb;
    if (dg_bits_out) *dg_bits_out = dg;
    return r.max_len <= 32767u && tb + dg + 1 <= 31;  // the top bit stays clear, so no key equals the empty marker
}

void launch_pairgen(const PairgenArgs &a, int slots_log2, cudaStream_t s, LaunchCounter *lc) {
    const uint32_t n_work = a.reads.n_reads;  // (a listed pass usually finds its device-side count at zero and leaves at once)
    if (n_work == 0) return;
    const bool narrow = pairgen_narrow(a.reads, slots_log2);
    const size_t smem = pairgen_smem_bytes(slots_log2, narrow);
    void (*kern)(PairgenArgs, int) = a.flat ? (narrow ? pairgen_flat_kernel<true> : pairgen_flat_kernel<false>)
                                            : (narrow ? pairgen_group_kernel<true> : pairgen_group_kernel<false>);
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
    const uint32_t lq = a.reads.len[q], ql = lq & HLO_LEN_MASK;
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
            const uint2 h = a.hits[i];
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
