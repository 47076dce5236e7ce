// records.cu — tail of the chain on the device: per-pair keep-longest (filter_ovlp_inline.py:12-39),
// numeric canonicalisation (sfo2overlaps.py:37-50), the (idA, idB) order of `sort -k1,1n -k2,2n`
// (sfo2overlaps.py:53) and expansion of the compact verify records into hlo_rec.
//
// Canonical dedup rule: one record per unordered pair — largest blocklen, then
// fewest mismatches, then '+' before '-', then the smaller diagonal.  The reference keeps the first-seen
// longest inside each 1000-line chunk of minimap2's output order; that order is not reproducible
// without minimap2, so the canonical rule is global.
//
// No global sort: the output order is (cluster, idA, idB) with idA < idB numerically, and a read has
// only a dozen overlaps as idA.  So: (1) count the records per idA (one L2 atomic each; its return
// value is the record's slot inside its group), (2) exclusive scan, (3) scatter 8-byte (rank of idB,
// record index) entries into the groups, (4) one thread per record decides whether it is the record to
// keep for its pair (marks its entry, counts per idA), (5) scan, (6) one thread per kept record counts
// the kept entries of its group with a smaller idB and expands the record at that place.  Every pass is
// record-parallel and coalesced on the record array; the group loops (a dozen entries) run out of L1.
// Algorithmic bytes per record: 3 * 16 (Rec16 read by three passes) + 8 (entry) + 40 (out).
#include "kernels.cuh"

namespace hlo {

__device__ __forceinline__ bool rec_better(const Rec16 &x, const Rec16 &y) {
    const uint32_t bx = x.kb >> 16, by = y.kb >> 16;
    if (bx != by) return bx > by;
    const uint32_t kx = x.kb & HLO_KB_K_MASK, ky = y.kb & HLO_KB_K_MASK;
    if (kx != ky) return kx < ky;
    const uint32_t sx = x.ds & 1u, sy = y.ds & 1u;
    if (sx != sy) return sx < sy;
    return (x.ds >> 1) < (y.ds >> 1);
}

// Every pass takes the number of accepted records from the verify kernels' device counter (clamped to the buffer) and
// runs on a resident grid that strides: the host sizes nothing from a count it would have to wait for.
constexpr int RT_THREADS = 256;
constexpr uint32_t GB_MIN = 48;  // groups (records per idA) above this size are ranked by group_rank_big_kernel
static unsigned tail_grid(uint64_t n_cap) {
    const uint64_t want = (n_cap + RT_THREADS - 1) / RT_THREADS;
    return (unsigned)(want < 148ull * 8 ? (want ? want : 1) : 148ull * 8);
}
__device__ __forceinline__ uint64_t tail_n(const unsigned long long *n_ptr, uint64_t n_cap) {
    const unsigned long long n = *n_ptr;
    return n < n_cap ? n : n_cap;
}

// group = numeric rank (batch-global) of the record's idA; slot[i] = position of record i inside its group
__global__ void __launch_bounds__(RT_THREADS) group_count_kernel(const Rec16 *__restrict__ rec, const unsigned long long *__restrict__ n_ptr,
                                                                 uint64_t n_cap, DevReads r, uint32_t *__restrict__ cnt,
                                                                 uint32_t *__restrict__ slot) {
    const uint64_t n = tail_n(n_ptr, n_cap);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 qt = *reinterpret_cast<const uint2 *>(&rec[i]);
        const uint32_t nq = r.plain_names ? qt.x : __ldg(r.nrank + qt.x), nt = r.plain_names ? qt.y : __ldg(r.nrank + qt.y);
        slot[i] = atomicAdd(&cnt[min(nq, nt)], 1u);
    }
}
void launch_group_count(const Rec16 *rec, const unsigned long long *n_ptr, uint64_t n_cap, DevReads r, uint32_t *cnt, uint32_t *slot,
                        cudaStream_t s, LaunchCounter *lc) {
    if (n_cap == 0) return;
    group_count_kernel<<<tail_grid(n_cap), RT_THREADS, 0, s>>>(rec, n_ptr, n_cap, r, cnt, slot);
    if (lc) lc->total++;
}

// entry = (rank of idB, record index) at off[group] + slot; entg = its group
__global__ void __launch_bounds__(RT_THREADS) group_scatter_kernel(const Rec16 *__restrict__ rec, const unsigned long long *__restrict__ n_ptr,
                                                                   uint64_t n_cap, DevReads r, const uint32_t *__restrict__ off,
                                                                   const uint32_t *__restrict__ slot, uint2 *__restrict__ ent,
                                                                   uint32_t *__restrict__ entg) {
    const uint64_t n = tail_n(n_ptr, n_cap);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 qt = *reinterpret_cast<const uint2 *>(&rec[i]);
        const uint32_t nq = r.plain_names ? qt.x : __ldg(r.nrank + qt.x), nt = r.plain_names ? qt.y : __ldg(r.nrank + qt.y);
        const uint32_t g = min(nq, nt), p = off[g] + slot[i];
        HLO_DBG_ASSERT(g < r.n_reads && p < n && p < off[g + 1]);
        ent[p] = make_uint2(max(nq, nt), (uint32_t)i);
        entg[p] = g;
    }
}
void launch_group_scatter(const Rec16 *rec, const unsigned long long *n_ptr, uint64_t n_cap, DevReads r, const uint32_t *off,
                          const uint32_t *slot, uint2 *ent, uint32_t *entg, cudaStream_t s, LaunchCounter *lc) {
    if (n_cap == 0) return;
    group_scatter_kernel<<<tail_grid(n_cap), RT_THREADS, 0, s>>>(rec, n_ptr, n_cap, r, off, slot, ent, entg);
    if (lc) lc->total++;
}

// Pass A, one thread per ENTRY (entries are laid out group by group, so the lanes of a warp walk the same one
// or two groups and their loads coalesce): is my record the one to keep for its pair?  (No other entry of the
// group with the same idB is better.)  The answer goes to the top bit of the entry and is counted per group.
__global__ void __launch_bounds__(RT_THREADS) group_mark_kernel(const Rec16 *__restrict__ rec, const unsigned long long *__restrict__ n_ptr,
                                                                uint64_t n_cap, const uint32_t *__restrict__ off,
                                                                const uint32_t *__restrict__ entg, int dedup, uint2 *ent,
                                                                uint32_t *__restrict__ uniq, uint32_t *__restrict__ lt_out, int big_groups,
                                                                uint32_t *__restrict__ big_list, unsigned long long *__restrict__ big_cnt) {
    const uint64_t n = tail_n(n_ptr, n_cap);
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 me = ent[p];
        const uint32_t g = entg[p];
        bool keep = true;
        uint32_t lt = 0;  // entries of the group with a smaller idB: my place in the group when it holds no duplicate pair
        const uint32_t b = off[g], e = off[g + 1];
        if (big_groups && e - b > GB_MIN) {  // group_rank_big_kernel ranks the large groups (this loop is quadratic per group):
            if (p == b) big_list[atomicAdd(big_cnt, 1ULL)] = g;  // its first entry lists it
            continue;
        }
        for (uint32_t j = b; j < e; ++j) {
            const uint2 y = ent[j];
            lt += y.x < me.x;
            // same pair, better record (the mark bit another thread may have set meanwhile is masked off)
            if (dedup && y.x == me.x && j != p && rec_better(rec[y.y & 0x7fffffffu], rec[me.y & 0x7fffffffu])) keep = false;
        }
        lt_out[p] = lt;
        // pair_rule 1 (filter_trans_ovlp_inline_v3.py): the best row stands for its pair even when it failed the overhang
        // test — and then the pair yields nothing
        if (keep && (rec[me.y & 0x7fffffffu].kb & HLO_KB_OVERHANG_FAIL)) keep = false;
        if (keep) {
            ent[p].y = me.y | 0x80000000u;
            atomicAdd(&uniq[g], 1u);
        }
    }
}
// Groups of more than GB_MIN entries (deep clusters: BASELINE configs[3] has 137 records per idA, and the pass above is
// quadratic per group): one warp per group ranks the entries through a bitmap of the cluster's reads in shared memory —
// bit = "some entry has this idB"; the number of set bits below an entry's idB is the entry's final place, distinct idB
// counted once.  Linear in the group size.  A group that holds the same pair twice (the count of entries exceeds the set
// bits) is rare: its warp runs the quadratic rule.  Clusters of more than GB_MAX_CLUSTER reads do not fit the bitmap
// and stay with the pass above.  The large groups come as a list (group_mark_kernel's), one warp per list entry: deep
// clusters put their large groups next to each other (groups are in name order), and a warp that scanned a window of
// consecutive groups had them all to itself.
constexpr uint32_t GB_WARPS = 4, GB_MAX_CLUSTER = 65536, GB_WORDS = GB_MAX_CLUSTER / 32;
__global__ void __launch_bounds__(GB_WARPS * 32) group_rank_big_kernel(const Rec16 *__restrict__ rec, DevReads r, const uint32_t *__restrict__ off,
                                                                       const uint32_t *__restrict__ cluster_off, int dedup, uint2 *ent,
                                                                       uint32_t *__restrict__ uniq, uint32_t *__restrict__ lt_out,
                                                                       const uint32_t *__restrict__ big_list,
                                                                       const unsigned long long *__restrict__ big_cnt) {
    __shared__ uint32_t bm[GB_WARPS][GB_WORDS];  // bitmap, then (in place) the exclusive prefix of its popcounts per 32 words
    __shared__ uint32_t pre[GB_WARPS][GB_WORDS / 32 + 1];
    const unsigned lane = threadIdx.x & 31, wp = threadIdx.x >> 5, FULL = 0xffffffffu;
    const unsigned long long n_big = *big_cnt;
    for (unsigned long long bi = (unsigned long long)blockIdx.x * GB_WARPS + wp; bi < n_big; bi += (unsigned long long)gridDim.x * GB_WARPS) {
      {
        const uint32_t g = big_list[bi];
        HLO_DBG_ASSERT(g < r.n_reads && off[g + 1] - off[g] > GB_MIN);
        const uint32_t b = off[g], e = off[g + 1], sz = e - b;
        const uint32_t cb = r.cbase[g], N = cluster_off[r.cidx[g] + 1] - cb;  // ranks of a cluster's names are cb .. cb + N - 1
        const uint32_t nw = (N + 31) / 32;
        uint32_t *m = bm[wp];
        for (uint32_t i = lane; i < nw; i += 32) m[i] = 0u;
        __syncwarp();
        for (uint32_t j = b + lane; j < e; j += 32) {
            const uint32_t x = ent[j].x - cb;
            atomicOr(&m[x >> 5], 1u << (x & 31));
        }
        __syncwarp();
        // pre[k] = set bits in words [0, 32k)
        uint32_t run = 0;
        for (uint32_t k0 = 0; k0 < nw; k0 += 32) {
            const uint32_t c = k0 + lane < nw ? (uint32_t)__popc(m[k0 + lane]) : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, inc, d);
                if ((int)lane >= d) inc += o;
            }
            if (lane == 0) pre[wp][k0 >> 5] = run;
            run += __shfl_sync(FULL, inc, 31);
        }
        __syncwarp();
        const uint32_t distinct = run;
        if (!dedup || distinct == sz) {  // every pair once (or duplicates are wanted): the place of an entry = set bits below its idB
            uint32_t kept = 0;  // (pair_rule 1 drops the rows that failed the overhang test)
            for (uint32_t j = b + lane; j < e; j += 32) {
                const uint2 me = ent[j];
                const uint32_t x = me.x - cb, w = x >> 5;
                uint32_t lt = pre[wp][w >> 5];
                for (uint32_t k = w & ~31u; k < w; ++k) lt += __popc(m[k]);
                lt += __popc(m[w] & ((1u << (x & 31)) - 1u));
                lt_out[j] = lt;
                if (!(rec[me.y & 0x7fffffffu].kb & HLO_KB_OVERHANG_FAIL)) ent[j].y = me.y | 0x80000000u, ++kept;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) kept += __shfl_xor_sync(FULL, kept, d);
            if (lane == 0) uniq[g] = kept;
        } else {  // the same pair twice in the group: the quadratic rule of group_mark_kernel, one lane per entry
            uint32_t kept = 0;
            for (uint32_t p = b + lane; p < e; p += 32) {
                const uint2 me = ent[p];
                bool keep = true;
                uint32_t lt = 0;
                for (uint32_t j = b; j < e; ++j) {
                    const uint2 y = ent[j];
                    lt += y.x < me.x;
                    if (y.x == me.x && j != p && rec_better(rec[y.y & 0x7fffffffu], rec[me.y & 0x7fffffffu])) keep = false;
                }
                lt_out[p] = lt;
                if (keep && (rec[me.y & 0x7fffffffu].kb & HLO_KB_OVERHANG_FAIL)) keep = false;
                if (keep) ent[p].y = me.y | 0x80000000u, ++kept;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) kept += __shfl_xor_sync(FULL, kept, d);
            if (lane == 0) uniq[g] = kept;
        }
        __syncwarp();
      }
    }
}

void launch_group_mark(const Rec16 *rec, const unsigned long long *n_ptr, uint64_t n_cap, DevReads r, const uint32_t *cluster_off, const uint32_t *off,
                       const uint32_t *entg, int dedup, uint2 *ent, uint32_t *uniq, uint32_t *lt, uint32_t *big_list, unsigned long long *big_cnt,
                       cudaStream_t s, LaunchCounter *lc) {
    if (n_cap == 0) return;
    const int big = r.max_cluster <= GB_MAX_CLUSTER && r.n_reads > 0;
    group_mark_kernel<<<tail_grid(n_cap), RT_THREADS, 0, s>>>(rec, n_ptr, n_cap, off, entg, dedup, ent, uniq, lt, big, big_list, big_cnt);
    if (lc) lc->total++;
    if (big) {
        const uint32_t want = (r.n_reads + GB_WARPS - 1) / GB_WARPS;  // (an upper bound of the list; usually it is short and the grid leaves at once)
        group_rank_big_kernel<<<want < 148u * 6 ? want : 148u * 6, GB_WARPS * 32, 0, s>>>(rec, r, off, cluster_off, dedup, ent, uniq, lt, big_list, big_cnt);
        if (lc) lc->total++;
    }
}

// order-independent digest term of one output record (bench.py recomputes it on the host from the downloaded records)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t rec_digest(const hlo_rec &o) {
    const uint64_t w0 = (uint64_t)o.idA | (uint64_t)o.idB << 32;
    const uint64_t w1 = (uint64_t)(uint32_t)o.OHA | (uint64_t)(uint32_t)o.OHB << 32;
    const uint64_t w2 = (uint64_t)o.OLA | (uint64_t)o.K << 32 | (uint64_t)o.ori << 63;
    return mix64(w0 ^ mix64(w1 ^ mix64(w2 + 0x9e3779b97f4a7c15ULL)));
}

// Pass B, one thread per kept entry: its place inside the group = number of kept entries with a smaller idB
// (= the count pass A left when the group has no duplicate pair; otherwise counted here)
// (ties between kept entries cannot occur with dedup; without dedup the better record goes first); expand it there,
// in every form the caller asked for: hlo_rec (32 B), hlo_rec16, hlo_rec8, hlo_rec_aux.
__global__ void __launch_bounds__(RT_THREADS) group_expand_kernel(const Rec16 *__restrict__ rec, const unsigned long long *__restrict__ n_ptr,
                                                                  uint64_t n_cap, DevReads r, const uint32_t *__restrict__ off,
                                                                  const uint32_t *__restrict__ uoff, const uint32_t *__restrict__ entg,
                                                                  const uint2 *__restrict__ ent, const uint32_t *__restrict__ lt,
                                                                  ExpandOut o_) {
    const uint64_t n = tail_n(n_ptr, n_cap);
    uint64_t dg = 0;
    for (uint64_t pe = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; pe < n; pe += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 me = ent[pe];
        if (!(me.y >> 31)) continue;
        const uint32_t g = entg[pe];
        const uint32_t b = off[g], e = off[g + 1];
        const Rec16 x = rec[me.y & 0x7fffffffu];
        uint32_t rank = lt[pe];  // exact when every entry of the group was kept (no duplicate pair: the usual case)
        if (uoff[g + 1] - uoff[g] != e - b) {
            rank = 0;
            for (uint32_t j = b; j < e; ++j) {
                const uint2 y = ent[j];
                if (!(y.y >> 31) || j == pe) continue;
                rank += y.x < me.x || (y.x == me.x && rec_better(rec[y.y & 0x7fffffffu], x));
            }
        }
        const uint32_t p = uoff[g] + rank;
        HLO_DBG_ASSERT(p < uoff[g + 1] && p < n_cap && x.q < r.n_reads && x.t < r.n_reads);
        const uint4 mq = __ldg(r.meta + x.q), mt = __ldg(r.meta + x.t);  // (length, name, cluster, cluster base)
        const int ql = (int)mq.x, tl = (int)mt.x;
        const int diag = (int)(x.ds >> 1) - HLO_DIAG_BIAS;
        hlo_rec o;
        expand_record(mq.y, mt.y, ql, tl, diag, (int)(x.ds & 1u), x.kb & HLO_KB_K_MASK, &o);
        if (o_.out8)  // 8-byte form (hlo_rec8): ids below 2^16; OHB and OLA follow from OHA and the two read lengths
            reinterpret_cast<uint2 *>(o_.out8)[p] = make_uint2(o.idA | o.idB << 16, ((uint32_t)o.OHA & 0xffffu) | (o.K | o.ori << 15) << 16);
        if (o_.out16)  // 16-byte form (hlo_rec16): OLB == OLA, 16-bit overhangs and lengths
            reinterpret_cast<uint4 *>(o_.out16)[p] = make_uint4(o.idA, o.idB, ((uint32_t)o.OHA & 0xffffu) | (uint32_t)o.OHB << 16,
                                                                o.OLA | (o.K | o.ori << 15) << 16);
        if (o_.out) {
            uint4 *dst = reinterpret_cast<uint4 *>(o_.out + p);
            dst[0] = make_uint4(o.idA, o.idB, (uint32_t)o.OHA, (uint32_t)o.OHB);
            dst[1] = make_uint4(o.OLA, o.OLB, o.K, o.ori);
        }
        if (o_.aux) {
            hlo_rec_aux ax;
            ax.cluster = mq.z;
            ax.blocklen = x.kb >> 16;
            o_.aux[p] = ax;
        }
        if (o_.digest) dg += rec_digest(o);
    }
    if (o_.digest) {  // (sums commute: any order gives the same value)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) dg += __shfl_xor_sync(0xffffffffu, dg, d);
        if ((threadIdx.x & 31) == 0 && dg) atomicAdd(o_.digest, (unsigned long long)dg);
    }
}
void launch_group_expand(const Rec16 *rec, const unsigned long long *n_ptr, uint64_t n_cap, DevReads r, const uint32_t *off, const uint32_t *uoff,
                         const uint32_t *entg, const uint2 *ent, const uint32_t *lt, const ExpandOut &o, cudaStream_t s, LaunchCounter *lc) {
    if (n_cap == 0) return;
    group_expand_kernel<<<tail_grid(n_cap), RT_THREADS, 0, s>>>(rec, n_ptr, n_cap, r, off, uoff, entg, ent, lt, o);
    if (lc) lc->total++;
}

uint64_t host_rec_digest(const hlo_rec &o) { return rec_digest(o); }

// coff[c] = index of the first output record of cluster c = output offset of the cluster's first group (groups are
// numbered cluster by cluster: cluster base + numeric rank of idA); coff[n_clusters] = number of records
__global__ void cluster_offsets_kernel(const uint32_t *__restrict__ uoff, const uint32_t *__restrict__ cluster_off, uint32_t n_clusters,
                                       unsigned long long *__restrict__ coff) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= n_clusters) coff[c] = uoff[cluster_off[c]];
}
void launch_cluster_offsets(const uint32_t *uoff, const uint32_t *cluster_off, uint32_t n_clusters, unsigned long long *coff, cudaStream_t s,
                            LaunchCounter *lc) {
    cluster_offsets_kernel<<<(n_clusters + 1 + 255) / 256, 256, 0, s>>>(uoff, cluster_off, n_clusters, coff);
    if (lc) lc->total++;
}

}  // namespace hlo
