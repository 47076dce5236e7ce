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
    const uint32_t kx = x.kb & 0xffffu, ky = y.kb & 0xffffu;
    if (kx != ky) return kx < ky;
    const uint32_t sx = x.ds & 1u, sy = y.ds & 1u;
    if (sx != sy) return sx < sy;
    return (x.ds >> 1) < (y.ds >> 1);
}

// group = numeric rank (batch-global) of the record's idA; slot[i] = position of record i inside its group
__global__ void __launch_bounds__(256) group_count_kernel(const Rec16 *__restrict__ rec, uint64_t n, DevReads r,
                                                          uint32_t *__restrict__ cnt, uint32_t *__restrict__ slot) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 qt = *reinterpret_cast<const uint2 *>(&rec[i]);
    const uint32_t nq = r.plain_names ? qt.x : __ldg(r.nrank + qt.x), nt = r.plain_names ? qt.y : __ldg(r.nrank + qt.y);
    slot[i] = atomicAdd(&cnt[min(nq, nt)], 1u);
}
void launch_group_count(const Rec16 *rec, uint64_t n, DevReads r, uint32_t *cnt, uint32_t *slot, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    group_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rec, n, r, cnt, slot);
    if (lc) lc->total++;
}

// entry = (rank of idB, record index) at off[group] + slot; entg = its group
__global__ void __launch_bounds__(256) group_scatter_kernel(const Rec16 *__restrict__ rec, uint64_t n, DevReads r,
                                                            const uint32_t *__restrict__ off, const uint32_t *__restrict__ slot,
                                                            uint2 *__restrict__ ent, uint32_t *__restrict__ entg) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 qt = *reinterpret_cast<const uint2 *>(&rec[i]);
    const uint32_t nq = r.plain_names ? qt.x : __ldg(r.nrank + qt.x), nt = r.plain_names ? qt.y : __ldg(r.nrank + qt.y);
    const uint32_t g = min(nq, nt), p = off[g] + slot[i];
    ent[p] = make_uint2(max(nq, nt), (uint32_t)i);
    entg[p] = g;
}
void launch_group_scatter(const Rec16 *rec, uint64_t n, DevReads r, const uint32_t *off, const uint32_t *slot, uint2 *ent,
                          uint32_t *entg, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    group_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rec, n, r, off, slot, ent, entg);
    if (lc) lc->total++;
}

// Pass A, one thread per ENTRY (entries are laid out group by group, so the lanes of a warp walk the same one
// or two groups and their loads coalesce): is my record the one to keep for its pair?  (No other entry of the
// group with the same idB is better.)  The answer goes to the top bit of the entry and is counted per group.
__global__ void __launch_bounds__(256) group_mark_kernel(const Rec16 *__restrict__ rec, uint64_t n, const uint32_t *__restrict__ off,
                                                         const uint32_t *__restrict__ entg, int dedup, uint2 *ent,
                                                         uint32_t *__restrict__ uniq, uint32_t *__restrict__ lt_out) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint2 me = ent[p];
    const uint32_t g = entg[p];
    bool keep = true;
    uint32_t lt = 0;  // entries of the group with a smaller idB: my place in the group when it holds no duplicate pair
    const uint32_t b = off[g], e = off[g + 1];
    for (uint32_t j = b; j < e; ++j) {
        const uint2 y = ent[j];
        lt += y.x < me.x;
        if (dedup && y.x == me.x && j != p && rec_better(rec[y.y & 0x7fffffffu], rec[me.y])) keep = false;  // same pair, better record
    }
    lt_out[p] = lt;
    if (keep) {
        ent[p].y = me.y | 0x80000000u;
        atomicAdd(&uniq[g], 1u);
    }
}
void launch_group_mark(const Rec16 *rec, uint64_t n, const uint32_t *off, const uint32_t *entg, int dedup, uint2 *ent, uint32_t *uniq,
                       uint32_t *lt, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    group_mark_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rec, n, off, entg, dedup, ent, uniq, lt);
    if (lc) lc->total++;
}

// Pass B, one thread per kept entry: its place inside the group = number of kept entries with a smaller idB
// (= the count pass A left when the group has no duplicate pair; otherwise counted here)
// (ties between kept entries cannot occur with dedup; without dedup the better record goes first); expand it there.
__global__ void __launch_bounds__(256) group_expand_kernel(const Rec16 *__restrict__ rec, uint64_t n, DevReads r,
                                                           const uint32_t *__restrict__ off, const uint32_t *__restrict__ uoff,
                                                           const uint32_t *__restrict__ entg, const uint2 *__restrict__ ent,
                                                           const uint32_t *__restrict__ lt, hlo_rec *__restrict__ out, hlo_rec16 *__restrict__ out16,
                                                           hlo_rec_aux *__restrict__ aux) {
    const uint64_t pe = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pe >= n) return;
    const uint2 me = ent[pe];
    if (!(me.y >> 31)) return;
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
    const uint4 mq = __ldg(r.meta + x.q), mt = __ldg(r.meta + x.t);  // (length, name, cluster, cluster base)
    const int ql = (int)mq.x, tl = (int)mt.x;
    const int diag = (int)(x.ds >> 1) - HLO_DIAG_BIAS;
    hlo_rec o;
    expand_record(mq.y, mt.y, ql, tl, diag, (int)(x.ds & 1u), x.kb & 0xffffu, &o);
    if (out16) {  // 16-byte form (hlo_rec16): OLB == OLA, 16-bit overhangs and lengths
        reinterpret_cast<uint4 *>(out16)[p] = make_uint4(o.idA, o.idB, ((uint32_t)o.OHA & 0xffffu) | (uint32_t)o.OHB << 16,
                                                          o.OLA | (o.K | o.ori << 15) << 16);
    } else {
        uint4 *dst = reinterpret_cast<uint4 *>(out + p);
        dst[0] = make_uint4(o.idA, o.idB, (uint32_t)o.OHA, (uint32_t)o.OHB);
        dst[1] = make_uint4(o.OLA, o.OLB, o.K, o.ori);
    }
    if (aux) {
        hlo_rec_aux ax;
        ax.cluster = mq.z;
        ax.blocklen = x.kb >> 16;
        aux[p] = ax;
    }
}
void launch_group_expand(const Rec16 *rec, uint64_t n, DevReads r, const uint32_t *off, const uint32_t *uoff, const uint32_t *entg,
                         const uint2 *ent, const uint32_t *lt, hlo_rec *out, hlo_rec16 *out16, hlo_rec_aux *aux, cudaStream_t s,
                         LaunchCounter *lc) {
    if (n == 0) return;
    group_expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rec, n, r, off, uoff, entg, ent, lt, out, out16, aux);
    if (lc) lc->total++;
}

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
