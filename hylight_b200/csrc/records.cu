// records.cu — tail of the chain on the device: per-pair keep-longest (filter_ovlp_inline.py:12-39),
// numeric canonicalisation (sfo2overlaps.py:37-50), the (idA, idB) order of `sort -k1,1n -k2,2n`
// (sfo2overlaps.py:53) and expansion of the compact verify records into hlo_rec.
//
// Canonical dedup rule: one record per unordered pair — largest blocklen, then
// fewest mismatches, then '+' before '-', then the smaller diagonal.  The reference keeps the first-seen
// longest inside each 1000-line chunk of minimap2's output order; that order is not reproducible
// without minimap2, so the canonical rule is global.
#include "kernels.cuh"

namespace hlo {

__global__ void record_keys_kernel(const Rec16 *__restrict__ rec, uint64_t n, DevReads r, uint64_t *__restrict__ keys,
                                   uint32_t *__restrict__ vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t q = rec[i].q, t = rec[i].t;
    const uint32_t nq = r.nrank[q], nt = r.nrank[t];
    const uint32_t lo = nq < nt ? nq : nt, hi = nq < nt ? nt : nq;
    keys[i] = (uint64_t)lo << 32 | hi;
    vals[i] = (uint32_t)i;
}
void launch_record_keys(const Rec16 *rec, uint64_t n, DevReads r, uint64_t *keys, uint32_t *vals, cudaStream_t s,
                        LaunchCounter *lc) {
    if (n == 0) return;
    record_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rec, n, r, keys, vals);
    if (lc) lc->total++;
}

__device__ __forceinline__ bool rec_better(const Rec16 &x, const Rec16 &y) {
    const uint32_t bx = x.kb >> 16, by = y.kb >> 16;
    if (bx != by) return bx > by;
    const uint32_t kx = x.kb & 0xffffu, ky = y.kb & 0xffffu;
    if (kx != ky) return kx < ky;
    const uint32_t sx = x.ds & 1u, sy = y.ds & 1u;
    if (sx != sy) return sx < sy;
    return (x.ds >> 1) < (y.ds >> 1);
}

__global__ void dedup_flags_kernel(const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ svals,
                                   const Rec16 *__restrict__ rec, uint64_t n, int dedup, uint32_t *__restrict__ flags,
                                   uint32_t *__restrict__ best) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!dedup) {
        flags[i] = 1u, best[i] = svals[i];
        return;
    }
    const uint64_t key = skeys[i];
    if (i > 0 && skeys[i - 1] == key) {
        flags[i] = 0u, best[i] = 0u;
        return;
    }
    uint32_t bi = svals[i];
    Rec16 br = rec[bi];
    for (uint64_t j = i + 1; j < n && skeys[j] == key; ++j) {
        const uint32_t cj = svals[j];
        const Rec16 cr = rec[cj];
        if (rec_better(cr, br)) br = cr, bi = cj;
    }
    flags[i] = 1u, best[i] = bi;
}
void launch_dedup_flags(const uint64_t *skeys, const uint32_t *svals, const Rec16 *rec, uint64_t n, int dedup, uint32_t *flags,
                        uint32_t *best, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    dedup_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(skeys, svals, rec, n, dedup, flags, best);
    if (lc) lc->total++;
}

__global__ void expand_records_kernel(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ pos,
                                      const uint32_t *__restrict__ best, const Rec16 *__restrict__ rec, uint64_t n, DevReads r,
                                      hlo_rec *__restrict__ out, hlo_rec_aux *__restrict__ aux) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const Rec16 x = rec[best[i]];
    const int ql = (int)(r.len[x.q] & HLO_LEN_MASK), tl = (int)(r.len[x.t] & HLO_LEN_MASK);
    const int diag = (int)(x.ds >> 1) - HLO_DIAG_BIAS;
    hlo_rec o;
    expand_record(r.ids[x.q], r.ids[x.t], ql, tl, diag, (int)(x.ds & 1u), x.kb & 0xffffu, &o);
    const uint32_t p = pos[i];
    out[p] = o;
    hlo_rec_aux ax;
    ax.cluster = r.cidx[x.q];
    ax.blocklen = x.kb >> 16;
    aux[p] = ax;
}
void launch_expand_records(const uint32_t *flags, const uint32_t *pos, const uint32_t *best, const Rec16 *rec, uint64_t n,
                           DevReads r, hlo_rec *out, hlo_rec_aux *aux, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    expand_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(flags, pos, best, rec, n, r, out, aux);
    if (lc) lc->total++;
}

// coff[c] = index of the first output record of cluster c (records are sorted by cluster); coff[n_clusters] = n
__global__ void cluster_offsets_kernel(const hlo_rec_aux *__restrict__ aux, uint64_t n, uint32_t n_clusters,
                                       unsigned long long *__restrict__ coff) {
    const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    const long long cur = aux[o].cluster;
    const long long prev = o ? (long long)aux[o - 1].cluster : -1;
    for (long long c = prev + 1; c <= cur; ++c) coff[c] = o;
    if (o == n - 1)
        for (long long c = cur + 1; c <= (long long)n_clusters; ++c) coff[c] = n;
}
void launch_cluster_offsets(const hlo_rec_aux *aux, uint64_t n, uint32_t n_clusters, unsigned long long *coff, cudaStream_t s,
                            LaunchCounter *lc) {
    if (n == 0) {
        cudaMemsetAsync(coff, 0, ((size_t)n_clusters + 1) * sizeof(unsigned long long), s);
        return;
    }
    cluster_offsets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(aux, n, n_clusters, coff);
    if (lc) lc->total++;
}

}  // namespace hlo
