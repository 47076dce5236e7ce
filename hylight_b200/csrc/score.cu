// score.cu — the consumer's phred-weighted overlap score on the device.
//
// Restates tools/HaploConduct/src/EdgeCalculator.cpp:26-56 (score), :59-63 (phred_to_prob), :67-139
// (overlap_score) for the S-S case of compute_overlap (:199-221): read ID1 oriented by ORI1 and read ID2 by
// ORI2 (reverse complement = plane 1 of the packed rows, phreds read backwards), seq1[POS1 + i] against
// seq2[i] over L = min(L1 - POS1, L2); positions with an ambiguous base are skipped; the log-probabilities
// are accumulated in position order in FP64 with explicitly rounded operations (no FMA contraction), so the
// only difference from the CPU restatement (oracle/hlo_oracle.c:orc_overlap_score) is the last-ulp behaviour
// of log()/exp().  One thread per record; the phred -> probability table (pow(10, -Q/10) for every byte value,
// computed on the host with the reference's own expression) sits in shared memory.
#include "kernels.cuh"

namespace hlo {

__device__ __forceinline__ uint32_t row_code(const uint32_t *row, int j) { return (row[j >> 4] >> ((j & 15) * 2)) & 3u; }

__global__ void __launch_bounds__(128) score_records_kernel(DevReads r, const uint8_t *__restrict__ quals, const uint64_t *__restrict__ base_off,
                                                            const uint32_t *__restrict__ cluster_off, const hlo_rec *__restrict__ rec,
                                                            const uint32_t *__restrict__ rec_cluster, uint64_t n,
                                                            const double *__restrict__ prob_tab, double mismatch_thr, uint32_t min_read_len,
                                                            hlo_score *__restrict__ out) {
    __shared__ double s_prob[128];
    if (threadIdx.x < 128) s_prob[threadIdx.x] = prob_tab[threadIdx.x];
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const hlo_rec x = rec[i];
    const uint32_t cb = rec_cluster ? cluster_off[rec_cluster[i]] : 0u;
    // the 13-column fields of sfo2overlaps.py:150-200
    uint32_t id1, id2;
    int pos1;
    bool fw1, fw2;
    if (x.OHA >= 0) id1 = x.idA, id2 = x.idB, pos1 = x.OHA, fw1 = true, fw2 = x.ori == 0;
    else id1 = x.idB, id2 = x.idA, pos1 = -x.OHA, fw1 = x.ori == 0, fw2 = true;
    const uint32_t r1 = cb + id1, r2 = cb + id2;
    hlo_score res;
    res.score = 0.0, res.mismatch_rate = 1.0;
    if (r1 < r.n_reads && r2 < r.n_reads) {
        const uint32_t lf1 = r.len[r1], lf2 = r.len[r2];
        const int L1 = (int)(lf1 & HLO_LEN_MASK), L2 = (int)(lf2 & HLO_LEN_MASK);
        if (pos1 < L1 && (uint32_t)L1 >= min_read_len && (uint32_t)L2 >= min_read_len) {
            const bool amb = ((lf1 | lf2) & HLO_AMB_FLAG) != 0;
            const uint64_t plane = (uint64_t)r.n_reads * r.stride_w;
            const uint32_t *row1 = r.bases + (fw1 ? 0 : plane) + (uint64_t)r1 * r.stride_w;
            const uint32_t *row2 = r.bases + (fw2 ? 0 : plane) + (uint64_t)r2 * r.stride_w;
            const uint32_t *arow1 = r.amb + (fw1 ? 0 : plane) + (uint64_t)r1 * r.stride_w;
            const uint32_t *arow2 = r.amb + (fw2 ? 0 : plane) + (uint64_t)r2 * r.stride_w;
            const uint8_t *q1 = quals + base_off[r1], *q2 = quals + base_off[r2];
            const int L = min(L1 - pos1, L2);
            double total = 0.0, total_len = 0.0;
            int mm = 0;
            bool bad = false;
            for (int k = 0; k < L; ++k) {
                const int j1 = pos1 + k, j2 = k;
                if (amb && (row_code(arow1, j1) | row_code(arow2, j2))) continue;  // an N in either read: position skipped
                const double p1 = s_prob[(fw1 ? q1[j1] : q1[L1 - 1 - j1]) & 127], p2 = s_prob[(fw2 ? q2[j2] : q2[L2 - 1 - j2]) & 127];
                double p;
                if (row_code(row1, j1) == row_code(row2, j2)) {
                    // (1-p1)*(1-p2) + (p1*p2)/3.0
                    p = __dadd_rn(__dmul_rn(__dsub_rn(1.0, p1), __dsub_rn(1.0, p2)), __ddiv_rn(__dmul_rn(p1, p2), 3.0));
                } else {
                    // p1*(1-p2)/3.0 + p2*(1-p1)/3.0 + (2/9.0)*p1*p2
                    const double a = __ddiv_rn(__dmul_rn(p1, __dsub_rn(1.0, p2)), 3.0);
                    const double b = __ddiv_rn(__dmul_rn(p2, __dsub_rn(1.0, p1)), 3.0);
                    const double c = __dmul_rn(__dmul_rn(2 / 9.0, p1), p2);
                    p = __dadd_rn(__dadd_rn(a, b), c);
                    ++mm;
                }
                if (p < mismatch_thr) {  // an unacceptable mismatch: the whole overlap scores 0
                    bad = true;
                    break;
                }
                total = __dadd_rn(total, log(p));
                total_len += 1.0;
            }
            if (!bad && total_len > 0.0) {
                res.mismatch_rate = __ddiv_rn((double)(float)mm, total_len);
                res.score = exp(__dmul_rn(__ddiv_rn(1.0, total_len), total));
            }
        }
    }
    out[i] = res;
}

void launch_score_records(const DevReads &r, const uint8_t *quals, const uint64_t *base_off, const uint32_t *cluster_off, const hlo_rec *rec,
                          const uint32_t *rec_cluster, uint64_t n, const double *prob_tab, double mismatch_thr, uint32_t min_read_len,
                          hlo_score *out, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    score_records_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(r, quals, base_off, cluster_off, rec, rec_cluster, n, prob_tab,
                                                                     mismatch_thr, min_read_len, out);
    if (lc) lc->total++;
}

}  // namespace hlo
