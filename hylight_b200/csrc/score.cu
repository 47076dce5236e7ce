// score.cu — the consumer's phred-weighted overlap score on the device.
//
// Restates tools/HaploConduct/src/EdgeCalculator.cpp:26-56 (score), :59-63 (phred_to_prob), :67-139
// (overlap_score) for the S-S case of compute_overlap (:199-221): read ID1 oriented by ORI1 and read ID2 by
// ORI2 (reverse complement = plane 1 of the packed rows, phreds read backwards), seq1[POS1 + i] against
// seq2[i] over L = min(L1 - POS1, L2); positions with an ambiguous base are skipped; the log-probabilities
// are accumulated in position order in FP64 with explicitly rounded operations (no FMA contraction), so the
// only difference from the CPU restatement (oracle/hlo_oracle.c:orc_overlap_score) is the last-ulp behaviour
// of log()/exp().  One thread per record; the phred -> probability table (pow(10, -Q/10) for every byte value) is
// computed on the host with the reference's own expression.
#include "kernels.cuh"

namespace hlo {

__device__ __forceinline__ uint32_t row_code(const uint32_t *row, int j) { return (row[j >> 4] >> ((j & 15) * 2)) & 3u; }

// The probability that two observed bases agree in truth depends only on the two phreds and on whether the bases
// match: log p for every (match?, phred byte 1, phred byte 2) is tabulated per call by score_table_kernel — with
// exactly the operations the per-position code would use, so a lookup equals the computation bit for bit; an entry
// whose p is below the mismatch threshold holds NaN ("this overlap scores 0").  The record kernel keeps the part of the
// table for phred bytes 33..126 (94 x 94 x 2 doubles = 141 KB) in shared memory — a random 8-byte lookup per
// position costs one wavefront per distinct line on the global path, a few per warp in shared memory — and falls back
// to the table in global memory for other byte values.
constexpr int SC_Q0 = 33, SC_NQ = 94;
constexpr int SC_THREADS = 1024;

__global__ void __launch_bounds__(128) score_table_kernel(const double *__restrict__ prob_tab, double mismatch_thr,
                                                          double *__restrict__ tab) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2u * 128u * 128u) return;
    const double p1 = prob_tab[(i >> 7) & 127u], p2 = prob_tab[i & 127u];
    double p;
    if ((i >> 14) == 0) {
        // (1-p1)*(1-p2) + (p1*p2)/3.0
        p = __dadd_rn(__dmul_rn(__dsub_rn(1.0, p1), __dsub_rn(1.0, p2)), __ddiv_rn(__dmul_rn(p1, p2), 3.0));
    } else {
        // p1*(1-p2)/3.0 + p2*(1-p1)/3.0 + (2/9.0)*p1*p2
        const double a = __ddiv_rn(__dmul_rn(p1, __dsub_rn(1.0, p2)), 3.0);
        const double b = __ddiv_rn(__dmul_rn(p2, __dsub_rn(1.0, p1)), 3.0);
        const double c = __dmul_rn(__dmul_rn(2 / 9.0, p1), p2);
        p = __dadd_rn(__dadd_rn(a, b), c);
    }
    // EdgeCalculator::score returns log p <= 0, or a positive sentinel (1: N position, 2: unacceptable mismatch), and
    // overlap_score reads every other positive or NaN value as "unacceptable" too: that is what a phred byte below
    // '!' (error probability > 1, p outside [0, 1]) amounts to
    const double lp = log(p);
    tab[i] = (p < mismatch_thr || !(lp <= 0.0)) ? __longlong_as_double(0x7ff8000000000000LL) : lp;
}
void launch_score_table(const double *prob_tab, double mismatch_thr, double *tab, cudaStream_t s, LaunchCounter *lc) {
    score_table_kernel<<<2 * 128 * 128 / 128, 128, 0, s>>>(prob_tab, mismatch_thr, tab);
    if (lc) lc->total++;
}

// qflag[read] = 1 when every phred byte of the read lies in the range the shared-memory table covers
__global__ void __launch_bounds__(256) qual_flags_kernel(const uint8_t *__restrict__ quals, const uint64_t *__restrict__ base_off,
                                                         uint32_t n_reads, uint8_t *__restrict__ qflag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    bool ok = true;
    for (uint64_t o = base_off[i]; o < base_off[i + 1]; ++o) ok &= (uint32_t)(quals[o] - SC_Q0) < (uint32_t)SC_NQ;
    qflag[i] = ok;
}
void launch_qual_flags(const uint8_t *quals, const uint64_t *base_off, uint32_t n_reads, uint8_t *qflag, cudaStream_t s, LaunchCounter *lc) {
    if (n_reads == 0) return;
    qual_flags_kernel<<<(n_reads + 255) / 256, 256, 0, s>>>(quals, base_off, n_reads, qflag);
    if (lc) lc->total++;
}

// the 16 bytes at quals + A .. A + 15 (any alignment, A may be slightly negative: the buffer is padded on both sides)
__device__ __forceinline__ void load16(const uint8_t *quals, long long A, uint32_t X[4]) {
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(quals + (A & ~3LL));
    const uint32_t sh = ((uint32_t)A & 3u) * 8u;
    uint32_t W[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) W[i] = wp[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) X[i] = __funnelshift_r(W[i], W[i + 1], sh);
}

__global__ void __launch_bounds__(SC_THREADS, 1) score_records_kernel(DevReads r, const uint8_t *__restrict__ quals,
                                                                      const uint8_t *__restrict__ qflag, const uint64_t *__restrict__ base_off,
                                                                      const uint32_t *__restrict__ cluster_off, const hlo_rec *__restrict__ rec,
                                                                      const uint32_t *__restrict__ rec_cluster, int cluster_stride,
                                                                      const uint32_t *__restrict__ n_ptr, uint64_t n_cap,
                                                                      const double *__restrict__ tab, uint32_t min_read_len,
                                                                      hlo_score *__restrict__ out) {
    // n_ptr: the record count on the device (the fused path never brings it to the host first); else n_cap is the count
    uint64_t n = n_cap;
    if (n_ptr) n = *n_ptr < n_cap ? *n_ptr : n_cap;
    extern __shared__ double s_tab[];  // [2][SC_NQ][SC_NQ]
    for (uint32_t e = threadIdx.x; e < 2u * SC_NQ * SC_NQ; e += blockDim.x) {
        const uint32_t m = e / (SC_NQ * SC_NQ), rem = e % (SC_NQ * SC_NQ);
        s_tab[e] = tab[m << 14 | (rem / SC_NQ + SC_Q0) << 7 | (rem % SC_NQ + SC_Q0)];
    }
    __syncthreads();
    const double *t_match = s_tab - (SC_Q0 * SC_NQ + SC_Q0);  // indexed by phred byte 1 * SC_NQ + phred byte 2
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const hlo_rec x = rec[i];
        const uint32_t cb = rec_cluster ? cluster_off[rec_cluster[i * cluster_stride]] : 0u;  // (stride 2: the cluster field of hlo_rec_aux)
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
                const int L = min(L1 - pos1, L2);
                const long long b1 = (long long)base_off[r1], b2 = (long long)base_off[r2];
                double total = 0.0;
                int mm = 0, cnt = 0;
                if (!amb && qflag[r1] && qflag[r2]) {
                    // 16 positions per step: mismatch mask from the packed words, 16 phreds of each read from five aligned
                    // loads, one shared-memory lookup and one add per position, in position order.  A NaN entry (p below
                    // the threshold) poisons the sum, which is all the early exit of the reference amounts to.
                    for (int k0 = 0; k0 < L; k0 += 16) {
                        const int j1 = pos1 + k0;
                        const uint32_t sh1 = (uint32_t)(j1 & 15) * 2u;
                        const uint32_t w1 = __funnelshift_r(row1[j1 >> 4], row1[(j1 >> 4) + 1], sh1);  // (rows are padded)
                        const uint32_t d = w1 ^ row2[k0 >> 4];
                        const int np = min(16, L - k0);
                        uint32_t m16 = (d | d >> 1) & 0x55555555u;              // bit 2p: position p mismatches
                        if (np < 16) m16 &= (1u << (2 * np)) - 1u;
                        mm += __popc(m16);
                        cnt += np;
                        uint32_t X1[4], X2[4];
                        load16(quals, fw1 ? b1 + j1 : b1 + L1 - 1 - j1 - 15, X1);
                        load16(quals, fw2 ? b2 + k0 : b2 + L2 - 1 - k0 - 15, X2);
#pragma unroll
                        for (int t = 0; t < 16; ++t) {
                            if (t < np) {
                                const int bf = t, br = 15 - t;  // byte of position t in a forward / backward walk
                                const uint32_t qa = fw1 ? __byte_perm(X1[bf >> 2], 0u, 0x4440u | (bf & 3)) : __byte_perm(X1[br >> 2], 0u, 0x4440u | (br & 3));
                                const uint32_t qb = fw2 ? __byte_perm(X2[bf >> 2], 0u, 0x4440u | (bf & 3)) : __byte_perm(X2[br >> 2], 0u, 0x4440u | (br & 3));
                                const uint32_t idx = qa * SC_NQ + qb + ((m16 >> (2 * t)) & 1u) * (SC_NQ * SC_NQ);
                                total = __dadd_rn(total, t_match[idx]);
                            }
                        }
                    }
                } else {  // ambiguous bases or unusual phred bytes: position by position, table in global memory
                    const uint32_t *arow1 = r.amb + (fw1 ? 0 : plane) + (uint64_t)r1 * r.stride_w;
                    const uint32_t *arow2 = r.amb + (fw2 ? 0 : plane) + (uint64_t)r2 * r.stride_w;
                    for (int k = 0; k < L; ++k) {
                        const int j1 = pos1 + k;
                        if (amb && (row_code(arow1, j1) | row_code(arow2, k))) continue;  // an N in either read: position skipped
                        const uint32_t qa = quals[fw1 ? b1 + j1 : b1 + L1 - 1 - j1] & 127u, qb = quals[fw2 ? b2 + k : b2 + L2 - 1 - k] & 127u;
                        const uint32_t mis = row_code(row1, j1) != row_code(row2, k);
                        total = __dadd_rn(total, __ldg(tab + (mis << 14 | qa << 7 | qb)));
                        mm += (int)mis;
                        ++cnt;
                        if (total != total) break;
                    }
                }
                if (total == total && cnt > 0) {  // (NaN: an unacceptable mismatch, the overlap scores 0)
                    const double total_len = (double)cnt;
                    res.mismatch_rate = __ddiv_rn((double)(float)mm, total_len);
                    res.score = exp(__dmul_rn(__ddiv_rn(1.0, total_len), total));
                }
            }
        }
        out[i] = res;
    }
}

void launch_score_records(const DevReads &r, const uint8_t *quals, const uint8_t *qflag, const uint64_t *base_off, const uint32_t *cluster_off,
                          const hlo_rec *rec, const uint32_t *rec_cluster, int cluster_stride, const uint32_t *n_ptr, uint64_t n,
                          const double *tab, uint32_t min_read_len, hlo_score *out, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    const size_t smem = (size_t)2 * SC_NQ * SC_NQ * sizeof(double);
    cudaFuncSetAttribute(score_records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t want = (n + SC_THREADS - 1) / SC_THREADS;
    const unsigned grid = (unsigned)(want < (uint64_t)sms ? want : (uint64_t)sms);
    score_records_kernel<<<grid, SC_THREADS, smem, s>>>(r, quals, qflag, base_off, cluster_off, rec, rec_cluster, cluster_stride, n_ptr, n, tab,
                                                        min_read_len, out);
    if (lc) lc->total++;
}

}  // namespace hlo
