// verify.cu — base-level verification of candidates: ungapped compare on the candidate's diagonal with
// minimap2's end rule, then filter_ovlp_inline.py's thresholds, then stream compaction.
//
// Semantics (oracle/hlo_oracle.c:orc_extend + orc_filter_row; reference files:
// filter_ovlp_inline.py:58,73-87 thresholds/overhang, EdgeCalculator.cpp:86-132 ungapped compare with
// ambiguous positions skipped; minimap2 align.c SR mode for the end rule):
//   q' = query as aligned (reverse complement when strand = 1); q'[j] <-> t[j + D], j in [js, je).
//   An end where the QUERY ends first is extended to the read end (end bonus); an end where the TARGET
//   ends first is clipped at the earliest maximum of the cumulative (+A/-B/-N) score counted outward
//   from the outermost anchor.  K = mismatches in the block, blocklen = block columns - ambiguous.
//   Accept iff blocklen >= min_len, !(matches/blocklen < min_id), !(clip_l + clip_r > min(o, maplen*r)).
//
// verify_fast_kernel: 4 lanes per candidate, 8 candidates per warp.  Lane s loads bases [64s, 64s+64)
// of both rows with one 128-bit load each; the target words are re-aligned to the query with register
// rotation + 4-lane shuffles + funnel shifts; mismatches = popc((x | x>>1) & 0x5555..) of the XOR.
// The end rule only needs the outermost 32 bases of each target-limited end as long as the overlap holds
// at most bad_limit (21 for -A4 -B2) bad (mismatch/ambiguous) columns; anything else (or reads > 256 bp) is deferred to
// verify_exact_kernel, which walks base by base.  Accepted records leave through one warp-aggregated
// atomicAdd per warp iteration.
//
// Bound: L1/L2 sector bandwidth + integer issue (the rows of a cluster are L2 resident); algorithmic
// bytes per candidate = 16 (candidate) + 2*ceil(OL/4) (bases in the overlap) + 16 per accepted record
// (SURVEY.md §8d counts 32 B for the expanded record; the expansion happens in records.cu).
#include "kernels.cuh"

namespace hlo {

__device__ __forceinline__ uint32_t sel4(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t r) {
    const uint32_t lo = (r & 1u) ? x1 : x0, hi = (r & 1u) ? x3 : x2;
    return (r & 2u) ? hi : lo;
}

// 2-bit-granular mask of positions [lo, hi) of a 16-base word (bit 2p set for position p)
__device__ __forceinline__ uint32_t range_mask16(int lo, int hi) {
    lo = max(lo, 0), hi = min(hi, 16);
    const uint32_t mhi = hi >= 16 ? 0xffffffffu : (hi <= 0 ? 0u : ((1u << (2 * hi)) - 1u));
    const uint32_t mlo = lo >= 16 ? 0xffffffffu : ((1u << (2 * lo)) - 1u);
    return mhi & ~mlo & 0x55555555u;
}

// fetch word `w` (0..15, else 0) of the 16-word row spread over the 4 lanes of a group (lane s holds
// words 4s..4s+3 in r0..r3).  `w` must be uniform across the group.
__device__ __forceinline__ uint32_t group_word(unsigned gmask, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, int w) {
    const uint32_t v = sel4(r0, r1, r2, r3, (uint32_t)w & 3u);
    const uint32_t x = __shfl_sync(gmask, v, (w >> 2) & 3, 4);
    return (w >= 0 && w < 16) ? x : 0u;
}

// re-align a row held as 16 words over 4 lanes so that this lane gets bases [64s + D, 64s + D + 64)
__device__ __forceinline__ void realign_row(unsigned gmask, int sub, uint4 tv, int D, uint32_t out[4]) {
    const int a = D >> 4, b2 = (D & 15) * 2;
    const uint32_t ar = (uint32_t)a & 3u;
    uint32_t u[4];
    u[0] = sel4(tv.x, tv.y, tv.z, tv.w, ar);
    u[1] = sel4(tv.x, tv.y, tv.z, tv.w, (ar + 1u) & 3u);
    u[2] = sel4(tv.x, tv.y, tv.z, tv.w, (ar + 2u) & 3u);
    u[3] = sel4(tv.x, tv.y, tv.z, tv.w, (ar + 3u) & 3u);
    uint32_t x[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int w = 4 * sub + a + i;
        const uint32_t v = __shfl_sync(gmask, u[i & 3], (w >> 2) & 3, 4);
        x[i] = (w >= 0 && w < 16) ? v : 0u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = __funnelshift_r(x[i], x[i + 1], b2);
}

struct ClipResult {
    int clip, mm, amb;
};

// end rule on a 32-base window; bit 2(p-1) of `bad` set <=> the base at outward distance p (1 = the
// terminal base) is a mismatch or ambiguous; `ambw` marks the ambiguous ones.
__device__ __forceinline__ ClipResult clip_scan(uint64_t bad, uint64_t ambw, int A, int B, int N) {
    ClipResult r{0, 0, 0};
    int best = 0, cm = 0, ca = 0;
    while (bad) {
        const int bit = __ffsll((long long)bad) - 1;
        const int p = (bit >> 1) + 1;
        if ((ambw >> bit) & 1ull) ++ca; else ++cm;
        const int R = A * p - (A + B) * cm - (A + N) * ca;
        if (R <= best) best = R, r.clip = p, r.mm = cm, r.amb = ca;
        bad &= bad - 1;
    }
    return r;
}

__device__ __forceinline__ uint64_t brev_pairs64(uint64_t v) {
    // reverse the order of the 32 two-bit fields (flags sit on the even bit of each field)
    return __brevll(v) >> 1;
}

__global__ void __launch_bounds__(256) verify_fast_kernel(VerifyArgs a) {
    const unsigned lane = threadIdx.x & 31;
    const int sub = (int)(lane & 3u);
    const unsigned gmask = 0xFu << (lane & ~3u);
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint64_t n_groups_total = ((uint64_t)gridDim.x * blockDim.x) >> 2;
    const uint64_t g0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const uint64_t plane = (uint64_t)a.reads.n_reads * a.reads.stride_w;
    const uint32_t sw = a.reads.stride_w;
    const uint64_t n_round = (a.n + 7) & ~7ull;  // whole warps iterate together
    for (uint64_t c = g0; c < n_round; c += n_groups_total) {
        bool accept = false, defer = false;
        Rec16 rec;
        rec.q = rec.t = rec.ds = rec.kb = 0;
        if (c < a.n) {
            const uint4 cv = __ldg(reinterpret_cast<const uint4 *>(a.cands) + c);
            const uint32_t q = cv.x, t = cv.y;
            const int D = (int)cv.z;
            const uint32_t strand = cv.w & 1u;
            int aS = (int)((cv.w >> 1) & 0x7fffu), aE = (int)(cv.w >> 16);
            const uint32_t lq = __ldg(a.reads.len + q), lt = __ldg(a.reads.len + t);
            const int ql = (int)(lq & HLO_LEN_MASK), tl = (int)(lt & HLO_LEN_MASK);
            const bool ambf = ((lq | lt) & HLO_AMB_FLAG) != 0;
            const int js = D < 0 ? -D : 0, je = min(ql, tl - D);
            const int OL = je - js;
            if (ql > 256 || tl > 256) {
                defer = OL >= a.prm.min_len;
            } else if (OL >= a.prm.min_len) {
                const uint64_t qoff = (uint64_t)strand * plane + (uint64_t)q * sw, toff = (uint64_t)t * sw;
                const uint4 qv = __ldg(reinterpret_cast<const uint4 *>(a.reads.bases + qoff) + sub);
                const uint4 tv = __ldg(reinterpret_cast<const uint4 *>(a.reads.bases + toff) + sub);
                uint32_t tw[4];
                realign_row(gmask, sub, tv, D, tw);
                const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
                uint32_t mmw[4], amw[4] = {0u, 0u, 0u, 0u};
                int mm = 0, amb = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t d = qw[i] ^ tw[i];
                    const int base = 64 * sub + 16 * i;
                    const uint32_t vm = range_mask16(js - base, je - base);
                    mmw[i] = (d | (d >> 1)) & vm;
                }
                if (ambf) {
                    const uint4 qa = __ldg(reinterpret_cast<const uint4 *>(a.reads.amb + qoff) + sub);
                    const uint4 ta = __ldg(reinterpret_cast<const uint4 *>(a.reads.amb + toff) + sub);
                    uint32_t taw[4];
                    realign_row(gmask, sub, ta, D, taw);
                    const uint32_t qaw[4] = {qa.x, qa.y, qa.z, qa.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int base = 64 * sub + 16 * i;
                        amw[i] = (qaw[i] | taw[i]) & range_mask16(js - base, je - base);
                        mmw[i] &= ~amw[i];
                        amb += __popc(amw[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) mm += __popc(mmw[i]);
                mm += __shfl_xor_sync(gmask, mm, 1, 4);
                mm += __shfl_xor_sync(gmask, mm, 2, 4);
                if (ambf) {
                    amb += __shfl_xor_sync(gmask, amb, 1, 4);
                    amb += __shfl_xor_sync(gmask, amb, 2, 4);
                }
                const int o = a.prm.max_overhang;
                const bool left_t = D < 0, right_t = (tl - D) < ql;
                // cheap exact reject: the block keeps at least mm - o mismatches and is at most OL long
                if (mm <= (int)__ldg(a.tab.max_k_upto + OL) + o) {
                    int cl = 0, cr = 0, cmm = 0, camb = 0;
                    bool ok = true;
                    if ((mm + amb) > 0 && (left_t || right_t)) {
                        if (mm + amb > a.prm.bad_limit) {
                            ok = false, defer = true;
                        } else {
                            if (aE == 0) aS = je, aE = js;
                            aS = max(aS, js), aE = min(aE, je);
                            uint32_t bw[4], aw4[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) bw[i] = mmw[i] | amw[i], aw4[i] = amw[i];
                            if (left_t) {
                                const int W0 = js >> 4, sh = (js & 15) * 2;
                                const uint32_t b0 = group_word(gmask, bw[0], bw[1], bw[2], bw[3], W0);
                                const uint32_t b1 = group_word(gmask, bw[0], bw[1], bw[2], bw[3], W0 + 1);
                                const uint32_t b2 = group_word(gmask, bw[0], bw[1], bw[2], bw[3], W0 + 2);
                                uint64_t bad = (uint64_t)__funnelshift_r(b0, b1, sh) | (uint64_t)__funnelshift_r(b1, b2, sh) << 32;
                                uint64_t ambw = 0;
                                if (ambf) {
                                    const uint32_t a0 = group_word(gmask, aw4[0], aw4[1], aw4[2], aw4[3], W0);
                                    const uint32_t a1 = group_word(gmask, aw4[0], aw4[1], aw4[2], aw4[3], W0 + 1);
                                    const uint32_t a2 = group_word(gmask, aw4[0], aw4[1], aw4[2], aw4[3], W0 + 2);
                                    ambw = (uint64_t)__funnelshift_r(a0, a1, sh) | (uint64_t)__funnelshift_r(a1, a2, sh) << 32;
                                }
                                const int EL = aS - js;
                                if (EL < 32) bad &= (EL <= 0) ? 0ull : ((1ull << (2 * EL)) - 1ull);
                                const ClipResult r = clip_scan(bad, ambw, a.prm.a, a.prm.b, a.prm.n);
                                cl = r.clip, cmm += r.mm, camb += r.amb;
                            }
                            if (right_t) {
                                const int E0 = je - 32;
                                const int W0 = E0 >> 4, sh = (E0 & 15) * 2;
                                const uint32_t b0 = group_word(gmask, bw[0], bw[1], bw[2], bw[3], W0);
                                const uint32_t b1 = group_word(gmask, bw[0], bw[1], bw[2], bw[3], W0 + 1);
                                const uint32_t b2 = group_word(gmask, bw[0], bw[1], bw[2], bw[3], W0 + 2);
                                uint64_t bad = (uint64_t)__funnelshift_r(b0, b1, sh) | (uint64_t)__funnelshift_r(b1, b2, sh) << 32;
                                uint64_t ambw = 0;
                                if (ambf) {
                                    const uint32_t a0 = group_word(gmask, aw4[0], aw4[1], aw4[2], aw4[3], W0);
                                    const uint32_t a1 = group_word(gmask, aw4[0], aw4[1], aw4[2], aw4[3], W0 + 1);
                                    const uint32_t a2 = group_word(gmask, aw4[0], aw4[1], aw4[2], aw4[3], W0 + 2);
                                    ambw = (uint64_t)__funnelshift_r(a0, a1, sh) | (uint64_t)__funnelshift_r(a1, a2, sh) << 32;
                                }
                                // outward order: the terminal base je-1 becomes distance 1
                                bad = brev_pairs64(bad), ambw = brev_pairs64(ambw);
                                const int ER = je - aE;
                                if (ER < 32) bad &= (ER <= 0) ? 0ull : ((1ull << (2 * ER)) - 1ull);
                                const ClipResult r = clip_scan(bad, ambw, a.prm.a, a.prm.b, a.prm.n);
                                cr = r.clip, cmm += r.mm, camb += r.amb;
                            }
                        }
                    }
                    if (ok && cl + cr <= o) {
                        const int maplen = OL - cl - cr;
                        const int K = mm - cmm, n_ambi = amb - camb;
                        const int blocklen = maplen - n_ambi;
                        if (maplen >= 1 && blocklen >= a.prm.min_len && blocklen >= 1 &&
                            blocklen - K >= (int)__ldg(a.tab.min_matches + blocklen) &&
                            cl + cr <= (int)__ldg(a.tab.max_clip + maplen)) {
                            accept = true;
                            rec.q = q, rec.t = t;
                            rec.ds = ((uint32_t)(D + HLO_DIAG_BIAS) << 1) | strand;
                            rec.kb = (uint32_t)K | ((uint32_t)blocklen << 16);
                        }
                    }
                }
            }
        }
        // stream compaction: one atomicAdd per warp for the accepted, one for the deferred
        const bool lead = sub == 0;
        const unsigned acc_b = __ballot_sync(0xffffffffu, lead && accept);
        const unsigned def_b = __ballot_sync(0xffffffffu, lead && defer);
        if (acc_b) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&a.counters[0], (unsigned long long)__popc(acc_b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lead && accept)
                reinterpret_cast<uint4 *>(a.out)[base + __popc(acc_b & lt_mask)] = make_uint4(rec.q, rec.t, rec.ds, rec.kb);
        }
        if (def_b) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&a.counters[1], (unsigned long long)__popc(def_b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lead && defer) a.deferred[base + __popc(def_b & lt_mask)] = (uint32_t)c;
        }
    }
}

void launch_verify_fast(const VerifyArgs &a, cudaStream_t s, LaunchCounter *lc) {
    if (a.n == 0) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t want = (a.n * 4 + 255) / 256;
    uint64_t grid = (uint64_t)sms * 8 * 4;  // persistent-ish: a few waves of 256-thread blocks
    if (grid > want) grid = want;
    verify_fast_kernel<<<(unsigned)grid, 256, 0, s>>>(a);
    if (lc) lc->total++, lc->verify++;
}

// ------------------------------------------------------------------------------------------------
// exact path: one thread per candidate, base by base (any read length, any number of bad columns)

__device__ __forceinline__ int get_base(const uint32_t *row, int j) { return (int)((row[j >> 4] >> ((j & 15) * 2)) & 3u); }

__global__ void __launch_bounds__(128) verify_exact_kernel(VerifyArgs a, const uint32_t *__restrict__ idx, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t c = idx ? idx[i] : i;
    const hlo_cand cd = a.cands[c];
    const uint32_t q = cd.q, t = cd.t;
    const int D = cd.diag;
    const uint32_t strand = cd.strand_anchor & 1u;
    int aS = (int)((cd.strand_anchor >> 1) & 0x7fffu), aE = (int)(cd.strand_anchor >> 16);
    const uint32_t lq = a.reads.len[q], lt = a.reads.len[t];
    const int ql = (int)(lq & HLO_LEN_MASK), tl = (int)(lt & HLO_LEN_MASK);
    const bool ambf = ((lq | lt) & HLO_AMB_FLAG) != 0;
    const int js = D < 0 ? -D : 0, je = min(ql, tl - D);
    if (je <= js) return;
    const uint64_t plane = (uint64_t)a.reads.n_reads * a.reads.stride_w;
    const uint32_t *Q = a.reads.bases + (uint64_t)strand * plane + (uint64_t)q * a.reads.stride_w;
    const uint32_t *T = a.reads.bases + (uint64_t)t * a.reads.stride_w;
    const uint32_t *QA = a.reads.amb + (uint64_t)strand * plane + (uint64_t)q * a.reads.stride_w;
    const uint32_t *TA = a.reads.amb + (uint64_t)t * a.reads.stride_w;
    if (aE == 0) aS = je, aE = js;
    aS = max(aS, js), aE = min(aE, je);
    const int A = a.prm.a, B = a.prm.b, N = a.prm.n;
    int cl = 0, cr = 0;
    // column class: 0 match, 1 mismatch, 2 ambiguous
#define COLCLASS(j) ((ambf && (get_base(QA, (j)) | get_base(TA, (j) + D))) ? 2 : (get_base(Q, (j)) != get_base(T, (j) + D)))
    if (D < 0) {
        int s = 0, best = 0, best_n = 0, cnt = 0;
        for (int j = aS - 1; j >= js; --j) {
            const int cc = COLCLASS(j);
            s += cc == 0 ? A : (cc == 1 ? -B : -N);
            ++cnt;
            if (s > best) best = s, best_n = cnt;
        }
        cl = (aS - js) - best_n;
    }
    if (tl - D < ql) {
        int s = 0, best = 0, best_n = 0, cnt = 0;
        for (int j = aE; j < je; ++j) {
            const int cc = COLCLASS(j);
            s += cc == 0 ? A : (cc == 1 ? -B : -N);
            ++cnt;
            if (s > best) best = s, best_n = cnt;
        }
        cr = (je - aE) - best_n;
    }
    const int bs = js + cl, be = je - cr;
    int K = 0, n_ambi = 0;
    for (int j = bs; j < be; ++j) {
        const int cc = COLCLASS(j);
        K += cc == 1, n_ambi += cc == 2;
    }
#undef COLCLASS
    const int maplen = be - bs, blocklen = maplen - n_ambi;
    if (maplen < 1 || blocklen < 1 || blocklen < a.prm.min_len) return;
    if (blocklen - K < (int)a.tab.min_matches[blocklen]) return;
    if (cl + cr > (int)a.tab.max_clip[maplen]) return;
    const unsigned long long p = atomicAdd(&a.counters[0], 1ULL);
    Rec16 rec;
    rec.q = q, rec.t = t;
    rec.ds = ((uint32_t)(D + HLO_DIAG_BIAS) << 1) | strand;
    rec.kb = (uint32_t)K | ((uint32_t)blocklen << 16);
    a.out[p] = rec;
}

void launch_verify_exact(const VerifyArgs &a, const uint32_t *idx, uint64_t n, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    verify_exact_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a, idx, n);
    if (lc) lc->total++, lc->verify++;
}

}  // namespace hlo
