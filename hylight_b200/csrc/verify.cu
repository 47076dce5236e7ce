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
// verify_fast_kernel<W, FULLW>: ONE THREAD PER CANDIDATE.  The internal candidate (CandI) carries both read
// lengths, so the geometry is known before a row is touched:
//   1. overlaps shorter than min_len leave at once; pairs with an ambiguous base or with BOTH ends
//      target-limited (query contains target, unequal lengths) go to verify_exact_kernel;
//   2. a target-limited RIGHT end is moved to position 0 by switching both reads to their
//      reverse-complement plane (D' = tl - ql - D), so the end rule only ever looks at bit 0.. of the
//      mismatch words;
//   3. the read that starts first is X (compared from its base 0), the other is Y shifted by S = |D|:
//      NV LDG.128 per row (Y from the 16-byte vector that holds base S), a two-stage word select,
//      funnel shifts for the residual 2-bit offset, XOR, (x|x>>1)&0x5555.. fold, popcount;
//   4. thresholds from the host-built integer tables (shared memory), end rule on the outermost 32
//      bases; accepted 16-byte records are staged PER WARP in shared memory (positions from a ballot) and
//      leave with one global atomicAdd per ~100 records and coalesced stores (a global atomic per warp
//      iteration on one address was the top stall of the first version: profiles/r1b; a per-block stage
//      cost a barrier per iteration).  The next candidate is loaded one iteration ahead.
// (The north star sketches a warp per pair; one thread per candidate measured 3x faster — a 150-base row is
// three 16-byte vectors, too little for 32 lanes — and is what this file does.  DESIGN.md §0.)
// Bound: integer (ALU) pipe and L1 wavefronts (one per 16-byte vector of the randomly addressed row);
// rows are L2-resident.  Algorithmic bytes per candidate = 16 + 2*ceil(OL/4) (when OL >= min_len) + 16
// per accepted record (SURVEY.md §8d counts 32 B for the expanded record; the expansion is records.cu).
#include <algorithm>

#include "kernels.cuh"

namespace hlo {

struct ClipResult {
    int clip, mm;
};

// end rule on the outermost 32 bases of a target-limited end; bit 2(p-1) of (lo | hi<<32) set <=> the base at
// outward distance p (1 = the terminal base) is a mismatch.  Returns the clipped length and the
// mismatches inside it: the largest p that minimises R(p) = A*p - (A+B)*mismatches(p) over R <= 0.
__device__ __forceinline__ ClipResult clip_scan32(uint32_t lo, uint32_t hi, int A, int B) {
    ClipResult r{0, 0};
    int best = 0, cm = 0;
    while (lo) {
        const int p = ((__ffs((int)lo) - 1) >> 1) + 1;
        ++cm;
        const int R = A * p - (A + B) * cm;
        if (R <= best) best = R, r.clip = p, r.mm = cm;
        lo &= lo - 1;
    }
    while (hi) {
        const int p = ((__ffs((int)hi) - 1) >> 1) + 17;
        ++cm;
        const int R = A * p - (A + B) * cm;
        if (R <= best) best = R, r.clip = p, r.mm = cm;
        hi &= hi - 1;
    }
    return r;
}

// low n bits set, n clamped to [0, 32]
__device__ __forceinline__ uint32_t low_mask(int n) {
    n = max(n, 0);
    return __funnelshift_lc(0xffffffffu, 0u, (uint32_t)n);  // (0:ones) << min(n,32), upper word
}

constexpr int VF_THREADS = 128;
constexpr int VF_WARPS = VF_THREADS / 32;
constexpr int VF_STAGE = 128;  // records staged per WARP before one flush (one global atomic per flush)

// W = 32-bit data words per row actually compared (ceil(max_len / 16)); the row stride is NV = ceil(W/4)
// 16-byte vectors.  FULLW = words that every overlap of at least min_len bases covers completely (no mask).
// PR = hlo_params.pair_rule (1: rows that fail only the overhang test are emitted, flagged)
template <int W, int FULLW, bool PR>
__global__ void __launch_bounds__(VF_THREADS) verify_fast_kernel(VerifyArgs a) {
    constexpr int NV = (W + 3) / 4, NW = 4 * NV;
    __shared__ uint16_t s_tab[3 * 1024];
    __shared__ uint4 s_out[VF_WARPS][VF_STAGE];
    // thresholds into shared memory (3 tables, max_len + 1 <= 1024 entries of each; max_len = this batch's longest read)
    const int tl1 = a.tab.max_len + 1;
    for (int i = threadIdx.x; i < tl1; i += VF_THREADS)
        s_tab[i] = a.tab.min_matches[i], s_tab[tl1 + i] = a.tab.max_clip[i], s_tab[2 * tl1 + i] = a.tab.max_k_upto[i];
    __syncthreads();
    const uint16_t *t_min_matches = s_tab, *t_max_clip = s_tab + tl1, *t_max_k = s_tab + 2 * tl1;
    uint4 *w_out = s_out[threadIdx.x >> 5];
    uint32_t w_cnt = 0;  // records staged by this warp (warp-uniform)

    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint64_t n_threads = (uint64_t)gridDim.x * VF_THREADS;
    const uint64_t plane4 = (uint64_t)a.reads.n_reads * NV;  // uint4 per plane
    const uint4 *rows = reinterpret_cast<const uint4 *>(a.reads.bases);
    const uint4 *cands = reinterpret_cast<const uint4 *>(a.cands);
    // the candidate count is the pair generator's device counter (it may exceed the buffer: then the host reruns)
    const unsigned long long n_dev = *a.n_ptr;
    const uint64_t a_n = n_dev < a.n_cap ? n_dev : a.n_cap;
    const uint64_t n_round = (a_n + 31) & ~31ull;  // whole warps iterate together
    const int o = a.prm.max_overhang, min_len = a.prm.min_len, defer_mm = a.prm.defer_mm;
    uint64_t c = (uint64_t)blockIdx.x * VF_THREADS + threadIdx.x;
    if (c >= n_round) return;  // (after the only block barrier)
    uint4 cv_next = make_uint4(0u, 0u, 0u, 0u);
    if (c < a_n) cv_next = __ldg(cands + c);
    for (; c < n_round; c += n_threads) {
        bool accept = false, defer = false;
        uint4 rec = make_uint4(0u, 0u, 0u, 0u);
        const uint4 cv = cv_next;
        const bool live = c < a_n;
        if (c + n_threads < a_n) cv_next = __ldg(cands + c + n_threads);  // next iteration's candidate, in flight early
        if (live) {
            const uint32_t q = cv.x, t = cv.y;
            const int D = (int)(cv.z & 0xffffu) - 32768;
            const int ql = (int)((cv.z >> 16) & 0x7fffu), tl = (int)(cv.w & 0x7fffu);
            const uint32_t strand = cv.z >> 31;
            const int OL = min(ql, tl - D) - max(-D, 0);
            if (OL >= min_len) {
                const bool left_t = D < 0, right_t = (tl - D) < ql;
                if (((cv.w >> 29) & 1u) || (left_t && right_t)) {
                    defer = true;
                } else {
                    // orientation in which a target-limited end, if any, is the LEFT end
                    const uint32_t flip = right_t ? 1u : 0u;
                    const int Dn = flip ? tl - ql - D : D;
                    const uint4 *Qrow = rows + (uint64_t)(strand ^ flip) * plane4 + (uint64_t)q * NV;
                    const uint4 *Trow = rows + (uint64_t)flip * plane4 + (uint64_t)t * NV;
                    const bool neg = Dn < 0;  // then the target starts first and its end is the limited one
                    const int S = neg ? -Dn : Dn;
                    const uint4 *Xrow = neg ? Trow : Qrow;
                    const uint4 *Yrow = (neg ? Qrow : Trow) + (S >> 6);
                    // a shifted row read may run past its row: into the next rows, or (last row of the plane) the padding
                    HLO_DBG_ASSERT(q < a.reads.n_reads && t < a.reads.n_reads && (S >> 6) <= NV && NV <= HLO_PAD_WORDS / 4);
                    uint32_t x[NW], y[NW + 1];
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const uint4 xv = __ldg(Xrow + v);
                        x[4 * v] = xv.x, x[4 * v + 1] = xv.y, x[4 * v + 2] = xv.z, x[4 * v + 3] = xv.w;
                        // vectors of Y past the row are never compared (they lie beyond the overlap); reading
                        // them stays inside the padded allocation
                        const uint4 yv = __ldg(Yrow + v);
                        y[4 * v] = yv.x, y[4 * v + 1] = yv.y, y[4 * v + 2] = yv.z, y[4 * v + 3] = yv.w;
                    }
                    y[NW] = 0u;
                    // residual word shift (S/16 mod 4) by two select stages, then the 2-bit funnel shift
                    const bool w2 = (S & 32) != 0, w1 = (S & 16) != 0;
#pragma unroll
                    for (int i = 0; i + 2 <= NW; ++i) y[i] = w2 ? y[i + 2] : y[i];
#pragma unroll
                    for (int i = 0; i + 1 <= NW; ++i) y[i] = w1 ? y[i + 1] : y[i];
                    const uint32_t b2 = (uint32_t)(S & 15) * 2u;
                    uint32_t f[W];
                    int mm = 0;
#pragma unroll
                    for (int i = 0; i < W; ++i) {
                        const uint32_t ys = __funnelshift_r(y[i], y[i + 1], b2);
                        const uint32_t d = x[i] ^ ys;
                        f[i] = (d | (d >> 1)) & 0x55555555u;
                        if (i >= FULLW) f[i] &= low_mask(2 * OL - 32 * i);
                    }
#pragma unroll
                    for (int i = 0; i + 1 < W; i += 2) mm += __popc(f[i] | (f[i + 1] << 1));
                    if (W & 1) mm += __popc(f[W - 1]);
                    // cheap exact reject: the block is at most OL long and keeps at least mm - o mismatches.  pair_rule 1 also
                    // wants the rows that fail the overhang test, and those may clip away any number of mismatches — but never
                    // more than one per clipped base, and a block of min_len bases must remain: a row with more than
                    // max_k + (OL - min_len) mismatches cannot pass; between a window's worth and that bound the exact kernel decides
                    const int max_k = (int)t_max_k[OL];
                    if (mm <= max_k + (PR ? OL - min_len : o)) {
                        int cl = 0, cmm = 0;
                        bool ok = true;
                        if (mm >= defer_mm || (PR && mm > max_k + 33)) {
                            ok = false, defer = true;  // enough bad columns to cost a query-limited end its end bonus: the exact rule decides
                        } else if (mm > 0 && (left_t || right_t)) {
                            if (mm > a.prm.bad_limit) {
                                ok = false, defer = true;  // a 32-base window no longer decides the end rule
                            } else {
                                const int REG = (int)((flip ? (cv.w >> 22) : (cv.w >> 15)) & 0x7fu);
                                uint32_t lo = f[0], hi = W > 1 ? f[1] : 0u;
                                if (REG < 32) lo &= low_mask(2 * REG), hi &= low_mask(2 * REG - 32);
                                const ClipResult r = clip_scan32(lo, hi, a.prm.a, a.prm.b);
                                cl = r.clip, cmm = r.mm;
                            }
                        }
                        if (ok && (cl <= o || PR)) {
                            const int maplen = OL - cl;  // no ambiguous base on this path: blocklen = maplen
                            const int K = mm - cmm;
                            if (maplen >= min_len && maplen >= 1 && maplen - K >= (int)t_min_matches[maplen]) {
                                const bool oh_ok = cl <= o && cl <= (int)t_max_clip[maplen];
                                if (oh_ok || PR) {
                                    accept = true;
                                    rec = make_uint4(q, t, ((uint32_t)(D + HLO_DIAG_BIAS) << 1) | strand,
                                                     (uint32_t)K | (oh_ok ? 0u : HLO_KB_OVERHANG_FAIL) | ((uint32_t)maplen << 16));
                                }
                            }
                        }
                    }
                }
            }
        }
        // stream compaction: accepted records are staged per warp in shared memory (positions from a ballot, no
        // atomics) and a full stage leaves with ONE global atomicAdd and coalesced 16-byte stores; no block barrier
        const unsigned acc_b = __ballot_sync(0xffffffffu, accept);
        HLO_DBG_ASSERT(w_cnt + __popc(acc_b) <= VF_STAGE);
        if (accept) w_out[w_cnt + __popc(acc_b & lt_mask)] = rec;
        w_cnt += __popc(acc_b);
        const unsigned def_b = __ballot_sync(0xffffffffu, defer);
        if (def_b) {  // rare
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.def_counter, (unsigned long long)__popc(def_b));
            base = __shfl_sync(0xffffffffu, base, 0);
            HLO_DBG_ASSERT(base + __popc(def_b) <= a.n_cap);
            if (defer) a.deferred[base + __popc(def_b & lt_mask)] = (uint32_t)c;  // (one slot per candidate: cannot overflow)
        }
        const bool last = c + n_threads >= n_round;
        if (w_cnt > VF_STAGE - 32 || (last && w_cnt)) {  // the next iteration might not fit
            __syncwarp();
            unsigned long long gb = 0;
            if (lane == 0) gb = atomicAdd(a.acc_counter, (unsigned long long)w_cnt);
            gb = __shfl_sync(0xffffffffu, gb, 0);
            for (uint32_t i = lane; i < w_cnt; i += 32)
                if (gb + i < a.out_cap) reinterpret_cast<uint4 *>(a.out)[gb + i] = w_out[i];  // past the end: counted only
            __syncwarp();
            w_cnt = 0;
        }
    }
}

bool verify_fast_supported(const DevReads &r, const VerifyTables &tab) {
    return r.max_len >= 1 && r.max_len <= 320 && tab.max_len + 1 <= 1024 && r.n_reads > 0;
}

template <int W, int F, bool PR>
static void launch_vf3(const VerifyArgs &a, int sms, cudaStream_t s) {
    int per_sm = 8;  // persistent grid: as many blocks as are resident at once (each builds its tables once)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, verify_fast_kernel<W, F, PR>, VF_THREADS, 0);
    const uint64_t want = (a.n_cap + VF_THREADS - 1) / VF_THREADS;
    const uint64_t grid = std::min<uint64_t>(want, (uint64_t)sms * (per_sm > 0 ? per_sm : 1));
    verify_fast_kernel<W, F, PR><<<(unsigned)grid, VF_THREADS, 0, s>>>(a);
}
template <int W, int F>
static void launch_vf2(const VerifyArgs &a, int sms, cudaStream_t s) {
    if (a.prm.pair_rule) launch_vf3<W, F, true>(a, sms, s);
    else launch_vf3<W, F, false>(a, sms, s);
}
template <int W>
static void launch_vf(const VerifyArgs &a, int sms, cudaStream_t s) {
    constexpr int F3 = W >= 3 ? 3 : 0;
    if (a.prm.min_len >= 48 && F3) launch_vf2<W, F3>(a, sms, s);
    else launch_vf2<W, 0>(a, sms, s);
}

void launch_verify_fast(const VerifyArgs &a, cudaStream_t s, LaunchCounter *lc) {
    if (a.n_cap == 0) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int W = (int)((a.reads.max_len + 15) / 16);
    switch (W) {
#define VF_CASE(w) case w: launch_vf<w>(a, sms, s); break;
        VF_CASE(1) VF_CASE(2) VF_CASE(3) VF_CASE(4) VF_CASE(5) VF_CASE(6) VF_CASE(7) VF_CASE(8) VF_CASE(9) VF_CASE(10)
        VF_CASE(11) VF_CASE(12) VF_CASE(13) VF_CASE(14) VF_CASE(15) VF_CASE(16) VF_CASE(17) VF_CASE(18) VF_CASE(19)
        default: launch_vf<20>(a, sms, s); break;
#undef VF_CASE
    }
    if (lc) lc->total++, lc->verify++;
}

// ------------------------------------------------------------------------------------------------
// exact path: one thread per candidate, base by base (any read length, ambiguous bases, both ends
// target-limited, any number of bad columns)

__device__ __forceinline__ int get_base(const uint32_t *row, int j) { return (int)((row[j >> 4] >> ((j & 15) * 2)) & 3u); }

__device__ __forceinline__ void verify_exact_one(const VerifyArgs &a, uint64_t c) {
    const CandI cd = a.cands[c];
    const uint32_t q = cd.q, t = cd.t;
    const int D = (int)(cd.z & 0xffffu) - 32768;
    const uint32_t strand = cd.z >> 31;
    const int ql = (int)((cd.z >> 16) & 0x7fffu), tl = (int)(cd.w & 0x7fffu);
    int EL = (int)((cd.w >> 15) & 0x7fu), ER = (int)((cd.w >> 22) & 0x7fu);
    if (cd.w & HLO_CAND_FAR) {  // a region longer than the 7-bit fields hold: the true lengths
        const uint32_t far = a.cand_far[c];
        EL = (int)(far & 0xffffu), ER = (int)(far >> 16);
    }
    const bool ambf = ((cd.w >> 29) & 1u) != 0;
    const int js = D < 0 ? -D : 0, je = min(ql, tl - D);
    if (je <= js) return;
    HLO_DBG_ASSERT(q < a.reads.n_reads && t < a.reads.n_reads && (uint32_t)ql <= a.reads.stride_w * 16 && (uint32_t)tl <= a.reads.stride_w * 16 &&
                   je + D <= tl && js + D >= 0);
    const uint64_t plane = (uint64_t)a.reads.n_reads * a.reads.stride_w;
    const uint32_t *Q = a.reads.bases + (uint64_t)strand * plane + (uint64_t)q * a.reads.stride_w;
    const uint32_t *T = a.reads.bases + (uint64_t)t * a.reads.stride_w;
    const uint32_t *QA = a.reads.amb + (uint64_t)strand * plane + (uint64_t)q * a.reads.stride_w;
    const uint32_t *TA = a.reads.amb + (uint64_t)t * a.reads.stride_w;
    // the regions between the overlap ends and the outermost anchors
    const int aS = min(js + EL, je), aE = max(je - ER, js);
    const int A = a.prm.a, B = a.prm.b, N = a.prm.n;
    int cl = 0, cr = 0;
    // column class: 0 match, 1 mismatch, 2 ambiguous
#define COLCLASS(j) ((ambf && (get_base(QA, (j)) | get_base(TA, (j) + D))) ? 2 : (get_base(Q, (j)) != get_base(T, (j) + D)))
    // Each end is a ksw2 extension from the outermost anchor outward (oracle/hlo_oracle.c:orc_extend_end): s = cumulative
    // score, best = its largest value so far (strict >: the earliest position), stop when best - s > zdrop.  An end where
    // the QUERY ends first reaches the read end iff not z-dropped and s_end + end_bonus > best; every other end stops at
    // the position of best.
    {
        int s = 0, best = 0, best_n = 0, cnt = 0;
        bool dropped = false;
        for (int j = aS - 1; j >= js; --j) {
            const int cc = COLCLASS(j);
            s += cc == 0 ? A : (cc == 1 ? -B : -N);
            ++cnt;
            if (s > best) best = s, best_n = cnt;
            else if (best - s > a.prm.zdrop) {
                dropped = true;
                break;
            }
        }
        const bool reach = D >= 0 && !dropped && s + a.prm.end_bonus > best;
        cl = reach ? 0 : (aS - js) - best_n;
    }
    {
        int s = 0, best = 0, best_n = 0, cnt = 0;
        bool dropped = false;
        for (int j = aE; j < je; ++j) {
            const int cc = COLCLASS(j);
            s += cc == 0 ? A : (cc == 1 ? -B : -N);
            ++cnt;
            if (s > best) best = s, best_n = cnt;
            else if (best - s > a.prm.zdrop) {
                dropped = true;
                break;
            }
        }
        const bool reach = !(tl - D < ql) && !dropped && s + a.prm.end_bonus > best;
        cr = reach ? 0 : (je - aE) - best_n;
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
    if (maplen > a.tab.max_len) return;
    if (blocklen - K < (int)a.tab.min_matches[blocklen]) return;
    const bool oh_ok = cl + cr <= (int)a.tab.max_clip[maplen];
    if (!oh_ok && !a.prm.pair_rule) return;
    const unsigned long long p = atomicAdd(a.acc_counter, 1ULL);
    if (p >= a.out_cap) return;  // counted only
    Rec16 rec;
    rec.q = q, rec.t = t;
    rec.ds = ((uint32_t)(D + HLO_DIAG_BIAS) << 1) | strand;
    rec.kb = (uint32_t)K | (oh_ok ? 0u : HLO_KB_OVERHANG_FAIL) | ((uint32_t)blocklen << 16);
    a.out[p] = rec;
}

// idx == nullptr: candidates [0, min(*a.n_ptr, a.n_cap)); else the *a.def_counter candidates listed in idx
__global__ void __launch_bounds__(128) verify_exact_kernel(VerifyArgs a, const uint32_t *__restrict__ idx) {
    unsigned long long n = idx ? *a.def_counter : *a.n_ptr;
    if (n > a.n_cap) n = a.n_cap;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        verify_exact_one(a, idx ? idx[i] : i);
}

// all_cands: every candidate (option "verify_exact", long reads); else the deferred list of the fast kernel (usually a
// handful: a small grid that strides)
void launch_verify_exact(const VerifyArgs &a, bool all_cands, cudaStream_t s, LaunchCounter *lc) {
    if (a.n_cap == 0) return;
    const uint64_t want = all_cands ? (a.n_cap + 127) / 128 : 148 * 4;
    const unsigned grid = (unsigned)std::min<uint64_t>(want, 148 * 64);
    verify_exact_kernel<<<grid, 128, 0, s>>>(a, all_cands ? nullptr : a.deferred);
    if (lc) lc->total++, lc->verify++;
}

// hlo_cand (ABI) -> CandI, for caller-supplied candidate lists
__global__ void cands_to_internal_kernel(const hlo_cand *__restrict__ in, uint64_t n, DevReads r, CandI *__restrict__ out,
                                         uint32_t *__restrict__ far) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const hlo_cand c = in[i];
    const uint32_t lq = r.len[c.q], lt = r.len[c.t];
    out[i] = make_candi(c.q, c.t, c.diag, c.strand_anchor & 1u, lq & 0x7fffu, lt & 0x7fffu, (int)((c.strand_anchor >> 1) & 0x7fffu),
                        (int)(c.strand_anchor >> 16), (lq | lt) >> 31, far + i);
}
void launch_cands_to_internal(const hlo_cand *in, uint64_t n, const DevReads &r, CandI *out, uint32_t *far, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    cands_to_internal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, n, r, out, far);
    if (lc) lc->total++;
}

}  // namespace hlo
