// sketch.cu — (w,k) minimizer sketch, bit-exact with minimap2's mm_sketch() (non-HPC branch).
//
// minimap2 is not in /root/reference (un-vendored dependency, call site polyte.tune_params.py:510,
// "-k 21 -w 11"); the algorithm restated here is the one in oracle/hlo_oracle.c:orc_sketch, and the
// parity tests compare the two tuple for tuple: x = hash64(canonical k-mer) << 8 | span,
// y = rid << 32 | last_pos << 1 | strand.
//
// One thread per read (taken in name order, see below) runs the sequential window automaton exactly (ties, the first-window special
// case, ambiguous-base resets and the end-of-read flush are all order dependent), reading the read's
// packed row; a count pass sizes the output, a write pass fills it, so the output is in minimap2's
// emission order.  Integer-pipe bound: ~150 ALU ops per base; algorithmic bytes per read =
// ceil(L/4) in + 16 B per minimizer out.
#include <type_traits>

#include "kernels.cuh"

namespace hlo {

#define HLO_MAX_W HLO_MAX_W_PUBLIC
#ifndef HLO_SK_BLOCKS
#define HLO_SK_BLOCKS 6  // resident blocks of sketch_fast_kernel<11> per SM the register budget is set for (6: 79 registers, no spill)
#endif

// minimap2's hash64 (Thomas Wang's 64-bit mix, every step masked to 2k bits).  The shift-add steps are
// written as multiplications by constants (key + (key << 3) + (key << 8) = key * 265, ...): they run on
// the FMA pipe as IMADs and leave the ALU pipe, which bounds the sketch kernels, to the xor-shifts.
__device__ __forceinline__ uint64_t mm_hash64(uint64_t key, uint64_t mask) {
    key = (key * 2097151ULL - 1ULL) & mask;  // ~key + (key << 21)  (~key = -key - 1)
    key = key ^ key >> 24;
    key = (key * 265ULL) & mask;             // key + (key << 3) + (key << 8)
    key = key ^ key >> 14;
    key = (key * 21ULL) & mask;              // key + (key << 2) + (key << 4)
    key = key ^ key >> 28;
    key = (key * 2147483649ULL) & mask;      // key + (key << 31)
    return key;
}

struct MzOut {
    uint64_t *x;     // hash << 8 | span (nullptr: not wanted)
    uint2 *rr;       // (read, pos << 1 | strand)
    uint64_t *keys;  // sort key = cluster << hash_bits | hash, with vals = own index (nullptr: not wanted)
    uint32_t *vals;
    int hash_bits;
};

template <bool WRITE>
__device__ __forceinline__ void mz_push(uint32_t &n, uint64_t x, uint32_t rp, uint32_t rid, uint32_t base, const MzOut &o,
                                        uint64_t chi = 0) {
    if (WRITE) {
        if (o.x) o.x[base + n] = x;
        o.rr[base + n] = make_uint2(rid, rp);
        if (o.keys) o.keys[base + n] = (x >> 8) | chi, o.vals[base + n] = base + n;
    }
    ++n;
}

template <bool WRITE>
__device__ __forceinline__ void sketch_read(const DevReads &r, int w, int k, uint32_t *__restrict__ count,
                                            const uint32_t *__restrict__ off, const MzOut &o, uint32_t xi) {
    // read perm[xi]: the minimizer arrays are laid out in the strcmp order of the read names inside each
    // cluster, so that after the stable hash sort every seed bucket is ordered the way minimap2's -X rule
    // compares names (pairgen.cu)
    const uint32_t read = __ldg(r.perm + xi);
    const uint32_t lenf = r.len[read];
    const int len = (int)(lenf & HLO_LEN_MASK);
    const bool has_amb = (lenf & HLO_AMB_FLAG) != 0;
    const uint32_t *row = r.bases + (uint64_t)read * r.stride_w;
    const uint32_t *arow = r.amb + (uint64_t)read * r.stride_w;
    const uint64_t shift1 = 2 * (uint64_t)(k - 1), mask = (1ULL << 2 * k) - 1;
    const uint64_t INF = ~0ULL;
    uint64_t kmer0 = 0, kmer1 = 0;
    uint64_t buf_x[HLO_MAX_W];
    uint32_t buf_y[HLO_MAX_W];
    uint64_t min_x = INF;
    uint32_t min_y = 0xffffffffu;
    const uint32_t base = WRITE ? off[xi] : 0u;
    const uint64_t chi = WRITE && o.keys ? (uint64_t)r.cidx[read] << o.hash_bits : 0ull;  // cluster-major sort key
    uint32_t n = 0;
    int l = 0, buf_pos = 0, min_pos = 0;
#pragma unroll
    for (int j = 0; j < HLO_MAX_W; ++j) buf_x[j] = INF, buf_y[j] = 0xffffffffu;
    uint32_t word = 0, aword = 0;
    for (int i = 0; i < len; ++i) {
        if ((i & 15) == 0) {
            word = __ldg(row + (i >> 4));
            aword = has_amb ? __ldg(arow + (i >> 4)) : 0u;
        }
        const uint32_t c = word & 3u, amb = aword & 1u;
        word >>= 2, aword >>= 2;
        uint64_t info_x = INF;
        uint32_t info_y = 0xffffffffu;
        if (!amb) {
            const int kmer_span = l + 1 < k ? l + 1 : k;
            kmer0 = (kmer0 << 2 | (uint64_t)c) & mask;
            kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
            if (kmer0 == kmer1) continue;  // strand-symmetric k-mer
            const uint32_t z = kmer0 < kmer1 ? 0u : 1u;
            ++l;
            if (l >= k) {
                info_x = mm_hash64(z ? kmer1 : kmer0, mask) << 8 | (uint64_t)kmer_span;
                info_y = (uint32_t)i << 1 | z;
            }
        } else
            l = 0;
        buf_x[buf_pos] = info_x, buf_y[buf_pos] = info_y;
        if (l == w + k - 1 && min_x != INF) {  // first full window: identical k-mers were not stored yet
            for (int j = buf_pos + 1; j < w; ++j)
                if (min_x == buf_x[j] && buf_y[j] != min_y) mz_push<WRITE>(n, buf_x[j], buf_y[j], read, base, o, chi);
            for (int j = 0; j < buf_pos; ++j)
                if (min_x == buf_x[j] && buf_y[j] != min_y) mz_push<WRITE>(n, buf_x[j], buf_y[j], read, base, o, chi);
        }
        if (info_x <= min_x) {  // a new minimum; write the old one
            if (l >= w + k && min_x != INF) mz_push<WRITE>(n, min_x, min_y, read, base, o, chi);
            min_x = info_x, min_y = info_y, min_pos = buf_pos;
        } else if (buf_pos == min_pos) {  // the old minimum left the window
            if (l >= w + k - 1 && min_x != INF) mz_push<WRITE>(n, min_x, min_y, read, base, o, chi);
            min_x = INF;
            for (int j = buf_pos + 1; j < w; ++j)
                if (min_x >= buf_x[j]) min_x = buf_x[j], min_y = buf_y[j], min_pos = j;  // >= : rightmost smallest
            for (int j = 0; j <= buf_pos; ++j)
                if (min_x >= buf_x[j]) min_x = buf_x[j], min_y = buf_y[j], min_pos = j;
            if (l >= w + k - 1 && min_x != INF) {
                for (int j = buf_pos + 1; j < w; ++j)
                    if (min_x == buf_x[j] && min_y != buf_y[j]) mz_push<WRITE>(n, buf_x[j], buf_y[j], read, base, o, chi);
                for (int j = 0; j <= buf_pos; ++j)
                    if (min_x == buf_x[j] && min_y != buf_y[j]) mz_push<WRITE>(n, buf_x[j], buf_y[j], read, base, o, chi);
            }
        }
        if (++buf_pos == w) buf_pos = 0;
    }
    if (min_x != INF) mz_push<WRITE>(n, min_x, min_y, read, base, o, chi);
    if (!WRITE) count[xi] = n;
}

// list != nullptr: the reads listed there (x indices); their number is *n_list_ptr (a device counter, clamped to the
// batch size).  cap_mz: entries the output arrays hold — a batch with more minimizers writes nothing (the host sees the
// count and reruns with larger arrays).
template <bool WRITE>
__global__ void __launch_bounds__(128) sketch_kernel(DevReads r, int w, int k, uint32_t *__restrict__ count,
                                                     const uint32_t *__restrict__ off, MzOut o,
                                                     const uint32_t *__restrict__ list, const unsigned long long *__restrict__ n_list_ptr,
                                                     uint32_t cap_mz) {
    uint32_t n = r.n_reads;
    if (list) {
        const unsigned long long nl = *n_list_ptr;
        n = nl < (unsigned long long)r.n_reads ? (uint32_t)nl : r.n_reads;
    }
    if (WRITE && off[r.n_reads] > cap_mz) return;
    for (uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x; ti < n; ti += gridDim.x * blockDim.x)
        sketch_read<WRITE>(r, w, k, count, off, o, list ? list[ti] : ti);
}

void launch_sketch_count(const DevReads &r, int w, int k, uint32_t *count, cudaStream_t s, LaunchCounter *lc) {
    if (r.n_reads == 0) return;
    MzOut o{nullptr, nullptr, nullptr, nullptr, 0};
    sketch_kernel<false><<<(r.n_reads + 127) / 128, 128, 0, s>>>(r, w, k, count, nullptr, o, nullptr, nullptr, 0u);
    if (lc) lc->total++;
}

void launch_sketch_write(const DevReads &r, int w, int k, const uint32_t *off, uint64_t *mz_x, uint2 *mz_rr, uint64_t *keys, uint32_t *vals,
                         int hash_bits, const uint32_t *list, const unsigned long long *n_list_ptr, uint32_t cap_mz, cudaStream_t s,
                         LaunchCounter *lc) {
    if (r.n_reads == 0) return;
    MzOut o{mz_x, mz_rr, keys, vals, hash_bits};
    // a listed launch usually has nothing to do (no read overflowed its strip): a small grid that strides
    const unsigned grid = list ? 148u : (r.n_reads + 127) / 128;
    sketch_kernel<true><<<grid, 128, 0, s>>>(r, w, k, nullptr, off, o, list, n_list_ptr, cap_mz);
    if (lc) lc->total++;
}

// ------------------------------------------------------------------------------------------------
// fast path: window size WW known at compile time and k odd (no k-mer equals its own reverse complement,
// so the "strand-symmetric k-mer" skip of mm_sketch never fires and buffer slot = position mod w).
// The step loop is unrolled WW times: the ring buffer lives in registers with static indices, the rescan
// when the minimum leaves the window is a straight-line select chain, and nothing touches local memory.
// ONE pass: the automaton writes its minimizers into a per-read scratch strip of `cap` entries and counts
// them; after the prefix sum a gather kernel compacts the strips (and writes the sort keys on the way).
// A read that emits more than `cap` minimizers is listed and redone by the generic kernel.
// compile-time loop: the body sees its index as a constant expression, so the window arrays below are
// indexed by constants from the first optimisation pass on and live in registers (a `#pragma unroll` loop
// left them in local memory)
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// The complete automaton (ties included) for ONE read, same strip output: the rerun of a read in which the lean
// automaton of sketch_fast_kernel met equal keys.  It runs in a kernel of its own (sketch_exact_list_kernel) over the
// list of such reads: it is the rare path and must not set the register budget of the common one.
template <int WW>
__device__ __forceinline__ uint32_t sketch_strip_exact(const DevReads &r, int k, uint32_t cap, uint64_t *__restrict__ sx,
                                                    uint32_t *__restrict__ srp, uint32_t xi) {
    const uint32_t read = __ldg(r.perm + xi);
    const uint32_t lenf = __ldg(r.len + read);
    const int len = (int)(lenf & HLO_LEN_MASK);
    const bool has_amb = (lenf & HLO_AMB_FLAG) != 0;
    const uint32_t *row = r.bases + (uint64_t)read * r.stride_w;
    const uint32_t *arow = r.amb + (uint64_t)read * r.stride_w;
    const uint64_t shift1 = 2 * (uint64_t)(k - 1), mask = (1ULL << 2 * k) - 1;
    const uint64_t INF = ~0ULL;
    uint64_t kmer0 = 0, kmer1 = 0;
    // ring buffer of the last WW entries: slot = step mod WW, a compile-time constant in the unrolled body
    uint64_t bx[WW];
    uint32_t by[WW];
    static_for<0, WW>([&](auto J) { bx[J] = INF, by[J] = 0xffffffffu; });
    // mm_sketch's `min` is at every step the RIGHTMOST SMALLEST entry of the window.  Instead of rescanning the
    // window when the minimum leaves it (a per-lane event: the warp would run the rescan with a handful of active
    // lanes at almost every step), the window minimum is kept by block decomposition: steps are grouped in
    // blocks of WW (one pass of the unrolled body); suf[j] = rightmost smallest of the previous block's slots
    // j..WW-1, computed for all j at the block boundary (the same step for every lane), and the minimum of the
    // current block's slots 0..s runs along.  Window minimum after step s = the smaller of the two, ties to the
    // running one (it lies to the right).
    uint64_t sufx[WW];
    uint32_t sufy[WW];
    uint32_t suf_tie = 0;          // bit j: the suffix minimum of j.. occurs more than once in j..WW-1
    uint32_t sufs[WW];             // its slot
    static_for<0, WW>([&](auto J) { sufx[J] = INF, sufy[J] = 0xffffffffu, sufs[J] = WW - 1; });  // all-INF window: rightmost slot
    uint64_t px = INF;             // running minimum of the current block
    uint32_t py = 0xffffffffu;
    int pslot = 0;
    bool ptie = false;
    uint64_t min_x = INF;          // mm_sketch's `min`
    uint32_t min_y = 0xffffffffu;
    int min_slot = 0, l = 0;
    uint32_t n = 0;
    const uint64_t sbase = (uint64_t)xi * cap;
    auto emit = [&](uint64_t X, uint32_t Y) {
        if (n < cap) sx[sbase + n] = X, srp[sbase + n] = Y;
        ++n;
    };
    uint32_t word = 0, aword = 0;
    for (int i0 = 0; i0 < len; i0 += WW) {
        if (i0 > 0) {  // block boundary: suffix minima of the block just finished (the ring holds exactly it)
            uint64_t cx = INF;
            uint32_t cy = 0xffffffffu, cs = WW - 1, tie = 0;
            suf_tie = 0;
            static_for<0, WW>([&](auto JJ) {
                constexpr int j = WW - 1 - JJ;
                if (j == WW - 1 || bx[j] < cx) cx = bx[j], cy = by[j], cs = j, tie = 0;  // strict <: the rightmost stays
                else if (bx[j] == cx) tie = 1;
                sufx[j] = cx, sufy[j] = cy, sufs[j] = cs;
                suf_tie |= tie << j;
            });
        }
        static_for<0, WW>([&](auto SS) {
            constexpr int s = SS;
            const int i = i0 + s;
            if (i < len) {
                if ((i & 15) == 0) {
                    word = __ldg(row + (i >> 4));
                    aword = has_amb ? __ldg(arow + (i >> 4)) : 0u;
                }
                const uint32_t c = word & 3u, amb = aword & 1u;
                word >>= 2, aword >>= 2;
                uint64_t info_x = INF;
                uint32_t info_y = 0xffffffffu;
                if (!amb) {
                    const int kmer_span = l + 1 < k ? l + 1 : k;
                    kmer0 = (kmer0 << 2 | (uint64_t)c) & mask;
                    kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
                    const uint32_t z = kmer0 < kmer1 ? 0u : 1u;  // k odd: kmer0 != kmer1 always
                    ++l;
                    if (l >= k) {
                        info_x = mm_hash64(z ? kmer1 : kmer0, mask) << 8 | (uint64_t)kmer_span;
                        info_y = (uint32_t)i << 1 | z;
                    }
                } else
                    l = 0;
                bx[s] = info_x, by[s] = info_y;
                if (l == WW + k - 1 && min_x != INF) {  // first full window: identical k-mers were not stored yet
                    static_for<s + 1, WW>([&](auto J) {
                        if (min_x == bx[J] && by[J] != min_y) emit(bx[J], by[J]);
                    });
                    static_for<0, s>([&](auto J) {
                        if (min_x == bx[J] && by[J] != min_y) emit(bx[J], by[J]);
                    });
                }
                // running minimum of the current block (<=: the rightmost smallest)
                if (s == 0 || info_x <= px) ptie = (s != 0) && info_x == px, px = info_x, py = info_y, pslot = s;
                // the events of mm_sketch, decided against the OLD minimum
                const bool ev_a = info_x <= min_x;                 // a new minimum
                const bool ev_b = !ev_a && min_slot == s;          // the old minimum left the window
                if (((ev_a && l >= WW + k) || (ev_b && l >= WW + k - 1)) && min_x != INF) emit(min_x, min_y);
                // the new window minimum
                bool tie;
                if (s == WW - 1 || px <= sufx[s + 1 < WW ? s + 1 : 0]) {
                    tie = ptie || (s != WW - 1 && px == sufx[s + 1 < WW ? s + 1 : 0]);
                    min_x = px, min_y = py, min_slot = pslot;
                } else {
                    constexpr int j = s + 1 < WW ? s + 1 : 0;
                    tie = (suf_tie >> j) & 1u;
                    min_x = sufx[j], min_y = sufy[j], min_slot = (int)sufs[j];
                }
                if (ev_b && tie && l >= WW + k - 1 && min_x != INF) {  // rare: other entries share the new minimum
                    auto dup = [&](auto J) {
                        if (min_x == bx[J] && min_y != by[J]) emit(bx[J], by[J]);
                    };
                    static_for<s + 1, WW>(dup);
                    static_for<0, s + 1>(dup);
                }
            }
        });
    }
    if (min_x != INF) emit(min_x, min_y);
    return n;
}


// mm_sketch emits an entry that equals the window minimum but is not "the" minimum only through two special
// blocks (first full window; after the minimum left the window).  Both need two EQUAL valid entries inside one
// window — the same canonical k-mer twice within w + k bases, i.e. low-complexity sequence.  The fast kernel does
// not carry those blocks: it raises a sticky flag at the three places where equal values can meet (running
// minimum, running minimum against the previous block's suffix minimum, suffix scan), and a flagged read is redone
// on the spot by the complete automaton (sketch_strip_exact).  Invalid entries (the first k-1 positions,
// positions after an N) get distinct markers ~i (bit 63 set, a newer one is smaller), so they never tie, never win
// against a valid entry and keep mm_sketch's "rightmost" choice among themselves.

// The window is ORDERED on 32-bit keys: hk = the top 31 bits of the hash (hash >> lo_bits, lo_bits = max(2k - 31, 0)); the
// low bits ride along in the upper half of the position word (y = low << 16 | pos << 1 | strand; pos < 2^15).  hk_a < hk_b
// implies hash_a < hash_b, and two entries with EQUAL hk are treated like equal hashes: the read goes to the complete
// automaton, which compares the full values (a false alarm once in 2^31 comparisons).  Every compare and select of the
// window logic is then one 32-bit operation instead of a 64-bit pair, and the ring and the suffix minima take 4 instead of
// 6 registers per window slot.
template <int WW, bool AMB>
__device__ __forceinline__ void sketch_fast_body(const DevReads &r, int k, uint32_t cap, uint64_t *__restrict__ sx,
                                                 uint32_t *__restrict__ srp, uint32_t *__restrict__ count,
                                                 uint32_t *__restrict__ ovf_list, unsigned long long *__restrict__ ovf_cnt,
                                                 uint32_t *__restrict__ tie_list, unsigned long long *__restrict__ tie_cnt) {
    const uint32_t xi = blockIdx.x * blockDim.x + threadIdx.x;
    if (xi >= r.n_reads) return;
    const uint32_t read = __ldg(r.perm + xi);
    const uint32_t lenf = __ldg(r.len + read);
    const int len = (int)(lenf & HLO_LEN_MASK);
    const bool has_amb = AMB && (lenf & HLO_AMB_FLAG) != 0;
    const uint32_t *row = r.bases + (uint64_t)read * r.stride_w;
    const uint32_t *arow = r.amb + (uint64_t)read * r.stride_w;
    const uint64_t shift1 = 2 * (uint64_t)(k - 1), mask = (1ULL << 2 * k) - 1;
    const int lo_bits = 2 * k > 31 ? 2 * k - 31 : 0;
    const uint32_t lo_mask = (1u << lo_bits) - 1u;
    uint64_t kmer0 = 0, kmer1 = 0;
    // ring buffer of the last WW entries: slot = step mod WW, a compile-time constant in the unrolled body.
    // mm_sketch's `min` is at every step the RIGHTMOST SMALLEST entry of the window; it is kept by block
    // decomposition: steps are grouped in blocks of WW (one pass of the unrolled body); suf[j] = rightmost smallest
    // of the previous block's slots j..WW-1, computed for all j at the block boundary (the same step for every
    // lane), and the minimum of the current block's slots 0..s runs along.  Window minimum after step s = the
    // smaller of the two, ties to the running one (it lies to the right).
    // Invalid entries (first k-1 positions, positions after an N) get distinct markers 0xffffffff - i (bit 31 set, a newer
    // one is smaller): they never tie, never win against a valid entry and keep mm_sketch's "rightmost" choice.
    uint32_t bx[WW], sufx[WW], by[WW], sufy[WW];
    // (before the first block boundary the suffix minima are invalid entries that equal no marker and not each other)
    static_for<0, WW>([&](auto J) { bx[J] = 0xffffffffu, by[J] = 0xffffffffu, sufx[J] = 0x80000000u | (uint32_t)J, sufy[J] = 0xffffffffu; });
    uint32_t px = 0xffffffffu, min_x = 0xffffffffu;  // running minimum of the current block; mm_sketch's `min`
    uint32_t py = 0xffffffffu, min_y = 0xffffffffu;
    bool tie = false;
    int l = 0;
    uint32_t n = 0;
    const uint64_t sbase = (uint64_t)xi * cap;
    auto emit = [&](uint32_t X, uint32_t Y) {  // back to minimap2's words: x = hash << 8 | span, position word = pos << 1 | strand
        if (n < cap) sx[sbase + n] = (((uint64_t)X << lo_bits) | (uint64_t)(Y >> 16)) << 8 | (uint64_t)k, srp[sbase + n] = Y & 0xffffu;
        ++n;
    };
    auto valid = [](uint32_t X) { return (int)X >= 0; };
    uint32_t word = 0, aword = 0;
    for (int i0 = 0; i0 < len; i0 += WW) {
        if (i0 > 0) {  // block boundary: suffix minima of the block just finished (the ring holds exactly it)
            uint32_t cx = bx[WW - 1], cy = by[WW - 1];
            sufx[WW - 1] = cx, sufy[WW - 1] = cy;
            static_for<1, WW>([&](auto JJ) {
                constexpr int j = WW - 1 - JJ;
                tie |= bx[j] == cx;
                if (bx[j] < cx) cx = bx[j], cy = by[j];  // strict <: the rightmost stays
                sufx[j] = cx, sufy[j] = cy;
            });
        }
        static_for<0, WW>([&](auto SS) {
            constexpr int s = SS;
            const int i = i0 + s;
            if (i < len) {
                if ((i & 15) == 0) {
                    word = __ldg(row + (i >> 4));
                    if (AMB) aword = has_amb ? __ldg(arow + (i >> 4)) : 0u;
                }
                const uint32_t c = word & 3u;
                word >>= 2;
                bool amb = false;
                if (AMB) amb = aword & 1u, aword >>= 2;
                uint32_t info_x = 0xffffffffu - (uint32_t)i;  // invalid marker
                uint32_t info_y = 0xffffffffu;
                if (!amb) {
                    kmer0 = (kmer0 << 2 | (uint64_t)c) & mask;
                    kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
                    const uint32_t z = kmer0 < kmer1 ? 0u : 1u;  // k odd: kmer0 != kmer1 always
                    ++l;
                    if (l >= k) {
                        const uint64_t h = mm_hash64(z ? kmer1 : kmer0, mask);
                        info_x = (uint32_t)(h >> lo_bits);
                        info_y = ((uint32_t)h & lo_mask) << 16 | (uint32_t)i << 1 | z;
                    }
                } else
                    l = 0;
                bx[s] = info_x, by[s] = info_y;
                // running minimum of the current block (<=: the rightmost smallest)
                if (s == 0) px = info_x, py = info_y;
                else {
                    tie |= info_x == px;
                    if (info_x <= px) px = info_x, py = info_y;
                }
                // the events of mm_sketch, decided against the OLD minimum
                const bool ev_a = info_x <= min_x;                                            // a new minimum
                const bool ev_b = !ev_a && ((min_y & 0xffffu) >> 1) == (uint32_t)(i - WW);    // the old minimum left the window
                if (((ev_a && l >= WW + k) || (ev_b && l >= WW + k - 1)) && valid(min_x)) emit(min_x, min_y);
                // the new window minimum
                if (s == WW - 1) min_x = px, min_y = py;
                else {
                    constexpr int j = s + 1 < WW ? s + 1 : 0;
                    tie |= px == sufx[j];
                    if (px <= sufx[j]) min_x = px, min_y = py;
                    else min_x = sufx[j], min_y = sufy[j];
                }
            }
        });
    }
    if (valid(min_x)) emit(min_x, min_y);
    if (tie) {  // equal keys met: the complete automaton decides (sketch_exact_list_kernel, right behind this kernel)
        tie_list[atomicAdd(tie_cnt, 1ULL)] = xi;
        count[xi] = 0;
        return;
    }
    count[xi] = n;
    if (n > cap) {  // more minimizers than the strip holds: the generic kernel writes this read
        const unsigned long long p = atomicAdd(ovf_cnt, 1ULL);
        ovf_list[p] = xi;
    }
}

template <int WW>
__global__ void __launch_bounds__(128) sketch_exact_list_kernel(DevReads r, int k, uint32_t cap, uint64_t *__restrict__ sx,
                                                                uint32_t *__restrict__ srp, uint32_t *__restrict__ count,
                                                                uint32_t *__restrict__ ovf_list, unsigned long long *__restrict__ ovf_cnt,
                                                                const uint32_t *__restrict__ tie_list,
                                                                const unsigned long long *__restrict__ tie_cnt) {
    const unsigned long long nl = *tie_cnt;
    const uint32_t n_list = nl < (unsigned long long)r.n_reads ? (uint32_t)nl : r.n_reads;
    for (uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x; ti < n_list; ti += gridDim.x * blockDim.x) {
        const uint32_t xi = tie_list[ti];
        const uint32_t n = sketch_strip_exact<WW>(r, k, cap, sx, srp, xi);
        count[xi] = n;
        if (n > cap) {
            const unsigned long long p = atomicAdd(ovf_cnt, 1ULL);
            ovf_list[p] = xi;
        }
    }
}

// A batch without ambiguous bases (the packer's device counter says so; the host never looks) runs the body that
// does not touch the ambiguity plane.  The branch is uniform over the grid.
template <int WW>
__global__ void __launch_bounds__(128, WW <= 11 ? HLO_SK_BLOCKS : 3) sketch_fast_kernel(DevReads r, int k, uint32_t cap, uint64_t *__restrict__ sx,
                                                             uint32_t *__restrict__ srp, uint32_t *__restrict__ count,
                                                             uint32_t *__restrict__ ovf_list, unsigned long long *__restrict__ ovf_cnt,
                                                             uint32_t *__restrict__ tie_list, unsigned long long *__restrict__ tie_cnt,
                                                             const uint32_t *__restrict__ amb_count) {
    if (*amb_count) sketch_fast_body<WW, true>(r, k, cap, sx, srp, count, ovf_list, ovf_cnt, tie_list, tie_cnt);
    else sketch_fast_body<WW, false>(r, k, cap, sx, srp, count, ovf_list, ovf_cnt, tie_list, tie_cnt);
}

// k <= 23: the hash bits below the 31-bit ordering key (2k - 31 <= 15) fit next to the 16-bit position word
bool sketch_fast_supported(int w, int k) { return (k & 1) && k >= 1 && k <= 23 && (w == 11 || w == 5 || w == 19); }

template <int WW>
static void launch_sketch_fast_w(const DevReads &r, int k, uint32_t cap, uint64_t *sx, uint32_t *srp, uint32_t *count, uint32_t *ovf_list,
                                 unsigned long long *ovf_cnt, uint32_t *tie_list, unsigned long long *tie_cnt, const uint32_t *amb_count,
                                 cudaStream_t s) {
    const unsigned grid = (r.n_reads + 127) / 128;
    sketch_fast_kernel<WW><<<grid, 128, 0, s>>>(r, k, cap, sx, srp, count, ovf_list, ovf_cnt, tie_list, tie_cnt, amb_count);
    // the reads the lean automaton gave up on (usually none: a small grid that strides over the device-side list)
    sketch_exact_list_kernel<WW><<<grid < 296u ? grid : 296u, 128, 0, s>>>(r, k, cap, sx, srp, count, ovf_list, ovf_cnt, tie_list, tie_cnt);
}

void launch_sketch_fast(const DevReads &r, int w, int k, uint32_t cap, uint64_t *sx, uint32_t *srp, uint32_t *count,
                        uint32_t *ovf_list, unsigned long long *ovf_cnt, uint32_t *tie_list, unsigned long long *tie_cnt,
                        const uint32_t *amb_count, cudaStream_t s, LaunchCounter *lc) {
    if (r.n_reads == 0) return;
    if (w == 11) launch_sketch_fast_w<11>(r, k, cap, sx, srp, count, ovf_list, ovf_cnt, tie_list, tie_cnt, amb_count, s);
    else if (w == 5) launch_sketch_fast_w<5>(r, k, cap, sx, srp, count, ovf_list, ovf_cnt, tie_list, tie_cnt, amb_count, s);
    else launch_sketch_fast_w<19>(r, k, cap, sx, srp, count, ovf_list, ovf_cnt, tie_list, tie_cnt, amb_count, s);
    if (lc) lc->total += 2;
}

// strips -> dense arrays in name order, plus the (cluster, hash | index) pairs the radix sort starts from.
// 8 lanes per read walk its strip (a strip holds ~22 of `cap` entries); strips of reads that overflowed are
// skipped (the generic kernel writes those reads afterwards).
__global__ void __launch_bounds__(256) sketch_compact_kernel(DevReads r, uint32_t cap, const uint64_t *__restrict__ sx,
                                                             const uint32_t *__restrict__ srp, const uint32_t *__restrict__ count,
                                                             const uint32_t *__restrict__ off, uint64_t *__restrict__ mz_x,
                                                             uint2 *__restrict__ mz_rr,
                                                             uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, int hash_bits,
                                                             uint32_t cap_mz) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t xi = (uint32_t)(gid >> 3), l = (uint32_t)(gid & 7);
    if (xi >= r.n_reads) return;
    if (off[r.n_reads] > cap_mz) return;  // more minimizers than the arrays hold: the host reruns the batch with larger ones
    const uint32_t n = count[xi];
    if (n > cap) return;  // written by the generic kernel
    const uint32_t read = r.perm[xi];
    const uint64_t chi = (uint64_t)r.cidx[read] << hash_bits;  // cluster-major: a cluster's buckets stay together
    const uint32_t o = off[xi];
    const uint64_t sb = (uint64_t)xi * cap;
    for (uint32_t j = l; j < n; j += 8) {
        const uint32_t dst = o + j;
        HLO_DBG_ASSERT(dst < cap_mz && dst < off[r.n_reads]);
        const uint64_t x = sx[sb + j];
        if (mz_x) mz_x[dst] = x;
        mz_rr[dst] = make_uint2(read, srp[sb + j]);
        keys[dst] = (x >> 8) | chi;
        vals[dst] = dst;
    }
}

void launch_sketch_compact(const DevReads &r, uint32_t cap, const uint64_t *sx, const uint32_t *srp, const uint32_t *count,
                           const uint32_t *off, uint64_t *mz_x, uint2 *mz_rr, uint64_t *keys, uint32_t *vals,
                           int hash_bits, uint32_t cap_mz, cudaStream_t s, LaunchCounter *lc) {
    const uint64_t total = (uint64_t)r.n_reads * 8;
    if (total == 0) return;
    sketch_compact_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(r, cap, sx, srp, count, off, mz_x, mz_rr, keys, vals, hash_bits,
                                                                          cap_mz);
    if (lc) lc->total++;
}

__global__ void mz_assemble_kernel(const uint64_t *__restrict__ x, const uint2 *__restrict__ rr, uint64_t n,
                                   hlo_mz *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hlo_mz m;
    m.x = x[i];
    m.y = (uint64_t)rr[i].x << 32 | rr[i].y;
    out[i] = m;
}

void launch_mz_assemble(const uint64_t *mz_x, const uint2 *mz_rr, uint64_t n, hlo_mz *out, cudaStream_t s, LaunchCounter *lc) {
    if (n == 0) return;
    mz_assemble_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mz_x, mz_rr, n, out);
    if (lc) lc->total++;
}

}  // namespace hlo
