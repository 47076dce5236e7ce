// common.cuh — shared device/host definitions of libhlovlp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hlovlp.h"

// -DHLO_DEBUG_BOUNDS (make debug -> libhlovlp_dbg.so): device-side asserts on the indices the kernels compute from device
// data (counters, offsets, hash slots).  A failed assert traps the kernel and the next CUDA call reports it; the GPU
// test-suite runs once under this build (HLO_LIB=.../libhlovlp_dbg.so), profiles/r2_debug_bounds.log.
#ifdef HLO_DEBUG_BOUNDS
#include <assert.h>
#define HLO_DBG_ASSERT(c) assert(c)
#else
#define HLO_DBG_ASSERT(c) ((void)0)
#endif

namespace hlo {

// ---------------------------------------------------------------------------------------------
// Device-resident batch of packed reads.
//
// Layout in HBM: every read has one row of `stride_w` = 4*nv 32-bit words per
// strand (nv 16-byte vectors; 150 bp -> 48 B, 250 bp -> 64 B), 16 bases per word, base i at bits
// [2(i%16), 2(i%16)+2) of word i/16 (little-endian in the word); A=0 C=1 G=2 T=3, ambiguous bases are
// stored as 0 and flagged in the parallel `amb` plane with the same geometry (01 = ambiguous).
// Plane 0 holds the read as given, plane 1 its reverse complement, so the verify kernel always
// compares two forward-packed rows (and can move a target-limited overlap end to position 0 by
// switching both reads to the other plane).  Rows are 16-byte aligned: a thread reads a row with nv
// LDG.128.  The allocation is padded by PAD_WORDS words on both sides (shifted vector loads may run
// past a row; what they fetch there is masked out).
struct DevReads {
    uint32_t n_reads = 0, n_clusters = 0;
    uint32_t stride_w = 0;  // words per row = 4 * nv
    uint32_t nv = 0;        // 16-byte vectors per row
    uint32_t max_len = 0;
    uint32_t max_cluster = 0;  // reads in the largest cluster
    uint64_t n_bases = 0;
    int has_amb = 0;        // any ambiguous base in the batch
    uint32_t *len = nullptr;    // [n] read length | (has ambiguous base ? 0x80000000 : 0)
    uint32_t *cbase = nullptr;  // [n] index of the first read of the read's cluster
    uint32_t *cidx = nullptr;   // [n] cluster index
    uint32_t *ids = nullptr;    // [n] decimal name
    uint32_t *nrank = nullptr;  // [n] cbase + numeric rank of the name inside the cluster
    int plain_names = 0;        // names are 0..N-1 in file order: nrank[i] == i, no lookup needed
    uint4 *meta = nullptr;      // [n] (length, name, cluster, cluster base): one gather where a record needs them all
    uint32_t *perm = nullptr;   // [n] x -> read: reads of a cluster in strcmp order of their decimal names
    uint32_t *bases = nullptr;  // [2][n][stride_w]
    uint32_t *amb = nullptr;    // [2][n][stride_w]
};
#define HLO_PAD_WORDS 64
#define HLO_PAD_CANARY 0xA5  // byte the padding is filled with; nothing may ever WRITE there (hlo_debug_check_reads)

// One block of 64-bit device counters per context, zeroed once per batch and read back ONCE at its end: every stage
// sizes its work from the counts its predecessor left here, so the host never waits between stages.
enum Ctr : int {
    C_AMB = 0,      // ambiguous bases in the batch (pack)
    C_SK_OVF,       // reads the single-pass sketch kernel handed to the generic one
    C_MZ,           // minimizers (low word written by the scan)
    C_COLL,         // low word: mixed prefix groups repaired; high word: 1 when one was too large (full sort needed)
    C_CAND,         // candidates (may exceed the buffer: then only counted)
    C_HITS,         // seed hits
    C_PG_OVF,       // [PG_PASSES] reads whose table overflowed in pass p
    C_PG_WORK = C_PG_OVF + 5,  // [PG_PASSES] work-stealing counters of the pair generator passes
    C_ACC = C_PG_WORK + 5,     // accepted records
    C_DEF,          // candidates deferred to the exact verify kernel
    C_NOUT,         // records after the per-pair dedup (low word written by the scan)
    C_FLAGS,        // HLO_FLAG_*
    C_DIGEST,       // order-independent sum of a 64-bit mix of every output record
    C_MIDOCC_LO,    // smallest / largest per-cluster occurrence cut-off of the batch
    C_MIDOCC_HI,
    C_SK_TIE,       // reads in which the lean sketch automaton met equal keys: redone by the complete one
    C_BIG_GROUPS,   // idA groups too large for the quadratic rule of the records tail (listed for group_rank_big_kernel)
    C_COUNT = 32
};
#define HLO_PG_PASSES 5
#define HLO_FLAG_MIDOCC_AMBIGUOUS 1ull  // a cluster's occurrence quantile fell into the clamp bin of the histogram

#define HLO_MAX_W_PUBLIC 32  // window sizes the sketch kernel keeps in its per-thread ring
#define HLO_LEN_MASK 0x7fffffffu
#define HLO_AMB_FLAG 0x80000000u

// internal compact verified record (16 B): what the verify kernels emit
struct Rec16 {
    uint32_t q, t;
    uint32_t ds;  // (diag + 2^30) << 1 | strand
    uint32_t kb;  // K | blocklen << 16
};
#define HLO_DIAG_BIAS (1 << 30)

// Internal candidate (16 B) = hlo_cand + both read lengths, so that the verify kernel knows the overlap
// geometry before it touches a row:
//   z = (diag + 32768) | ql << 16 | strand << 31
//   w = tl | EL << 15 | ER << 22 | amb << 29
// EL / ER = number of bases between the overlap's left / right end and the outermost anchor k-mer,
// capped at 127 (the end rule never looks further inside the supported domain);
// amb = either read holds an ambiguous base.
struct CandI {
    uint32_t q, t, z, w;
};
#define HLO_REGION_CAP 127
#define HLO_CAND_FAR (1u << 30)  // CandI.w: an end region is longer than HLO_REGION_CAP; the true lengths are in the side array
// *far (when not null) receives el | er << 16, the uncapped distances from the overlap ends to the outermost anchors, for
// a candidate flagged HLO_CAND_FAR: the exact verify kernel extends from the true anchors (an extension that starts 127
// bases from the end instead of, say, 300 can stop elsewhere when mismatches lie in between); the fast kernel only ever
// looks at whether a region covers its 32-base window.
__host__ __device__ __forceinline__ CandI make_candi(uint32_t q, uint32_t t, int diag, uint32_t strand, uint32_t ql, uint32_t tl,
                                                     int a_start, int a_end, uint32_t amb, uint32_t *far = nullptr) {
    const int js = diag < 0 ? -diag : 0;
    const int je = (int)ql < (int)tl - diag ? (int)ql : (int)tl - diag;
    if (a_end == 0) a_start = je, a_end = js;  // no anchor information: each end region is the whole overlap
    if (a_start < js) a_start = js;
    if (a_end > je) a_end = je;
    int el = a_start - js, er = je - a_end;
    el = el < 0 ? 0 : el, er = er < 0 ? 0 : er;
    const bool is_far = el > HLO_REGION_CAP || er > HLO_REGION_CAP;
    if (is_far && far) *far = (uint32_t)el | (uint32_t)er << 16;
    el = el > HLO_REGION_CAP ? HLO_REGION_CAP : el;
    er = er > HLO_REGION_CAP ? HLO_REGION_CAP : er;
    CandI c;
    c.q = q, c.t = t;
    c.z = (uint32_t)(diag + 32768) | ql << 16 | strand << 31;
    c.w = tl | (uint32_t)el << 15 | (uint32_t)er << 22 | (amb & 1u) << 29 | (is_far && far ? HLO_CAND_FAR : 0u);
    return c;
}

// device tables for the python-double threshold tests (filter_ovlp_inline.py:58,83)
struct VerifyTables {
    const uint16_t *min_matches;  // [max_len+1] smallest matches with !(matches/blocklen < min_identity)
    const uint16_t *max_clip;     // [max_len+1] largest overhang with !(overhang > min(o, maplen*r))
    const uint16_t *max_k_upto;   // [max_len+1] running max of b - min_matches[b]
    int32_t max_len;
};

struct VerifyParams {
    int32_t k, a, b, n;  // k-mer, match, mismatch, ambiguous scores
    int32_t min_len, max_overhang;
    int32_t bad_limit;  // fast verify kernel: largest number of bad columns for which a 32-base end window is exact
    int32_t end_bonus, zdrop;  // ksw2 extension: a query-limited end reaches the read end iff s_end + end_bonus > max; stop at max - s > zdrop
    int32_t defer_mm;   // fast verify kernel: from this many mismatches on, a query-limited end might not reach the read end: exact kernel
    int32_t pair_rule;  // 1: rows that pass length + identity but fail the overhang test are emitted with HLO_KB_OVERHANG_FAIL
};
#define HLO_KB_OVERHANG_FAIL 0x8000u  // bit 15 of Rec16.kb (K uses bits 0..14): the row failed the overhang test (pair_rule 1)
#define HLO_KB_K_MASK 0x7fffu

// ---------------------------------------------------------------------------------------------
// decimal-string order of two unsigned integers == strcmp(str(a), str(b)) < 0  (minimap2 -X skips a
// seed when strcmp(qname, tname) > 0; names are the integers written by rename_fas.py:48-57)
__host__ __device__ __forceinline__ int dec_digits(uint32_t v) {
    return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6
         : v < 10000000u ? 7 : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}
__host__ __device__ __forceinline__ uint64_t pow10_u64(int e) {
    uint64_t r = 1;
    for (int i = 0; i < e; ++i) r *= 10;
    return r;
}
__host__ __device__ __forceinline__ bool name_less(uint32_t a, uint32_t b) {
    const int da = dec_digits(a), db = dec_digits(b);
    if (da == db) return a < b;
    if (da < db) {  // a is the shorter string: pad it, a prefix sorts first
        const uint64_t as = (uint64_t)a * pow10_u64(db - da);
        return as <= (uint64_t)b;
    }
    const uint64_t bs = (uint64_t)b * pow10_u64(da - db);
    return (uint64_t)a < bs;
}

// ---------------------------------------------------------------------------------------------
// Record expansion: minimap22sfo.py:47-74 + sfo2overlaps.py:37-50 from (ql, tl, diag, strand, K).
// With PAF coordinates qs/qe/ts/te of ANY block on the diagonal:
//   '+': OHA = qstart - sstart = -diag            OHB = slen - sstart - (qlen - qstart) = tl - ql - diag
//   '-': OHA = qstart - (slen - send) = ql-tl+diag OHB = send - (qlen - qstart) = diag
__host__ __device__ __forceinline__ void expand_record(uint32_t qid, uint32_t tid, int ql, int tl, int diag, int strand,
                                                       uint32_t K, hlo_rec *o) {
    int OHA, OHB;
    if (!strand) OHA = -diag, OHB = tl - ql - diag;
    else OHA = ql - tl + diag, OHB = diag;
    const int OLA = OHA >= 0 ? ((ql - OHA) < tl ? (ql - OHA) : tl) : ((tl + OHA) < ql ? (tl + OHA) : ql);
    uint32_t idA = qid, idB = tid;
    if (name_less(idB, idA)) {  // minimap22sfo.py:62 (string compare); never fires for -X candidates
        uint32_t t = idA; idA = idB; idB = t;
        if (!strand) OHA = -OHA, OHB = -OHB;
        else { int x = OHA; OHA = OHB; OHB = x; }
    }
    if (idA > idB) {  // sfo2overlaps.py:41-47 flip_N / flip_I
        uint32_t t = idA; idA = idB; idB = t;
        if (!strand) OHA = -OHA, OHB = -OHB;
        else { int x = OHA; OHA = OHB; OHB = x; }
    }
    o->idA = idA, o->idB = idB, o->OHA = OHA, o->OHB = OHB;
    o->OLA = (uint32_t)OLA, o->OLB = (uint32_t)OLA, o->K = K, o->ori = (uint32_t)strand;
}

}  // namespace hlo
