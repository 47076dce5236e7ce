// kernels.cuh — host-callable launchers of the sm_100a kernels (one .cu per stage).
#pragma once
#include "common.cuh"

namespace hlo {

struct LaunchCounter {
    uint64_t total = 0, verify = 0;
};

// ---- pack.cu --------------------------------------------------------------------------------
// ASCII -> 2-bit rows (both strands) + ambiguity planes.  amb_count (device u32) receives the number
// of ambiguous bases.  len[] must hold plain lengths on entry; the kernel ORs in HLO_AMB_FLAG.
void launch_pack(const uint8_t *ascii, const uint64_t *base_off, DevReads r, uint32_t *amb_count, cudaStream_t s,
                 LaunchCounter *lc);

// per-read metadata (len, cbase, cidx, ids, nrank, perm) on the device for names 0..N-1 in file order
void launch_read_meta(const uint64_t *base_off, const uint32_t *cluster_off, DevReads r, cudaStream_t s, LaunchCounter *lc);

// packed input (hlo_batch.packed): per-read metadata + words per read (wcount, n_reads + 1 entries, the last 0); then,
// with woff = their exclusive scan, the rows of both strands and the listed ambiguous bases
void launch_packed_meta(const uint32_t *read_len, const uint32_t *cluster_off, DevReads r, uint32_t *wcount, cudaStream_t s, LaunchCounter *lc);
void launch_unpack2b(const uint32_t *packed, const uint32_t *woff, const unsigned long long *amb, uint64_t n_amb, DevReads r,
                     uint32_t *amb_count, cudaStream_t s, LaunchCounter *lc);

// ---- sketch.cu ------------------------------------------------------------------------------
// pass 0: count[r] = number of minimizers of read r;  pass 1: write them at off[r].
void launch_sketch_count(const DevReads &r, int w, int k, uint32_t *count, cudaStream_t s, LaunchCounter *lc);
// list != nullptr: only the *n_list_ptr listed reads (x indices; the count is a device counter)
// mz_x == nullptr: hash words not wanted; keys != nullptr: also the sort pairs (cluster << hash_bits | hash, own index)
// cap_mz: entries the output arrays hold (a batch with more writes nothing; the host reruns it)
void launch_sketch_write(const DevReads &r, int w, int k, const uint32_t *off, uint64_t *mz_x, uint2 *mz_rr, uint64_t *keys, uint32_t *vals,
                         int hash_bits, const uint32_t *list, const unsigned long long *n_list_ptr, uint32_t cap_mz, cudaStream_t s,
                         LaunchCounter *lc);
// single-pass fast path (static window, k odd): per-read strips of `cap` entries + counts, then a gather
bool sketch_fast_supported(int w, int k);
// tie_list / tie_cnt: reads in which equal window keys met; a second kernel reruns them with the complete automaton
void launch_sketch_fast(const DevReads &r, int w, int k, uint32_t cap, uint64_t *sx, uint32_t *srp, uint32_t *count,
                        uint32_t *ovf_list, unsigned long long *ovf_cnt, uint32_t *tie_list, unsigned long long *tie_cnt,
                        const uint32_t *amb_count, cudaStream_t s, LaunchCounter *lc);
void launch_sketch_compact(const DevReads &r, uint32_t cap, const uint64_t *sx, const uint32_t *srp, const uint32_t *count,
                           const uint32_t *off, uint64_t *mz_x, uint2 *mz_rr, uint64_t *keys, uint32_t *vals, int hash_bits,
                           uint32_t cap_mz, cudaStream_t s, LaunchCounter *lc);
void launch_mz_assemble(const uint64_t *mz_x, const uint2 *mz_rr, uint64_t n, hlo_mz *out, cudaStream_t s, LaunchCounter *lc);

// ---- scan_sort.cu ---------------------------------------------------------------------------
// Exclusive prefix sum of n u32 values (in may alias out).  `tmp` must hold scan_tmp_words(n) u32.
// If total != nullptr, *total (device u32) receives the sum of all inputs.
size_t scan_tmp_words(uint64_t n);
void launch_exclusive_scan(const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *tmp, uint32_t *total, cudaStream_t s,
                           LaunchCounter *lc);

// Stable LSD radix sort of (u64 key, u32 value) pairs on key bits [begin_bit, end_bit), applied to every segment of the
// array on its own (segment c = elements [elem_off[seg_off[c]], elem_off[seg_off[c + 1]]): the minimizers of cluster c,
// elem_off = per-read minimizer offsets, seg_off = the cluster offsets, both on the device).  Segments need no key
// bits and no pass.  keys/vals are double buffers; returns which buffer (0/1) holds the result.  n is an upper bound
// of the element count (it sizes the grid and `tmp`, sort_segmented_tmp_words(n, n_seg) u32): the real extents come
// from the device arrays.
size_t sort_segmented_tmp_words(uint64_t n, uint32_t n_seg);
int launch_radix_sort_segmented(uint64_t *keys[2], uint32_t *vals[2], uint64_t n, int begin_bit, int end_bit, const uint32_t *elem_off,
                                const uint32_t *seg_off, uint32_t n_seg, uint32_t *tmp, cudaStream_t s, LaunchCounter *lc);

// Equal keys adjacent + stable is all the index needs: after a sort on key bits [shift, ...) only, list the groups
// (runs of equal key >> shift) that mix different keys and order those by the full key.  list: n_cap/2 + 1 words;
// state (2 device words, zeroed by the caller): [0] groups fixed, [1] = 1 when a group was too large (redo with a full sort).
// *n_ptr = element count (device).
void launch_collision_fix(uint64_t *keys, uint32_t *vals, uint64_t n_cap, const uint32_t *n_ptr, int shift, uint32_t *list, uint32_t *state,
                          cudaStream_t s, LaunchCounter *lc);

// pairs (key, index) in any order -> keys_out[index] = key, vals_out[index] = index
void launch_unpermute_pairs(const uint64_t *keys, const uint32_t *vals, uint64_t n_cap, const uint32_t *n_ptr, uint64_t *keys_out,
                            uint32_t *vals_out, cudaStream_t s, LaunchCounter *lc);

// up to 64 words from device memory to (mapped) host memory: the read-back path for counters
void launch_copy_words(const uint32_t *src, uint32_t *dst, uint32_t n, cudaStream_t s, LaunchCounter *lc);

// ---- pairgen.cu -----------------------------------------------------------------------------
// index build on the hash-sorted minimizers
// sort key = cluster index << hash_bits | minimizer hash (cluster-major: buckets of one cluster stay together)
void launch_hash_keys(const uint64_t *mz_x, const uint2 *mz_rr, const uint32_t *cidx, int hash_bits, uint64_t n_cap, const uint32_t *n_ptr,
                      uint64_t *keys, uint32_t *vals, cudaStream_t s, LaunchCounter *lc);
// hits[i] = the i-th grouped entry — (read, pos<<1|strand), 8 bytes, or with dg_bits > 0 the 32-bit word
// (read - cluster_off[cluster]) << (dg_bits + 1) | strand << dg_bits | pos; range[m] = (i + 1, bucket end) of minimizer m, or (0, 0)
// when its bucket holds more than the cluster's occurrence cut-off midocc[cluster] — fixed_mid_occ when > 0, else
// minimap2's data-derived value from the occurrence histogram hist[n_clusters][bins] (both filled here).
// skeys: equal keys adjacent (the grouped order; cluster = key >> hash_bits), svals: their m.
size_t bucket_hist_words(uint32_t n_clusters, uint32_t bins);  // u32 words of `hist`
void launch_bucket_index(const uint64_t *skeys, const uint32_t *svals, const uint2 *mz_rr, uint64_t n_cap, const uint32_t *n_ptr, int hash_bits,
                         int fixed_mid_occ, uint32_t n_clusters, uint32_t *midocc, uint32_t *hist, uint32_t bins, unsigned long long *counters,
                         const uint32_t *cluster_off, int dg_bits, void *hits, uint2 *range, cudaStream_t s, LaunchCounter *lc);

struct PairgenArgs {
    DevReads reads;
    const uint32_t *mz_off;  // [n_reads+1] minimizers of the x-th read (name order)
    const uint2 *mz_rr;      // [n_mz] (read, pos<<1|strand), read order
    const uint2 *range;      // [n_mz] hits [s,e) minimizer m seeds (the part of its bucket behind its own entry)
    const uint2 *hits;       // [n_mz] (read, pos<<1|strand) in hash order, or
    const uint32_t *hits32 = nullptr;  // ... (read - cluster base) << (dg_bits + 1) | strand << dg_bits | pos (whenever the fast path is possible)
    int k, gap_min;
    int dg_bits;             // bits of (diagonal + max_len) in the 32-bit table key (>= 5)
    int external;            // 1: emit hlo_cand (exact anchor span), 0: emit CandI
    void *cands;             // output, 16 B per candidate
    uint32_t *cand_far = nullptr;  // [cand_cap] true end-region lengths of the candidates flagged HLO_CAND_FAR (internal form only)
    uint64_t cand_cap;
    unsigned long long *cand_counter, *hit_counter;  // candidates (counted past cand_cap too), seed hits
    unsigned long long *ovf_counter;                 // reads whose table overflowed in this pass
    unsigned long long *work_counter;                // work-stealing counter of this pass
    const uint32_t *work = nullptr;  // fast path only: x indices of the reads to do (nullptr: all reads)
    const unsigned long long *n_work_ptr = nullptr;  // ... and their number: the overflow counter of the previous pass
    uint32_t *overflow_reads;      // x indices of the reads whose shared-memory table overflowed
    uint32_t overflow_cap;
};
// fast path: 16 lanes per query read, seed table in shared memory (2^slots_log2 slots per read)
bool pairgen_fast_supported(const DevReads &r, int *dg_bits_out);
size_t pairgen_smem_bytes(int slots_log2);
void launch_pairgen(const PairgenArgs &a, int slots_log2, cudaStream_t s, LaunchCounter *lc);
// exact generic path: one block per listed read (list == nullptr: reads 0..n_list-1), 64-bit keys, table in HBM
void launch_generic_hit_counts(const PairgenArgs &a, const uint32_t *list, uint32_t n_list, uint32_t *counts, cudaStream_t s,
                               LaunchCounter *lc);
void launch_pairgen_generic(const PairgenArgs &a, const uint32_t *list, uint32_t n_list, unsigned long long *gtable_keys,
                            uint32_t *gtable_mn, uint32_t *gtable_mx, const uint64_t *gtable_off, cudaStream_t s,
                            LaunchCounter *lc);

// ---- verify.cu ------------------------------------------------------------------------------
struct VerifyArgs {
    DevReads reads;
    VerifyTables tab;
    VerifyParams prm;
    const CandI *cands;
    const uint32_t *cand_far = nullptr;  // side array of the candidates flagged HLO_CAND_FAR (same index)
    const unsigned long long *n_ptr;  // candidates: a device counter (the pair generator's), clamped to n_cap
    uint64_t n_cap;
    Rec16 *out;                    // accepted records (unordered)
    uint64_t out_cap;              // records `out` holds; more are only counted
    unsigned long long *acc_counter, *def_counter;  // accepted, deferred
    uint32_t *deferred;            // candidate indices for the exact path (n_cap entries)
};
bool verify_fast_supported(const DevReads &r, const VerifyTables &tab);
void launch_verify_fast(const VerifyArgs &a, cudaStream_t s, LaunchCounter *lc);
// exact per-candidate path: every candidate, or the candidates the fast kernel deferred
void launch_verify_exact(const VerifyArgs &a, bool all_cands, cudaStream_t s, LaunchCounter *lc);
void launch_cands_to_internal(const hlo_cand *in, uint64_t n, const DevReads &r, CandI *out, uint32_t *far, cudaStream_t s, LaunchCounter *lc);

// ---- records.cu -----------------------------------------------------------------------------
// group the accepted records by idA, sort the small groups by idB, keep the best record of every pair.
// *n_ptr = accepted records (device counter), clamped to n_cap.
struct ExpandOut {
    hlo_rec *out = nullptr;        // each output form is written when its pointer is set
    hlo_rec16 *out16 = nullptr;
    hlo_rec8 *out8 = nullptr;
    hlo_rec_aux *aux = nullptr;
    unsigned long long *digest = nullptr;
};
void launch_group_count(const Rec16 *rec, const unsigned long long *n_ptr, uint64_t n_cap, DevReads r, uint32_t *cnt, uint32_t *slot,
                        cudaStream_t s, LaunchCounter *lc);
void launch_group_scatter(const Rec16 *rec, const unsigned long long *n_ptr, uint64_t n_cap, DevReads r, const uint32_t *off,
                          const uint32_t *slot, uint2 *ent, uint32_t *entg, cudaStream_t s, LaunchCounter *lc);
void launch_group_mark(const Rec16 *rec, const unsigned long long *n_ptr, uint64_t n_cap, DevReads r, const uint32_t *cluster_off, const uint32_t *off,
                       const uint32_t *entg, int dedup, uint2 *ent, uint32_t *uniq, uint32_t *lt, uint32_t *big_list, unsigned long long *big_cnt,
                       cudaStream_t s, LaunchCounter *lc);  // big_list: n_reads words, *big_cnt zeroed by the caller
void launch_group_expand(const Rec16 *rec, const unsigned long long *n_ptr, uint64_t n_cap, DevReads r, const uint32_t *off, const uint32_t *uoff,
                         const uint32_t *entg, const uint2 *ent, const uint32_t *lt, const ExpandOut &o, cudaStream_t s, LaunchCounter *lc);
void launch_cluster_offsets(const uint32_t *uoff, const uint32_t *cluster_off, uint32_t n_clusters, unsigned long long *coff, cudaStream_t s,
                            LaunchCounter *lc);
uint64_t host_rec_digest(const hlo_rec &o);  // the digest term of one record, on the host

// ---- score.cu -------------------------------------------------------------------------------
// log p for (match?, phred byte 1, phred byte 2): 2 * 128 * 128 doubles from the 128-entry phred -> probability table; NaN where
// p < mismatch_thr
void launch_score_table(const double *prob_tab, double mismatch_thr, double *tab, cudaStream_t s, LaunchCounter *lc);
void launch_qual_flags(const uint8_t *quals, const uint64_t *base_off, uint32_t n_reads, uint8_t *qflag, cudaStream_t s, LaunchCounter *lc);
// quals: padded by 32 readable bytes in front and behind
// rec_cluster[i * cluster_stride] = cluster of record i (nullptr: one cluster); n_ptr != nullptr: the record count is
// *n_ptr on the device, clamped to n
void launch_score_records(const DevReads &r, const uint8_t *quals, const uint8_t *qflag, const uint64_t *base_off, const uint32_t *cluster_off,
                          const hlo_rec *rec, const uint32_t *rec_cluster, int cluster_stride, const uint32_t *n_ptr, uint64_t n,
                          const double *tab, uint32_t min_read_len, hlo_score *out, cudaStream_t s, LaunchCounter *lc);

}  // namespace hlo
