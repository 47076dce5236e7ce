/*
 * hlovlp.h — C ABI of libhlovlp, the B200-native short-read overlap stage for HyLight.
 *
 * The reference has no FFI for this path: it is a chain of processes glued by files
 * (minimap2 | cut | filter_ovlp_inline.py > overlaps.paf ; minimap22sfo.py ; sfo2overlaps.py),
 * called from /root/reference/script/polyte.tune_params.py:488-529 (preprocessing_denovo) and
 * :534-579 (run_sfo).  This header is the boundary a maintainer would bind instead of that chain
 * (see INTEGRATION.md for the ctypes stub that replaces the os.system() calls at
 * polyte.tune_params.py:510-517 and :549-557).
 *
 * Conventions: every function returns 0 or a negative hlo_status; nothing throws or exits.
 * Inputs are caller-owned plain host pointers, only read during the call (hlo_submit: until the
 * matching hlo_wait returns).  Outputs are ctx-owned host buffers valid until the next call on the
 * same ctx (or hlo_ctx_destroy).
 * A ctx is bound to one CUDA device and is single-threaded; distinct ctxs are independent.
 * There is no CPU fallback: without a CUDA device hlo_ctx_create fails with HLO_ENODEV.
 */
#ifndef HLOVLP_H
#define HLOVLP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HLO_ABI_VERSION 2

typedef enum hlo_status {
    HLO_OK = 0,
    HLO_EINVAL = -1,       /* bad argument */
    HLO_ENODEV = -2,       /* no usable CUDA device */
    HLO_ECUDA = -3,        /* CUDA runtime error (hlo_last_error has the text) */
    HLO_ENOMEM = -4,       /* host or device allocation failed */
    HLO_EUNSUPPORTED = -5, /* parameter combination outside the supported domain */
    HLO_ERANGE = -6,       /* a size exceeds a documented limit */
    HLO_EIO = -7           /* file error (text emitters) */
} hlo_status;

/*
 * Parameters of the chain.  Defaults (hlo_params_default) are the literal values of the
 * reference call site, polyte.tune_params.py:500-510:
 *   minimap2 --sr -X -c -k 21 -w 11 -s 60 -m 30 -n 2 -r 0 -A 4 -B 2 --end-bonus 100
 *   filter_ovlp_inline.py <min_overlap_len=60> 0.98 1 0.8
 */
typedef struct hlo_params {
    int32_t k;               /* -k 21 */
    int32_t w;               /* -w 11 */
    int32_t min_seeds;       /* -n 2   (1 or 2 supported) */
    int32_t min_chain_score; /* -m 30  (must be <= 2k) */
    int32_t mid_occ;         /* 0 = minimap2's default: per index (= per cluster) the (1 - 2e-4) quantile of the occurrences of
                                the distinct minimizers, + 1, clamped to [10, 10^6] (mm_idx_cal_max_occ; the reference command
                                passes the flag --sr, not the preset -x sr, so nothing fixes it); > 0 = that cut-off (-f INT).
                                Minimizers occurring more often seed nothing. */
    int32_t match_score;     /* -A 4 */
    int32_t mismatch_pen;    /* -B 2 */
    int32_t ambi_pen;        /* --score-N default 1 */
    int32_t end_bonus;       /* --end-bonus 100 (domain check only) */
    int32_t min_dp_score;    /* -s 60 (domain check only) */
    int32_t min_ovlp_len;    /* filter_ovlp_inline.py argv[1]; 60 first pass, 50 later passes */
    int32_t max_overhang;    /* filter_ovlp_inline.py argv[3] = 1 (0..15) */
    double min_identity;     /* filter_ovlp_inline.py argv[2] = 0.98 */
    double overhang_ratio;   /* filter_ovlp_inline.py argv[4] = 0.8 */
    int32_t zdrop;           /* -z 400 (minimap2 default): an end extension stops where the score has fallen this far below its
                                maximum; only the exact verify kernel ever meets it */
    int32_t pair_rule;       /* which row of a read pair survives:
                                0 = filter_ovlp_inline.py:12-39,42-91 (short-read pass): rows failing the overhang test are dropped
                                    first, then the longest row of the pair is kept;
                                1 = filter_trans_ovlp_inline_v3.py:49-77 (HyLight.py:309-311, contig stage b): the first row of the
                                    pair that passes length and identity stands for the pair, and the pair is dropped when THAT
                                    row fails the overhang test.
                                Both scripts work on 1000-line chunks of minimap2's output order; the library applies the rule
                                per cluster with "first" / "longest" = the canonical best row. */
} hlo_params;

/*
 * One chain of minimap2 = one candidate: query read q (the lexicographically smaller decimal
 * name, minimap2 -X), target read t, relative strand, diagonal and the span of the chained
 * anchors.  q and t are read indices inside the hlo_reads object (batch-global).
 *   diag : t_pos - q'_pos, q' = the query as aligned (reverse-complemented when strand = 1),
 *          i.e. q'[j] is compared with t[j + diag].
 *   strand_anchor : bit 0 = strand (0 '+', 1 '-'); bits 1..15 = start of the first anchor
 *          k-mer in q' coordinates; bits 16..31 = end (exclusive) of the last anchor k-mer.
 *          Bits 16..31 == 0 means "no anchor information": the end rule then scans the whole
 *          overlap (identical results inside the supported domain).
 */
typedef struct hlo_cand {
    uint32_t q, t;
    int32_t diag;
    uint32_t strand_anchor;
} hlo_cand; /* 16 bytes */

#define HLO_CAND_PACK(strand, a_start, a_end) \
    ((uint32_t)((strand) & 1u) | ((uint32_t)(a_start) << 1) | ((uint32_t)(a_end) << 16))

/*
 * One verified overlap = the content of one line of sfo2overlaps.py's sorted 10-field
 * temporary file (sfo2overlaps.py:37-50): ids numerically ordered (idA < idB), SFO fields
 * after flip_N / flip_I.  ori: 0 = 'N' (same strand), 1 = 'I'.
 * idA/idB are read ids local to their cluster (the decimal names written by rename_fas.py).
 */
typedef struct hlo_rec {
    uint32_t idA, idB;
    int32_t OHA, OHB;
    uint32_t OLA, OLB;
    uint32_t K;
    uint32_t ori;
} hlo_rec; /* 32 bytes */

/* The same record in 16 bytes (option "compact"): half the device->host traffic of the result, which is what
 * bounds the end-to-end rate once several GPUs share a host.  Valid because reads are at most 32767 bases:
 * OLB == OLA always, |OHA|, |OHB| < 32768, K < 32768. */
typedef struct hlo_rec16 {
    uint32_t idA, idB;
    int16_t OHA, OHB;
    uint16_t OLA;
    uint16_t K_ori; /* K | ori << 15 */
} hlo_rec16; /* 16 bytes */

static inline void hlo_rec16_expand(const hlo_rec16 *c, hlo_rec *r) {
    r->idA = c->idA, r->idB = c->idB, r->OHA = c->OHA, r->OHB = c->OHB;
    r->OLA = c->OLA, r->OLB = c->OLA, r->K = c->K_ori & 0x7fffu, r->ori = c->K_ori >> 15;
}

/* The same record in 8 bytes (option "compact" = 2): a quarter of hlo_rec's device->host traffic.  Valid when every
 * read name of the batch is below 65536 (a cluster of HyLight's --size 15000 has 30000 reads).  OHB and OLA are
 * not stored: with the two read lengths, which the caller has, minimap22sfo.py:51-59 gives
 *   OHB = OHA + len(idB) - len(idA)      OLA = OLB = OHA >= 0 ? min(len(idA) - OHA, len(idB)) : min(len(idB) + OHA, len(idA))
 * (both hold before and after the flips of sfo2overlaps.py:112-122, which swap the reads together with the fields). */
typedef struct hlo_rec8 {
    uint16_t idA, idB;
    int16_t OHA;
    uint16_t K_ori; /* K | ori << 15 */
} hlo_rec8; /* 8 bytes */

static inline void hlo_rec8_expand(const hlo_rec8 *c, uint32_t lenA, uint32_t lenB, hlo_rec *r) {
    const int32_t a = c->OHA, la = (int32_t)lenA, lb = (int32_t)lenB;
    const int32_t ol = a >= 0 ? ((la - a) < lb ? (la - a) : lb) : ((lb + a) < la ? (lb + a) : la);
    r->idA = c->idA, r->idB = c->idB, r->OHA = a, r->OHB = a + lb - la;
    r->OLA = (uint32_t)ol, r->OLB = (uint32_t)ol, r->K = c->K_ori & 0x7fffu, r->ori = c->K_ori >> 15;
}

/* Extra per-record facts the reference computes on the way (PAF columns 10/11) */
typedef struct hlo_rec_aux {
    uint32_t cluster; /* cluster index inside the batch */
    uint32_t blocklen; /* PAF col 11 */
} hlo_rec_aux;

struct hlo_score;
typedef struct hlo_recs {
    const hlo_rec *rec;         /* n records, sorted by (cluster, idA, idB, ...) */
    const hlo_rec_aux *aux;     /* n entries, parallel to rec */
    uint64_t n;
    const uint64_t *cluster_off; /* n_clusters + 1 offsets into rec (NULL for hlo_verify) */
    uint32_t n_clusters;
    const hlo_rec16 *rec16;     /* with option "compact" = 1 (or 2 when a name of the batch is >= 65536): the n records in 16-byte form, rec == NULL */
    const hlo_rec8 *rec8;       /* with option "compact" = 2: the n records in 8-byte form, rec == rec16 == NULL */
    const struct hlo_score *score; /* with option "scores" = 1, qualities in the batch and read names 0..N-1 in file order (the
                                      consumer indexes its reads by name, FastqStorage.cpp:109-117): n entries parallel to the
                                      records; else NULL */
} hlo_recs;

/* One batch of clusters on the host, the input of hlo_submit (and of hlo_overlap_host through its arguments). */
typedef struct hlo_batch {
    const char *bases;           /* all sequences concatenated, ASCII (ACGTUN, any case; other letters are ambiguous) */
    const char *quals;           /* phred+33, same offsets as bases; NULL unless option "scores" is set */
    const uint64_t *base_off;    /* n_reads + 1 offsets into bases */
    uint32_t n_reads;
    const uint32_t *cluster_off; /* n_clusters + 1 offsets into the read list (NULL = one cluster) */
    uint32_t n_clusters;
    const uint32_t *ids;         /* n_reads decimal names (NULL = position inside the cluster, what rename_fas.py writes) */
    /* Alternative to bases / base_off for callers that keep their reads 2-bit packed (a quarter of the upload; hlo_pack2b
     * produces this form): packed != NULL and bases == NULL.  Read r occupies ceil(read_len[r] / 16) consecutive 32-bit
     * words, reads back to back in read order; base i of a read sits in bits [2(i % 16), 2(i % 16) + 2) of its word i / 16,
     * A=0 C=1 G=2 T/U=3; an ambiguous base is stored as 0 and listed in amb as read << 32 | position (any order).
     * quals must be NULL in this form. */
    const uint32_t *packed;
    const uint32_t *read_len;    /* n_reads lengths */
    const uint64_t *amb;         /* n_amb entries, or NULL */
    uint64_t n_amb;
} hlo_batch;

/* minimizer tuple, same two words as minimap2's mm128_t: x = hash<<8 | span, y = rid<<32 | pos<<1 | strand */
typedef struct hlo_mz {
    uint64_t x, y;
} hlo_mz;

/* per-stage device timings of the last hlo_overlap_batch / hlo_verify call, milliseconds (CUDA events) */
typedef struct hlo_timing {
    float h2d_ms, pack_ms, sketch_ms, sort_ms, index_ms, pairgen_ms, verify_ms, dedup_ms, d2h_ms, total_ms;
    float verify_fast_ms;  /* the one-thread-per-candidate verify kernel alone (one launch) */
    float sort_fallback;   /* 1 = a mixed hash-prefix group was too large to repair and the index was redone on all hash bits */
    uint64_t n_reads, n_bases, n_minimizers, n_seed_hits, n_candidates, n_accepted, n_records;
    uint64_t n_deferred;     /* candidates that took the exact slow verify path */
    uint64_t n_overflow_reads; /* query reads whose seed table overflowed shared memory */
    uint64_t verify_launches, total_launches;
    uint64_t digest;         /* option "digest": sum over the output records of a 64-bit mix of their fields (order independent) */
    uint64_t n_reruns;       /* times the batch was enqueued again because a device buffer sized from an estimate was too small */
    uint64_t n_syncs;        /* host waits on the device during the call */
    uint32_t mid_occ_lo, mid_occ_hi; /* smallest / largest per-cluster occurrence cut-off used */
    float score_ms;          /* the fused score kernel (option "scores") */
    uint32_t n_sketch_reruns; /* reads the lean sketch automaton handed to the complete one (equal window keys met) */
} hlo_timing;

typedef struct hlo_ctx hlo_ctx;
typedef struct hlo_reads hlo_reads;

void hlo_params_default(hlo_params *p);
int hlo_abi_version(void);
const char *hlo_strerror(int status);
const char *hlo_last_error(const hlo_ctx *ctx);

/* Number of visible CUDA devices (0 if none / driver missing). */
int hlo_device_count(void);

int hlo_ctx_create(int device, const hlo_params *params, hlo_ctx **out);
void hlo_ctx_destroy(hlo_ctx *ctx);

/* Page-locked host memory for the input buffers (plain malloc'd memory works too, but the
 * host->device copy of pageable memory is staged and slower). */
void *hlo_host_alloc(size_t bytes);
void hlo_host_free(void *p);

/*
 * Upload and 2-bit-pack a batch of clusters.  Replaces rename_fas.py:48-75 + fastq2fasta.py:32-40
 * + minimap2's own sequence encoding: reads of cluster c are seq[cluster_off[c] .. cluster_off[c+1]),
 * read i of a cluster gets the decimal name i (rename_fas.py) unless ids != NULL gives it.
 *   bases      : all sequences concatenated, ASCII (ACGTUN, any case; other letters are ambiguous)
 *   base_off   : n_reads + 1 offsets into bases
 *   cluster_off: n_clusters + 1 offsets into the read list (NULL = one cluster)
 *   ids        : n_reads decimal names (NULL = position inside the cluster)
 */
int hlo_pack_reads(hlo_ctx *ctx, const char *bases, const uint64_t *base_off, uint32_t n_reads,
                   const uint32_t *cluster_off, uint32_t n_clusters, const uint32_t *ids, hlo_reads **out);
void hlo_reads_free(hlo_reads *reads);

/* Parity entry points ----------------------------------------------------------------------- */

/* (w,k) minimizers of every read, in read order, exactly as minimap2's mm_sketch emits them. */
int hlo_sketch(hlo_ctx *ctx, const hlo_reads *reads, const hlo_mz **mz, const uint64_t **read_off, uint64_t *n);

/* Chains (candidates) of every query read; sorted by (q, t, strand, diag). */
int hlo_candidates(hlo_ctx *ctx, const hlo_reads *reads, const hlo_cand **cands, uint64_t *n);

/*
 * Verify a caller-supplied candidate list (the reference's own candidate list in the parity tests):
 * base-level ungapped extension with minimap2's end rule, filter_ovlp_inline.py:58,80-87 thresholds,
 * minimap22sfo.py:47-74 geometry and sfo2overlaps.py:41-49 canonicalisation.  One record per accepted
 * candidate, no per-pair dedup, sorted by (q, t, strand, diag) of the candidate.
 */
int hlo_verify(hlo_ctx *ctx, const hlo_reads *reads, const hlo_cand *cands, uint64_t n, hlo_recs *out);

/* Full path --------------------------------------------------------------------------------- */

/* sketch -> sort -> seed/chain -> verify -> per-pair keep-longest -> sorted records, per cluster. */
int hlo_overlap_batch(hlo_ctx *ctx, const hlo_reads *reads, hlo_recs *out);

/* pack + overlap in one call from host buffers (the e2e entry the bench times) = hlo_submit + hlo_wait. */
int hlo_overlap_host(hlo_ctx *ctx, const char *bases, const uint64_t *base_off, uint32_t n_reads,
                     const uint32_t *cluster_off, uint32_t n_clusters, hlo_recs *out);

/*
 * The same, asynchronous: hlo_submit enqueues the upload and every kernel of the chain on the ctx
 * stream and returns without waiting for the device; hlo_wait blocks until the records are on the host.  No stage
 * waits for the host in between: buffers are sized from estimates learned from earlier batches, counts stay in a
 * device counter block, and hlo_wait reruns the batch with larger buffers in the rare case one was too small
 * (hlo_timing.n_reruns).  One batch in flight per ctx; a caller overlaps batches by cycling over two or three
 * contexts.  The input buffers must stay valid and unchanged until hlo_wait returns.
 */
int hlo_submit(hlo_ctx *ctx, const hlo_batch *batch);
int hlo_wait(hlo_ctx *ctx, hlo_recs *out);

/* Host-side helper: ASCII reads -> the packed form of hlo_batch (plain C loop; a caller with many cores splits the
 * read list).  words: sum over reads of ceil(len / 16) u32; read_len: n_reads; amb: room for amb_cap entries — the
 * number of ambiguous bases found is returned in *n_amb, HLO_ERANGE when it exceeds amb_cap.  Returns the number of
 * words written in *n_words. */
int hlo_pack2b(const char *bases, const uint64_t *base_off, uint32_t n_reads, uint32_t *words, uint64_t *n_words, uint32_t *read_len,
               uint64_t *amb, uint64_t amb_cap, uint64_t *n_amb);

int hlo_get_timing(const hlo_ctx *ctx, hlo_timing *t);

/* Next step of the path: what the consumer computes for every line of the .overlaps
 * file, tools/HaploConduct/src/EdgeCalculator.cpp:26-139 (score, overlap_score) in the S-S case of
 * compute_overlap (:199-233): orient read ID1 by ORI1 and read ID2 by ORI2 (sequence reverse-complemented,
 * phreds reversed), compare seq1[POS1 + i] with seq2[i] over L = min(L1 - POS1, L2), skip positions with an
 * ambiguous base, per position p(same strain) from the two phred error probabilities, score = exp(mean log p),
 * mismatch_rate = mismatches / compared positions; score 0 when a position has p < mismatch_prob, when a read is
 * shorter than min_read_len, or when nothing was compared.  FP64, accumulated in position order.
 *   quals : phred+33 strings, concatenated with the same offsets as the bases given to hlo_pack_reads
 *   rec_cluster : cluster index of every record (hlo_rec_aux.cluster), or NULL when the batch has one cluster
 * Only for read names 0..N-1 in file order (HLO_EUNSUPPORTED otherwise). */
typedef struct hlo_score {
    double score;
    double mismatch_rate;
} hlo_score;
int hlo_score_records(hlo_ctx *ctx, const hlo_reads *reads, const char *quals, const hlo_rec *rec, const uint32_t *rec_cluster,
                      uint64_t n, double mismatch_prob, uint32_t min_read_len, const hlo_score **out);
/* ViralQuasispecies --mismatch / --min_read_len for the fused score of hlo_submit (option "scores"): the score kernel
 * then runs on the device-resident records before they are downloaded, and hlo_recs.score comes back with them. */
int hlo_set_score_params(hlo_ctx *ctx, double mismatch_prob, uint32_t min_read_len);

/* Tuning / test knobs.  "verify_exact" (0|1): route every candidate through the base-by-base verify
 * kernel; "pairgen_slots_log2" (5..10): slots of the per-read shared-memory seed table;
 * "repack" (0|1): hlo_overlap_batch re-runs the 2-bit packer from the ASCII bases kept in HBM (so a
 * device-resident measurement covers the packer); "d2h" (0|1): copy the records back to the host
 * (with 0 hlo_overlap_batch returns n only; rec/aux/cluster_off are NULL); "aux" (0|1): also copy the
 * hlo_rec_aux array back (with 0 hlo_recs.aux is NULL: a caller that only writes the files needs rec and
 * cluster_off); "compact" (0|1|2): hlo_overlap_batch / hlo_overlap_host / hlo_wait return hlo_recs.rec16
 * (16-byte records) or hlo_recs.rec8 (8-byte records) instead of rec; "scores" (0|1): fused overlap score (see
 * hlo_set_score_params); "digest" (0|1): hlo_timing.digest; "sketch_generic" / "pairgen_generic" (0|1): force the
 * generic (any w, k / 64-bit keyed) kernels instead of the specialised ones; "sort_segmented" (1|0): sort every
 * cluster's minimizers as its own segment (0: one segment with the cluster index in the key); "sort_prefix" (1|0):
 * sort only a prefix of the hash and repair the groups in which two hashes share it (0: all hash bits);
 * "sort_hash_bits" (0 = automatic): length of that prefix; "cap_scale_pct" (default 100): scales the first-batch
 * buffer estimates (tests use a small value to force the rerun path).  None of them changes a result. */
int hlo_set_option(hlo_ctx *ctx, const char *name, int64_t value);

/* Memory-safety check: the packed rows of a read set are padded on both sides because shifted vector loads may READ past
 * a row; nothing may ever WRITE there.  The padding is filled with a canary byte when the reads are packed; this returns
 * HLO_OK while it is intact (reads == NULL: the read set of the last hlo_submit on this ctx), HLO_ECUDA when a kernel wrote
 * into it.  Waits for the ctx stream. */
int hlo_debug_check_reads(hlo_ctx *ctx, const hlo_reads *reads);

/* The CUDA stream (cudaStream_t) every call on this ctx enqueues its work on, so that a caller can
 * bracket calls with its own CUDA events. */
void *hlo_stream(const hlo_ctx *ctx);

/* Text emitters (host): byte-identical formatting to the reference scripts ------------------- */

/* 8-field SFO lines (minimap22sfo.py:73) in the string-ordered form of quirk D8. */
int hlo_write_sfo(const hlo_rec *rec, uint64_t n, const char *path);
/* 13-field SAVAGE .overlaps lines (sfo2overlaps.py:150-200), written to a temp file then renamed. */
int hlo_write_overlaps(const hlo_rec *rec, uint64_t n, const char *path);
/* hlo_write_overlaps / hlo_write_sfo for 16-byte records */
int hlo_write_overlaps16(const hlo_rec16 *rec, uint64_t n, const char *path);
int hlo_write_sfo16(const hlo_rec16 *rec, uint64_t n, const char *path);
/* ... and for 8-byte records: len_by_id[id] = length of the read named id (n_ids entries) */
int hlo_write_overlaps8(const hlo_rec8 *rec, uint64_t n, const uint32_t *len_by_id, uint32_t n_ids, const char *path);
int hlo_write_sfo8(const hlo_rec8 *rec, uint64_t n, const uint32_t *len_by_id, uint32_t n_ids, const char *path);
/*
 * Binary hand-off past the text reader of the consumer (tools/HaploConduct/src/EdgeCalculator.cpp:569-604;
 * SURVEY.md §8f row 4): a 32-byte header {"HLOB", u32 version = 1, u32 record bytes = 16, u32 flags (bit 0: scores
 * follow), u64 n, u64 reserved} then n hlo_rec16, then (flag bit 0) n hlo_score.  INTEGRATION.md §5 holds the reader
 * patch.  score may be NULL.  Written to a temp file, then renamed.
 */
int hlo_write_binary16(const hlo_rec16 *rec, const struct hlo_score *score, uint64_t n, const char *path);
/* The same lines with two more columns, hlo_score.score and .mismatch_rate (%.17g): the file a HaploConduct whose
 * reader takes the precomputed score would consume (15 fields; the stock reader wants exactly 13). */
int hlo_write_overlaps_scored(const hlo_rec *rec, const struct hlo_score *score, uint64_t n, const char *path);
/* Format one 13-field line into buf (>= 128 bytes); returns its length. */
int hlo_format_overlap_line(const hlo_rec *rec, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* HLOVLP_H */
