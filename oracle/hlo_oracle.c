/*
 * hlo_oracle.c — CPU restatement of HyLight's short-read overlap chain.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libhlovlp) never does.
 *
 * What is restated, and from where:
 *   (1) minimap2's arithmetic for the flag set of polyte.tune_params.py:510
 *       ("--sr -X -c -k 21 -w 11 -s 60 -m 30 -n 2 -r 0 -A 4 -B 2 --end-bonus 100").
 *       minimap2 is a third-party dependency that is ABSENT from /root/reference and whose version the
 *       reference does not pin (README.md:18,26: "conda install -c bioconda minimap2").  This file
 *       restates its published algorithm (sketch.c mm_sketch, index/seed collection with -X,
 *       same-diagonal chaining with -r 0, the SR-mode ungapped fill and the ksw2 end rule) as
 *       described in SURVEY.md Appendix B.  PARITY UNPINNED for this part: no minimap2 binary or
 *       golden PAF exists in the reference or in this image.
 *   (2) script/filter_ovlp_inline.py:12-39,42-91,94-106, script/minimap22sfo.py:33-75,
 *       script/sfo2overlaps.py:31-53,112-122,150-200.  PINNED: tests/golden/ holds outputs of the
 *       unmodified reference scripts (tests/golden/make_golden.py) and tests/test_oracle_golden.py
 *       checks this file against them byte for byte.
 *   (3) the consumer's ungapped comparison semantics, tools/HaploConduct/src/EdgeCalculator.cpp:67-139
 *       (L = min(L1-pos, L2), N positions skipped, mismatches counted).
 */
#define _GNU_SOURCE
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hlovlp.h"

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                                */

static const uint8_t *nt4_table(void) {
    static uint8_t tab[256];
    static int init = 0;
    if (!init) {
        memset(tab, 4, sizeof tab);
        tab['A'] = tab['a'] = 0;
        tab['C'] = tab['c'] = 1;
        tab['G'] = tab['g'] = 2;
        tab['T'] = tab['t'] = 3;
        tab['U'] = tab['u'] = 3;
        init = 1;
    }
    return tab;
}

void orc_free(void *p) { free(p); }

void orc_params_default(hlo_params *p) {
    p->k = 21;
    p->w = 11;
    p->min_seeds = 2;
    p->min_chain_score = 30;
    p->mid_occ = 0; /* minimap2's default: derived from the index (orc_mid_occ) */
    p->match_score = 4;
    p->mismatch_pen = 2;
    p->ambi_pen = 1;
    p->end_bonus = 100;
    p->min_dp_score = 60;
    p->min_ovlp_len = 60;
    p->max_overhang = 1;
    p->min_identity = 0.98;
    p->overhang_ratio = 0.8;
    p->zdrop = 400;
    p->pair_rule = 0;
}

/* strcmp() of the decimal spellings of a and b — the order minimap2 -X uses on read names
 * (skip_seed: "cmp = strcmp(qname, tname); if (NO_DUAL && cmp > 0) skip"); names are the integers
 * written by rename_fas.py:48-57. */
int orc_name_cmp(uint32_t a, uint32_t b) {
    char sa[16], sb[16];
    snprintf(sa, sizeof sa, "%u", a);
    snprintf(sb, sizeof sb, "%u", b);
    int c = strcmp(sa, sb);
    return c < 0 ? -1 : (c > 0 ? 1 : 0);
}

/* ------------------------------------------------------------------------------------------ */
/* (1a) minimizer sketch — minimap2 sketch.c, mm_sketch(), non-HPC branch                        */

static inline uint64_t mm_hash64(uint64_t key, uint64_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

uint64_t orc_hash64(uint64_t key, uint64_t mask) { return mm_hash64(key, mask); }

/* Emits the minimizers of one sequence in minimap2's order.  Returns the number written (stops
 * writing, but keeps counting, when cap is reached). */
uint64_t orc_sketch(const char *str, int len, int w, int k, uint32_t rid, hlo_mz *out, uint64_t cap) {
    const uint8_t *nt4 = nt4_table();
    const uint64_t shift1 = 2 * (uint64_t)(k - 1), mask = (1ULL << 2 * k) - 1;
    uint64_t kmer[2] = {0, 0};
    hlo_mz buf[256], min = {UINT64_MAX, UINT64_MAX};
    uint64_t n = 0;
    int i, j, l, buf_pos, min_pos, kmer_span = 0;
    if (len <= 0 || w <= 0 || w >= 256 || k <= 0 || k > 28) return 0;
#define PUSH(v)                      \
    do {                             \
        if (n < cap) out[n] = (v);   \
        ++n;                         \
    } while (0)
    memset(buf, 0xff, (size_t)w * 16);
    for (i = l = buf_pos = min_pos = 0; i < len; ++i) {
        int c = nt4[(uint8_t)str[i]];
        hlo_mz info = {UINT64_MAX, UINT64_MAX};
        if (c < 4) {
            int z;
            kmer_span = l + 1 < k ? l + 1 : k;
            kmer[0] = (kmer[0] << 2 | (uint64_t)c) & mask;
            kmer[1] = (kmer[1] >> 2) | (3ULL ^ (uint64_t)c) << shift1;
            if (kmer[0] == kmer[1]) continue; /* strand-symmetric k-mer: skipped before ++l and before the buffer write */
            z = kmer[0] < kmer[1] ? 0 : 1;
            ++l;
            if (l >= k && kmer_span < 256) {
                info.x = mm_hash64(kmer[z], mask) << 8 | (uint64_t)kmer_span;
                info.y = (uint64_t)rid << 32 | (uint32_t)i << 1 | (uint64_t)z;
            }
        } else
            l = 0, kmer_span = 0;
        buf[buf_pos] = info;
        if (l == w + k - 1 && min.x != UINT64_MAX) { /* first full window: identical k-mers were not stored yet */
            for (j = buf_pos + 1; j < w; ++j)
                if (min.x == buf[j].x && buf[j].y != min.y) PUSH(buf[j]);
            for (j = 0; j < buf_pos; ++j)
                if (min.x == buf[j].x && buf[j].y != min.y) PUSH(buf[j]);
        }
        if (info.x <= min.x) { /* a new minimum; write the old one */
            if (l >= w + k && min.x != UINT64_MAX) PUSH(min);
            min = info, min_pos = buf_pos;
        } else if (buf_pos == min_pos) { /* the old minimum left the window */
            if (l >= w + k - 1 && min.x != UINT64_MAX) PUSH(min);
            for (j = buf_pos + 1, min.x = UINT64_MAX; j < w; ++j)
                if (min.x >= buf[j].x) min = buf[j], min_pos = j; /* >= : the rightmost smallest wins */
            for (j = 0; j <= buf_pos; ++j)
                if (min.x >= buf[j].x) min = buf[j], min_pos = j;
            if (l >= w + k - 1 && min.x != UINT64_MAX) {
                for (j = buf_pos + 1; j < w; ++j)
                    if (min.x == buf[j].x && min.y != buf[j].y) PUSH(buf[j]);
                for (j = 0; j <= buf_pos; ++j)
                    if (min.x == buf[j].x && min.y != buf[j].y) PUSH(buf[j]);
            }
        }
        if (++buf_pos == w) buf_pos = 0;
    }
    if (min.x != UINT64_MAX) PUSH(min);
#undef PUSH
    return n;
}

/* Sketch of every read.  *mz_out is malloc'd (orc_free); read_off has n_reads+1 entries. */
int orc_sketch_all(const char *bases, const uint64_t *base_off, uint32_t n_reads, int w, int k, hlo_mz **mz_out,
                   uint64_t *read_off) {
    uint64_t total = 0, cap = 0;
    hlo_mz *mz = NULL;
    for (uint32_t r = 0; r < n_reads; ++r) {
        int len = (int)(base_off[r + 1] - base_off[r]);
        uint64_t need = total + (uint64_t)(len > 0 ? len : 0) + 1;
        if (need > cap) {
            cap = need * 2 + 1024;
            hlo_mz *nm = (hlo_mz *)realloc(mz, cap * sizeof *mz);
            if (!nm) {
                free(mz);
                return HLO_ENOMEM;
            }
            mz = nm;
        }
        read_off[r] = total;
        total += orc_sketch(bases + base_off[r], len, w, k, r, mz + total, cap - total);
    }
    read_off[n_reads] = total;
    *mz_out = mz;
    return HLO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* (1b) seeds -> same-diagonal chains = candidates                                               */

typedef struct {
    uint64_t h; /* x >> 8 */
    uint64_t y;
} idx_ent;

static int idx_cmp(const void *a, const void *b) {
    const idx_ent *p = (const idx_ent *)a, *q = (const idx_ent *)b;
    if (p->h != q->h) return p->h < q->h ? -1 : 1;
    return p->y < q->y ? -1 : (p->y > q->y ? 1 : 0);
}

static int u32_cmp(const void *a, const void *b) {
    const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/*
 * The occurrence cut-off minimap2 derives from an index when -f / --mid-occ is not given (options.c
 * mm_mapopt_update + index.c mm_idx_cal_max_occ, defaults mid_occ_frac = 2e-4f, min_mid_occ = 10,
 * max_mid_occ = 1000000): with a[] = the occurrence counts of the n DISTINCT minimizers,
 *     thres = ks_ksmall(n, a, (uint32_t)((1. - f) * n)) + 1        (k-th smallest, k counted from 0)
 * then raised to min_mid_occ and capped at max_mid_occ.  The reference command line
 * (polyte.tune_params.py:510) passes the long flag --sr, which sets MM_F_SR only — not the preset
 * "-x sr" — so nothing fixes mid_occ and every cluster (one index per minimap2 process) gets its own.
 * idx: the cluster's (hash, y) entries sorted by hash.
 */
static int orc_mid_occ_sorted(const idx_ent *idx, uint64_t nm) {
    if (nm == 0) return 10;
    uint32_t *a = (uint32_t *)malloc(nm * sizeof *a);
    uint64_t n = 0;
    if (!a) return 10;
    for (uint64_t i = 0; i < nm;) {
        uint64_t j = i + 1;
        while (j < nm && idx[j].h == idx[i].h) ++j;
        a[n++] = (uint32_t)(j - i);
        i = j;
    }
    qsort(a, n, sizeof *a, u32_cmp);
    const float f = 2e-4f;
    const uint32_t k = (uint32_t)((1. - f) * n);
    long long thres = (long long)a[k < n ? k : n - 1] + 1;
    free(a);
    if (thres < 10) thres = 10;
    if (thres > 1000000) thres = 1000000;
    return (int)thres;
}

typedef struct {
    uint32_t t;
    uint32_t strand;
    int32_t diag;
    int32_t qp;
} anchor_t;

static int anchor_cmp(const void *a, const void *b) {
    const anchor_t *p = (const anchor_t *)a, *q = (const anchor_t *)b;
    if (p->t != q->t) return p->t < q->t ? -1 : 1;
    if (p->strand != q->strand) return p->strand < q->strand ? -1 : 1;
    if (p->diag != q->diag) return p->diag < q->diag ? -1 : 1;
    return p->qp < q->qp ? -1 : (p->qp > q->qp ? 1 : 0);
}

/*
 * Candidates of a batch of clusters.  For every query read (file order) collect the seed hits of its
 * minimizers inside its own cluster, drop minimizers with more than mid_occ occurrences (P->mid_occ > 0:
 * that value; else minimap2's data-derived cut-off of the cluster, orc_mid_occ_sorted; the
 * high-occurrence rescue selects round(span/500)=0 seeds for reads <= 250 bp), apply
 * -X (no self hits; a seed is skipped when strcmp(qname,tname) > 0), convert to anchors, and chain the
 * anchors of one (target, strand, diagonal) — the only chains "-r 0" (bw = 0) admits.  Chain score =
 * k + sum(min(k, gap)); kept iff #anchors >= -n and score >= -m.
 * Output sorted by (q, t, strand, diag); q/t are batch-global read indices.
 */
int orc_candidates2(const char *bases, const uint64_t *base_off, uint32_t n_reads, const uint32_t *cluster_off,
                    uint32_t n_clusters, const uint32_t *ids, const hlo_params *P, hlo_cand **out, uint64_t *n_out,
                    uint64_t *n_seed_hits, uint32_t *mid_occ_out /* n_clusters entries, or NULL */) {
    uint32_t oc[2];
    if (!cluster_off) {
        oc[0] = 0, oc[1] = n_reads;
        cluster_off = oc;
        n_clusters = 1;
    }
    const int k = P->k, w = P->w;
    uint64_t *roff = (uint64_t *)malloc(((size_t)n_reads + 1) * sizeof *roff);
    hlo_mz *mz = NULL;
    hlo_cand *cands = NULL;
    uint64_t n_c = 0, cap_c = 0, hits_total = 0;
    anchor_t *anc = NULL;
    size_t cap_a = 0;
    int rc = HLO_OK;
    if (!roff) return HLO_ENOMEM;
    rc = orc_sketch_all(bases, base_off, n_reads, w, k, &mz, roff);
    if (rc) {
        free(roff);
        return rc;
    }
    for (uint32_t c = 0; c < n_clusters && rc == HLO_OK; ++c) {
        const uint32_t r0 = cluster_off[c], r1 = cluster_off[c + 1];
        const uint64_t m0 = roff[r0], m1 = roff[r1], nm = m1 - m0;
        idx_ent *idx = (idx_ent *)malloc((nm ? nm : 1) * sizeof *idx);
        if (!idx) {
            rc = HLO_ENOMEM;
            break;
        }
        for (uint64_t i = 0; i < nm; ++i) idx[i].h = mz[m0 + i].x >> 8, idx[i].y = mz[m0 + i].y;
        qsort(idx, nm, sizeof *idx, idx_cmp);
        const int mid_occ = P->mid_occ > 0 ? P->mid_occ : orc_mid_occ_sorted(idx, nm);
        if (mid_occ_out) mid_occ_out[c] = (uint32_t)mid_occ;
        for (uint32_t q = r0; q < r1 && rc == HLO_OK; ++q) {
            const int ql = (int)(base_off[q + 1] - base_off[q]);
            const uint32_t qid = ids ? ids[q] : q - r0;
            size_t na = 0;
            for (uint64_t m = roff[q]; m < roff[q + 1]; ++m) {
                const uint64_t h = mz[m].x >> 8;
                const int span = (int)(mz[m].x & 0xff);
                const uint32_t qpos = (uint32_t)mz[m].y >> 1, qz = (uint32_t)mz[m].y & 1;
                /* bucket = equal range of h */
                uint64_t lo = 0, hi = nm;
                while (lo < hi) {
                    uint64_t mid = (lo + hi) >> 1;
                    if (idx[mid].h < h) lo = mid + 1; else hi = mid;
                }
                uint64_t s = lo;
                hi = nm;
                while (lo < hi) {
                    uint64_t mid = (lo + hi) >> 1;
                    if (idx[mid].h <= h) lo = mid + 1; else hi = mid;
                }
                uint64_t e = lo;
                if ((int64_t)(e - s) > (int64_t)mid_occ) continue;
                for (uint64_t j = s; j < e; ++j) {
                    const uint32_t t = (uint32_t)(idx[j].y >> 32);
                    if (t == q) continue; /* -X: no self hits (self chains would be dropped at filter_ovlp_inline.py:26) */
                    const uint32_t tid = ids ? ids[t] : t - r0;
                    if (orc_name_cmp(qid, tid) > 0) continue; /* -X: each unordered pair once, query = smaller name */
                    ++hits_total;
                    if (na == cap_a) {
                        cap_a = cap_a ? cap_a * 2 : 4096;
                        anchor_t *nn = (anchor_t *)realloc(anc, cap_a * sizeof *anc);
                        if (!nn) {
                            rc = HLO_ENOMEM;
                            break;
                        }
                        anc = nn;
                    }
                    const uint32_t tz = (uint32_t)idx[j].y & 1;
                    const int32_t tp = (int32_t)((uint32_t)idx[j].y >> 1);
                    anchor_t *a = &anc[na++];
                    a->t = t;
                    a->strand = qz != tz;
                    a->qp = a->strand ? ql - ((int32_t)qpos + 1 - span) - 1 : (int32_t)qpos;
                    a->diag = tp - a->qp;
                }
                if (rc) break;
            }
            if (rc) break;
            qsort(anc, na, sizeof *anc, anchor_cmp);
            for (size_t i = 0; i < na;) {
                size_t j = i + 1;
                int score = k, cnt = 1;
                while (j < na && anc[j].t == anc[i].t && anc[j].strand == anc[i].strand && anc[j].diag == anc[i].diag) {
                    int gap = anc[j].qp - anc[j - 1].qp;
                    if (gap > 0) { /* identical anchors cannot occur; keep the DP's dq > 0 rule anyway */
                        score += gap < k ? gap : k;
                        ++cnt;
                    }
                    ++j;
                }
                if (cnt >= P->min_seeds && score >= P->min_chain_score) {
                    if (n_c == cap_c) {
                        cap_c = cap_c ? cap_c * 2 : 1 << 16;
                        hlo_cand *nn = (hlo_cand *)realloc(cands, cap_c * sizeof *cands);
                        if (!nn) {
                            rc = HLO_ENOMEM;
                            break;
                        }
                        cands = nn;
                    }
                    int a_start = anc[i].qp - k + 1, a_end = anc[j - 1].qp + 1;
                    hlo_cand *cd = &cands[n_c++];
                    cd->q = q, cd->t = anc[i].t, cd->diag = anc[i].diag;
                    cd->strand_anchor = HLO_CAND_PACK(anc[i].strand, a_start, a_end);
                }
                i = j;
            }
        }
        free(idx);
    }
    free(anc);
    free(mz);
    free(roff);
    if (rc) {
        free(cands);
        return rc;
    }
    *out = cands;
    *n_out = n_c;
    if (n_seed_hits) *n_seed_hits = hits_total;
    return HLO_OK;
}

int orc_candidates(const char *bases, const uint64_t *base_off, uint32_t n_reads, const uint32_t *cluster_off,
                   uint32_t n_clusters, const uint32_t *ids, const hlo_params *P, hlo_cand **out, uint64_t *n_out,
                   uint64_t *n_seed_hits) {
    return orc_candidates2(bases, base_off, n_reads, cluster_off, n_clusters, ids, P, out, n_out, n_seed_hits, NULL);
}

/* the cut-off alone, per cluster (tests compare it with hlo_timing.mid_occ_lo / _hi) */
int orc_mid_occ(const char *bases, const uint64_t *base_off, uint32_t n_reads, const uint32_t *cluster_off, uint32_t n_clusters,
                const hlo_params *P, uint32_t *mid_occ_out) {
    hlo_cand *c = NULL;
    uint64_t n = 0, h = 0;
    int rc = orc_candidates2(bases, base_off, n_reads, cluster_off, n_clusters, NULL, P, &c, &n, &h, mid_occ_out);
    free(c);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* (1c) base-level extension of one candidate -> PAF row                                         */

typedef struct orc_paf {
    uint32_t q, t;        /* batch-global read indices */
    uint32_t qid, tid;    /* decimal names */
    int32_t ql, qs, qe;   /* PAF cols 2-4 (query coordinates on the original query strand) */
    int32_t strand;       /* 0 '+', 1 '-' */
    int32_t tl, ts, te;   /* PAF cols 7-9 */
    int32_t matches;      /* PAF col 10 */
    int32_t blocklen;     /* PAF col 11 */
    int32_t dp_max;       /* peak local score over the block (mm_update_extra) */
    int32_t clip_l, clip_r, n_ambi;
    int32_t valid;        /* 0: geometry gives no overlap */
    int32_t zdrop_mid;    /* 1: inside the block the score falls more than zdrop below an earlier maximum: minimap2 would
                             split the alignment there (mm_align1's z-drop test); reported, not modelled */
} orc_paf;

static inline int comp_base(int c) { return c < 4 ? 3 - c : 4; }

/*
 * Ungapped extension with minimap2's end rule (align.c mm_align1 in SR mode + ksw2 extension):
 *  - between first and last anchor: ungapped ("if (is_sr) perform ungapped alignment");
 *  - each end is a ksw2 EXTENSION from the outermost anchor outward: cumulative score s (+A match, -B mismatch,
 *    -score_N ambiguous), max = its largest value so far (strict >: the earliest position wins), and the extension
 *    stops early when max - s > zdrop (ksw_apply_zdrop; off the diagonal there is nothing: bw = 0);
 *  - an end where the QUERY ends first reaches the query end iff it was not z-dropped and mqe + end_bonus > max
 *    (mqe = s at the query end; ksw2 "reach_end"); otherwise, and always where the TARGET ends first (the query end is
 *    never reached, mqe = -inf), the block ends at the position of max.
 *    (For the reads of the short-read pass — 150-250 bp at >= 98 % — B * mismatches < end_bonus, so a query-limited end
 *    always reaches the read end; the general rule matters for contigs)
 * matches/blocklen exclude ambiguous columns (mm_update_extra).
 */
static int orc_extend_end(const int *sc, int n, int query_limited, int end_bonus, int zdrop) {
    /* sc[0..n): column scores from the anchor outward; returns the number of columns kept */
    int s = 0, best = 0, best_n = 0, dropped = 0;
    for (int i = 0; i < n; ++i) {
        s += sc[i];
        if (s > best) best = s, best_n = i + 1;
        else if (best - s > zdrop) {
            dropped = 1;
            break;
        }
    }
    if (query_limited && !dropped && s + end_bonus > best) return n;
    return best_n;
}

void orc_extend(const char *bases, const uint64_t *base_off, const hlo_params *P, const hlo_cand *cd, uint32_t qid,
                uint32_t tid, orc_paf *o) {
    const uint8_t *nt4 = nt4_table();
    const char *qs_ = bases + base_off[cd->q], *ts_ = bases + base_off[cd->t];
    const int ql = (int)(base_off[cd->q + 1] - base_off[cd->q]), tl = (int)(base_off[cd->t + 1] - base_off[cd->t]);
    const int strand = (int)(cd->strand_anchor & 1u), D = cd->diag;
    int a_start = (int)((cd->strand_anchor >> 1) & 0x7fffu), a_end = (int)(cd->strand_anchor >> 16);
    const int js = D < 0 ? -D : 0, je = ql < tl - D ? ql : tl - D;
    memset(o, 0, sizeof *o);
    o->q = cd->q, o->t = cd->t, o->qid = qid, o->tid = tid, o->ql = ql, o->tl = tl, o->strand = strand;
    if (je <= js) return;
    if (a_end == 0) a_start = je, a_end = js; /* no anchor information: each end region is the whole overlap */
    if (a_start < js) a_start = js;
    if (a_end > je) a_end = je;
#define QB(j) (strand ? comp_base(nt4[(uint8_t)qs_[ql - 1 - (j)]]) : nt4[(uint8_t)qs_[(j)]])
#define TB(j) (nt4[(uint8_t)ts_[(j) + D]])
#define SC(a, b) (((a) > 3 || (b) > 3) ? -P->ambi_pen : ((a) == (b) ? P->match_score : -P->mismatch_pen))
    int clip_l = 0, clip_r = 0;
    int *sc = (int *)malloc((size_t)(je - js + 1) * sizeof *sc);
    {   /* left end: target-limited iff D < 0 (then js = -D > 0 = where the target starts), else the query starts at 0 */
        const int n = a_start - js;
        for (int i = 0; i < n; ++i) {
            const int j = a_start - 1 - i, a = QB(j), b = TB(j);
            sc[i] = SC(a, b);
        }
        clip_l = n - orc_extend_end(sc, n, D >= 0, P->end_bonus, P->zdrop);
    }
    {   /* right end: target-limited iff tl - D < ql */
        const int n = je - a_end;
        for (int i = 0; i < n; ++i) {
            const int j = a_end + i, a = QB(j), b = TB(j);
            sc[i] = SC(a, b);
        }
        clip_r = n - orc_extend_end(sc, n, !(tl - D < ql), P->end_bonus, P->zdrop);
    }
    free(sc);
    const int bs = js + clip_l, be = je - clip_r;
    int n_ambi = 0, n_diff = 0, s = 0, dp_max = 0, run_max = 0, run = 0, zmid = 0;
    for (int j = bs; j < be; ++j) {
        int a = QB(j), b = TB(j);
        if (a > 3 || b > 3) ++n_ambi;
        else if (a != b) ++n_diff;
        s += SC(a, b);
        if (s < 0) s = 0;
        else if (s > dp_max) dp_max = s;
        run += SC(a, b);
        if (run > run_max) run_max = run;
        else if (run_max - run > P->zdrop) zmid = 1;
    }
#undef QB
#undef TB
#undef SC
    o->valid = 1;
    o->clip_l = clip_l, o->clip_r = clip_r, o->n_ambi = n_ambi;
    o->blocklen = (be - bs) - n_ambi;
    o->matches = o->blocklen - n_diff;
    o->dp_max = dp_max;
    o->zdrop_mid = zmid;
    if (strand) o->qs = ql - be, o->qe = ql - bs;
    else o->qs = bs, o->qe = be;
    o->ts = bs + D, o->te = be + D;
}

int orc_extend_all(const char *bases, const uint64_t *base_off, const uint32_t *read_cluster_base, const uint32_t *ids,
                   const hlo_params *P, const hlo_cand *cands, uint64_t n, orc_paf *out) {
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t q = cands[i].q, t = cands[i].t;
        uint32_t qid = ids ? ids[q] : q - (read_cluster_base ? read_cluster_base[q] : 0);
        uint32_t tid = ids ? ids[t] : t - (read_cluster_base ? read_cluster_base[t] : 0);
        orc_extend(bases, base_off, P, &cands[i], qid, tid, &out[i]);
    }
    return HLO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* (2a) filter_ovlp_inline.py                                                                    */

/* rm_intermatch, filter_ovlp_inline.py:42-91 (and the same two tests of filter_trans_ovlp_inline_v3.py:49-77).
 * Returns bit 0: the row passes length and identity; bit 1: ... and the overhang test. */
int orc_filter_row2(const orc_paf *r, const hlo_params *P) {
    if (!r->valid) return 0;
    const int l1 = r->ql, b1 = r->qs, e1 = r->qe, l2 = r->tl;
    if (r->blocklen < P->min_ovlp_len) return 0;
    if (r->blocklen == 0) return 0; /* python would raise ZeroDivisionError; no row with blocklen 0 is an overlap */
    if ((double)r->matches / (double)r->blocklen < P->min_identity) return 0;
    int b2, e2;
    if (r->strand) b2 = l2 - r->te, e2 = l2 - r->ts;
    else b2 = r->ts, e2 = r->te;
    const int overhang = (b1 < b2 ? b1 : b2) + ((l1 - e1) < (l2 - e2) ? (l1 - e1) : (l2 - e2));
    const int maplen = (e1 - b1) > (e2 - b2) ? (e1 - b1) : (e2 - b2);
    const double x = (double)maplen * P->overhang_ratio;
    /* python: overhang > min(o, maplen * r); min() returns x only when x < o */
    if (x < (double)P->max_overhang) {
        if ((double)overhang > x) return 1;
    } else if (overhang > P->max_overhang)
        return 1;
    return 3;
}

/* Returns 1 if filter_ovlp_inline.py's rm_intermatch keeps the row. */
int orc_filter_row(const orc_paf *r, const hlo_params *P) { return orc_filter_row2(r, P) == 3; }

/* ------------------------------------------------------------------------------------------ */
/* (2b) minimap22sfo.py:47-74 + sfo2overlaps.py:37-50 (flip_N :112-117, flip_I :119-122)          */

typedef struct orc_sfo {
    uint32_t idA, idB; /* as minimap22sfo.py writes them (string-ordered) */
    int32_t ori;       /* 0 'N', 1 'I' */
    int32_t OHA, OHB, OLA, OLB, K;
} orc_sfo;

static int str_gt_u32(uint32_t a, uint32_t b) { return orc_name_cmp(a, b) > 0; }

void orc_paf_to_sfo(const orc_paf *r, orc_sfo *s) {
    const int qlen = r->ql, qstart = r->qs, slen = r->tl, sstart = r->ts, send = r->te;
    int OHA, OHB, OLA;
    s->idA = r->qid, s->idB = r->tid, s->ori = r->strand;
    if (!r->strand) {
        OHA = qstart - sstart;
        OHB = slen - sstart - (qlen - qstart);
    } else {
        OHA = qstart - (slen - send);
        OHB = send - (qlen - qstart);
    }
    if (OHA >= 0) OLA = (qlen - OHA) < slen ? (qlen - OHA) : slen;
    else OLA = (slen + OHA) < qlen ? (slen + OHA) : qlen;
    if (str_gt_u32(s->idA, s->idB)) { /* minimap22sfo.py:62 compares the id STRINGS */
        uint32_t tmp = s->idA;
        s->idA = s->idB, s->idB = tmp;
        if (!r->strand) OHA = -OHA, OHB = -OHB;
        else {
            int t2 = OHA;
            OHA = OHB, OHB = t2;
        }
    }
    s->OHA = OHA, s->OHB = OHB, s->OLA = OLA, s->OLB = OLA, s->K = r->blocklen - r->matches;
}

/* sfo2overlaps.py:37-50 with num_pairs = 0: numeric canonicalisation -> hlo_rec */
void orc_sfo_to_rec(const orc_sfo *s, hlo_rec *o) {
    if (s->idA > s->idB) {
        o->idA = s->idB, o->idB = s->idA;
        if (s->ori) o->OHA = s->OHB, o->OHB = s->OHA;  /* flip_I */
        else o->OHA = -s->OHA, o->OHB = -s->OHB;       /* flip_N */
        o->OLA = (uint32_t)s->OLB, o->OLB = (uint32_t)s->OLA;
    } else {
        o->idA = s->idA, o->idB = s->idB, o->OHA = s->OHA, o->OHB = s->OHB;
        o->OLA = (uint32_t)s->OLA, o->OLB = (uint32_t)s->OLB;
    }
    o->K = (uint32_t)s->K;
    o->ori = (uint32_t)s->ori;
}

/* the SFO line minimap22sfo.py would have written for a record (inverse of the flip when the
 * string order of the ids differs from the numeric order) */
void orc_rec_to_sfo(const hlo_rec *r, orc_sfo *s) {
    s->ori = (int32_t)r->ori, s->K = (int32_t)r->K;
    if (str_gt_u32(r->idA, r->idB)) {
        s->idA = r->idB, s->idB = r->idA;
        if (r->ori) s->OHA = r->OHB, s->OHB = r->OHA;
        else s->OHA = -r->OHA, s->OHB = -r->OHB;
        s->OLA = (int32_t)r->OLB, s->OLB = (int32_t)r->OLA;
    } else {
        s->idA = r->idA, s->idB = r->idB, s->OHA = r->OHA, s->OHB = r->OHB;
        s->OLA = (int32_t)r->OLA, s->OLB = (int32_t)r->OLB;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* (2c) sort + 13-column lines, sfo2overlaps.py:53, :150-200                                      */

int orc_format_tmp_line(const hlo_rec *r, char *buf, size_t cap) {
    /* idA idB sfo_idA sfo_idB ori OHA OHB OLA OLB K  — the 10-field line of sfo2overlaps.py:47-49 */
    /* fields 3.. are the (possibly flipped) SFO tuple: after a flip the tuple's ids are swapped too, so with
     * num_pairs = 0 fields 3,4 always repeat fields 1,2 */
    return snprintf(buf, cap, "%u\t%u\t%u\t%u\t%c\t%d\t%d\t%u\t%u\t%u", r->idA, r->idB, r->idA, r->idB,
                    r->ori ? 'I' : 'N', r->OHA, r->OHB, r->OLA, r->OLB, r->K);
}

int orc_format_overlap_line(const hlo_rec *r, char *buf, size_t cap) {
    /* get_s_s_overlap, sfo2overlaps.py:150-200 */
    const int OHA = r->OHA, OHB = r->OHB, OLA = (int)r->OLA, OLB = (int)r->OLB;
    const int ovlen = OLA < OLB ? OLA : OLB;
    int readlenA, readlenB;
    uint32_t id1, id2;
    int pos1;
    char ori = r->ori ? '-' : '+', ori1, ori2;
    if (OHA >= 0) {
        if (OHB >= 0) readlenA = OLA + OHA, readlenB = OLB + OHB;
        else readlenA = OLA + OHA - OHB, readlenB = OLB;
        id1 = r->idA, id2 = r->idB, pos1 = OHA, ori1 = '+', ori2 = ori;
    } else {
        if (OHB >= 0) readlenA = OLA, readlenB = -OHA + OLB + OHB;
        else readlenA = OLA - OHB, readlenB = -OHA + OLB;
        id1 = r->idB, id2 = r->idA, pos1 = -OHA, ori1 = ori, ori2 = '+';
    }
    const int minreadlen = readlenA < readlenB ? readlenA : readlenB;
    /* python 3: min(round(100*ovlen/minreadlen), 100) — true division, round-half-even on the double */
    double v = (double)(100 * (long long)ovlen) / (double)minreadlen;
    int save = fegetround();
    fesetround(FE_TONEAREST);
    long long perc = llrint(v);
    fesetround(save);
    if (perc > 100) perc = 100;
    return snprintf(buf, cap, "%u\t%u\t%d\t-\t-\t%c\t%c\t%lld\t-\t%d\t-\ts\ts", id1, id2, pos1, ori1, ori2, perc, ovlen);
}

/* `sort -k1,1n -k2,2n -k3,3n -k4,4n` on the 10-field lines under LC_ALL=C: four numeric keys, then
 * the whole line bytewise as last resort.  Fields 3,4 equal fields 1,2 (see orc_format_tmp_line). */
static int rec_line_cmp(const void *a, const void *b) {
    const hlo_rec *p = (const hlo_rec *)a, *q = (const hlo_rec *)b;
    if (p->idA != q->idA) return p->idA < q->idA ? -1 : 1;
    if (p->idB != q->idB) return p->idB < q->idB ? -1 : 1;
    char la[160], lb[160];
    orc_format_tmp_line(p, la, sizeof la);
    orc_format_tmp_line(q, lb, sizeof lb);
    return strcmp(la, lb);
}

static int rec_equal(const hlo_rec *p, const hlo_rec *q) { return memcmp(p, q, sizeof *p) == 0; }

/* sort + uniq in place; returns the new count */
uint64_t orc_sort_uniq_recs(hlo_rec *r, uint64_t n) {
    if (n == 0) return 0;
    qsort(r, n, sizeof *r, rec_line_cmp);
    uint64_t m = 1;
    for (uint64_t i = 1; i < n; ++i)
        if (!rec_equal(&r[i], &r[m - 1])) r[m++] = r[i];
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* (2d) per-pair dedup                                                                           */

typedef struct {
    orc_paf row;
    uint64_t order;
} dd_ent;

/* canonical rule: per unordered pair keep the row with the largest blocklen; ties: more
 * matches, then '+' before '-', then the smaller diagonal.  (filter_ovlp_inline.py:30-37 keeps the
 * first-seen longest inside each 1000-line chunk of minimap2's output order.) */
static int paf_better(const orc_paf *a, int da, const orc_paf *b, int db) {
    if (a->blocklen != b->blocklen) return a->blocklen > b->blocklen;
    if (a->matches != b->matches) return a->matches > b->matches;
    if (a->strand != b->strand) return a->strand < b->strand;
    return da < db;
}

static int dd_cmp(const void *a, const void *b) {
    const dd_ent *p = (const dd_ent *)a, *q = (const dd_ent *)b;
    if (p->row.q != q->row.q) return p->row.q < q->row.q ? -1 : 1;
    if (p->row.t != q->row.t) return p->row.t < q->row.t ? -1 : 1;
    return p->order < q->order ? -1 : (p->order > q->order ? 1 : 0);
}

/*
 * Full chain for a batch: candidates -> extension -> filter -> canonical per-pair dedup -> records
 * sorted per cluster by sfo2overlaps.py's sort.  rec_out malloc'd; cluster_rec_off has n_clusters+1.
 * If cands_in != NULL it is used instead of the oracle's own candidate list.
 */
int orc_overlap_batch(const char *bases, const uint64_t *base_off, uint32_t n_reads, const uint32_t *cluster_off,
                      uint32_t n_clusters, const uint32_t *ids, const hlo_params *P, int dedup, hlo_rec **rec_out,
                      uint32_t **blocklen_out, uint64_t *cluster_rec_off, uint64_t *stats /* [4]: hits, cands, accepted, records */) {
    uint32_t oc[2];
    if (!cluster_off) {
        oc[0] = 0, oc[1] = n_reads;
        cluster_off = oc;
        n_clusters = 1;
    }
    hlo_cand *cands = NULL;
    uint64_t n_c = 0, hits = 0;
    int rc = orc_candidates(bases, base_off, n_reads, cluster_off, n_clusters, ids, P, &cands, &n_c, &hits);
    if (rc) return rc;
    uint32_t *rbase = (uint32_t *)malloc(((size_t)n_reads + 1) * sizeof *rbase);
    uint32_t *rclu = (uint32_t *)malloc(((size_t)n_reads + 1) * sizeof *rclu);
    dd_ent *rows = (dd_ent *)malloc((n_c ? n_c : 1) * sizeof *rows);
    if (!rbase || !rclu || !rows) {
        free(cands), free(rbase), free(rclu), free(rows);
        return HLO_ENOMEM;
    }
    for (uint32_t c = 0; c < n_clusters; ++c)
        for (uint32_t r = cluster_off[c]; r < cluster_off[c + 1]; ++r) rbase[r] = cluster_off[c], rclu[r] = c;
    uint64_t n_acc = 0;
    for (uint64_t i = 0; i < n_c; ++i) {
        orc_paf row;
        uint32_t q = cands[i].q, t = cands[i].t;
        orc_extend(bases, base_off, P, &cands[i], ids ? ids[q] : q - rbase[q], ids ? ids[t] : t - rbase[t], &row);
        const int pass = orc_filter_row2(&row, P);
        /* pair_rule 0: only rows that pass everything compete for their pair; pair_rule 1: every row that passes length
         * and identity competes, and the winner's overhang test decides the pair (row.valid = 2 marks a failed test) */
        if (P->pair_rule ? !(pass & 1) : pass != 3) continue;
        if (pass != 3) row.valid = 2;
        rows[n_acc].row = row;
        rows[n_acc].order = i;
        ++n_acc;
    }
    /* dedup per unordered pair: the candidate list holds each pair with a fixed query (-X), so (q,t) identifies it */
    uint64_t n_keep = n_acc;
    if (dedup) {
        qsort(rows, n_acc, sizeof *rows, dd_cmp);
        n_keep = 0;
        for (uint64_t i = 0; i < n_acc;) {
            uint64_t j = i + 1, best = i;
            while (j < n_acc && rows[j].row.q == rows[i].row.q && rows[j].row.t == rows[i].row.t) {
                if (paf_better(&rows[j].row, cands[rows[j].order].diag, &rows[best].row, cands[rows[best].order].diag)) best = j;
                ++j;
            }
            if (rows[best].row.valid == 1) rows[n_keep++] = rows[best]; /* (pair_rule 1: the winner failed the overhang test) */
            i = j;
        }
    } else if (P->pair_rule) {
        n_keep = 0;
        for (uint64_t i = 0; i < n_acc; ++i)
            if (rows[i].row.valid == 1) rows[n_keep++] = rows[i];
    }
    hlo_rec *rec = (hlo_rec *)malloc((n_keep ? n_keep : 1) * sizeof *rec);
    uint32_t *blk = (uint32_t *)malloc((n_keep ? n_keep : 1) * sizeof *blk);
    if (!rec || !blk) {
        free(cands), free(rbase), free(rclu), free(rows), free(rec), free(blk);
        return HLO_ENOMEM;
    }
    /* bucket per cluster (rows are in q order = cluster order), then sort each cluster's records */
    memset(cluster_rec_off, 0, ((size_t)n_clusters + 1) * sizeof *cluster_rec_off);
    for (uint64_t i = 0; i < n_keep; ++i) ++cluster_rec_off[rclu[rows[i].row.q] + 1];
    for (uint32_t c = 0; c < n_clusters; ++c) cluster_rec_off[c + 1] += cluster_rec_off[c];
    {
        uint64_t *cur = (uint64_t *)malloc(((size_t)n_clusters + 1) * sizeof *cur);
        if (!cur) return HLO_ENOMEM;
        memcpy(cur, cluster_rec_off, ((size_t)n_clusters + 1) * sizeof *cur);
        for (uint64_t i = 0; i < n_keep; ++i) {
            orc_sfo s;
            uint64_t p = cur[rclu[rows[i].row.q]]++;
            orc_paf_to_sfo(&rows[i].row, &s);
            orc_sfo_to_rec(&s, &rec[p]);
            /* stash blocklen in OLB's slot? no: keep a parallel array, reordered below */
            blk[p] = (uint32_t)rows[i].row.blocklen;
        }
        free(cur);
    }
    uint64_t n_total = 0;
    for (uint32_t c = 0; c < n_clusters; ++c) {
        uint64_t s = cluster_rec_off[c], e = cluster_rec_off[c + 1], n = e - s;
        /* sort records together with blocklen: sort an index permutation through a temp pair array */
        typedef struct { hlo_rec r; uint32_t b; } pr;
        pr *tmp = (pr *)malloc((n ? n : 1) * sizeof *tmp);
        if (!tmp) return HLO_ENOMEM;
        for (uint64_t i = 0; i < n; ++i) tmp[i].r = rec[s + i], tmp[i].b = blk[s + i];
        qsort(tmp, n, sizeof *tmp, rec_line_cmp); /* hlo_rec is the first member */
        uint64_t m = 0;
        for (uint64_t i = 0; i < n; ++i) {
            if (m && rec_equal(&tmp[i].r, &rec[n_total + m - 1])) continue; /* uniq */
            rec[n_total + m] = tmp[i].r, blk[n_total + m] = tmp[i].b;
            ++m;
        }
        free(tmp);
        cluster_rec_off[c] = n_total;
        n_total += m;
    }
    /* cluster_rec_off[c] now holds starts; fix up the terminator */
    cluster_rec_off[n_clusters] = n_total;
    if (stats) stats[0] = hits, stats[1] = n_c, stats[2] = n_acc, stats[3] = n_total;
    free(cands), free(rbase), free(rclu), free(rows);
    *rec_out = rec;
    if (blocklen_out) *blocklen_out = blk; else free(blk);
    return HLO_OK;
}

/* Verification oracle on a given candidate list (parity gate 1): one record per accepted candidate,
 * in candidate order; accept[i] tells which candidates passed; paf_out (optional) gets every row. */
int orc_verify(const char *bases, const uint64_t *base_off, uint32_t n_reads, const uint32_t *cluster_off,
               uint32_t n_clusters, const uint32_t *ids, const hlo_params *P, const hlo_cand *cands, uint64_t n,
               hlo_rec *rec_out, uint32_t *blocklen_out, uint8_t *accept, orc_paf *paf_out, uint64_t *n_acc_out) {
    uint32_t oc[2];
    if (!cluster_off) {
        oc[0] = 0, oc[1] = n_reads;
        cluster_off = oc;
        n_clusters = 1;
    }
    uint32_t *rbase = (uint32_t *)malloc(((size_t)n_reads + 1) * sizeof *rbase);
    if (!rbase) return HLO_ENOMEM;
    for (uint32_t c = 0; c < n_clusters; ++c)
        for (uint32_t r = cluster_off[c]; r < cluster_off[c + 1]; ++r) rbase[r] = cluster_off[c];
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; ++i) {
        orc_paf row;
        uint32_t q = cands[i].q, t = cands[i].t;
        orc_extend(bases, base_off, P, &cands[i], ids ? ids[q] : q - rbase[q], ids ? ids[t] : t - rbase[t], &row);
        if (paf_out) paf_out[i] = row;
        int keep = orc_filter_row(&row, P);
        if (accept) accept[i] = (uint8_t)keep;
        if (!keep) continue;
        orc_sfo s;
        orc_paf_to_sfo(&row, &s);
        orc_sfo_to_rec(&s, &rec_out[m]);
        if (blocklen_out) blocklen_out[m] = (uint32_t)row.blocklen;
        ++m;
    }
    free(rbase);
    *n_acc_out = m;
    return HLO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* (2e) text I/O used to pin this file against the reference scripts                              */

/* Write PAF rows (12 columns, the `cut -f 1-12` of polyte.tune_params.py:511). */
int orc_write_paf(const orc_paf *rows, uint64_t n, const char *path) {
    FILE *f = fopen(path, "w");
    if (!f) return HLO_EIO;
    for (uint64_t i = 0; i < n; ++i) {
        const orc_paf *r = &rows[i];
        if (!r->valid) continue;
        fprintf(f, "%u\t%d\t%d\t%d\t%c\t%u\t%d\t%d\t%d\t%d\t%d\t60\n", r->qid, r->ql, r->qs, r->qe, r->strand ? '-' : '+',
                r->tid, r->tl, r->ts, r->te, r->matches, r->blocklen);
    }
    return fclose(f) ? HLO_EIO : HLO_OK;
}

int orc_write_sfo(const hlo_rec *rec, uint64_t n, const char *path) {
    FILE *f = fopen(path, "w");
    if (!f) return HLO_EIO;
    for (uint64_t i = 0; i < n; ++i) {
        orc_sfo s;
        orc_rec_to_sfo(&rec[i], &s);
        fprintf(f, "%u\t%u\t%c\t%d\t%d\t%d\t%d\t%d\n", s.idA, s.idB, s.ori ? 'I' : 'N', s.OHA, s.OHB, s.OLA, s.OLB, s.K);
    }
    return fclose(f) ? HLO_EIO : HLO_OK;
}

int orc_write_overlaps(const hlo_rec *rec, uint64_t n, const char *path) {
    FILE *f = fopen(path, "w");
    char line[160], prev[160];
    if (!f) return HLO_EIO;
    prev[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        orc_format_overlap_line(&rec[i], line, sizeof line);
        if (i && strcmp(line, prev) == 0) continue; /* the final `uniq`, sfo2overlaps.py:107 */
        fputs(line, f);
        fputc('\n', f);
        strcpy(prev, line);
    }
    return fclose(f) ? HLO_EIO : HLO_OK;
}

/*
 * filter_ovlp_inline.py as a whole on a list of PAF rows in the GIVEN order (chunks of 1000 input
 * lines, rm_intermatch then rm_dupovlp with first-seen key order, :94-106).  keep_idx receives the
 * indices of surviving rows in output order; returns their number.
 */
uint64_t orc_filter_chunked(const orc_paf *rows, uint64_t n, const hlo_params *P, uint64_t *keep_idx) {
    uint64_t n_out = 0;
    uint64_t *sel = (uint64_t *)malloc(1000 * sizeof *sel);
    for (uint64_t c0 = 0; c0 < n; c0 += 1000) {
        uint64_t c1 = c0 + 1000 < n ? c0 + 1000 : n, ns = 0;
        for (uint64_t i = c0; i < c1; ++i) {
            if (!orc_filter_row(&rows[i], P)) continue;
            if (rows[i].qid == rows[i].tid) continue;
            uint32_t a = rows[i].qid, b = rows[i].tid;
            uint64_t j;
            for (j = 0; j < ns; ++j) {
                const orc_paf *o = &rows[sel[j]];
                if ((o->qid == a && o->tid == b) || (o->qid == b && o->tid == a)) break;
            }
            if (j == ns) sel[ns++] = i;
            else if (rows[i].blocklen > rows[sel[j]].blocklen) sel[j] = i;
        }
        for (uint64_t j = 0; j < ns; ++j) keep_idx[n_out++] = sel[j];
    }
    free(sel);
    return n_out;
}

/*
 * filter_trans_ovlp_inline_v3.py (HyLight.py:309-311 pipes minimap2 into it with "-len 90 -iden 0.99 -oh 2 -sfo") on a
 * list of PAF rows in the GIVEN order: per chunk of 1000 input lines (:33-37), a row that fails length / identity or
 * is a self hit is skipped (:49-50); the FIRST remaining row of an unordered pair registers the pair (:52-55) and
 * later rows of the pair are skipped; the registered row is then dropped when overhang > min(o, maplen * 0.8)
 * (:57-71; the ratio 0.8 is literal in the script).  keep_idx: surviving rows in output order.
 */
uint64_t orc_filter_chunked_v3(const orc_paf *rows, uint64_t n, const hlo_params *P, uint64_t *keep_idx) {
    uint64_t n_out = 0;
    uint64_t *seen = (uint64_t *)malloc(1000 * sizeof *seen);
    hlo_params Q = *P;
    Q.overhang_ratio = 0.8;
    for (uint64_t c0 = 0; c0 < n; c0 += 1000) {
        uint64_t c1 = c0 + 1000 < n ? c0 + 1000 : n, ns = 0;
        for (uint64_t i = c0; i < c1; ++i) {
            const int pass = orc_filter_row2(&rows[i], &Q);
            if (!(pass & 1)) continue;
            if (rows[i].qid == rows[i].tid) continue;
            uint32_t a = rows[i].qid, b = rows[i].tid;
            uint64_t j;
            for (j = 0; j < ns; ++j) {
                const orc_paf *o = &rows[seen[j]];
                if ((o->qid == a && o->tid == b) || (o->qid == b && o->tid == a)) break;
            }
            if (j < ns) continue;
            seen[ns++] = i;
            if (pass == 3) keep_idx[n_out++] = i;
        }
    }
    free(seen);
    return n_out;
}

/* the SFO lines that script prints with -sfo (its own minimap22sfo(), :83-102): no swap of the ids */
int orc_write_sfo_v3(const orc_paf *rows, const uint64_t *idx, uint64_t n, const char *path) {
    FILE *f = fopen(path, "w");
    if (!f) return HLO_EIO;
    for (uint64_t i = 0; i < n; ++i) {
        const orc_paf *r = &rows[idx ? idx[i] : i];
        int OHA, OHB, OLA;
        if (!r->strand) OHA = r->qs - r->ts, OHB = r->tl - r->ts - (r->ql - r->qs);
        else OHA = r->qs - (r->tl - r->te), OHB = r->te - (r->ql - r->qs);
        if (OHA >= 0) OLA = (r->ql - OHA) < r->tl ? (r->ql - OHA) : r->tl;
        else OLA = (r->tl + OHA) < r->ql ? (r->tl + OHA) : r->ql;
        fprintf(f, "%u\t%u\t%c\t%d\t%d\t%d\t%d\t%d\n", r->qid, r->tid, r->strand ? 'I' : 'N', OHA, OHB, OLA, OLA, r->blocklen - r->matches);
    }
    return fclose(f) ? HLO_EIO : HLO_OK;
}

/* Convert a list of PAF rows (already filtered) straight to sorted, uniq'd records: minimap22sfo.py +
 * sfo2overlaps.py.  rec must have room for n entries; returns the count. */
uint64_t orc_paf_rows_to_recs(const orc_paf *rows, const uint64_t *idx, uint64_t n, hlo_rec *rec) {
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const orc_paf *r = &rows[idx ? idx[i] : i];
        orc_sfo s;
        orc_paf_to_sfo(r, &s);
        orc_sfo_to_rec(&s, &rec[m]);
        if (rec[m].idA == rec[m].idB) continue; /* sfo2overlaps.py:69 */
        ++m;
    }
    return orc_sort_uniq_recs(rec, m);
}

size_t orc_sizeof_paf(void) { return sizeof(orc_paf); }

/* ------------------------------------------------------------------------------------------ */
/* (3) the consumer's phred-weighted overlap score: EdgeCalculator.cpp:26-56 (score), :59-63        */
/*     (phred_to_prob), :67-139 (overlap_score), S-S orientation handling :199-221                  */

static double orc_phred_to_prob(int phred) { return pow(10, -phred / 10.0); }

/* score(): returns log p (<= 0), 1 for an N position, 2 for an unacceptable mismatch */
static double orc_score(char nt1, char nt2, double p1, double p2, double mismatch_thr, int *mismatch_count) {
    double p;
    if (nt1 == 'N' || nt2 == 'N') return 1;
    else if (nt1 == nt2) p = (1 - p1) * (1 - p2) + (p1 * p2) / 3.0;
    else {
        p = p1 * (1 - p2) / 3.0 + p2 * (1 - p1) / 3.0 + (2 / 9.0) * p1 * p2;
        ++*mismatch_count;
    }
    if (p < mismatch_thr) return 2;
    return log(p);
}

static char orc_up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
static char orc_comp(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N'; }

/* one record -> (score, mismatch_rate), reads named 0..N-1 inside cluster `cb` (first read of the cluster) */
void orc_overlap_score(const char *bases, const char *quals, const uint64_t *base_off, uint32_t cb, const hlo_rec *r,
                       double mismatch_thr, uint32_t min_read_len, double *score_out, double *mismatch_rate_out) {
    /* the 13-column fields of sfo2overlaps.py:150-200 */
    const int OHA = r->OHA;
    uint32_t id1, id2;
    int pos1, ori1, ori2; /* ori: 1 = '+' */
    if (OHA >= 0) id1 = r->idA, id2 = r->idB, pos1 = OHA, ori1 = 1, ori2 = !r->ori;
    else id1 = r->idB, id2 = r->idA, pos1 = -OHA, ori1 = !r->ori, ori2 = 1;
    const char *s1 = bases + base_off[cb + id1], *s2 = bases + base_off[cb + id2];
    const char *q1 = quals + base_off[cb + id1], *q2 = quals + base_off[cb + id2];
    const int L1 = (int)(base_off[cb + id1 + 1] - base_off[cb + id1]), L2 = (int)(base_off[cb + id2 + 1] - base_off[cb + id2]);
    *mismatch_rate_out = 1.0;
    *score_out = 0.0;
    if (pos1 >= L1) return;
    if ((uint32_t)L1 < min_read_len || (uint32_t)L2 < min_read_len) return;
    const int L = (L1 - pos1) < L2 ? (L1 - pos1) : L2;
    double total_score = 0.0, total_len = 0.0;
    int mismatch_count = 0;
    for (int i = 0; i < L; ++i) {
        const int j1 = pos1 + i, j2 = i;
        /* FastqStorage upper-cases on load (FastqStorage.cpp:123); anything but ACGT is treated as N here */
        char n1 = ori1 ? orc_up(s1[j1]) : orc_comp(orc_up(s1[L1 - 1 - j1]));
        char n2 = ori2 ? orc_up(s2[j2]) : orc_comp(orc_up(s2[L2 - 1 - j2]));
        if (n1 == 'U') n1 = 'T';
        if (n2 == 'U') n2 = 'T';
        if (n1 != 'A' && n1 != 'C' && n1 != 'G' && n1 != 'T') n1 = 'N';
        if (n2 != 'A' && n2 != 'C' && n2 != 'G' && n2 != 'T') n2 = 'N';
        const int Q1 = (int)(ori1 ? q1[j1] : q1[L1 - 1 - j1]) - 33, Q2 = (int)(ori2 ? q2[j2] : q2[L2 - 1 - j2]) - 33;
        const double s = orc_score(n1, n2, orc_phred_to_prob(Q1), orc_phred_to_prob(Q2), mismatch_thr, &mismatch_count);
        if (s <= 0) total_score += s, total_len += 1;
        else if (s == 1) continue;
        else return; /* unacceptable mismatch: score 0, mismatch_rate stays 1 */
    }
    if (total_len == 0) return;
    *mismatch_rate_out = (float)mismatch_count / total_len;
    *score_out = exp((1.0 / total_len) * total_score);
}

void orc_score_records(const char *bases, const char *quals, const uint64_t *base_off, const uint32_t *cluster_off,
                       const hlo_rec *rec, const uint32_t *rec_cluster, uint64_t n, double mismatch_thr, uint32_t min_read_len,
                       double *out /* 2 per record */) {
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t cb = (cluster_off && rec_cluster) ? cluster_off[rec_cluster[i]] : 0u;
        orc_overlap_score(bases, quals, base_off, cb, &rec[i], mismatch_thr, min_read_len, &out[2 * i], &out[2 * i + 1]);
    }
}
