// ctx.cu — C ABI of libhlovlp (include/hlovlp.h): context, packing, stage orchestration, timing.
//
// One ctx = one CUDA device + one stream.  A batch is ENQUEUED as a whole — upload, 2-bit packing, sketch, segmented
// radix sort, bucket index, pair generation (all table sizes), verify, records tail, optional score — without the host
// waiting for the device in between: every stage takes its element count from a block of device counters (common.cuh:
// Ctr) its predecessor left there, grids are sized from upper bounds and stride, and buffers are sized from estimates
// learned from earlier batches.  hlo_wait reads the counter block once; a buffer that turned out too small makes it
// grow the estimate and enqueue the batch again (hlo_timing.n_reruns: first batches of a new workload only).  The
// one thing that needs a host decision inside a batch is the size of the record download.
// Stage boundaries are marked with CUDA events so that hlo_get_timing reports device time per stage.
// There is deliberately no CPU code path for any stage: every entry point fails with HLO_ENODEV / HLO_ECUDA when the
// device is missing.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace hlo;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr, cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        else p = nullptr;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr, cap = 0;
    }
    template <class T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr, cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        else p = nullptr;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr, cap = 0;
    }
    template <class T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

enum Stage { ST_BEGIN = 0, ST_H2D, ST_PACK, ST_SKETCH, ST_SORT, ST_INDEX, ST_PAIRGEN, ST_VERIFY, ST_DEDUP, ST_SCORE, ST_D2H, ST_COUNT };

}  // namespace

struct hlo_reads {
    hlo_ctx *ctx = nullptr;
    int device = -1;
    DevReads d;
    DevBuf b_ascii, b_off, b_coff, b_len, b_cbase, b_cidx, b_ids, b_nrank, b_perm, b_meta, b_bases, b_amb, b_qual, b_ambcount;
    std::vector<uint32_t> h_len, h_ids, h_cidx, h_cluster_off, h_perm;
    bool packed_input = false; // the batch arrived 2-bit packed: b_ascii holds the packed words, b_off the words-per-read scan (no ASCII, no base offsets)
    DevBuf b_rlen, b_amblist;
    uint32_t max_id = 0;       // largest read name of the batch (decides whether 8-byte records can hold the names)
    bool has_quals = false;    // b_qual holds the phred strings (32 bytes of padding in front, then a flag byte per read)
    float h2d_ms = 0.f, pack_ms = 0.f;
    void release() {
        b_ascii.release(), b_off.release(), b_coff.release(), b_len.release(), b_cbase.release(), b_cidx.release(), b_ids.release();
        b_nrank.release(), b_perm.release(), b_meta.release(), b_bases.release(), b_amb.release(), b_qual.release(), b_ambcount.release();
        b_rlen.release(), b_amblist.release();
    }
};

struct hlo_ctx {
    int device = 0;
    hlo_params P;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[ST_COUNT + 1] = {};
    cudaEvent_t ev_vf[2] = {}, ev_compute_done = nullptr;
    std::string err;
    LaunchCounter lc;
    hlo_timing tm;
    // options
    int pairgen_slots_log2 = 8;
    int pairgen_slots_auto = 0;  // raised by batches whose reads overflowed the tables
    int verify_exact = 0, repack = 0, d2h = 1, aux_out = 1, compact = 0, pairgen_generic = 0, sketch_generic = 0, sort_prefix = 1,
        sort_hash_bits = 0, sort_segmented = 1, scores = 0, digest = 0, cap_scale_pct = 100, serialize_compute = 1;
    int pg_dg_bits = 0;  // > 0: b_hits holds 32-bit hit words with this many diagonal bits (the shared-memory pair generator will run)
    double score_mismatch_prob = 0.0;
    uint32_t score_min_read_len = 0;
    // what earlier batches needed, per unit of the quantity the estimate scales with (0 = nothing learned yet)
    double need_mz_per_base = 0.0, need_c_per_mz = 0.0, need_a_per_mz = 0.0;
    // the batch in flight
    uint64_t cap_mz = 0, cap_c = 0, cap_a = 0;   // entries the minimizer / candidate / record arrays of this enqueue hold
    uint64_t min_mz = 0, min_c = 0, min_a = 0;   // lower bounds a rerun imposes
    bool force_full_sort = false;                // rerun after a hash-prefix group was too large to repair
    bool prefix_used = false;
    int n_pg_passes = 0;                         // shared-memory pair generator passes enqueued (table sizes slots0 ..)
    bool fast_pairgen = false;
    bool exact_all = false;                      // every candidate of this batch goes through the exact verify kernel
    bool inflight = false;
    int rec_form = 0;                            // 0 hlo_rec, 1 hlo_rec16, 2 hlo_rec8: what this batch downloads
    bool want_score = false;
    unsigned long long *mapped_host = nullptr, *mapped_dev = nullptr;  // the counter block as the host sees it (mapped pinned memory)
    unsigned long long hc[C_COUNT] = {};         // ... copied out after the wait
    // verify tables
    DevBuf b_tab;
    int tab_len = -1;
    VerifyTables tab{};
    // workspaces
    DevBuf b_cnt, b_mzoff, b_mzx, b_mzrr, b_keys[2], b_vals[2], b_tmp, b_flags, b_bid, b_range, b_hits, b_midocc, b_hist, b_seg1;
    DevBuf b_cands, b_candfar, b_counters, b_ovf, b_ovf2, b_ovfcnt, b_gkeys, b_gmn, b_gmx, b_goff;
    DevBuf b_rec16, b_deferred, b_best, b_pos, b_slot, b_ent, b_entg, b_out, b_outc, b_aux, b_coff, b_mz, b_candx, b_sx, b_srp;
    DevBuf b_qual, b_screc, b_scclu, b_scout, b_prob;
    PinBuf h_score;
    bool prob_ready = false;
    double prob_thr = -1.0;  // mismatch threshold the log-p table on the device was built for
    PinBuf h_out, h_aux, h_coff, h_mz, h_roff, h_cands, h_rec16, h_small;
    hlo_reads scratch_reads;
    uint64_t n_mz = 0;  // minimizers of the last finished run
};

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            char buf__[512];                                                                          \
            snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            ctx->err = buf__;                                                                         \
            cudaGetLastError();                                                                       \
            return e__ == cudaErrorMemoryAllocation ? HLO_ENOMEM : HLO_ECUDA;                         \
        }                                                                                             \
    } while (0)

#define FAIL(code, ...)                              \
    do {                                             \
        char buf__[512];                             \
        snprintf(buf__, sizeof buf__, __VA_ARGS__);  \
        ctx->err = buf__;                            \
        return (code);                               \
    } while (0)

// The compute sections of the contexts of one device run one after the other, in submission order, without any host
// lock: a submit makes its stream wait for the event the previous submit on that device recorded behind its last
// kernel.  Uploads (before the wait) and record downloads (after the event) of one context then overlap another's
// kernels, instead of three batches sharing the SMs kernel by kernel and reaching the PCIe link together.
static std::mutex g_order_mutex[64];
static cudaEvent_t g_last_compute[64] = {};

static inline unsigned long long *ctr(hlo_ctx *ctx, int i) { return ctx->b_counters.as<unsigned long long>() + i; }

// waits for the stream and copies the counter block out of the mapped page
static int sync_counters(hlo_ctx *ctx) {
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->tm.n_syncs++;
    memcpy(ctx->hc, ctx->mapped_host, sizeof ctx->hc);
    return HLO_OK;
}
static int enqueue_counter_readback(hlo_ctx *ctx) {
    launch_copy_words(ctx->b_counters.as<uint32_t>(), reinterpret_cast<uint32_t *>(ctx->mapped_dev), 2 * C_COUNT, ctx->stream, &ctx->lc);
    CK(cudaGetLastError());
    return HLO_OK;
}

extern "C" {

void hlo_params_default(hlo_params *p) {
    if (!p) return;
    p->k = 21, p->w = 11, p->min_seeds = 2, p->min_chain_score = 30, p->mid_occ = 0;
    p->match_score = 4, p->mismatch_pen = 2, p->ambi_pen = 1, p->end_bonus = 100, p->min_dp_score = 60;
    p->min_ovlp_len = 60, p->max_overhang = 1, p->min_identity = 0.98, p->overhang_ratio = 0.8;
    p->zdrop = 400, p->pair_rule = 0;
}

int hlo_abi_version(void) { return HLO_ABI_VERSION; }

const char *hlo_strerror(int s) {
    switch (s) {
        case HLO_OK: return "ok";
        case HLO_EINVAL: return "invalid argument";
        case HLO_ENODEV: return "no usable CUDA device (libhlovlp has no CPU fallback)";
        case HLO_ECUDA: return "CUDA runtime error";
        case HLO_ENOMEM: return "out of memory";
        case HLO_EUNSUPPORTED: return "parameter combination outside the supported domain";
        case HLO_ERANGE: return "size exceeds a documented limit";
        case HLO_EIO: return "I/O error";
        default: return "unknown status";
    }
}

const char *hlo_last_error(const hlo_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int hlo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void *hlo_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void hlo_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

static int check_params(const hlo_params &P, std::string &why) {
    char b[256];
#define BAD(...)                         \
    do {                                 \
        snprintf(b, sizeof b, __VA_ARGS__); \
        why = b;                         \
        return HLO_EUNSUPPORTED;         \
    } while (0)
    if (P.k < 1 || P.k > 28) BAD("k=%d outside [1,28]", P.k);
    if (P.w < 1 || P.w > HLO_MAX_W_PUBLIC) BAD("w=%d outside [1,%d]", P.w, HLO_MAX_W_PUBLIC);
    if (P.min_seeds < 1 || P.min_seeds > 2) BAD("min_seeds=%d: only 1 or 2 reduce to the min/max anchor test", P.min_seeds);
    if (P.min_chain_score > 2 * P.k) BAD("min_chain_score=%d > 2k: chain score no longer equals the anchor span", P.min_chain_score);
    if (P.mid_occ < 0) BAD("mid_occ must be >= 0 (0 = minimap2's data-derived cut-off)");
    if (P.match_score < 1 || P.mismatch_pen < 0 || P.ambi_pen < 0) BAD("scores must be A>=1, B>=0, N>=0");
    if (P.min_ovlp_len < 1 || P.min_ovlp_len > 32767) BAD("min_ovlp_len=%d outside [1,32767]", P.min_ovlp_len);
    if (P.max_overhang < 0 || P.max_overhang > 15) BAD("max_overhang=%d outside [0,15]", P.max_overhang);
    if (!(P.min_identity > 0.0 && P.min_identity <= 1.0)) BAD("min_identity must be in (0,1]");
    if (!(P.overhang_ratio >= 0.0)) BAD("overhang_ratio must be >= 0");
    if (P.zdrop < 0 || P.end_bonus < 0) BAD("zdrop and end_bonus must be >= 0");
    if (P.pair_rule != 0 && P.pair_rule != 1) BAD("pair_rule must be 0 (filter_ovlp_inline.py) or 1 (filter_trans_ovlp_inline_v3.py)");
    // minimap2's own post-alignment filters (-s min_dp_max, mlen >= -m) must never bind on an accepted overlap
    const double worst = (double)P.min_ovlp_len * ((double)P.match_score - (double)(P.match_score + P.mismatch_pen) * (1.0 - P.min_identity));
    if (worst < (double)P.min_dp_score) BAD("-s %d could drop overlaps the filter accepts (min dp score %.1f)", P.min_dp_score, worst);
    if ((double)P.min_ovlp_len * P.min_identity < (double)P.min_chain_score) BAD("-m %d could drop accepted overlaps", P.min_chain_score);
#undef BAD
    return HLO_OK;
}

int hlo_ctx_create(int device, const hlo_params *params, hlo_ctx **out) {
    if (!out) return HLO_EINVAL;
    *out = nullptr;
    int n = hlo_device_count();
    if (n <= 0 || device < 0 || device >= n) return HLO_ENODEV;
    hlo_ctx *ctx = new (std::nothrow) hlo_ctx();
    if (!ctx) return HLO_ENOMEM;
    if (params) ctx->P = *params;
    else hlo_params_default(&ctx->P);
    std::string why;
    int rc = check_params(ctx->P, why);
    if (rc) {
        fprintf(stderr, "hlo_ctx_create: %s\n", why.c_str());
        delete ctx;
        return rc;
    }
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return HLO_ENODEV;
    }
    for (int i = 0; i <= ST_COUNT; ++i) cudaEventCreate(&ctx->ev[i]);
    cudaEventCreate(&ctx->ev_vf[0]), cudaEventCreate(&ctx->ev_vf[1]);
    cudaEventCreateWithFlags(&ctx->ev_compute_done, cudaEventDisableTiming);
    if (cudaHostAlloc((void **)&ctx->mapped_host, sizeof(unsigned long long) * C_COUNT, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&ctx->mapped_dev, ctx->mapped_host, 0) != cudaSuccess ||
        ctx->b_counters.ensure(sizeof(unsigned long long) * C_COUNT) != cudaSuccess) {
        cudaGetLastError();
        hlo_ctx_destroy(ctx);
        return HLO_ENOMEM;
    }
    memset(&ctx->tm, 0, sizeof ctx->tm);
    ctx->scratch_reads.ctx = ctx;
    if (const char *e = getenv("HLO_PAIRGEN_SLOTS_LOG2")) {  // tuning knob for processes that cannot call hlo_set_option (bench.py, the host binary)
        const int v = atoi(e);
        if (v >= 5 && v <= 10) ctx->pairgen_slots_log2 = v;
    }
    *out = ctx;
    return HLO_OK;
}

void hlo_ctx_destroy(hlo_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    {
        std::lock_guard<std::mutex> g(g_order_mutex[ctx->device & 63]);
        if (g_last_compute[ctx->device & 63] == ctx->ev_compute_done) g_last_compute[ctx->device & 63] = nullptr;
    }
    DevBuf *bufs[] = {&ctx->b_tab, &ctx->b_cnt, &ctx->b_mzoff, &ctx->b_mzx, &ctx->b_mzrr, &ctx->b_keys[0], &ctx->b_keys[1], &ctx->b_vals[0],
                      &ctx->b_vals[1], &ctx->b_tmp, &ctx->b_flags, &ctx->b_bid, &ctx->b_range, &ctx->b_hits, &ctx->b_midocc, &ctx->b_hist,
                      &ctx->b_seg1, &ctx->b_cands, &ctx->b_candfar, &ctx->b_counters, &ctx->b_ovf, &ctx->b_ovf2, &ctx->b_ovfcnt, &ctx->b_gkeys, &ctx->b_gmn,
                      &ctx->b_gmx, &ctx->b_goff, &ctx->b_rec16, &ctx->b_deferred, &ctx->b_best, &ctx->b_pos, &ctx->b_slot, &ctx->b_ent,
                      &ctx->b_entg, &ctx->b_out, &ctx->b_outc, &ctx->b_aux, &ctx->b_coff, &ctx->b_mz, &ctx->b_candx, &ctx->b_sx, &ctx->b_srp,
                      &ctx->b_qual, &ctx->b_screc, &ctx->b_scclu, &ctx->b_scout, &ctx->b_prob};
    for (DevBuf *b : bufs) b->release();
    PinBuf *pins[] = {&ctx->h_score, &ctx->h_out, &ctx->h_aux, &ctx->h_coff, &ctx->h_mz, &ctx->h_roff, &ctx->h_cands, &ctx->h_rec16, &ctx->h_small};
    for (PinBuf *b : pins) b->release();
    ctx->scratch_reads.release();
    for (int i = 0; i <= ST_COUNT; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 2; ++i)
        if (ctx->ev_vf[i]) cudaEventDestroy(ctx->ev_vf[i]);
    if (ctx->ev_compute_done) cudaEventDestroy(ctx->ev_compute_done);
    if (ctx->mapped_host) cudaFreeHost(ctx->mapped_host);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// verify tables: the reference's double-precision threshold tests as integer tables
// (filter_ovlp_inline.py:58 "float(a[9]) / float(a[10]) < min_identity", :83 "overhang > min(o, maplen * r)")

static int build_tables(hlo_ctx *ctx, int max_len) {
    if (max_len <= ctx->tab_len) return HLO_OK;
    const int L = std::max(max_len, 256);
    std::vector<uint16_t> h((size_t)3 * (L + 1));
    uint16_t *mm = h.data(), *mc = mm + (L + 1), *mk = mc + (L + 1);
    const double id = ctx->P.min_identity, r = ctx->P.overhang_ratio;
    const int o = ctx->P.max_overhang;
    int run = 0;
    for (int b = 0; b <= L; ++b) {
        int m;
        if (b == 0) m = 1;
        else {
            m = (int)std::ceil(id * (double)b);
            if (m < 0) m = 0;
            while (m > 0 && !((double)(m - 1) / (double)b < id)) --m;
            while (m <= b && ((double)m / (double)b < id)) ++m;
        }
        mm[b] = (uint16_t)m;
        const double x = (double)b * r;
        int c;
        if (x < (double)o) c = (int)std::floor(x);  // python: min(o, x) == x -> keep iff !(overhang > x)
        else c = o;
        mc[b] = (uint16_t)std::max(c, 0);
        if (b >= 1 && b - m > run) run = b - m;
        mk[b] = (uint16_t)run;
    }
    CK(ctx->b_tab.ensure(h.size() * sizeof(uint16_t)));
    CK(cudaMemcpyAsync(ctx->b_tab.p, h.data(), h.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // (h is a local; happens when a ctx first meets a longer read, not per batch)
    ctx->tab.min_matches = ctx->b_tab.as<uint16_t>();
    ctx->tab.max_clip = ctx->tab.min_matches + (L + 1);
    ctx->tab.max_k_upto = ctx->tab.max_clip + (L + 1);
    ctx->tab.max_len = L;
    ctx->tab_len = L;
    return HLO_OK;
}

static VerifyParams verify_params(const hlo_ctx *ctx) {
    VerifyParams v;
    v.k = ctx->P.k, v.a = ctx->P.match_score, v.b = ctx->P.mismatch_pen, v.n = ctx->P.ambi_pen;
    v.min_len = ctx->P.min_ovlp_len, v.max_overhang = ctx->P.max_overhang;
    const int wmax = ctx->P.match_score + std::max(ctx->P.mismatch_pen, ctx->P.ambi_pen);
    v.bad_limit = (33 * ctx->P.match_score + wmax - 1) / wmax - 1;  // A*33 - wmax*m > 0
    if (v.bad_limit < 0) v.bad_limit = 0;
    v.end_bonus = ctx->P.end_bonus, v.zdrop = ctx->P.zdrop, v.pair_rule = ctx->P.pair_rule;
    // a query-limited end keeps its end bonus (and is not z-dropped) while the score cannot fall min(end_bonus, zdrop + 1)
    // below its maximum: B per mismatch.  The fast kernel holds no ambiguous base.
    const int room = std::min(ctx->P.end_bonus, ctx->P.zdrop + 1);
    v.defer_mm = ctx->P.mismatch_pen > 0 ? (room + ctx->P.mismatch_pen - 1) / ctx->P.mismatch_pen : 0x7fffffff;
    if (room <= 0) v.defer_mm = 0;
    return v;
}

// ------------------------------------------------------------------------------------------------
// packing

// host copies of the per-read metadata, needed only by the parity entry points (hlo_sketch, hlo_verify);
// when the metadata was built on the device they are fetched on first use
static int ensure_host_meta(hlo_ctx *ctx, const hlo_reads *Rc) {
    hlo_reads *R = const_cast<hlo_reads *>(Rc);
    const uint32_t n = R->d.n_reads;
    if (R->h_len.size() == n && R->h_perm.size() == n) return HLO_OK;
    R->h_len.resize(n), R->h_ids.resize(n), R->h_cidx.resize(n), R->h_perm.resize(n);
    if (n) {
        cudaStream_t s = ctx->stream;
        CK(cudaMemcpyAsync(R->h_len.data(), R->d.len, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(R->h_ids.data(), R->d.ids, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(R->h_cidx.data(), R->d.cidx, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(R->h_perm.data(), R->d.perm, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (uint32_t i = 0; i < n; ++i) R->h_len[i] &= HLO_LEN_MASK;
    }
    return HLO_OK;
}

// Upload + 2-bit packing of a batch, enqueued on the ctx stream; nothing here waits for the device (names other than
// 0..N-1 go through pageable host vectors, whose copies are staged before cudaMemcpyAsync returns).
static int pack_into(hlo_ctx *ctx, hlo_reads *R, const hlo_batch &B) {
    const char *bases = B.bases;
    const uint64_t *base_off = B.base_off;
    const uint32_t n_reads = B.n_reads;
    const uint32_t *cluster_off = B.cluster_off;
    uint32_t n_clusters = B.n_clusters;
    const uint32_t *ids = B.ids;
    const bool packed = B.packed != nullptr && !bases;
    if (packed) {
        if (!B.read_len && n_reads) FAIL(HLO_EINVAL, "packed batch without read_len");
        if (B.quals) FAIL(HLO_EUNSUPPORTED, "qualities (fused score) need the ASCII form of the batch");
        if (ids) FAIL(HLO_EUNSUPPORTED, "explicit read names need the ASCII form of the batch");
        if (B.n_amb && !B.amb) FAIL(HLO_EINVAL, "n_amb without amb");
    } else if (!bases || !base_off) FAIL(HLO_EINVAL, "hlo_pack_reads: null input");
    uint32_t one[2] = {0, n_reads};
    if (!cluster_off) cluster_off = one, n_clusters = 1;
    if (cluster_off[0] != 0 || cluster_off[n_clusters] != n_reads) FAIL(HLO_EINVAL, "cluster_off must span [0, n_reads]");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    R->ctx = ctx, R->device = ctx->device;
    R->h_cluster_off.assign(cluster_off, cluster_off + n_clusters + 1);
    R->h_len.clear(), R->h_ids.clear(), R->h_cidx.clear(), R->h_perm.clear();
    // one lean pass over the offsets: monotone, length limit, longest read (sizes the rows)
    uint32_t max_len = 0;
    uint64_t packed_words = 0, packed_bases = 0;
    if (packed) {
        uint32_t worst = 0;
        for (uint32_t i = 0; i < n_reads; ++i) {
            const uint32_t l = B.read_len[i];
            worst = l > worst ? l : worst;
            packed_words += (l + 15u) >> 4, packed_bases += l;
        }
        if (worst > 32767) FAIL(HLO_ERANGE, "a read of %u bp exceeds the limit of 32767 (anchor fields of hlo_cand)", worst);
        if (packed_words > 0xfffffff0ull) FAIL(HLO_ERANGE, "packed batch too large");
        max_len = worst;
    } else {
        uint64_t worst = 0;
        bool mono = true;
        for (uint32_t i = 0; i < n_reads; ++i) {
            mono &= base_off[i + 1] >= base_off[i];
            const uint64_t l = base_off[i + 1] - base_off[i];
            worst = l > worst ? l : worst;
        }
        if (!mono) FAIL(HLO_EINVAL, "base_off is not monotone");
        if (worst > 32767) FAIL(HLO_ERANGE, "a read of %llu bp exceeds the limit of 32767 (anchor fields of hlo_cand)", (unsigned long long)worst);
        max_len = (uint32_t)worst;
    }
    uint32_t max_cluster = 0;
    for (uint32_t c = 0; c < n_clusters; ++c) {
        if (cluster_off[c + 1] < cluster_off[c]) FAIL(HLO_EINVAL, "cluster_off not monotone");
        max_cluster = std::max(max_cluster, cluster_off[c + 1] - cluster_off[c]);
    }
    // names 0..N-1 in file order (what rename_fas.py writes): all per-read metadata is derived on the device
    bool plain = true;
    if (ids)
        for (uint32_t c = 0; c < n_clusters && plain; ++c)
            for (uint32_t i = cluster_off[c]; i < cluster_off[c + 1] && plain; ++i) plain = ids[i] == i - cluster_off[c];
    R->max_id = max_cluster ? max_cluster - 1 : 0;
    std::vector<uint32_t> cbase, nrank;
    if (!plain) {
        R->h_len.resize(n_reads), R->h_ids.resize(n_reads), R->h_cidx.resize(n_reads), R->h_perm.resize(n_reads);
        cbase.resize(n_reads), nrank.resize(n_reads);
        for (uint32_t i = 0; i < n_reads; ++i) R->h_len[i] = (uint32_t)(base_off[i + 1] - base_off[i]);
        uint32_t mx = 0;
        for (uint32_t c = 0; c < n_clusters; ++c) {
            const uint32_t r0 = cluster_off[c], r1 = cluster_off[c + 1], N = r1 - r0;
            for (uint32_t i = r0; i < r1; ++i) cbase[i] = r0, R->h_cidx[i] = c, R->h_ids[i] = ids[i], mx = std::max(mx, ids[i]);
            std::vector<uint32_t> ord(N);
            for (uint32_t i = 0; i < N; ++i) ord[i] = r0 + i;
            std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return ids[a] < ids[b] || (ids[a] == ids[b] && a < b); });
            for (uint32_t i = 0; i + 1 < N; ++i)
                if (ids[ord[i]] == ids[ord[i + 1]]) FAIL(HLO_EINVAL, "duplicate id %u in cluster %u", ids[ord[i]], c);
            for (uint32_t i = 0; i < N; ++i) nrank[ord[i]] = r0 + i;
            // strcmp order of the decimal names = the order minimap2 -X compares read names in
            std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return name_less(ids[a], ids[b]); });
            for (uint32_t x = 0; x < N; ++x) R->h_perm[r0 + x] = ord[x];
        }
        R->max_id = mx;
    }
    int rc = build_tables(ctx, (int)max_len);
    if (rc) return rc;
    const uint64_t n_bases = packed ? packed_bases : base_off[n_reads];
    R->packed_input = packed;
    const uint32_t words = (max_len + 15) / 16;
    const uint32_t nv = std::max(1u, (words + 3u) / 4u);
    const uint32_t stride_w = 4u * nv;
    DevReads &d = R->d;
    d.n_reads = n_reads, d.n_clusters = n_clusters, d.stride_w = stride_w, d.nv = nv, d.max_len = max_len, d.n_bases = n_bases;
    d.max_cluster = max_cluster;
    d.plain_names = plain ? 1 : 0;
    const size_t plane_words = (size_t)n_reads * stride_w;
    CK(R->b_ascii.ensure(packed ? packed_words * 4 + 16 : n_bases + 16));
    CK(R->b_off.ensure(((size_t)n_reads + 1) * 8));
    if (packed) {
        CK(R->b_rlen.ensure((size_t)n_reads * 4 + 4));
        CK(R->b_amblist.ensure(B.n_amb * 8 + 8));
        CK(ctx->b_tmp.ensure(scan_tmp_words((uint64_t)n_reads + 1) * 4));
    }
    CK(R->b_coff.ensure(((size_t)n_clusters + 1) * 4));
    CK(R->b_len.ensure((size_t)n_reads * 4 + 4));
    CK(R->b_cbase.ensure((size_t)n_reads * 4 + 4));
    CK(R->b_cidx.ensure((size_t)n_reads * 4 + 4));
    CK(R->b_ids.ensure((size_t)n_reads * 4 + 4));
    CK(R->b_nrank.ensure((size_t)n_reads * 4 + 4));
    CK(R->b_meta.ensure((size_t)n_reads * 16 + 16));
    CK(R->b_perm.ensure((size_t)n_reads * 4 + 4));
    CK(R->b_bases.ensure((plane_words * 2 + 2 * HLO_PAD_WORDS) * 4));
    CK(R->b_amb.ensure((plane_words * 2 + 2 * HLO_PAD_WORDS) * 4));
    CK(R->b_ambcount.ensure(16));  // ambiguous bases of THIS read set: picks the sketch kernel's variant on the device
    d.len = R->b_len.as<uint32_t>(), d.cbase = R->b_cbase.as<uint32_t>(), d.cidx = R->b_cidx.as<uint32_t>();
    d.ids = R->b_ids.as<uint32_t>(), d.nrank = R->b_nrank.as<uint32_t>(), d.perm = R->b_perm.as<uint32_t>();
    d.meta = R->b_meta.as<uint4>();
    d.bases = R->b_bases.as<uint32_t>() + HLO_PAD_WORDS, d.amb = R->b_amb.as<uint32_t>() + HLO_PAD_WORDS;
    R->has_quals = false;
    if (B.quals) CK(R->b_qual.ensure(n_bases + 64 + (size_t)n_reads + 32));  // 32 bytes of padding on both sides, then a flag per read
    CK(cudaEventRecord(ctx->ev[ST_BEGIN], s));
    if (n_reads && packed) {
        CK(cudaMemcpyAsync(R->b_ascii.p, B.packed, packed_words * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(R->b_rlen.p, B.read_len, (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
        if (B.n_amb) CK(cudaMemcpyAsync(R->b_amblist.p, B.amb, B.n_amb * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(R->b_coff.p, R->h_cluster_off.data(), ((size_t)n_clusters + 1) * 4, cudaMemcpyHostToDevice, s));
    } else if (n_reads) {
        CK(cudaMemcpyAsync(R->b_ascii.p, bases, n_bases, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(R->b_off.p, base_off, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(R->b_coff.p, R->h_cluster_off.data(), ((size_t)n_clusters + 1) * 4, cudaMemcpyHostToDevice, s));
        if (B.quals) {
            uint8_t *dq = R->b_qual.as<uint8_t>() + 32;
            CK(cudaMemsetAsync(R->b_qual.p, 0, 32, s));
            CK(cudaMemsetAsync(dq + n_bases, 0, 32, s));
            CK(cudaMemcpyAsync(dq, B.quals, n_bases, cudaMemcpyHostToDevice, s));
            R->has_quals = true;
        }
        if (!plain) {
            CK(cudaMemcpyAsync(d.len, R->h_len.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d.cbase, cbase.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d.cidx, R->h_cidx.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d.ids, R->h_ids.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d.nrank, nrank.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            std::vector<uint4> meta(n_reads);
            for (uint32_t i = 0; i < n_reads; ++i) meta[i] = make_uint4(R->h_len[i], R->h_ids[i], R->h_cidx[i], cbase[i]);
            CK(cudaMemcpyAsync(d.meta, meta.data(), (size_t)n_reads * 16, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d.perm, R->h_perm.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));  // (locals; only batches whose names are not 0..N-1 come here)
        }
    }
    CK(cudaMemsetAsync(ctr(ctx, 0), 0, sizeof(unsigned long long) * C_COUNT, s));
    CK(cudaMemsetAsync(ctr(ctx, C_MIDOCC_LO), 0xff, sizeof(unsigned long long), s));
    CK(cudaMemsetAsync(R->b_ambcount.p, 0, 16, s));
    for (DevBuf *b : {&R->b_bases, &R->b_amb}) {  // canaries in the padding in front of and behind the rows (hlo_debug_check_reads)
        CK(cudaMemsetAsync(b->p, HLO_PAD_CANARY, HLO_PAD_WORDS * 4, s));
        CK(cudaMemsetAsync(b->as<uint32_t>() + HLO_PAD_WORDS + plane_words * 2, HLO_PAD_CANARY, HLO_PAD_WORDS * 4, s));
    }
    CK(cudaEventRecord(ctx->ev[ST_H2D], s));
    if (packed && n_reads) {
        // words per read -> first word of every read (scan) -> rows of both strands; the listed ambiguous bases on top
        uint32_t *wcount = R->b_off.as<uint32_t>(), *woff = wcount + (n_reads + 1);
        launch_packed_meta(R->b_rlen.as<uint32_t>(), R->b_coff.as<uint32_t>(), d, wcount, s, &ctx->lc);
        launch_exclusive_scan(wcount, woff, (uint64_t)n_reads + 1, ctx->b_tmp.as<uint32_t>(), nullptr, s, &ctx->lc);
        launch_unpack2b(R->b_ascii.as<uint32_t>(), woff, R->b_amblist.as<unsigned long long>(), B.n_amb, d, R->b_ambcount.as<uint32_t>(), s,
                        &ctx->lc);
    } else {
        if (plain && n_reads) launch_read_meta(R->b_off.as<uint64_t>(), R->b_coff.as<uint32_t>(), d, s, &ctx->lc);
        launch_pack(R->b_ascii.as<uint8_t>(), R->b_off.as<uint64_t>(), d, R->b_ambcount.as<uint32_t>(), s, &ctx->lc);
    }
    if (R->has_quals)
        launch_qual_flags(R->b_qual.as<uint8_t>() + 32, R->b_off.as<uint64_t>(), n_reads,
                          R->b_qual.as<uint8_t>() + 32 + ((n_bases + 32 + 15) & ~(uint64_t)15), s, &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_PACK], s));
    CK(cudaEventRecord(ctx->ev[ST_COUNT], s));  // (ST_PACK is recorded again when a batch is rerun; this one stays)
    return HLO_OK;
}

extern "C" int hlo_pack_reads(hlo_ctx *ctx, const char *bases, const uint64_t *base_off, uint32_t n_reads,
                              const uint32_t *cluster_off, uint32_t n_clusters, const uint32_t *ids, hlo_reads **out) {
    if (!ctx || !out) return HLO_EINVAL;
    *out = nullptr;
    if (ctx->inflight) FAIL(HLO_EINVAL, "a submitted batch is in flight on this ctx: call hlo_wait first");
    hlo_reads *R = new (std::nothrow) hlo_reads();
    if (!R) return HLO_ENOMEM;
    hlo_batch B{};
    B.bases = bases, B.base_off = base_off, B.n_reads = n_reads, B.cluster_off = cluster_off, B.n_clusters = n_clusters, B.ids = ids;
    ctx->lc = LaunchCounter();
    int rc = pack_into(ctx, R, B);
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {  // the caller's buffers are free again when this returns
        cudaGetLastError();
        ctx->err = "hlo_pack_reads: device error while packing";
        rc = HLO_ECUDA;
    }
    if (!rc) {
        cudaEventElapsedTime(&R->h2d_ms, ctx->ev[ST_BEGIN], ctx->ev[ST_H2D]);
        cudaEventElapsedTime(&R->pack_ms, ctx->ev[ST_H2D], ctx->ev[ST_COUNT]);
        cudaGetLastError();
    }
    if (rc) {
        R->release();
        delete R;
        return rc;
    }
    *out = R;
    return HLO_OK;
}

extern "C" void hlo_reads_free(hlo_reads *R) {
    if (!R) return;
    if (R->device >= 0) cudaSetDevice(R->device);  // R->ctx may already be gone
    R->release();
    delete R;
}

// ------------------------------------------------------------------------------------------------
// stages (all enqueue-only)

static int bits_for(uint64_t v) {
    int b = 1;
    while (b < 63 && (1ull << b) < v) ++b;
    return b;
}

static uint32_t next_pow2_u32(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// buffer sizes of this enqueue: estimates (first batches) or what earlier batches needed plus slack, never below
// what a rerun found out
static void plan_caps(hlo_ctx *ctx, const hlo_reads *R) {
    const DevReads &d = R->d;
    const double scale = std::max(ctx->cap_scale_pct, 1) / 100.0;
    const uint64_t slack = std::max<uint64_t>(1024, (uint64_t)(65536 * std::min(scale, 1.0)));
    // a priori: 1.3x the expected minimizer density 2/(w+1), 2 candidates and 1 record per minimizer; once a batch has
    // been seen: what it needed + 15 % (grids and scans are sized by these bounds, so a tight one is also a fast one)
    const double dens = ctx->need_mz_per_base > 0 ? ctx->need_mz_per_base * 1.15 : 2.6 / (ctx->P.w + 1) * scale;
    uint64_t mz = (uint64_t)((double)d.n_bases * dens) + d.n_reads / 8 + slack;
    mz = std::max(mz, ctx->min_mz);
    mz = std::min<uint64_t>(mz, 0xfffffff0ull);
    uint64_t c = (uint64_t)((double)mz * (ctx->need_c_per_mz > 0 ? ctx->need_c_per_mz * 1.15 : 2.0 * scale)) + slack;
    uint64_t a = (uint64_t)((double)mz * (ctx->need_a_per_mz > 0 ? ctx->need_a_per_mz * 1.15 : 1.0 * scale)) + slack;
    c = std::max(c, ctx->min_c), a = std::max(a, ctx->min_a);
    a = std::min(a, c);  // every record comes from a candidate
    a = std::min<uint64_t>(a, 0x7ffffff0ull);  // records.cu keeps a mark in bit 31 of a record index
    ctx->cap_mz = mz, ctx->cap_c = c, ctx->cap_a = a;
}

__global__ void mz_clear_kernel(uint32_t *mz_off, uint32_t n_off, const unsigned long long *counters, uint64_t cap_mz) {
    if ((uint32_t)counters[C_MZ] <= cap_mz) return;  // (low word: the scan writes 32 bits)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_off; i += gridDim.x * blockDim.x) mz_off[i] = 0u;
}

// sketch.  Leaves mz_off (n_reads+1, name order), mz_rr (+ mz_x when wanted) on the device; with the single-pass
// fast kernel also the radix sort's input pairs.  A batch with more minimizers than cap_mz leaves an EMPTY sketch
// (all offsets zero) behind, so that every later stage is a no-op; C_MZ keeps the true count for the rerun.
static int enqueue_sketch(hlo_ctx *ctx, const hlo_reads *R, bool want_x, bool *keys_ready) {
    const DevReads &d = R->d;
    cudaStream_t s = ctx->stream;
    const uint32_t n = d.n_reads;
    if ((uint64_t)d.n_bases > 0xffffffffull) FAIL(HLO_ERANGE, "batch of %llu bases exceeds the 2^32 minimizer index range", (unsigned long long)d.n_bases);
    const uint64_t cap = ctx->cap_mz;
    CK(ctx->b_cnt.ensure(((size_t)n + 1) * 4));
    CK(ctx->b_mzoff.ensure(((size_t)n + 2) * 4));
    CK(ctx->b_tmp.ensure(scan_tmp_words((uint64_t)n + 1) * 4));
    CK(ctx->b_ovf.ensure((size_t)n * 4 + 4));
    CK(ctx->b_ovf2.ensure((size_t)n * 4 + 4));  // (the pair generator's second list: free until then)
    uint64_t *mz_x = nullptr;
    const bool fast = sketch_fast_supported(ctx->P.w, ctx->P.k) && !ctx->sketch_generic && n > 0;
    if (want_x || !fast) {
        CK(ctx->b_mzx.ensure(cap * 8 + 8));
        mz_x = ctx->b_mzx.as<uint64_t>();
    }
    CK(ctx->b_mzrr.ensure(cap * 8 + 8));
    for (int i = 0; i < 2; ++i) {
        CK(ctx->b_keys[i].ensure(cap * 8 + 8));
        CK(ctx->b_vals[i].ensure(cap * 4 + 4));
    }
    *keys_ready = false;
    uint32_t scap = 0;
    if (fast) {
        // strip capacity: 3.5x the expected density 2/(w+1) plus slack, never more than one per k-mer
        // (sized for reads of at most 1024 bp: in a batch that mixes short reads with a few contigs the long ones overflow
        // their strip and take the generic kernel, instead of every read paying for a contig-sized strip)
        const uint32_t cap_len = std::min<uint32_t>(d.max_len, 1024u);
        const uint32_t kmers = cap_len >= (uint32_t)ctx->P.k ? cap_len - (uint32_t)ctx->P.k + 1 : 0;
        scap = std::min<uint32_t>(std::max<uint32_t>(kmers, 1u), (uint32_t)(7.0 * kmers / (ctx->P.w + 1)) + 8u);
        CK(ctx->b_sx.ensure((size_t)n * scap * 8 + 8));
        CK(ctx->b_srp.ensure((size_t)n * scap * 4 + 4));
        CK(cudaMemsetAsync(ctx->b_cnt.as<uint32_t>() + n, 0, 4, s));
        launch_sketch_fast(d, ctx->P.w, ctx->P.k, scap, ctx->b_sx.as<uint64_t>(), ctx->b_srp.as<uint32_t>(), ctx->b_cnt.as<uint32_t>(),
                           ctx->b_ovf.as<uint32_t>(), ctr(ctx, C_SK_OVF), ctx->b_ovf2.as<uint32_t>(), ctr(ctx, C_SK_TIE),
                           R->b_ambcount.as<uint32_t>(), s, &ctx->lc);
    } else {
        CK(cudaMemsetAsync(ctx->b_cnt.p, 0, ((size_t)n + 1) * 4, s));
        launch_sketch_count(d, ctx->P.w, ctx->P.k, ctx->b_cnt.as<uint32_t>(), s, &ctx->lc);
    }
    // scanning n+1 counts (the last is 0) puts the total at mz_off[n] (and into the counter block)
    launch_exclusive_scan(ctx->b_cnt.as<uint32_t>(), ctx->b_mzoff.as<uint32_t>(), (uint64_t)n + 1, ctx->b_tmp.as<uint32_t>(),
                          reinterpret_cast<uint32_t *>(ctr(ctx, C_MZ)), s, &ctx->lc);
    if (fast) {
        launch_sketch_compact(d, scap, ctx->b_sx.as<uint64_t>(), ctx->b_srp.as<uint32_t>(), ctx->b_cnt.as<uint32_t>(),
                              ctx->b_mzoff.as<uint32_t>(), mz_x, ctx->b_mzrr.as<uint2>(), ctx->b_keys[0].as<uint64_t>(),
                              ctx->b_vals[0].as<uint32_t>(), 2 * ctx->P.k, (uint32_t)cap, s, &ctx->lc);
        *keys_ready = true;
        // reads with more minimizers than a strip holds: the generic kernel writes them in place (the list is usually empty)
        launch_sketch_write(d, ctx->P.w, ctx->P.k, ctx->b_mzoff.as<uint32_t>(), mz_x, ctx->b_mzrr.as<uint2>(), ctx->b_keys[0].as<uint64_t>(),
                            ctx->b_vals[0].as<uint32_t>(), 2 * ctx->P.k, ctx->b_ovf.as<uint32_t>(), ctr(ctx, C_SK_OVF), (uint32_t)cap, s, &ctx->lc);
    } else if (n) {
        launch_sketch_write(d, ctx->P.w, ctx->P.k, ctx->b_mzoff.as<uint32_t>(), mz_x, ctx->b_mzrr.as<uint2>(), nullptr, nullptr, 0, nullptr, nullptr,
                            (uint32_t)cap, s, &ctx->lc);
    }
    if (n) {
        mz_clear_kernel<<<64, 256, 0, s>>>(ctx->b_mzoff.as<uint32_t>(), n + 1, ctr(ctx, 0), cap);
        ctx->lc.total++;
    }
    CK(cudaGetLastError());
    return HLO_OK;
}

// hash sort + bucket index.  Leaves range[] (per minimizer, read order) and hits[] (hash order).
static int enqueue_index(hlo_ctx *ctx, const hlo_reads *R, bool keys_ready) {
    cudaStream_t s = ctx->stream;
    const DevReads &d = R->d;
    const uint64_t cap = ctx->cap_mz;
    const uint32_t *n_ptr = ctx->b_mzoff.as<uint32_t>() + d.n_reads;  // the minimizer count, on the device
    // key = cluster index << 2k | hash: 2k hash bits + the bits of the cluster count
    const int hash_bits = 2 * ctx->P.k;
    const int key_bits = hash_bits + (d.n_clusters > 1 ? bits_for(d.n_clusters) : 0);
    if (key_bits > 64) FAIL(HLO_ERANGE, "%u clusters with k=%d need a %d-bit sort key (limit 64): use smaller batches", d.n_clusters, ctx->P.k, key_bits);
    const bool segmented = ctx->sort_segmented != 0;
    const uint32_t n_seg = segmented ? d.n_clusters : 1u;
    CK(ctx->b_tmp.ensure(std::max(sort_segmented_tmp_words(cap, n_seg), scan_tmp_words(cap + 1)) * 4));
    CK(ctx->b_flags.ensure(cap * 4 + 8));
    CK(ctx->b_range.ensure(cap * 8 + 8));
    CK(ctx->b_hits.ensure(cap * 8 + 8));
    CK(ctx->b_midocc.ensure((size_t)d.n_clusters * 4 + 4));
    // occurrence histogram: a bucket can hold at most as many entries as its cluster has minimizers; sizes past the
    // clamp bin only matter when the cut-off quantile itself lands there (then the batch is refused)
    // (a tiny cluster of low-complexity reads can put every minimizer into one bucket: at least min(reads x length, 4096) bins)
    const uint32_t bins = next_pow2_u32((uint32_t)std::min<uint64_t>(
        65536ull, std::max<uint64_t>((uint64_t)d.max_cluster * 4 + 64, std::min<uint64_t>((uint64_t)d.max_cluster * d.max_len, 4096ull))));
    if (ctx->P.mid_occ <= 0) CK(ctx->b_hist.ensure(bucket_hist_words(d.n_clusters, bins) * 4));
    uint64_t *keys[2] = {ctx->b_keys[0].as<uint64_t>(), ctx->b_keys[1].as<uint64_t>()};
    uint32_t *vals[2] = {ctx->b_vals[0].as<uint32_t>(), ctx->b_vals[1].as<uint32_t>()};
    if (!keys_ready)
        launch_hash_keys(ctx->b_mzx.as<uint64_t>(), ctx->b_mzrr.as<uint2>(), d.cidx, hash_bits, cap, n_ptr, keys[0], vals[0], s, &ctx->lc);
    // (1) The minimizer array is cluster-major already, so every cluster is sorted as its own segment: the cluster bits
    //     of the key cost no pass.  (2) Grouping needs equal keys adjacent and stable, not a total order: only the top
    //     24 hash bits are sorted (3 passes instead of 6) and the rare groups in which two hashes share that prefix are
    //     repaired afterwards.  Option sort_segmented = 0: one segment, the cluster index is part of the sorted key.
    const uint32_t *seg_off = R->b_coff.as<uint32_t>();
    if (!segmented) {
        CK(ctx->b_seg1.ensure(8));
        CK(ctx->h_small.ensure(64));
        uint32_t *hs = ctx->h_small.as<uint32_t>();
        hs[0] = 0, hs[1] = d.n_reads;
        CK(cudaMemcpyAsync(ctx->b_seg1.p, hs, 8, cudaMemcpyHostToDevice, s));
        seg_off = ctx->b_seg1.as<uint32_t>();
    }
    const int top = segmented ? hash_bits : key_bits;  // sorted bits end here, and this many exist below
    int sort_bits = ctx->sort_hash_bits > 0 ? (segmented ? 0 : key_bits - hash_bits) + ctx->sort_hash_bits
                                            : (segmented ? 24 : ((key_bits - hash_bits + 19 + 7) / 8) * 8);
    const bool prefix = !ctx->force_full_sort && ctx->sort_prefix && sort_bits < top;
    ctx->prefix_used = prefix;
    if (!prefix) sort_bits = top;
    const int cur = launch_radix_sort_segmented(keys, vals, cap, top - sort_bits, top, ctx->b_mzoff.as<uint32_t>(), seg_off, n_seg,
                                                ctx->b_tmp.as<uint32_t>(), s, &ctx->lc);
    if (prefix)
        launch_collision_fix(keys[cur], vals[cur], cap, n_ptr, top - sort_bits, ctx->b_flags.as<uint32_t>(),
                             reinterpret_cast<uint32_t *>(ctr(ctx, C_COLL)), s, &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_SORT], s));
    // the hit array in the form its consumer wants: 32-bit words in the key layout of the shared-memory pair generator
    // whenever that one can run, (read, pos<<1|strand) pairs for the HBM-table kernel otherwise
    int dg_bits = 0;
    ctx->pg_dg_bits = pairgen_fast_supported(d, &dg_bits) && !ctx->pairgen_generic ? dg_bits : 0;
    launch_bucket_index(keys[cur], vals[cur], ctx->b_mzrr.as<uint2>(), cap, n_ptr, hash_bits, ctx->P.mid_occ, d.n_clusters,
                        ctx->b_midocc.as<uint32_t>(), ctx->b_hist.as<uint32_t>(), bins, ctr(ctx, 0), R->b_coff.as<uint32_t>(), ctx->pg_dg_bits,
                        ctx->b_hits.p, ctx->b_range.as<uint2>(), s, &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_INDEX], s));
    return HLO_OK;
}

constexpr int PAIRGEN_WIDE_SLOTS_LOG2 = 11;

static PairgenArgs pairgen_args(hlo_ctx *ctx, const hlo_reads *R, int external) {
    const DevReads &d = R->d;
    PairgenArgs a;
    a.reads = d;
    a.mz_off = ctx->b_mzoff.as<uint32_t>(), a.mz_rr = ctx->b_mzrr.as<uint2>();
    a.range = ctx->b_range.as<uint2>(), a.hits = ctx->b_hits.as<uint2>();
    a.k = ctx->P.k;
    a.gap_min = std::max(std::max(ctx->P.min_chain_score - ctx->P.k, ctx->P.min_seeds - 1), 0);
    a.external = external;
    a.hits32 = ctx->pg_dg_bits ? ctx->b_hits.as<uint32_t>() : nullptr;
    a.cands = ctx->b_cands.p, a.cand_cap = ctx->cap_c;
    a.cand_far = ctx->b_candfar.as<uint32_t>();
    a.cand_counter = ctr(ctx, C_CAND), a.hit_counter = ctr(ctx, C_HITS);
    a.ovf_counter = ctr(ctx, C_PG_OVF), a.work_counter = ctr(ctx, C_PG_WORK);
    a.overflow_reads = ctx->b_ovf.as<uint32_t>(), a.overflow_cap = d.n_reads;
    a.dg_bits = ctx->pg_dg_bits;
    return a;
}

// candidates into b_cands (CandI, or hlo_cand when external).  The shared-memory pair generator runs once for every
// read and then once per larger table size for the reads whose table overflowed — every pass takes its work list
// and its length from the device, so the later passes usually start, find nothing to do, and end.  What overflows
// even 2048 slots (and every batch whose keys do not fit 32 bits) is left to the HBM-table kernel, which needs
// host-side sizing and therefore runs from hlo_wait.
static int enqueue_pairgen(hlo_ctx *ctx, const hlo_reads *R, int external) {
    cudaStream_t s = ctx->stream;
    const DevReads &d = R->d;
    CK(ctx->b_ovf.ensure((size_t)d.n_reads * 4 + 4));
    CK(ctx->b_ovf2.ensure((size_t)d.n_reads * 4 + 4));
    CK(ctx->b_cands.ensure(ctx->cap_c * 16 + 16));
    CK(ctx->b_candfar.ensure(ctx->cap_c * 4 + 16));
    PairgenArgs a = pairgen_args(ctx, R, external);
    ctx->fast_pairgen = ctx->pg_dg_bits > 0 && d.n_reads > 0;
    ctx->n_pg_passes = 0;
    if (!ctx->fast_pairgen) return HLO_OK;
    const int slots0 = std::min(std::max(ctx->pairgen_slots_log2, ctx->pairgen_slots_auto), PAIRGEN_WIDE_SLOTS_LOG2);
    uint32_t *lists[2] = {ctx->b_ovf.as<uint32_t>(), ctx->b_ovf2.as<uint32_t>()};
    int p = 0;
    for (int sl = slots0; sl <= PAIRGEN_WIDE_SLOTS_LOG2 && p < HLO_PG_PASSES; ++sl, ++p) {
        PairgenArgs ap = a;
        ap.ovf_counter = ctr(ctx, C_PG_OVF + p), ap.work_counter = ctr(ctx, C_PG_WORK + p);
        ap.overflow_reads = lists[p & 1];
        if (p > 0) ap.work = lists[(p - 1) & 1], ap.n_work_ptr = ctr(ctx, C_PG_OVF + p - 1);
        launch_pairgen(ap, sl, s, &ctx->lc);
    }
    ctx->n_pg_passes = p;
    CK(cudaGetLastError());
    return HLO_OK;
}

// The HBM-table pair generator for the reads listed on the device (n_list of them; list == nullptr: every read).
// Host-sized tables: this is the slow, synchronous path of batches with extremely deep reads.
static int run_pairgen_generic(hlo_ctx *ctx, const hlo_reads *R, int external, const uint32_t *list, uint32_t n_list) {
    cudaStream_t s = ctx->stream;
    PairgenArgs a = pairgen_args(ctx, R, external);
    if (!n_list) return HLO_OK;
    CK(ctx->b_ovfcnt.ensure((size_t)n_list * 4 + 4));
    launch_generic_hit_counts(a, list, n_list, ctx->b_ovfcnt.as<uint32_t>(), s, &ctx->lc);
    std::vector<uint32_t> cnt(n_list);
    CK(cudaMemcpyAsync(cnt.data(), ctx->b_ovfcnt.p, (size_t)n_list * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    ctx->tm.n_syncs++;
    const uint64_t slot_budget = 1ull << 28;  // 4 GiB of tables per chunk
    uint32_t i0 = 0;
    std::vector<uint64_t> off;
    std::vector<uint32_t> xs;
    while (i0 < n_list) {
        off.assign(1, 0);
        uint32_t i1 = i0;
        while (i1 < n_list) {
            const uint64_t sl = next_pow2_u32(std::max(2u * cnt[i1], 64u));
            if (i1 > i0 && off.back() + sl > slot_budget) break;
            off.push_back(off.back() + sl);
            ++i1;
        }
        const uint64_t tot = off.back();
        CK(ctx->b_gkeys.ensure(tot * 8 + 8));
        CK(ctx->b_gmn.ensure(tot * 4 + 4));
        CK(ctx->b_gmx.ensure(tot * 4 + 4));
        CK(ctx->b_goff.ensure(off.size() * 8));
        CK(cudaMemcpyAsync(ctx->b_goff.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(ctx->b_gkeys.p, 0xff, tot * 8, s));
        CK(cudaMemsetAsync(ctx->b_gmn.p, 0xff, tot * 4, s));
        CK(cudaMemsetAsync(ctx->b_gmx.p, 0, tot * 4, s));
        // a chunk is addressed by offsetting the list (or, without a list, by a list of x indices)
        const uint32_t *chunk_list = list ? list + i0 : nullptr;
        if (!list && i0 > 0) {
            xs.resize(i1 - i0);
            for (uint32_t i = i0; i < i1; ++i) xs[i - i0] = i;
            CK(ctx->b_ovfcnt.ensure((size_t)xs.size() * 4 + 4));
            CK(cudaMemcpyAsync(ctx->b_ovfcnt.p, xs.data(), xs.size() * 4, cudaMemcpyHostToDevice, s));
            chunk_list = ctx->b_ovfcnt.as<uint32_t>();
        }
        launch_pairgen_generic(a, chunk_list, i1 - i0, ctx->b_gkeys.as<unsigned long long>(), ctx->b_gmn.as<uint32_t>(),
                               ctx->b_gmx.as<uint32_t>(), ctx->b_goff.as<uint64_t>(), s, &ctx->lc);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(s));
        ctx->tm.n_syncs++;
        i0 = i1;
    }
    return HLO_OK;
}

// verify the candidates in b_cands (their count: the device counter, clamped to n_cap) -> Rec16 in b_rec16
static int enqueue_verify(hlo_ctx *ctx, const hlo_reads *R, uint64_t n_cap) {
    cudaStream_t s = ctx->stream;
    CK(ctx->b_rec16.ensure(ctx->cap_a * sizeof(Rec16) + 16));
    CK(ctx->b_deferred.ensure(n_cap * 4 + 4));
    CK(cudaMemsetAsync(ctr(ctx, C_ACC), 0, 2 * sizeof(unsigned long long), s));  // C_ACC, C_DEF (a second verify after the generic pair generator starts over)
    VerifyArgs a;
    a.reads = R->d, a.tab = ctx->tab, a.prm = verify_params(ctx);
    a.cands = ctx->b_cands.as<CandI>(), a.cand_far = ctx->b_candfar.as<uint32_t>();
    a.n_ptr = ctr(ctx, C_CAND), a.n_cap = n_cap;
    a.out = ctx->b_rec16.as<Rec16>(), a.out_cap = ctx->cap_a;
    a.acc_counter = ctr(ctx, C_ACC), a.def_counter = ctr(ctx, C_DEF);
    a.deferred = ctx->b_deferred.as<uint32_t>();
    // the shared-memory threshold tables of the fast kernel hold 1023 entries: decided by THIS batch's longest read,
    // not by the longest read the ctx ever saw (the tables only grow)
    VerifyTables tab_batch = ctx->tab;
    tab_batch.max_len = std::max<int32_t>((int32_t)R->d.max_len, 1);
    const bool force_exact = ctx->verify_exact != 0 || !verify_fast_supported(R->d, tab_batch);
    ctx->exact_all = force_exact;
    if (force_exact) {
        launch_verify_exact(a, true, s, &ctx->lc);
    } else {
        VerifyArgs af = a;
        af.tab.max_len = tab_batch.max_len;  // the kernel stages max_len + 1 entries of each of the three tables
        CK(cudaEventRecord(ctx->ev_vf[0], s));
        launch_verify_fast(af, s, &ctx->lc);
        CK(cudaEventRecord(ctx->ev_vf[1], s));
        launch_verify_exact(a, false, s, &ctx->lc);  // the deferred candidates (a device-side count, usually a handful)
    }
    CK(cudaGetLastError());
    return HLO_OK;
}

// per-pair keep-longest + (idA, idB) order + expansion: group by idA, sort the small groups (records.cu)
static int enqueue_tail(hlo_ctx *ctx, const hlo_reads *R) {
    cudaStream_t s = ctx->stream;
    const DevReads &d = R->d;
    const uint32_t ng = d.n_reads;
    const uint64_t cap = ctx->cap_a;
    const unsigned long long *n_ptr = ctr(ctx, C_ACC);
    CK(ctx->b_bid.ensure(((size_t)ng + 2) * 4 * 4));  // records per idA | group offsets | distinct idB per idA | output offsets
    CK(ctx->b_slot.ensure(cap * 4 + 4));              // slot of a record inside its group, then "entries with a smaller idB"
    CK(ctx->b_ent.ensure(cap * 8 + 8));               // (rank of idB, record index) entries
    CK(ctx->b_entg.ensure(cap * 4 + 4));              // group of an entry
    CK(ctx->b_tmp.ensure(scan_tmp_words((uint64_t)ng + 1) * 4));
    CK(ctx->b_coff.ensure(((size_t)d.n_clusters + 2) * 8));
    const bool need_rec = ctx->rec_form == 0 || ctx->want_score;
    if (need_rec) CK(ctx->b_out.ensure(cap * sizeof(hlo_rec) + 32));
    if (ctx->rec_form != 0) CK(ctx->b_outc.ensure(cap * (ctx->rec_form == 1 ? sizeof(hlo_rec16) : sizeof(hlo_rec8)) + 32));
    const bool need_aux = ctx->aux_out || ctx->want_score;
    if (need_aux) CK(ctx->b_aux.ensure(cap * sizeof(hlo_rec_aux) + 32));
    uint32_t *g_cnt = ctx->b_bid.as<uint32_t>(), *g_off = g_cnt + (ng + 2), *g_uniq = g_off + (ng + 2), *g_uoff = g_uniq + (ng + 2);
    uint2 *g_ent = ctx->b_ent.as<uint2>();
    uint32_t *g_entg = ctx->b_entg.as<uint32_t>(), *g_slot = ctx->b_slot.as<uint32_t>();
    const Rec16 *rec = ctx->b_rec16.as<Rec16>();
    CK(cudaMemsetAsync(g_cnt, 0, ((size_t)ng + 2) * 4, s));
    CK(cudaMemsetAsync(g_uniq, 0, ((size_t)ng + 2) * 4, s));
    launch_group_count(rec, n_ptr, cap, d, g_cnt, g_slot, s, &ctx->lc);
    launch_exclusive_scan(g_cnt, g_off, (uint64_t)ng + 1, ctx->b_tmp.as<uint32_t>(), nullptr, s, &ctx->lc);
    launch_group_scatter(rec, n_ptr, cap, d, g_off, g_slot, g_ent, g_entg, s, &ctx->lc);
    uint32_t *g_lt = g_slot;  // (the slots are consumed by the scatter)
    CK(cudaMemsetAsync(ctr(ctx, C_BIG_GROUPS), 0, sizeof(unsigned long long), s));  // (the tail may run twice per batch)
    // (the per-group counts are consumed by the scan: their array takes the list of large groups)
    launch_group_mark(rec, n_ptr, cap, d, R->b_coff.as<uint32_t>(), g_off, g_entg, 1, g_ent, g_uniq, g_lt, g_cnt, ctr(ctx, C_BIG_GROUPS), s, &ctx->lc);
    launch_exclusive_scan(g_uniq, g_uoff, (uint64_t)ng + 1, ctx->b_tmp.as<uint32_t>(), reinterpret_cast<uint32_t *>(ctr(ctx, C_NOUT)), s, &ctx->lc);
    ExpandOut eo;
    if (need_rec) eo.out = ctx->b_out.as<hlo_rec>();
    if (ctx->rec_form == 1) eo.out16 = ctx->b_outc.as<hlo_rec16>();
    if (ctx->rec_form == 2) eo.out8 = ctx->b_outc.as<hlo_rec8>();
    if (need_aux) eo.aux = ctx->b_aux.as<hlo_rec_aux>();
    if (ctx->digest) {
        CK(cudaMemsetAsync(ctr(ctx, C_DIGEST), 0, sizeof(unsigned long long), s));  // (the tail runs twice after the HBM-table pair generator)
        eo.digest = ctr(ctx, C_DIGEST);
    }
    launch_group_expand(rec, n_ptr, cap, d, g_off, g_uoff, g_entg, g_ent, g_lt, eo, s, &ctx->lc);
    if (d.n_reads) launch_cluster_offsets(g_uoff, R->b_coff.as<uint32_t>(), d.n_clusters, ctx->b_coff.as<unsigned long long>(), s, &ctx->lc);
    else CK(cudaMemsetAsync(ctx->b_coff.p, 0, ((size_t)d.n_clusters + 1) * 8, s));
    CK(cudaGetLastError());
    return HLO_OK;
}

static int ensure_score_table(hlo_ctx *ctx, double mismatch_prob) {
    cudaStream_t s = ctx->stream;
    if (!ctx->prob_ready) {  // phred_to_prob (EdgeCalculator.cpp:59-63) for every byte value, the reference's own expression
        CK(ctx->h_small.ensure(128 * sizeof(double)));
        double *tab = ctx->h_small.as<double>();
        for (int b = 0; b < 128; ++b) tab[b] = pow(10, -(b - 33) / 10.0);
        CK(ctx->b_prob.ensure(128 * sizeof(double) + 2 * 128 * 128 * sizeof(double)));
        CK(cudaMemcpyAsync(ctx->b_prob.p, tab, 128 * sizeof(double), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));  // (once per ctx)
        ctx->prob_ready = true;
        ctx->prob_thr = -1.0;
    }
    if (ctx->prob_thr != mismatch_prob) {
        launch_score_table(ctx->b_prob.as<double>(), mismatch_prob, ctx->b_prob.as<double>() + 128, s, &ctx->lc);
        ctx->prob_thr = mismatch_prob;
    }
    return HLO_OK;
}

// the consumer's overlap score (score.cu) on the records the tail just left in HBM
static int enqueue_score(hlo_ctx *ctx, const hlo_reads *R) {
    cudaStream_t s = ctx->stream;
    const DevReads &d = R->d;
    int rc = ensure_score_table(ctx, ctx->score_mismatch_prob);
    if (rc) return rc;
    CK(ctx->b_scout.ensure(ctx->cap_a * sizeof(hlo_score) + 16));
    const uint8_t *dq = R->b_qual.as<uint8_t>() + 32;
    const uint8_t *qflag = dq + ((d.n_bases + 32 + 15) & ~(uint64_t)15);
    launch_score_records(d, dq, qflag, R->b_off.as<uint64_t>(), R->b_coff.as<uint32_t>(), ctx->b_out.as<hlo_rec>(),
                         reinterpret_cast<const uint32_t *>(ctx->b_aux.p), 2, reinterpret_cast<const uint32_t *>(ctr(ctx, C_NOUT)), ctx->cap_a,
                         ctx->b_prob.as<double>() + 128, ctx->score_min_read_len, ctx->b_scout.as<hlo_score>(), s, &ctx->lc);
    CK(cudaGetLastError());
    return HLO_OK;
}

// sketch -> index -> pair generation
static int enqueue_front(hlo_ctx *ctx, const hlo_reads *R, int external) {
    cudaStream_t s = ctx->stream;
    bool keys_ready = false;
    int rc = enqueue_sketch(ctx, R, false, &keys_ready);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[ST_SKETCH], s));
    rc = enqueue_index(ctx, R, keys_ready);
    if (rc) return rc;
    rc = enqueue_pairgen(ctx, R, external);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[ST_PAIRGEN], s));
    return HLO_OK;
}

// verify -> tail -> score -> counter block to the host page
static int enqueue_back(hlo_ctx *ctx, const hlo_reads *R) {
    cudaStream_t s = ctx->stream;
    int rc = enqueue_verify(ctx, R, ctx->cap_c);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[ST_VERIFY], s));
    rc = enqueue_tail(ctx, R);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[ST_DEDUP], s));
    if (ctx->want_score) {
        rc = enqueue_score(ctx, R);
        if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->ev[ST_SCORE], s));
    return enqueue_counter_readback(ctx);
}

static void reset_counters_for_rerun(hlo_ctx *ctx) {
    cudaMemsetAsync(ctr(ctx, 0), 0, sizeof(unsigned long long) * C_COUNT, ctx->stream);
    cudaMemsetAsync(ctr(ctx, C_MIDOCC_LO), 0xff, sizeof(unsigned long long), ctx->stream);
}

// After a wait: did an estimate-sized buffer overflow?  Updates the learned needs; returns true when the front must
// be enqueued again.
static bool front_needs_rerun(hlo_ctx *ctx, const hlo_reads *R) {
    const unsigned long long *hc = ctx->hc;
    const uint64_t mz = (uint32_t)hc[C_MZ];
    bool rerun = false;
    if (R->d.n_bases) ctx->need_mz_per_base = std::max(ctx->need_mz_per_base, (double)mz / (double)R->d.n_bases);
    if (mz > ctx->cap_mz) {
        ctx->min_mz = mz + mz / 16 + 1024;
        return true;  // (nothing behind the sketch ran on real data)
    }
    if (ctx->prefix_used && (hc[C_COLL] >> 32)) {  // rare: heavily repeated minimizers share a hash prefix — sort on all bits instead
        ctx->force_full_sort = true;
        ctx->tm.sort_fallback = 1.0f;
        return true;
    }
    if (mz) ctx->need_c_per_mz = std::max(ctx->need_c_per_mz, (double)hc[C_CAND] / (double)mz);
    if (hc[C_CAND] > ctx->cap_c) {
        ctx->min_c = hc[C_CAND] + hc[C_CAND] / 16 + 1024;
        rerun = true;
    }
    return rerun;
}

static void start_compute_section(hlo_ctx *ctx) {
    if (!ctx->serialize_compute) return;
    std::lock_guard<std::mutex> g(g_order_mutex[ctx->device & 63]);
    cudaEvent_t prev = g_last_compute[ctx->device & 63];
    if (prev && prev != ctx->ev_compute_done) cudaStreamWaitEvent(ctx->stream, prev, 0);
}
static void end_compute_section(hlo_ctx *ctx) {
    cudaEventRecord(ctx->ev_compute_done, ctx->stream);
    if (!ctx->serialize_compute) return;
    std::lock_guard<std::mutex> g(g_order_mutex[ctx->device & 63]);
    g_last_compute[ctx->device & 63] = ctx->ev_compute_done;
}

static void begin_batch(hlo_ctx *ctx, const hlo_reads *R) {
    ctx->min_mz = ctx->min_c = ctx->min_a = 0;
    ctx->force_full_sort = false;
    // 8-byte records need every read name below 2^16; else the 16-byte form
    ctx->rec_form = ctx->compact == 2 ? (R->max_id < 65536u ? 2 : 1) : (ctx->compact ? 1 : 0);
    ctx->want_score = ctx->scores && R->has_quals && R->d.plain_names;
    plan_caps(ctx, R);
}

static void finish_timing(hlo_ctx *ctx, const hlo_reads *R, int last_stage) {
    auto el = [&](int a, int b) {
        float v = 0.f;
        if (cudaEventElapsedTime(&v, ctx->ev[a], ctx->ev[b]) != cudaSuccess) {
            cudaGetLastError();
            v = 0.f;
        }
        return v;
    };
    hlo_timing &t = ctx->tm;
    t.h2d_ms = R->h2d_ms, t.pack_ms = R->pack_ms;
    if (last_stage >= ST_SKETCH) t.sketch_ms = el(ST_PACK, ST_SKETCH);
    if (last_stage >= ST_SORT) t.sort_ms = el(ST_SKETCH, ST_SORT);
    if (last_stage >= ST_INDEX) t.index_ms = el(ST_SORT, ST_INDEX);
    if (last_stage >= ST_PAIRGEN) t.pairgen_ms = el(ST_INDEX, ST_PAIRGEN);
    if (last_stage >= ST_VERIFY) t.verify_ms = el(ST_PAIRGEN, ST_VERIFY);
    if (last_stage >= ST_DEDUP) t.dedup_ms = el(ST_VERIFY, ST_DEDUP);
    if (last_stage >= ST_SCORE) t.score_ms = el(ST_DEDUP, ST_SCORE);
    if (last_stage >= ST_D2H) t.d2h_ms = el(ST_SCORE, ST_D2H);
    t.total_ms = t.h2d_ms + t.pack_ms + t.sketch_ms + t.sort_ms + t.index_ms + t.pairgen_ms + t.verify_ms + t.dedup_ms + t.score_ms + t.d2h_ms;
    t.n_reads = R->d.n_reads, t.n_bases = R->d.n_bases, t.n_minimizers = ctx->n_mz;
    t.total_launches = ctx->lc.total, t.verify_launches = ctx->lc.verify;
    if (last_stage >= ST_VERIFY && ctx->lc.verify) {
        float v = 0.f;
        if (cudaEventElapsedTime(&v, ctx->ev_vf[0], ctx->ev_vf[1]) == cudaSuccess) t.verify_fast_ms = v;
        else cudaGetLastError();
    }
}

static void reset_timing(hlo_ctx *ctx) {
    memset(&ctx->tm, 0, sizeof ctx->tm);
    ctx->lc = LaunchCounter();
}

// Runs the front (sketch .. pair generation) to completion, rerunning when a buffer was too small and handing
// overflowing reads to the HBM-table kernel.  On return hc[] holds the counters and b_cands every candidate.
// with_back: also enqueue verify + tail behind the front, so that the usual batch needs ONE wait.
static int complete_front(hlo_ctx *ctx, const hlo_reads *R, int external, bool with_back, bool already_enqueued) {
    cudaStream_t s = ctx->stream;
    for (int attempt = 0; attempt < 6; ++attempt) {
        if (!(already_enqueued && attempt == 0)) {
            if (attempt > 0) {
                ctx->tm.n_reruns++;
                plan_caps(ctx, R);
                reset_counters_for_rerun(ctx);
                CK(cudaEventRecord(ctx->ev[ST_PACK], s));
            }
            int rc = enqueue_front(ctx, R, external);
            if (rc) return rc;
            if (with_back) rc = enqueue_back(ctx, R);
            else rc = enqueue_counter_readback(ctx);
            if (rc) return rc;
            end_compute_section(ctx);
        }
        int rc = sync_counters(ctx);
        if (rc) return rc;
        const unsigned long long *hc = ctx->hc;
        if (hc[C_FLAGS] & HLO_FLAG_MIDOCC_AMBIGUOUS)
            FAIL(HLO_EUNSUPPORTED, "a cluster's minimizer occurrences are too extreme for the automatic occurrence cut-off "
                                   "(the quantile fell into the histogram's clamp bin): set mid_occ explicitly");
        if (front_needs_rerun(ctx, R)) continue;
        // reads the shared-memory tables could not hold (or all of them, when the fast kernel does not apply)
        uint32_t n_list = R->d.n_reads;
        const uint32_t *list = nullptr;
        uint64_t first_overflow = 0;
        if (ctx->fast_pairgen) {
            const int last = ctx->n_pg_passes - 1;
            n_list = (uint32_t)std::min<unsigned long long>(hc[C_PG_OVF + last], R->d.n_reads);
            list = (last & 1) ? ctx->b_ovf2.as<uint32_t>() : ctx->b_ovf.as<uint32_t>();
            first_overflow = hc[C_PG_OVF];
            // a batch that overflowed a lot: the following ones start with larger tables
            const int slots0 = std::min(std::max(ctx->pairgen_slots_log2, ctx->pairgen_slots_auto), PAIRGEN_WIDE_SLOTS_LOG2);
            // (only overflow drives this: larger tables cost occupancy — config 2 runs 1.7x slower at 512 slots — and a
            // crowded but not overflowing table measured the same at both sizes, BASELINE configs[3]: 13.4 ms)
            if (first_overflow > R->d.n_reads / 16 && ctx->pairgen_slots_auto < 10) ctx->pairgen_slots_auto = std::min(slots0 + 1, 10);
        }
        ctx->tm.n_overflow_reads = ctx->fast_pairgen ? first_overflow : 0;
        if (n_list) {
            rc = run_pairgen_generic(ctx, R, external, list, n_list);
            if (rc) return rc;
            rc = enqueue_counter_readback(ctx);
            if (rc) return rc;
            rc = sync_counters(ctx);
            if (rc) return rc;
            if (front_needs_rerun(ctx, R)) continue;
            if (with_back) {  // the records behind the first verify missed the candidates just added: once more
                rc = enqueue_back(ctx, R);
                if (rc) return rc;
                end_compute_section(ctx);
                rc = sync_counters(ctx);
                if (rc) return rc;
            }
        }
        ctx->n_mz = (uint32_t)ctx->hc[C_MZ];
        ctx->tm.n_sketch_reruns = (uint32_t)ctx->hc[C_SK_TIE];
        ctx->tm.n_seed_hits = ctx->hc[C_HITS];
        ctx->tm.n_candidates = ctx->hc[C_CAND];
        ctx->tm.mid_occ_lo = ctx->hc[C_MIDOCC_LO] > 0xffffffffull ? 0u : (uint32_t)ctx->hc[C_MIDOCC_LO];
        ctx->tm.mid_occ_hi = (uint32_t)ctx->hc[C_MIDOCC_HI];
        return HLO_OK;
    }
    FAIL(HLO_ECUDA, "buffer sizing did not converge");
}

// the whole chain on packed reads; out may be nullptr with option d2h = 0
static int run_batch(hlo_ctx *ctx, const hlo_reads *R, hlo_recs *out, bool already_enqueued) {
    cudaStream_t s = ctx->stream;
    const DevReads &d = R->d;
    bool enqueued = already_enqueued;
    for (int attempt = 0; attempt < 6; ++attempt) {
        int rc = complete_front(ctx, R, 0, true, enqueued);
        enqueued = false;
        if (rc) return rc;
        const unsigned long long *hc = ctx->hc;
        const uint64_t mz = ctx->n_mz;
        if (mz) ctx->need_a_per_mz = std::max(ctx->need_a_per_mz, (double)hc[C_ACC] / (double)mz);
        if (hc[C_ACC] > ctx->cap_a) {  // more records than the estimate: grow, and the next pass enqueues front and back again
            if (hc[C_ACC] >= 0x7ffffff0ull) FAIL(HLO_ERANGE, "more than 2^31 accepted overlaps in one batch: use smaller batches");
            ctx->min_a = hc[C_ACC] + hc[C_ACC] / 16 + 1024;
            ctx->min_c = std::max(ctx->min_c, ctx->min_a);
            ctx->tm.n_reruns++;
            plan_caps(ctx, R);
            reset_counters_for_rerun(ctx);
            CK(cudaEventRecord(ctx->ev[ST_PACK], s));
            continue;
        }
        ctx->tm.n_accepted = hc[C_ACC];
        ctx->tm.n_deferred = ctx->exact_all ? hc[C_CAND] : hc[C_DEF];
        ctx->tm.digest = hc[C_DIGEST];
        const uint64_t n_out = (uint32_t)hc[C_NOUT];
        ctx->tm.n_records = n_out;
        if (ctx->d2h && out) {
            const size_t rb = ctx->rec_form == 0 ? sizeof(hlo_rec) : ctx->rec_form == 1 ? sizeof(hlo_rec16) : sizeof(hlo_rec8);
            CK(ctx->h_out.ensure(n_out * rb + 32));
            CK(ctx->h_coff.ensure(((size_t)d.n_clusters + 2) * 8));
            CK(cudaMemcpyAsync(ctx->h_out.p, ctx->rec_form == 0 ? ctx->b_out.p : ctx->b_outc.p, n_out * rb, cudaMemcpyDeviceToHost, s));
            if (ctx->aux_out) {
                CK(ctx->h_aux.ensure(n_out * sizeof(hlo_rec_aux) + 32));
                CK(cudaMemcpyAsync(ctx->h_aux.p, ctx->b_aux.p, n_out * sizeof(hlo_rec_aux), cudaMemcpyDeviceToHost, s));
            }
            if (ctx->want_score) {
                CK(ctx->h_score.ensure(n_out * sizeof(hlo_score) + 32));
                CK(cudaMemcpyAsync(ctx->h_score.p, ctx->b_scout.p, n_out * sizeof(hlo_score), cudaMemcpyDeviceToHost, s));
            }
            CK(cudaMemcpyAsync(ctx->h_coff.p, ctx->b_coff.p, ((size_t)d.n_clusters + 1) * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaEventRecord(ctx->ev[ST_D2H], s));
            CK(cudaStreamSynchronize(s));
            ctx->tm.n_syncs++;
        } else {
            CK(cudaEventRecord(ctx->ev[ST_D2H], s));
            CK(cudaEventSynchronize(ctx->ev[ST_D2H]));
        }
        if (out) {
            memset(out, 0, sizeof *out);
            if (ctx->d2h) {
                if (ctx->rec_form == 0) out->rec = ctx->h_out.as<hlo_rec>();
                else if (ctx->rec_form == 1) out->rec16 = ctx->h_out.as<hlo_rec16>();
                else out->rec8 = ctx->h_out.as<hlo_rec8>();
                out->aux = ctx->aux_out ? ctx->h_aux.as<hlo_rec_aux>() : nullptr;
                out->score = ctx->want_score ? ctx->h_score.as<hlo_score>() : nullptr;
                out->cluster_off = ctx->h_coff.as<uint64_t>();
            }
            out->n = n_out, out->n_clusters = d.n_clusters;
        }
        finish_timing(ctx, R, ST_D2H);
        return HLO_OK;
    }
    FAIL(HLO_ECUDA, "buffer sizing did not converge");
}

extern "C" {

int hlo_sketch(hlo_ctx *ctx, const hlo_reads *R, const hlo_mz **mz, const uint64_t **read_off, uint64_t *n) {
    if (!ctx || !R || !mz || !read_off || !n) return HLO_EINVAL;
    if (ctx->inflight) FAIL(HLO_EINVAL, "a submitted batch is in flight on this ctx: call hlo_wait first");
    CK(cudaSetDevice(ctx->device));
    reset_timing(ctx);
    cudaStream_t s = ctx->stream;
    begin_batch(ctx, R);
    uint64_t tot = 0;
    for (int attempt = 0;; ++attempt) {
        if (attempt > 4) FAIL(HLO_ECUDA, "buffer sizing did not converge");
        if (attempt) plan_caps(ctx, R), ctx->tm.n_reruns++;
        reset_counters_for_rerun(ctx);
        CK(cudaEventRecord(ctx->ev[ST_PACK], s));
        bool keys_ready = false;
        int rc = enqueue_sketch(ctx, R, true, &keys_ready);
        if (rc) return rc;
        rc = enqueue_counter_readback(ctx);
        if (rc) return rc;
        rc = sync_counters(ctx);
        if (rc) return rc;
        tot = (uint32_t)ctx->hc[C_MZ];
        if (tot <= ctx->cap_mz) break;
        ctx->min_mz = tot + 1024;
    }
    ctx->n_mz = tot;
    CK(ctx->b_mz.ensure(tot * sizeof(hlo_mz) + 16));
    launch_mz_assemble(ctx->b_mzx.as<uint64_t>(), ctx->b_mzrr.as<uint2>(), tot, ctx->b_mz.as<hlo_mz>(), s,
                       &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_SKETCH], s));
    // the device arrays are in name order (sketch.cu); hand them back in the caller's read order
    int rc = ensure_host_meta(ctx, R);
    if (rc) return rc;
    CK(ctx->h_mz.ensure(tot * sizeof(hlo_mz) + 16));
    CK(ctx->h_roff.ensure(((size_t)R->d.n_reads + 1) * 8));
    std::vector<uint32_t> off32((size_t)R->d.n_reads + 1);
    std::vector<hlo_mz> tmp((size_t)tot);
    CK(cudaMemcpyAsync(tmp.data(), ctx->b_mz.p, tot * sizeof(hlo_mz), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(off32.data(), ctx->b_mzoff.p, off32.size() * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    uint64_t *ro = ctx->h_roff.as<uint64_t>();
    hlo_mz *hm = ctx->h_mz.as<hlo_mz>();
    {
        const uint32_t nr = R->d.n_reads;
        std::vector<uint32_t> xof(nr);  // read -> x
        for (uint32_t x = 0; x < nr; ++x) xof[R->h_perm[x]] = x;
        uint64_t w = 0;
        for (uint32_t r = 0; r < nr; ++r) {
            ro[r] = w;
            const uint32_t x = xof[r];
            for (uint32_t m = off32[x]; m < off32[x + 1]; ++m) hm[w++] = tmp[m];
        }
        ro[nr] = w;
    }
    *mz = ctx->h_mz.as<hlo_mz>(), *read_off = ro, *n = tot;
    finish_timing(ctx, R, ST_SKETCH);
    return HLO_OK;
}

int hlo_candidates(hlo_ctx *ctx, const hlo_reads *R, const hlo_cand **cands, uint64_t *n) {
    if (!ctx || !R || !cands || !n) return HLO_EINVAL;
    if (ctx->inflight) FAIL(HLO_EINVAL, "a submitted batch is in flight on this ctx: call hlo_wait first");
    CK(cudaSetDevice(ctx->device));
    reset_timing(ctx);
    begin_batch(ctx, R);
    reset_counters_for_rerun(ctx);
    CK(cudaEventRecord(ctx->ev[ST_PACK], ctx->stream));
    int rc = complete_front(ctx, R, 1, false, false);
    if (rc) return rc;
    const uint64_t n_c = ctx->hc[C_CAND];
    CK(ctx->h_cands.ensure(n_c * sizeof(hlo_cand) + 16));
    CK(cudaMemcpyAsync(ctx->h_cands.p, ctx->b_cands.p, n_c * sizeof(hlo_cand), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    hlo_cand *h = ctx->h_cands.as<hlo_cand>();
    std::sort(h, h + n_c, [](const hlo_cand &x, const hlo_cand &y) {
        if (x.q != y.q) return x.q < y.q;
        if (x.t != y.t) return x.t < y.t;
        const uint32_t sx = x.strand_anchor & 1u, sy = y.strand_anchor & 1u;
        if (sx != sy) return sx < sy;
        return x.diag < y.diag;
    });
    *cands = h, *n = n_c;
    finish_timing(ctx, R, ST_PAIRGEN);
    return HLO_OK;
}

int hlo_verify(hlo_ctx *ctx, const hlo_reads *R, const hlo_cand *cands, uint64_t n, hlo_recs *out) {
    if (!ctx || !R || (!cands && n) || !out) return HLO_EINVAL;
    if (ctx->inflight) FAIL(HLO_EINVAL, "a submitted batch is in flight on this ctx: call hlo_wait first");
    CK(cudaSetDevice(ctx->device));
    reset_timing(ctx);
    cudaStream_t s = ctx->stream;
    {
        const int rcm = ensure_host_meta(ctx, R);
        if (rcm) return rcm;
    }
    // A candidate is trusted by the kernels (its diagonal and anchors are packed next to the read lengths and steer row
    // reads): everything is checked here
    for (uint64_t i = 0; i < n; ++i) {
        const hlo_cand &c = cands[i];
        if (c.q >= R->d.n_reads || c.t >= R->d.n_reads) FAIL(HLO_EINVAL, "candidate %llu names a read outside the batch", (unsigned long long)i);
        const int ql = (int)R->h_len[c.q], tl = (int)R->h_len[c.t];
        if (!(c.diag > -ql && c.diag < tl) && !(ql == 0 || tl == 0))
            FAIL(HLO_EINVAL, "candidate %llu: diagonal %d leaves no overlap between reads of %d and %d bases", (unsigned long long)i, c.diag, ql, tl);
        if ((ql == 0 || tl == 0) && (c.diag <= -32768 || c.diag >= 32768)) FAIL(HLO_EINVAL, "candidate %llu: diagonal out of range", (unsigned long long)i);
        const int a_start = (int)((c.strand_anchor >> 1) & 0x7fffu), a_end = (int)(c.strand_anchor >> 16);
        if (a_end != 0 && (a_start > a_end || a_end > ql))
            FAIL(HLO_EINVAL, "candidate %llu: anchor span [%d, %d) outside the query (%d bases)", (unsigned long long)i, a_start, a_end, ql);
    }
    ctx->rec_form = 0, ctx->want_score = false;
    ctx->cap_c = n, ctx->cap_a = n;
    CK(ctx->b_candx.ensure(n * sizeof(hlo_cand) + 16));
    CK(ctx->b_cands.ensure(n * sizeof(CandI) + 16));
    CK(ctx->b_candfar.ensure(n * 4 + 16));
    CK(ctx->h_small.ensure(64));
    CK(cudaMemsetAsync(ctr(ctx, 1), 0, sizeof(unsigned long long) * (C_COUNT - 1), s));
    *ctx->h_small.as<unsigned long long>() = n;
    CK(cudaMemcpyAsync(ctr(ctx, C_CAND), ctx->h_small.p, 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->b_candx.p, cands, n * sizeof(hlo_cand), cudaMemcpyHostToDevice, s));
    launch_cands_to_internal(ctx->b_candx.as<hlo_cand>(), n, R->d, ctx->b_cands.as<CandI>(), ctx->b_candfar.as<uint32_t>(), s, &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_PAIRGEN], s));
    int rc = enqueue_verify(ctx, R, n);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[ST_VERIFY], s));
    rc = enqueue_counter_readback(ctx);
    if (rc) return rc;
    rc = sync_counters(ctx);
    if (rc) return rc;
    const uint64_t n_acc = ctx->hc[C_ACC];
    CK(ctx->h_rec16.ensure(n_acc * sizeof(Rec16) + 16));
    CK(cudaMemcpyAsync(ctx->h_rec16.p, ctx->b_rec16.p, n_acc * sizeof(Rec16), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    Rec16 *h = ctx->h_rec16.as<Rec16>();
    std::sort(h, h + n_acc, [](const Rec16 &x, const Rec16 &y) {
        if (x.q != y.q) return x.q < y.q;
        if (x.t != y.t) return x.t < y.t;
        if ((x.ds & 1u) != (y.ds & 1u)) return (x.ds & 1u) < (y.ds & 1u);
        if (x.ds != y.ds) return x.ds < y.ds;
        return x.kb < y.kb;
    });
    CK(ctx->h_out.ensure(n_acc * sizeof(hlo_rec) + 16));
    CK(ctx->h_aux.ensure(n_acc * sizeof(hlo_rec_aux) + 16));
    hlo_rec *ro = ctx->h_out.as<hlo_rec>();
    hlo_rec_aux *ao = ctx->h_aux.as<hlo_rec_aux>();
    uint64_t n_keep = 0;
    for (uint64_t i = 0; i < n_acc; ++i) {
        const Rec16 &x = h[i];
        if (x.kb & HLO_KB_OVERHANG_FAIL) continue;  // (pair_rule 1 keeps such rows for the pair decision; a verified record they are not)
        expand_record(R->h_ids[x.q], R->h_ids[x.t], (int)R->h_len[x.q], (int)R->h_len[x.t], (int)(x.ds >> 1) - HLO_DIAG_BIAS,
                      (int)(x.ds & 1u), x.kb & HLO_KB_K_MASK, &ro[n_keep]);
        ao[n_keep].cluster = R->h_cidx[x.q], ao[n_keep].blocklen = x.kb >> 16;
        ++n_keep;
    }
    memset(out, 0, sizeof *out);
    out->rec = ro, out->aux = ao, out->n = n_keep, out->n_clusters = R->d.n_clusters;
    ctx->tm.n_candidates = n, ctx->tm.n_accepted = n_keep, ctx->tm.n_records = n_keep;
    ctx->tm.n_deferred = ctx->exact_all ? n : ctx->hc[C_DEF];
    {
        float v = 0.f;
        if (cudaEventElapsedTime(&v, ctx->ev[ST_PAIRGEN], ctx->ev[ST_VERIFY]) == cudaSuccess) ctx->tm.verify_ms = v;
        else cudaGetLastError();
        if (ctx->lc.verify && cudaEventElapsedTime(&v, ctx->ev_vf[0], ctx->ev_vf[1]) == cudaSuccess) ctx->tm.verify_fast_ms = v;
        else cudaGetLastError();
        ctx->tm.total_ms = ctx->tm.verify_ms;
        ctx->tm.n_reads = R->d.n_reads, ctx->tm.n_bases = R->d.n_bases;
        ctx->tm.total_launches = ctx->lc.total, ctx->tm.verify_launches = ctx->lc.verify;
    }
    return HLO_OK;
}

int hlo_overlap_batch(hlo_ctx *ctx, const hlo_reads *R, hlo_recs *out) {
    if (!ctx || !R || !out) return HLO_EINVAL;
    if (ctx->inflight) FAIL(HLO_EINVAL, "a submitted batch is in flight on this ctx: call hlo_wait first");
    CK(cudaSetDevice(ctx->device));
    reset_timing(ctx);
    cudaStream_t s = ctx->stream;
    const DevReads &d = R->d;
    begin_batch(ctx, R);
    start_compute_section(ctx);
    reset_counters_for_rerun(ctx);
    bool repacked = false;
    if (ctx->repack && d.n_reads && !R->packed_input) {
        // the packer again, from the ASCII bases still resident in HBM (results are identical; the
        // ambiguity flags it ORs into len[] are idempotent)
        CK(cudaMemsetAsync(R->b_ambcount.p, 0, 16, s));
        CK(cudaEventRecord(ctx->ev[ST_H2D], s));
        launch_pack(R->b_ascii.as<uint8_t>(), R->b_off.as<uint64_t>(), d, R->b_ambcount.as<uint32_t>(), s, &ctx->lc);
        CK(cudaGetLastError());
        repacked = true;
    }
    CK(cudaEventRecord(ctx->ev[ST_PACK], s));
    int rc = run_batch(ctx, R, out, false);
    if (rc) return rc;
    if (repacked) {  // this call ran the packer itself: report its time instead of the upload's
        float v = 0.f;
        if (cudaEventElapsedTime(&v, ctx->ev[ST_H2D], ctx->ev[ST_PACK]) != cudaSuccess) cudaGetLastError(), v = 0.f;
        ctx->tm.total_ms += v - ctx->tm.pack_ms - ctx->tm.h2d_ms;
        ctx->tm.pack_ms = v, ctx->tm.h2d_ms = 0.f;
    }
    return HLO_OK;
}

int hlo_submit(hlo_ctx *ctx, const hlo_batch *batch) {
    if (!ctx || !batch) return HLO_EINVAL;
    if (ctx->inflight) FAIL(HLO_EINVAL, "a submitted batch is already in flight on this ctx: call hlo_wait first");
    CK(cudaSetDevice(ctx->device));
    reset_timing(ctx);
    hlo_reads *R = &ctx->scratch_reads;
    if (ctx->scores && !batch->quals) FAIL(HLO_EINVAL, "option scores is set but the batch carries no qualities");
    int rc = pack_into(ctx, R, *batch);
    if (rc) return rc;
    if (ctx->scores && !R->d.plain_names) FAIL(HLO_EUNSUPPORTED, "the fused score needs read names 0..N-1 in file order");
    begin_batch(ctx, R);
    start_compute_section(ctx);  // (the uploads above are already enqueued: they overlap the previous batch's kernels)
    rc = enqueue_front(ctx, R, 0);
    if (!rc) rc = enqueue_back(ctx, R);
    if (rc) return rc;
    end_compute_section(ctx);
    ctx->inflight = true;
    return HLO_OK;
}

int hlo_wait(hlo_ctx *ctx, hlo_recs *out) {
    if (!ctx || !out) return HLO_EINVAL;
    if (!ctx->inflight) FAIL(HLO_EINVAL, "hlo_wait without hlo_submit");
    CK(cudaSetDevice(ctx->device));
    ctx->inflight = false;
    hlo_reads *R = &ctx->scratch_reads;
    int rc = run_batch(ctx, R, out, true);
    if (rc) return rc;
    cudaEventElapsedTime(&R->h2d_ms, ctx->ev[ST_BEGIN], ctx->ev[ST_H2D]);
    cudaEventElapsedTime(&R->pack_ms, ctx->ev[ST_H2D], ctx->ev[ST_COUNT]);
    cudaGetLastError();
    ctx->tm.total_ms += R->h2d_ms + R->pack_ms - ctx->tm.h2d_ms - ctx->tm.pack_ms;
    ctx->tm.h2d_ms = R->h2d_ms, ctx->tm.pack_ms = R->pack_ms;
    return HLO_OK;
}

int hlo_overlap_host(hlo_ctx *ctx, const char *bases, const uint64_t *base_off, uint32_t n_reads, const uint32_t *cluster_off,
                     uint32_t n_clusters, hlo_recs *out) {
    if (!ctx || !out) return HLO_EINVAL;
    hlo_batch B{};
    B.bases = bases, B.base_off = base_off, B.n_reads = n_reads, B.cluster_off = cluster_off, B.n_clusters = n_clusters;
    int rc = hlo_submit(ctx, &B);
    if (rc) return rc;
    return hlo_wait(ctx, out);
}

int hlo_set_score_params(hlo_ctx *ctx, double mismatch_prob, uint32_t min_read_len) {
    if (!ctx) return HLO_EINVAL;
    ctx->score_mismatch_prob = mismatch_prob, ctx->score_min_read_len = min_read_len;
    return HLO_OK;
}

int hlo_set_option(hlo_ctx *ctx, const char *name, int64_t value) {
    if (!ctx || !name) return HLO_EINVAL;
    struct Opt {
        const char *name;
        int *field;
        int64_t lo, hi;
    };
    const Opt opts[] = {{"verify_exact", &ctx->verify_exact, 0, 1},       {"sketch_generic", &ctx->sketch_generic, 0, 1},
                        {"sort_segmented", &ctx->sort_segmented, 0, 1},   {"sort_prefix", &ctx->sort_prefix, 0, 1},
                        {"sort_hash_bits", &ctx->sort_hash_bits, 0, 64},  {"pairgen_generic", &ctx->pairgen_generic, 0, 1},
                        {"repack", &ctx->repack, 0, 1},                   {"compact", &ctx->compact, 0, 2},
                        {"aux", &ctx->aux_out, 0, 1},                     {"d2h", &ctx->d2h, 0, 1},
                        {"scores", &ctx->scores, 0, 1},                   {"digest", &ctx->digest, 0, 1},
                        {"cap_scale_pct", &ctx->cap_scale_pct, 1, 100000}, {"serialize_compute", &ctx->serialize_compute, 0, 1}};
    for (const Opt &o : opts)
        if (!strcmp(name, o.name)) {
            if (value < o.lo || value > o.hi) FAIL(HLO_ERANGE, "%s=%lld outside [%lld,%lld]", name, (long long)value, (long long)o.lo, (long long)o.hi);
            *o.field = (int)value;
            return HLO_OK;
        }
    if (!strcmp(name, "pairgen_slots_log2")) {
        if (value < 5 || value > 10) FAIL(HLO_ERANGE, "pairgen_slots_log2=%lld outside [5,10]", (long long)value);
        ctx->pairgen_slots_log2 = (int)value;
        ctx->pairgen_slots_auto = 0;
        return HLO_OK;
    }
    if (!strcmp(name, "forget_sizes")) {  // drop what earlier batches taught the buffer estimates
        ctx->need_mz_per_base = ctx->need_c_per_mz = ctx->need_a_per_mz = 0.0;
        return HLO_OK;
    }
    FAIL(HLO_EINVAL, "unknown option %s", name);
}

int hlo_score_records(hlo_ctx *ctx, const hlo_reads *R, const char *quals, const hlo_rec *rec, const uint32_t *rec_cluster, uint64_t n,
                      double mismatch_prob, uint32_t min_read_len, const hlo_score **out) {
    if (!ctx || !R || !quals || (!rec && n) || !out) return HLO_EINVAL;
    if (ctx->inflight) FAIL(HLO_EINVAL, "a submitted batch is in flight on this ctx: call hlo_wait first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const DevReads &d = R->d;
    if (d.n_clusters > 1 && !rec_cluster) FAIL(HLO_EINVAL, "hlo_score_records: rec_cluster is required for a multi-cluster batch");
    {   // records name reads by their decimal names: only names 0..N-1 in file order are resolved here
        const int rcm = ensure_host_meta(ctx, R);
        if (rcm) return rcm;
        for (uint32_t c = 0; c < d.n_clusters; ++c)
            for (uint32_t i = R->h_cluster_off[c]; i < R->h_cluster_off[c + 1]; ++i)
                if (R->h_ids[i] != i - R->h_cluster_off[c]) FAIL(HLO_EUNSUPPORTED, "hlo_score_records needs read names 0..N-1 in file order");
    }
    reset_timing(ctx);
    int rc = ensure_score_table(ctx, mismatch_prob);
    if (rc) return rc;
    CK(ctx->b_qual.ensure(d.n_bases + 64 + (size_t)d.n_reads + 32));  // 32 bytes of padding on both sides, then a flag per read
    CK(ctx->b_screc.ensure(n * sizeof(hlo_rec) + 32));
    CK(ctx->b_scclu.ensure(n * 4 + 4));
    CK(ctx->b_scout.ensure(n * sizeof(hlo_score) + 16));
    CK(ctx->h_score.ensure(n * sizeof(hlo_score) + 16));
    CK(cudaEventRecord(ctx->ev[ST_BEGIN], s));
    uint8_t *d_qual = ctx->b_qual.as<uint8_t>() + 32, *d_qflag = d_qual + ((d.n_bases + 32 + 15) & ~(uint64_t)15);
    CK(cudaMemsetAsync(ctx->b_qual.p, 0, 32, s));
    CK(cudaMemsetAsync(d_qual + d.n_bases, 0, 32, s));
    CK(cudaMemcpyAsync(d_qual, quals, d.n_bases, cudaMemcpyHostToDevice, s));
    if (n) CK(cudaMemcpyAsync(ctx->b_screc.p, rec, n * sizeof(hlo_rec), cudaMemcpyHostToDevice, s));
    if (n && rec_cluster) CK(cudaMemcpyAsync(ctx->b_scclu.p, rec_cluster, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ctx->ev[ST_H2D], s));
    launch_qual_flags(d_qual, R->b_off.as<uint64_t>(), d.n_reads, d_qflag, s, &ctx->lc);
    launch_score_records(d, d_qual, d_qflag, R->b_off.as<uint64_t>(), R->b_coff.as<uint32_t>(), ctx->b_screc.as<hlo_rec>(),
                         rec_cluster ? ctx->b_scclu.as<uint32_t>() : nullptr, 1, nullptr, n, ctx->b_prob.as<double>() + 128, min_read_len,
                         ctx->b_scout.as<hlo_score>(), s, &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_VERIFY], s));
    if (n) CK(cudaMemcpyAsync(ctx->h_score.p, ctx->b_scout.p, n * sizeof(hlo_score), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[ST_D2H], s));
    CK(cudaStreamSynchronize(s));
    {  // hlo_get_timing after this call: h2d_ms = uploads (qualities, records), verify_ms = the score kernel, d2h_ms, total_ms
        float v = 0.f;
        if (cudaEventElapsedTime(&v, ctx->ev[ST_BEGIN], ctx->ev[ST_H2D]) == cudaSuccess) ctx->tm.h2d_ms = v;
        if (cudaEventElapsedTime(&v, ctx->ev[ST_H2D], ctx->ev[ST_VERIFY]) == cudaSuccess) ctx->tm.verify_ms = v;
        if (cudaEventElapsedTime(&v, ctx->ev[ST_VERIFY], ctx->ev[ST_D2H]) == cudaSuccess) ctx->tm.d2h_ms = v;
        if (cudaEventElapsedTime(&v, ctx->ev[ST_BEGIN], ctx->ev[ST_D2H]) == cudaSuccess) ctx->tm.total_ms = v;
        ctx->tm.n_records = n, ctx->tm.n_reads = d.n_reads, ctx->tm.n_bases = d.n_bases;
        ctx->tm.total_launches = ctx->lc.total;
    }
    *out = ctx->h_score.as<hlo_score>();
    return HLO_OK;
}

int hlo_debug_check_reads(hlo_ctx *ctx, const hlo_reads *Rc) {
    if (!ctx) return HLO_EINVAL;
    const hlo_reads *R = Rc ? Rc : &ctx->scratch_reads;
    if (!R->b_bases.p || !R->b_amb.p) return HLO_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t plane_words = (size_t)R->d.n_reads * R->d.stride_w;
    uint32_t h[4][HLO_PAD_WORDS];
    const DevBuf *bufs[2] = {&R->b_bases, &R->b_amb};
    for (int i = 0; i < 2; ++i) {
        CK(cudaMemcpyAsync(h[2 * i], bufs[i]->p, HLO_PAD_WORDS * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h[2 * i + 1], bufs[i]->as<uint32_t>() + HLO_PAD_WORDS + plane_words * 2, HLO_PAD_WORDS * 4, cudaMemcpyDeviceToHost,
                           ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    const uint32_t want = 0x01010101u * HLO_PAD_CANARY;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < HLO_PAD_WORDS; ++j)
            if (h[i][j] != want)
                FAIL(HLO_ECUDA, "canary overwritten: %s padding of the %s plane, word %d = %08x", (i & 1) ? "trailing" : "leading",
                     i < 2 ? "base" : "ambiguity", j, h[i][j]);
    return HLO_OK;
}

void *hlo_stream(const hlo_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int hlo_get_timing(const hlo_ctx *ctx, hlo_timing *t) {
    if (!ctx || !t) return HLO_EINVAL;
    *t = ctx->tm;
    return HLO_OK;
}

}  // extern "C"
