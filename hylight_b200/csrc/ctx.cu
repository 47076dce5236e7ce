// ctx.cu — C ABI of libhlovlp (include/hlovlp.h): context, packing, stage orchestration, timing.
//
// One ctx = one CUDA device + one stream.  All device work of a call is enqueued on the ctx stream;
// stage boundaries are marked with CUDA events so that hlo_get_timing reports device time per stage.
// There is deliberately no CPU code path for any stage: every entry point fails with HLO_ENODEV /
// HLO_ECUDA when the device is missing.
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

enum Stage { ST_BEGIN = 0, ST_H2D, ST_PACK, ST_SKETCH, ST_SORT, ST_INDEX, ST_PAIRGEN, ST_VERIFY, ST_DEDUP, ST_D2H, ST_COUNT };

}  // namespace

struct hlo_reads {
    hlo_ctx *ctx = nullptr;
    int device = -1;
    DevReads d;
    DevBuf b_ascii, b_off, b_coff, b_len, b_cbase, b_cidx, b_ids, b_nrank, b_perm, b_meta, b_bases, b_amb, b_ambcount;
    std::vector<uint32_t> h_len, h_ids, h_cidx, h_cluster_off, h_perm;
    float h2d_ms = 0.f, pack_ms = 0.f;
    void release() {
        b_ascii.release(), b_off.release(), b_coff.release(), b_len.release(), b_cbase.release(), b_cidx.release(), b_ids.release();
        b_nrank.release(), b_perm.release(), b_meta.release(), b_bases.release(), b_amb.release(), b_ambcount.release();
    }
};

struct hlo_ctx {
    int device = 0;
    hlo_params P;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[ST_COUNT + 1] = {};
    std::string err;
    LaunchCounter lc;
    hlo_timing tm;
    int pairgen_slots_log2 = 8;
    int pairgen_slots_auto = 0;  // raised by batches whose reads overflowed the tables
    int verify_exact = 0, repack = 0, d2h = 1, aux_out = 1, compact = 0, pairgen_generic = 0, sketch_generic = 0, sort_prefix = 1, sort_hash_bits = 0, sort_segmented = 1;
    bool prefix_used = false, prefix_failed = false;
    int sorted_buf = 0;  // which of b_keys / b_vals holds the grouped pairs of the last run_index
    bool keys_ready = false;
    uint32_t *mapped_host = nullptr, *mapped_dev = nullptr;  // 256 B of mapped pinned memory (fetch_small)
    cudaEvent_t ev_vf[2] = {};
    // verify tables
    DevBuf b_tab;
    int tab_len = -1;
    VerifyTables tab{};
    // workspaces
    DevBuf b_cnt, b_mzoff, b_mzx, b_mzrr, b_keys[2], b_vals[2], b_tmp, b_flags, b_bid, b_bstart, b_range, b_hits;
    DevBuf b_cands, b_counters, b_ovf, b_ovf2, b_ovfcnt, b_gkeys, b_gmn, b_gmx, b_goff;
    DevBuf b_rec16, b_deferred, b_best, b_pos, b_out, b_aux, b_coff, b_mz, b_candx, b_sx, b_srp, b_qual, b_screc, b_scclu, b_scout, b_prob;
    PinBuf h_score;
    bool prob_ready = false;
    PinBuf h_out, h_aux, h_coff, h_mz, h_roff, h_cands, h_rec16, h_small;
    std::vector<uint64_t> v_coff;
    hlo_reads scratch_reads;
    // sketch state of the last run (device)
    uint64_t n_mz = 0;
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

// Small device -> host read-backs (counters) go through a kernel that stores into mapped pinned memory, not
// through the copy engine: a 4-byte cudaMemcpy queues behind another context's multi-hundred-MB record
// download on the same engine and stalls the whole pipeline for milliseconds.
#define FETCH(dst, dev_src, nbytes)                                              \
    do {                                                                         \
        int rc__ = fetch_small(ctx, (dst), (dev_src), (nbytes));                 \
        if (rc__) return rc__;                                                   \
    } while (0)

#define FAIL(code, ...)                              \
    do {                                             \
        char buf__[512];                             \
        snprintf(buf__, sizeof buf__, __VA_ARGS__);  \
        ctx->err = buf__;                            \
        return (code);                               \
    } while (0)

// Kernels of different contexts on one device are serialised (one compute section at a time), so that with
// two or more host workers per GPU one worker's uploads/downloads overlap another's kernels instead of every
// worker sharing the SMs and then the PCIe link in lockstep (measured: 18.5 -> ~11 ms per 32-cluster step).
static std::mutex g_compute_mutex[64];

// copies nbytes (multiple of 4, <= 256) from device memory to host memory and waits for it
static int fetch_small(hlo_ctx *ctx, void *dst, const void *dev_src, size_t nbytes) {
    if (nbytes > 256 || (nbytes & 3)) {
        ctx->err = "fetch_small: bad size";
        return HLO_EINVAL;
    }
    launch_copy_words(reinterpret_cast<const uint32_t *>(dev_src), ctx->mapped_dev, (uint32_t)(nbytes / 4), ctx->stream, &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(dst, ctx->mapped_host, nbytes);
    return HLO_OK;
}

extern "C" {

void hlo_params_default(hlo_params *p) {
    if (!p) return;
    p->k = 21, p->w = 11, p->min_seeds = 2, p->min_chain_score = 30, p->mid_occ = 1000;
    p->match_score = 4, p->mismatch_pen = 2, p->ambi_pen = 1, p->end_bonus = 100, p->min_dp_score = 60;
    p->min_ovlp_len = 60, p->max_overhang = 1, p->min_identity = 0.98, p->overhang_ratio = 0.8;
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
    if (P.mid_occ < 1) BAD("mid_occ must be >= 1");
    if (P.match_score < 1 || P.mismatch_pen < 0 || P.ambi_pen < 0) BAD("scores must be A>=1, B>=0, N>=0");
    if (P.min_ovlp_len < 1 || P.min_ovlp_len > 32767) BAD("min_ovlp_len=%d outside [1,32767]", P.min_ovlp_len);
    if (P.max_overhang < 0 || P.max_overhang > 15) BAD("max_overhang=%d outside [0,15]", P.max_overhang);
    if (!(P.min_identity > 0.0 && P.min_identity <= 1.0)) BAD("min_identity must be in (0,1]");
    if (!(P.overhang_ratio >= 0.0)) BAD("overhang_ratio must be >= 0");
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
    if (cudaHostAlloc((void **)&ctx->mapped_host, 256, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&ctx->mapped_dev, ctx->mapped_host, 0) != cudaSuccess) {
        cudaGetLastError();
        hlo_ctx_destroy(ctx);
        return HLO_ENOMEM;
    }
    memset(&ctx->tm, 0, sizeof ctx->tm);
    ctx->scratch_reads.ctx = ctx;
    *out = ctx;
    return HLO_OK;
}

void hlo_ctx_destroy(hlo_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->b_tab, &ctx->b_cnt, &ctx->b_mzoff, &ctx->b_mzx, &ctx->b_mzrr, &ctx->b_keys[0],
                      &ctx->b_keys[1], &ctx->b_vals[0], &ctx->b_vals[1], &ctx->b_tmp, &ctx->b_flags, &ctx->b_bid, &ctx->b_bstart,
                      &ctx->b_range, &ctx->b_hits, &ctx->b_cands, &ctx->b_counters, &ctx->b_ovf, &ctx->b_ovf2, &ctx->b_ovfcnt, &ctx->b_gkeys,
                      &ctx->b_gmn, &ctx->b_gmx, &ctx->b_goff, &ctx->b_rec16, &ctx->b_deferred, &ctx->b_best, &ctx->b_pos,
                      &ctx->b_out, &ctx->b_aux, &ctx->b_coff, &ctx->b_mz, &ctx->b_candx, &ctx->b_sx, &ctx->b_srp, &ctx->b_qual, &ctx->b_screc, &ctx->b_scclu, &ctx->b_scout, &ctx->b_prob};
    for (DevBuf *b : bufs) b->release();
    PinBuf *pins[] = {&ctx->h_score, &ctx->h_out, &ctx->h_aux, &ctx->h_coff, &ctx->h_mz, &ctx->h_roff, &ctx->h_cands, &ctx->h_rec16, &ctx->h_small};
    for (PinBuf *b : pins) b->release();
    ctx->scratch_reads.release();
    for (int i = 0; i <= ST_COUNT; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 2; ++i)
        if (ctx->ev_vf[i]) cudaEventDestroy(ctx->ev_vf[i]);
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
    CK(cudaStreamSynchronize(ctx->stream));
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

static int pack_into(hlo_ctx *ctx, hlo_reads *R, const char *bases, const uint64_t *base_off, uint32_t n_reads,
                     const uint32_t *cluster_off, uint32_t n_clusters, const uint32_t *ids) {
    if (!bases || !base_off) FAIL(HLO_EINVAL, "hlo_pack_reads: null input");
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
    {
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
    std::vector<uint32_t> cbase, nrank;
    if (!plain) {
        R->h_len.resize(n_reads), R->h_ids.resize(n_reads), R->h_cidx.resize(n_reads), R->h_perm.resize(n_reads);
        cbase.resize(n_reads), nrank.resize(n_reads);
        for (uint32_t i = 0; i < n_reads; ++i) R->h_len[i] = (uint32_t)(base_off[i + 1] - base_off[i]);
        for (uint32_t c = 0; c < n_clusters; ++c) {
            const uint32_t r0 = cluster_off[c], r1 = cluster_off[c + 1], N = r1 - r0;
            for (uint32_t i = r0; i < r1; ++i) cbase[i] = r0, R->h_cidx[i] = c, R->h_ids[i] = ids[i];
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
    }
    // the ungapped model equals minimap2's banded extension only while an accepted overlap cannot hold a
    // stretch that triggers z-drop / loses the end bonus
    {
        const int lim = std::min(ctx->P.end_bonus, 100);
        if ((1.0 - ctx->P.min_identity) * (double)max_len * (double)std::max(ctx->P.mismatch_pen, 1) >= (double)lim && max_len > 0)
            FAIL(HLO_EUNSUPPORTED, "reads of %u bp at min_identity %.3f leave the domain where the ungapped end rule is exact", max_len,
                 ctx->P.min_identity);
    }
    const uint64_t n_bases = base_off[n_reads];
    const uint32_t words = (max_len + 15) / 16;
    const uint32_t nv = std::max(1u, (words + 3u) / 4u);
    const uint32_t stride_w = 4u * nv;
    DevReads &d = R->d;
    d.n_reads = n_reads, d.n_clusters = n_clusters, d.stride_w = stride_w, d.nv = nv, d.max_len = max_len, d.n_bases = n_bases;
    d.max_cluster = max_cluster;
    d.plain_names = plain ? 1 : 0;
    const size_t plane_words = (size_t)n_reads * stride_w;
    CK(R->b_ascii.ensure(n_bases + 16));
    CK(R->b_off.ensure(((size_t)n_reads + 1) * 8));
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
    CK(R->b_ambcount.ensure(16));
    d.len = R->b_len.as<uint32_t>(), d.cbase = R->b_cbase.as<uint32_t>(), d.cidx = R->b_cidx.as<uint32_t>();
    d.ids = R->b_ids.as<uint32_t>(), d.nrank = R->b_nrank.as<uint32_t>(), d.perm = R->b_perm.as<uint32_t>();
    d.meta = R->b_meta.as<uint4>();
    d.bases = R->b_bases.as<uint32_t>() + HLO_PAD_WORDS, d.amb = R->b_amb.as<uint32_t>() + HLO_PAD_WORDS;
    CK(cudaEventRecord(ctx->ev[ST_BEGIN], s));
    if (n_reads) {
        CK(cudaMemcpyAsync(R->b_ascii.p, bases, n_bases, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(R->b_off.p, base_off, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(R->b_coff.p, cluster_off, ((size_t)n_clusters + 1) * 4, cudaMemcpyHostToDevice, s));
        if (!plain) {
            CK(cudaMemcpyAsync(d.len, R->h_len.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d.cbase, cbase.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d.cidx, R->h_cidx.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d.ids, R->h_ids.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d.nrank, nrank.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
            std::vector<uint4> meta(n_reads);
            for (uint32_t i = 0; i < n_reads; ++i) meta[i] = make_uint4(R->h_len[i], R->h_ids[i], R->h_cidx[i], cbase[i]);
            CK(cudaMemcpyAsync(d.meta, meta.data(), (size_t)n_reads * 16, cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));  // (meta is a local)
            CK(cudaMemcpyAsync(d.perm, R->h_perm.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, s));
        }
    }
    CK(cudaMemsetAsync(R->b_ambcount.p, 0, 16, s));
    CK(cudaEventRecord(ctx->ev[ST_H2D], s));
    if (plain && n_reads) launch_read_meta(R->b_off.as<uint64_t>(), R->b_coff.as<uint32_t>(), d, s, &ctx->lc);
    launch_pack(R->b_ascii.as<uint8_t>(), R->b_off.as<uint64_t>(), d, R->b_ambcount.as<uint32_t>(), s, &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_PACK], s));
    uint32_t amb = 0;
    FETCH(&amb, R->b_ambcount.p, 4);  // the wait also keeps the caller's host buffers alive until the copies are done
    d.has_amb = amb != 0;
    CK(cudaEventElapsedTime(&R->h2d_ms, ctx->ev[ST_BEGIN], ctx->ev[ST_H2D]));
    CK(cudaEventElapsedTime(&R->pack_ms, ctx->ev[ST_H2D], ctx->ev[ST_PACK]));
    int rc = build_tables(ctx, (int)max_len);
    if (rc) return rc;
    return HLO_OK;
}

extern "C" int hlo_pack_reads(hlo_ctx *ctx, const char *bases, const uint64_t *base_off, uint32_t n_reads,
                              const uint32_t *cluster_off, uint32_t n_clusters, const uint32_t *ids, hlo_reads **out) {
    if (!ctx || !out) return HLO_EINVAL;
    *out = nullptr;
    hlo_reads *R = new (std::nothrow) hlo_reads();
    if (!R) return HLO_ENOMEM;
    int rc = pack_into(ctx, R, bases, base_off, n_reads, cluster_off, n_clusters, ids);
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
// stages

// sketch.  Leaves mz_off (n_reads+1, name order), mz_x, mz_rp, mz_rid on the device; with the single-pass
// fast kernel also the radix sort's input pairs (keys_ready).
static int run_sketch(hlo_ctx *ctx, const hlo_reads *R, bool want_x) {
    const DevReads &d = R->d;
    cudaStream_t s = ctx->stream;
    const uint32_t n = d.n_reads;
    if ((uint64_t)d.n_bases > 0xffffffffull) FAIL(HLO_ERANGE, "batch of %llu bases exceeds the 2^32 minimizer index range", (unsigned long long)d.n_bases);
    CK(ctx->b_cnt.ensure(((size_t)n + 1) * 4));
    CK(ctx->b_mzoff.ensure(((size_t)n + 2) * 4));
    CK(ctx->b_tmp.ensure(scan_tmp_words((uint64_t)n + 1) * 4));
    CK(ctx->b_counters.ensure(16 * 8));
    ctx->keys_ready = false;
    const bool fast = sketch_fast_supported(ctx->P.w, ctx->P.k) && !ctx->sketch_generic && n > 0;
    uint32_t cap = 0;
    if (fast) {
        // strip capacity: 3.5x the expected density 2/(w+1) plus slack, never more than one per k-mer
        // (sized for reads of at most 1024 bp: in a batch that mixes short reads with a few contigs the long ones overflow
        // their strip and take the generic kernel, instead of every read paying for a contig-sized strip)
        const uint32_t cap_len = std::min<uint32_t>(d.max_len, 1024u);
        const uint32_t kmers = cap_len >= (uint32_t)ctx->P.k ? cap_len - (uint32_t)ctx->P.k + 1 : 0;
        cap = std::min<uint32_t>(std::max<uint32_t>(kmers, 1u), (uint32_t)(7.0 * kmers / (ctx->P.w + 1)) + 8u);
        CK(ctx->b_sx.ensure((size_t)n * cap * 8 + 8));
        CK(ctx->b_srp.ensure((size_t)n * cap * 4 + 4));
        CK(ctx->b_ovf.ensure((size_t)n * 4 + 4));
        CK(cudaMemsetAsync(ctx->b_cnt.as<uint32_t>() + n, 0, 4, s));
        CK(cudaMemsetAsync(ctx->b_counters.as<unsigned long long>() + 10, 0, 16, s));
        launch_sketch_fast(d, ctx->P.w, ctx->P.k, cap, ctx->b_sx.as<uint64_t>(), ctx->b_srp.as<uint32_t>(), ctx->b_cnt.as<uint32_t>(),
                           ctx->b_ovf.as<uint32_t>(), ctx->b_counters.as<unsigned long long>() + 10, s, &ctx->lc);
    } else {
        CK(cudaMemsetAsync(ctx->b_cnt.p, 0, ((size_t)n + 1) * 4, s));
        launch_sketch_count(d, ctx->P.w, ctx->P.k, ctx->b_cnt.as<uint32_t>(), s, &ctx->lc);
    }
    // scanning n+1 counts (the last is 0) puts the total at mz_off[n]
    launch_exclusive_scan(ctx->b_cnt.as<uint32_t>(), ctx->b_mzoff.as<uint32_t>(), (uint64_t)n + 1, ctx->b_tmp.as<uint32_t>(),
                          ctx->b_counters.as<uint32_t>() + 22, s, &ctx->lc);  // grand total -> low word of counters[11]
    CK(cudaGetLastError());
    uint32_t total = 0;
    unsigned long long n_ovf = 0;
    if (fast) {  // one fetch for both: the scan leaves the total right behind the sketch overflow counter
        unsigned long long two[2] = {0, 0};
        FETCH(two, ctx->b_counters.as<unsigned long long>() + 10, 16);
        n_ovf = two[0], total = (uint32_t)two[1];
    } else {
        FETCH(&total, ctx->b_counters.as<uint32_t>() + 22, 4);
    }
    ctx->n_mz = total;
    // the hash words themselves are only wanted by hlo_sketch: the index is built from the sort keys
    uint64_t *mz_x = nullptr;
    if (want_x || !fast) {
        CK(ctx->b_mzx.ensure((size_t)total * 8 + 8));
        mz_x = ctx->b_mzx.as<uint64_t>();
    }
    CK(ctx->b_mzrr.ensure((size_t)total * 8 + 8));
    if (fast) {
        for (int i = 0; i < 2; ++i) {
            CK(ctx->b_keys[i].ensure((size_t)total * 8 + 8));
            CK(ctx->b_vals[i].ensure((size_t)total * 4 + 4));
        }
        launch_sketch_compact(d, cap, ctx->b_sx.as<uint64_t>(), ctx->b_srp.as<uint32_t>(), ctx->b_cnt.as<uint32_t>(),
                              ctx->b_mzoff.as<uint32_t>(), mz_x, ctx->b_mzrr.as<uint2>(), ctx->b_keys[0].as<uint64_t>(),
                              ctx->b_vals[0].as<uint32_t>(), 2 * ctx->P.k, s, &ctx->lc);
        ctx->keys_ready = true;
        if (n_ovf)  // reads with more minimizers than a strip holds, or low-complexity reads: the generic kernel writes them in place
            launch_sketch_write(d, ctx->P.w, ctx->P.k, ctx->b_mzoff.as<uint32_t>(), mz_x, ctx->b_mzrr.as<uint2>(), ctx->b_keys[0].as<uint64_t>(),
                                ctx->b_vals[0].as<uint32_t>(), 2 * ctx->P.k, ctx->b_ovf.as<uint32_t>(), (uint32_t)n_ovf, s, &ctx->lc);
    } else {
        launch_sketch_write(d, ctx->P.w, ctx->P.k, ctx->b_mzoff.as<uint32_t>(), mz_x, ctx->b_mzrr.as<uint2>(), nullptr, nullptr, 0, nullptr, 0, s,
                            &ctx->lc);
    }
    CK(cudaGetLastError());
    return HLO_OK;
}

static int bits_for(uint64_t v) {
    int b = 1;
    while (b < 63 && (1ull << b) < v) ++b;
    return b;
}

// hash sort + bucket index.  Leaves range[] (per minimizer, read order) and hits[] (hash order).
static int run_index(hlo_ctx *ctx, const hlo_reads *R, bool allow_prefix) {
    cudaStream_t s = ctx->stream;
    const uint64_t n = ctx->n_mz;
    for (int i = 0; i < 2; ++i) {
        CK(ctx->b_keys[i].ensure(n * 8 + 8));
        CK(ctx->b_vals[i].ensure(n * 4 + 4));
    }
    CK(ctx->b_tmp.ensure(std::max(std::max(sort_tmp_words(n), sort_segmented_tmp_words(n, R->d.n_clusters)), scan_tmp_words(n + 1)) * 4));
    CK(ctx->b_flags.ensure(n * 4 + 4));
    CK(ctx->b_range.ensure(n * 8 + 8));
    CK(ctx->b_hits.ensure(n * 8 + 8));
    uint64_t *keys[2] = {ctx->b_keys[0].as<uint64_t>(), ctx->b_keys[1].as<uint64_t>()};
    uint32_t *vals[2] = {ctx->b_vals[0].as<uint32_t>(), ctx->b_vals[1].as<uint32_t>()};
    if (!allow_prefix) {
        // a redo after a prefix attempt: the (key, index) pairs are intact but permuted — put every pair back at its
        // index, into the other buffer, and sort from there
        const int from = ctx->sorted_buf;
        launch_unpermute_pairs(keys[from], vals[from], n, keys[from ^ 1], vals[from ^ 1], s, &ctx->lc);
        if (from == 0) std::swap(keys[0], keys[1]), std::swap(vals[0], vals[1]);
    } else if (!ctx->keys_ready) {
        launch_hash_keys(ctx->b_mzx.as<uint64_t>(), ctx->b_mzrr.as<uint2>(), R->d.cidx, 2 * ctx->P.k, n, keys[0], vals[0], s, &ctx->lc);
    }
    // key = cluster index << 2k | hash: 2k hash bits + the bits of the cluster count
    const int key_bits = 2 * ctx->P.k + (R->d.n_clusters > 1 ? bits_for(R->d.n_clusters) : 0);
    if (key_bits > 64) FAIL(HLO_ERANGE, "%u clusters with k=%d need a %d-bit sort key (limit 64): use smaller batches", R->d.n_clusters, ctx->P.k, key_bits);
    // (1) The minimizer array is cluster-major already, so every cluster is sorted as its own segment: the cluster bits
    //     of the key cost no pass.  (2) Grouping needs equal keys adjacent and stable, not a total order: only the top
    //     24 hash bits are sorted (3 passes instead of 6) and the rare groups in which two hashes share that prefix are
    //     repaired afterwards.
    const int hash_bits = 2 * ctx->P.k;
    const bool segmented = ctx->sort_segmented && n < (1ull << 32);
    const int top = segmented ? hash_bits : key_bits;          // sorted bits end here
    const int avail = segmented ? hash_bits : key_bits;        // ... and this many exist below
    int sort_bits = ctx->sort_hash_bits > 0 ? (segmented ? 0 : key_bits - hash_bits) + ctx->sort_hash_bits : (segmented ? 24 : ((key_bits - hash_bits + 19 + 7) / 8) * 8);
    const bool prefix = allow_prefix && ctx->sort_prefix && sort_bits < avail && n > 1;
    ctx->prefix_used = prefix;
    if (!prefix) sort_bits = avail;
    int cur;
    if (segmented)
        cur = launch_radix_sort_segmented(keys, vals, n, top - sort_bits, top, ctx->b_mzoff.as<uint32_t>(), R->b_coff.as<uint32_t>(), R->d.n_clusters,
                                          ctx->b_tmp.as<uint32_t>(), s, &ctx->lc);
    else
        cur = launch_radix_sort(keys, vals, n, top - sort_bits, top, ctx->b_tmp.as<uint32_t>(), s, &ctx->lc);
    if (prefix) {  // counters[12] = (groups repaired, 1 if one was too large: run_pairgen's fetch sees it and the index is redone)
        CK(cudaMemsetAsync(ctx->b_counters.as<unsigned long long>() + 12, 0, 8, s));
        launch_collision_fix(keys[cur], vals[cur], n, top - sort_bits, ctx->b_flags.as<uint32_t>(), ctx->b_counters.as<uint32_t>() + 24, s,
                             &ctx->lc);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_SORT], s));
    ctx->sorted_buf = keys[cur] == ctx->b_keys[0].as<uint64_t>() ? 0 : 1;
    launch_bucket_index(keys[cur], vals[cur], ctx->b_mzrr.as<uint2>(), n, ctx->P.mid_occ, ctx->b_hits.as<uint2>(), ctx->b_range.as<uint2>(), s,
                        &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_INDEX], s));
    return HLO_OK;
}

constexpr int PAIRGEN_WIDE_SLOTS_LOG2 = 11;

static uint32_t next_pow2_u32(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// candidates into b_cands (CandI, or hlo_cand when external); *n_out = their number
static int run_pairgen(hlo_ctx *ctx, const hlo_reads *R, int external, uint64_t *n_out) {
    cudaStream_t s = ctx->stream;
    const DevReads &d = R->d;
    CK(ctx->b_counters.ensure(16 * 8));
    CK(ctx->b_ovf.ensure((size_t)d.n_reads * 4 + 4));
    uint64_t cap = std::max<uint64_t>(ctx->b_cands.cap / 16, std::max<uint64_t>(ctx->n_mz * 3, 1 << 16));
    PairgenArgs a;
    a.reads = d;
    a.mz_off = ctx->b_mzoff.as<uint32_t>(), a.mz_rr = ctx->b_mzrr.as<uint2>();
    a.range = ctx->b_range.as<uint2>(), a.hits = ctx->b_hits.as<uint2>();
    a.k = ctx->P.k;
    a.gap_min = std::max(std::max(ctx->P.min_chain_score - ctx->P.k, ctx->P.min_seeds - 1), 0);
    a.external = external;
    a.counters = ctx->b_counters.as<unsigned long long>();
    a.overflow_reads = ctx->b_ovf.as<uint32_t>(), a.overflow_cap = d.n_reads;
    const bool fast = pairgen_fast_supported(d, &a.dg_bits) && !ctx->pairgen_generic;
    unsigned long long hc[13];
    bool wide_pass = false;
    const int slots0 = std::max(ctx->pairgen_slots_log2, ctx->pairgen_slots_auto);
    ctx->prefix_failed = false;
    if (ctx->prefix_used && !fast) {  // the fast path reads the prefix sort's flag with its own counters below
        unsigned long long flag = 0;
        FETCH(&flag, ctx->b_counters.as<unsigned long long>() + 12, 8);
        if (flag >> 32) {
            ctx->prefix_failed = true;
            *n_out = 0;
            return HLO_OK;
        }
    }
    for (int attempt = 0; attempt < 3; ++attempt) {
        CK(ctx->b_cands.ensure(cap * 16));
        a.cands = ctx->b_cands.p, a.cand_cap = cap;
        CK(cudaMemsetAsync(ctx->b_counters.p, 0, 64, s));
        uint32_t n_list = d.n_reads;        // reads the generic kernel has to do
        const uint32_t *list = nullptr;     // nullptr = all of them
        if (fast) {
            launch_pairgen(a, slots0, s, &ctx->lc);
            CK(cudaGetLastError());
            FETCH(hc, ctx->b_counters.p, sizeof hc);
            if (ctx->prefix_used && (hc[12] >> 32)) {  // a mixed prefix group was too large to repair: the caller redoes the index
                ctx->prefix_failed = true;
                *n_out = 0;
                return HLO_OK;
            }
            n_list = (uint32_t)hc[2], list = a.overflow_reads;
            const uint32_t first_overflow = n_list;
            // reads with more (target, strand, diagonal) chains than the table holds (deep coverage, longer reads) go through
            // the same kernel again with tables twice the size, up to 2048 slots, before the HBM-table kernel gets what is
            // left.  (Straight to the largest tables is slow: one block per SM and a 16 KB table to clear per read.)
            uint32_t *lists[2] = {ctx->b_ovf.as<uint32_t>(), nullptr};
            int from = 0;
            for (int sl = slots0 + 1; n_list && sl <= PAIRGEN_WIDE_SLOTS_LOG2; ++sl) {
                if (!lists[1]) {
                    CK(ctx->b_ovf2.ensure((size_t)d.n_reads * 4 + 4));
                    lists[1] = ctx->b_ovf2.as<uint32_t>();
                }
                PairgenArgs a2 = a;
                a2.work = lists[from], a2.n_work = n_list;
                a2.overflow_reads = lists[from ^ 1];
                CK(cudaMemsetAsync(ctx->b_counters.as<unsigned long long>() + 2, 0, 16, s));  // overflow count, work counter
                launch_pairgen(a2, sl, s, &ctx->lc);
                CK(cudaGetLastError());
                FETCH(hc, ctx->b_counters.p, sizeof hc);
                n_list = (uint32_t)hc[2], list = lists[from ^ 1];
                from ^= 1;
                wide_pass = true;
            }
            if (wide_pass) ctx->tm.n_overflow_reads = first_overflow;  // (reads that left the first table size)
            // a batch that overflowed a lot: the following ones start with larger tables
            if (attempt == 0 && first_overflow > d.n_reads / 16 && ctx->pairgen_slots_auto < 10) ctx->pairgen_slots_auto = slots0 + 1;
        }
        if (n_list) {  // exact path with tables in HBM, in chunks that bound the table memory
            CK(ctx->b_ovfcnt.ensure((size_t)n_list * 4 + 4));
            launch_generic_hit_counts(a, list, n_list, ctx->b_ovfcnt.as<uint32_t>(), s, &ctx->lc);
            std::vector<uint32_t> cnt(n_list);
            CK(cudaMemcpyAsync(cnt.data(), ctx->b_ovfcnt.p, (size_t)n_list * 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            const uint64_t slot_budget = 1ull << 28;  // 4 GiB of tables per chunk
            uint32_t i0 = 0;
            std::vector<uint64_t> off;
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
                    std::vector<uint32_t> xs(i1 - i0);
                    for (uint32_t i = i0; i < i1; ++i) xs[i - i0] = i;
                    CK(cudaMemcpyAsync(ctx->b_ovf.p, xs.data(), xs.size() * 4, cudaMemcpyHostToDevice, s));
                    CK(cudaStreamSynchronize(s));
                    chunk_list = ctx->b_ovf.as<uint32_t>();
                }
                launch_pairgen_generic(a, chunk_list, i1 - i0, ctx->b_gkeys.as<unsigned long long>(), ctx->b_gmn.as<uint32_t>(),
                                       ctx->b_gmx.as<uint32_t>(), ctx->b_goff.as<uint64_t>(), s, &ctx->lc);
                CK(cudaGetLastError());
                CK(cudaStreamSynchronize(s));
                i0 = i1;
            }
            FETCH(hc, ctx->b_counters.p, sizeof hc);
        }
        if (!wide_pass) ctx->tm.n_overflow_reads = fast ? n_list : 0;
        ctx->tm.n_seed_hits = hc[1];
        if (hc[0] <= cap) {
            *n_out = hc[0];
            return HLO_OK;
        }
        cap = hc[0] + hc[0] / 16 + 1024;  // buffer was too small: the kernels only counted past the end; redo
    }
    FAIL(HLO_ECUDA, "candidate buffer sizing did not converge");
}

// verify candidates (device) -> Rec16 in b_rec16; *n_acc = accepted
static int run_verify(hlo_ctx *ctx, const hlo_reads *R, const CandI *d_cands, uint64_t n, uint64_t *n_acc) {
    cudaStream_t s = ctx->stream;
    *n_acc = 0;
    CK(ctx->b_counters.ensure(8 * 8));
    CK(ctx->b_rec16.ensure(n * sizeof(Rec16) + 16));
    CK(ctx->b_deferred.ensure(n * 4 + 4));
    CK(cudaMemsetAsync(ctx->b_counters.p, 0, 64, s));
    VerifyArgs a;
    a.reads = R->d, a.tab = ctx->tab, a.prm = verify_params(ctx);
    a.cands = d_cands, a.n = n;
    a.out = ctx->b_rec16.as<Rec16>();
    a.counters = ctx->b_counters.as<unsigned long long>();
    a.deferred = ctx->b_deferred.as<uint32_t>();
    const bool force_exact = ctx->verify_exact != 0 || !verify_fast_supported(R->d, ctx->tab);
    unsigned long long hc[2] = {0, 0};
    if (force_exact) {
        launch_verify_exact(a, nullptr, n, s, &ctx->lc);
    } else {
        CK(cudaEventRecord(ctx->ev_vf[0], s));
        launch_verify_fast(a, s, &ctx->lc);
        CK(cudaEventRecord(ctx->ev_vf[1], s));
        CK(cudaGetLastError());
        FETCH(hc, ctx->b_counters.p, sizeof hc);
        if (n && cudaEventElapsedTime(&ctx->tm.verify_fast_ms, ctx->ev_vf[0], ctx->ev_vf[1]) != cudaSuccess) cudaGetLastError();
        if (hc[1]) launch_verify_exact(a, a.deferred, hc[1], s, &ctx->lc);
    }
    CK(cudaGetLastError());
    const unsigned long long n_def = hc[1];
    if (force_exact || n_def) FETCH(hc, ctx->b_counters.p, sizeof hc);  // else the count fetched after the fast kernel is final
    ctx->tm.n_deferred = force_exact ? n : n_def;
    *n_acc = hc[0];
    return HLO_OK;
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
    if (last_stage >= ST_D2H) t.d2h_ms = el(ST_DEDUP, ST_D2H);
    t.total_ms = t.h2d_ms + t.pack_ms + t.sketch_ms + t.sort_ms + t.index_ms + t.pairgen_ms + t.verify_ms + t.dedup_ms + t.d2h_ms;
    t.n_reads = R->d.n_reads, t.n_bases = R->d.n_bases, t.n_minimizers = ctx->n_mz;
    t.total_launches = ctx->lc.total, t.verify_launches = ctx->lc.verify;
}

static void reset_timing(hlo_ctx *ctx) {
    memset(&ctx->tm, 0, sizeof ctx->tm);
    ctx->lc = LaunchCounter();
}

extern "C" {

int hlo_sketch(hlo_ctx *ctx, const hlo_reads *R, const hlo_mz **mz, const uint64_t **read_off, uint64_t *n) {
    if (!ctx || !R || !mz || !read_off || !n) return HLO_EINVAL;
    CK(cudaSetDevice(ctx->device));
    reset_timing(ctx);
    cudaStream_t s = ctx->stream;
    CK(cudaEventRecord(ctx->ev[ST_PACK], s));
    int rc = run_sketch(ctx, R, true);
    if (rc) return rc;
    const uint64_t tot = ctx->n_mz;
    CK(ctx->b_mz.ensure(tot * sizeof(hlo_mz) + 16));
    launch_mz_assemble(ctx->b_mzx.as<uint64_t>(), ctx->b_mzrr.as<uint2>(), tot, ctx->b_mz.as<hlo_mz>(), s,
                       &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_SKETCH], s));
    // the device arrays are in name order (sketch.cu); hand them back in the caller's read order
    rc = ensure_host_meta(ctx, R);
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

static int stages_to_candidates(hlo_ctx *ctx, const hlo_reads *R, int external, uint64_t *n_c, bool mark_start = true) {
    cudaStream_t s = ctx->stream;
    if (mark_start) CK(cudaEventRecord(ctx->ev[ST_PACK], s));
    int rc = run_sketch(ctx, R, false);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[ST_SKETCH], s));
    rc = run_index(ctx, R, true);
    if (rc) return rc;
    rc = run_pairgen(ctx, R, external, n_c);
    if (rc) return rc;
    if (ctx->prefix_failed) {  // rare: two heavily repeated minimizers share a hash prefix — sort on all bits instead
        ctx->tm.sort_fallback = 1.0f;
        rc = run_index(ctx, R, false);
        if (rc) return rc;
        rc = run_pairgen(ctx, R, external, n_c);
        if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->ev[ST_PAIRGEN], s));
    ctx->tm.n_candidates = *n_c;
    return HLO_OK;
}

int hlo_candidates(hlo_ctx *ctx, const hlo_reads *R, const hlo_cand **cands, uint64_t *n) {
    if (!ctx || !R || !cands || !n) return HLO_EINVAL;
    CK(cudaSetDevice(ctx->device));
    reset_timing(ctx);
    uint64_t n_c = 0;
    int rc = stages_to_candidates(ctx, R, 1, &n_c);
    if (rc) return rc;
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
    CK(cudaSetDevice(ctx->device));
    reset_timing(ctx);
    cudaStream_t s = ctx->stream;
    {
        const int rcm = ensure_host_meta(ctx, R);
        if (rcm) return rcm;
    }
    for (uint64_t i = 0; i < n; ++i)
        if (cands[i].q >= R->d.n_reads || cands[i].t >= R->d.n_reads) FAIL(HLO_EINVAL, "candidate %llu names a read outside the batch", (unsigned long long)i);
    CK(ctx->b_candx.ensure(n * sizeof(hlo_cand) + 16));
    CK(ctx->b_cands.ensure(n * sizeof(CandI) + 16));
    CK(cudaMemcpyAsync(ctx->b_candx.p, cands, n * sizeof(hlo_cand), cudaMemcpyHostToDevice, s));
    launch_cands_to_internal(ctx->b_candx.as<hlo_cand>(), n, R->d, ctx->b_cands.as<CandI>(), s, &ctx->lc);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[ST_PAIRGEN], s));
    uint64_t n_acc = 0;
    int rc = run_verify(ctx, R, ctx->b_cands.as<CandI>(), n, &n_acc);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[ST_VERIFY], s));
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
    for (uint64_t i = 0; i < n_acc; ++i) {
        const Rec16 &x = h[i];
        expand_record(R->h_ids[x.q], R->h_ids[x.t], (int)R->h_len[x.q], (int)R->h_len[x.t], (int)(x.ds >> 1) - HLO_DIAG_BIAS,
                      (int)(x.ds & 1u), x.kb & 0xffffu, &ro[i]);
        ao[i].cluster = R->h_cidx[x.q], ao[i].blocklen = x.kb >> 16;
    }
    out->rec = ro, out->rec16 = nullptr, out->aux = ao, out->n = n_acc, out->cluster_off = nullptr, out->n_clusters = R->d.n_clusters;
    ctx->tm.n_candidates = n, ctx->tm.n_accepted = n_acc, ctx->tm.n_records = n_acc;
    {
        float v = 0.f;
        if (cudaEventElapsedTime(&v, ctx->ev[ST_PAIRGEN], ctx->ev[ST_VERIFY]) == cudaSuccess) ctx->tm.verify_ms = v;
        else cudaGetLastError();
        ctx->tm.total_ms = ctx->tm.verify_ms;
        ctx->tm.n_reads = R->d.n_reads, ctx->tm.n_bases = R->d.n_bases;
        ctx->tm.total_launches = ctx->lc.total, ctx->tm.verify_launches = ctx->lc.verify;
    }
    return HLO_OK;
}

int hlo_overlap_batch(hlo_ctx *ctx, const hlo_reads *R, hlo_recs *out) {
    if (!ctx || !R || !out) return HLO_EINVAL;
    CK(cudaSetDevice(ctx->device));
    reset_timing(ctx);
    cudaStream_t s = ctx->stream;
    const DevReads &d = R->d;
    uint64_t n_c = 0, n_acc = 0;
    float repack_ms = -1.f;
    std::unique_lock<std::mutex> compute_lock(g_compute_mutex[ctx->device & 63]);
    if (ctx->repack && d.n_reads) {
        // the packer again, from the ASCII bases still resident in HBM (results are identical; the
        // ambiguity flags it ORs into len[] are idempotent)
        CK(cudaEventRecord(ctx->ev[ST_H2D], s));
        launch_pack(R->b_ascii.as<uint8_t>(), R->b_off.as<uint64_t>(), d, R->b_ambcount.as<uint32_t>(), s, &ctx->lc);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[ST_PACK], s));
        repack_ms = 0.f;
    }
    int rc = stages_to_candidates(ctx, R, 0, &n_c, repack_ms < 0.f);
    if (rc) return rc;
    rc = run_verify(ctx, R, ctx->b_cands.as<CandI>(), n_c, &n_acc);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[ST_VERIFY], s));
    ctx->tm.n_accepted = n_acc;
    if (n_acc > 0xfffffff0ull) FAIL(HLO_ERANGE, "more than 2^32 accepted overlaps in one batch");
    // per-pair keep-longest + (idA, idB) order + expansion: group by idA, sort the small groups (records.cu)
    const uint32_t ng = d.n_reads;
    CK(ctx->b_flags.ensure(((size_t)ng + 2) * 4));   // records per idA
    CK(ctx->b_bid.ensure(((size_t)ng + 2) * 4));     // group offsets
    CK(ctx->b_best.ensure(((size_t)ng + 2) * 4));    // distinct idB per idA
    CK(ctx->b_pos.ensure(((size_t)ng + 2) * 4));     // output offsets
    CK(ctx->b_vals[0].ensure(n_acc * 4 + 4));        // slot of a record inside its group
    CK(ctx->b_keys[0].ensure(n_acc * 8 + 8));        // (rank of idB, record index) entries
    CK(ctx->b_vals[1].ensure(n_acc * 4 + 4));        // group of an entry
    CK(ctx->b_tmp.ensure(scan_tmp_words((uint64_t)ng + 1) * 4));
    CK(ctx->b_out.ensure(n_acc * sizeof(hlo_rec) + 32));
    CK(ctx->b_aux.ensure(n_acc * sizeof(hlo_rec_aux) + 32));
    CK(ctx->b_coff.ensure(((size_t)d.n_clusters + 2) * 8));
    uint32_t *g_cnt = ctx->b_flags.as<uint32_t>(), *g_off = ctx->b_bid.as<uint32_t>();
    uint32_t *g_uniq = ctx->b_best.as<uint32_t>(), *g_uoff = ctx->b_pos.as<uint32_t>();
    uint2 *g_ent = ctx->b_keys[0].as<uint2>();
    CK(cudaMemsetAsync(g_cnt, 0, ((size_t)ng + 1) * 4, s));
    CK(cudaMemsetAsync(g_uniq, 0, ((size_t)ng + 1) * 4, s));
    launch_group_count(ctx->b_rec16.as<Rec16>(), n_acc, d, g_cnt, ctx->b_vals[0].as<uint32_t>(), s, &ctx->lc);
    launch_exclusive_scan(g_cnt, g_off, (uint64_t)ng + 1, ctx->b_tmp.as<uint32_t>(), nullptr, s, &ctx->lc);
    uint32_t *g_entg = ctx->b_vals[1].as<uint32_t>();
    launch_group_scatter(ctx->b_rec16.as<Rec16>(), n_acc, d, g_off, ctx->b_vals[0].as<uint32_t>(), g_ent, g_entg, s, &ctx->lc);
    uint32_t *g_lt = ctx->b_vals[0].as<uint32_t>();  // (the slots are consumed by the scatter)
    launch_group_mark(ctx->b_rec16.as<Rec16>(), n_acc, g_off, g_entg, 1, g_ent, g_uniq, g_lt, s, &ctx->lc);
    uint32_t n_out32 = 0;
    launch_exclusive_scan(g_uniq, g_uoff, (uint64_t)ng + 1, ctx->b_tmp.as<uint32_t>(), ctx->b_counters.as<uint32_t>() + 8, s, &ctx->lc);
    launch_group_expand(ctx->b_rec16.as<Rec16>(), n_acc, d, g_off, g_uoff, g_entg, g_ent, g_lt, ctx->b_out.as<hlo_rec>(),
                        ctx->compact ? ctx->b_out.as<hlo_rec16>() : nullptr, ctx->aux_out ? ctx->b_aux.as<hlo_rec_aux>() : nullptr, s, &ctx->lc);
    if (d.n_reads) launch_cluster_offsets(g_uoff, R->b_coff.as<uint32_t>(), d.n_clusters, ctx->b_coff.as<unsigned long long>(), s, &ctx->lc);
    else CK(cudaMemsetAsync(ctx->b_coff.p, 0, ((size_t)d.n_clusters + 1) * 8, s));
    CK(cudaGetLastError());
    FETCH(&n_out32, ctx->b_counters.as<uint32_t>() + 8, 4);
    const uint64_t n_out = n_out32;
    CK(cudaEventRecord(ctx->ev[ST_DEDUP], s));
    compute_lock.unlock();  // the download below overlaps the next worker's kernels
    if (ctx->d2h) {
        CK(ctx->h_out.ensure(n_out * sizeof(hlo_rec) + 32));
        CK(ctx->h_aux.ensure(n_out * sizeof(hlo_rec_aux) + 32));
        CK(ctx->h_coff.ensure(((size_t)d.n_clusters + 2) * 8));
        CK(cudaMemcpyAsync(ctx->h_out.p, ctx->b_out.p, n_out * (ctx->compact ? sizeof(hlo_rec16) : sizeof(hlo_rec)), cudaMemcpyDeviceToHost, s));
        if (ctx->aux_out) CK(cudaMemcpyAsync(ctx->h_aux.p, ctx->b_aux.p, n_out * sizeof(hlo_rec_aux), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->h_coff.p, ctx->b_coff.p, ((size_t)d.n_clusters + 1) * 8, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaEventRecord(ctx->ev[ST_D2H], s));
    CK(cudaStreamSynchronize(s));
    if (ctx->d2h) {
        out->rec = ctx->compact ? nullptr : ctx->h_out.as<hlo_rec>();
        out->rec16 = ctx->compact ? ctx->h_out.as<hlo_rec16>() : nullptr;
        out->aux = ctx->aux_out ? ctx->h_aux.as<hlo_rec_aux>() : nullptr;
        out->cluster_off = ctx->h_coff.as<uint64_t>();
    } else {
        out->rec = nullptr, out->rec16 = nullptr, out->aux = nullptr, out->cluster_off = nullptr;
    }
    out->n = n_out, out->n_clusters = d.n_clusters;
    ctx->tm.n_records = n_out;
    finish_timing(ctx, R, ST_D2H);
    if (repack_ms >= 0.f) {  // this call ran the packer itself: report its time instead of the upload's
        float v = 0.f;
        if (cudaEventElapsedTime(&v, ctx->ev[ST_H2D], ctx->ev[ST_PACK]) != cudaSuccess) cudaGetLastError(), v = 0.f;
        ctx->tm.total_ms += v - ctx->tm.pack_ms - ctx->tm.h2d_ms;
        ctx->tm.pack_ms = v, ctx->tm.h2d_ms = 0.f;
    }
    return HLO_OK;
}

int hlo_overlap_host(hlo_ctx *ctx, const char *bases, const uint64_t *base_off, uint32_t n_reads, const uint32_t *cluster_off,
                     uint32_t n_clusters, hlo_recs *out) {
    if (!ctx || !out) return HLO_EINVAL;
    int rc = pack_into(ctx, &ctx->scratch_reads, bases, base_off, n_reads, cluster_off, n_clusters, nullptr);
    if (rc) return rc;
    return hlo_overlap_batch(ctx, &ctx->scratch_reads, out);
}

int hlo_set_option(hlo_ctx *ctx, const char *name, int64_t value) {
    if (!ctx || !name) return HLO_EINVAL;
    if (!strcmp(name, "verify_exact")) {
        ctx->verify_exact = value != 0;
        return HLO_OK;
    }
    if (!strcmp(name, "sort_onesweep")) {
        radix_sort_use_onesweep(value != 0);
        return HLO_OK;
    }
    if (!strcmp(name, "sketch_generic")) {
        ctx->sketch_generic = value != 0;
        return HLO_OK;
    }
    if (!strcmp(name, "sort_segmented")) {
        ctx->sort_segmented = value != 0;
        return HLO_OK;
    }
    if (!strcmp(name, "sort_prefix")) {
        ctx->sort_prefix = value != 0;
        return HLO_OK;
    }
    if (!strcmp(name, "sort_hash_bits")) {  // 0 = automatic
        if (value < 0 || value > 64) return HLO_EINVAL;
        ctx->sort_hash_bits = (int)value;
        return HLO_OK;
    }
    if (!strcmp(name, "pairgen_generic")) {
        ctx->pairgen_generic = value != 0;
        return HLO_OK;
    }
    if (!strcmp(name, "repack")) {
        ctx->repack = value != 0;
        return HLO_OK;
    }
    if (!strcmp(name, "compact")) {
        ctx->compact = value != 0;
        return HLO_OK;
    }
    if (!strcmp(name, "aux")) {
        ctx->aux_out = value != 0;
        return HLO_OK;
    }
    if (!strcmp(name, "d2h")) {
        ctx->d2h = value != 0;
        return HLO_OK;
    }
    if (!strcmp(name, "pairgen_slots_log2")) {
        if (value < 5 || value > 10) FAIL(HLO_ERANGE, "pairgen_slots_log2=%lld outside [5,10]", (long long)value);
        ctx->pairgen_slots_log2 = (int)value;
        ctx->pairgen_slots_auto = 0;
        return HLO_OK;
    }
    FAIL(HLO_EINVAL, "unknown option %s", name);
}

int hlo_score_records(hlo_ctx *ctx, const hlo_reads *R, const char *quals, const hlo_rec *rec, const uint32_t *rec_cluster, uint64_t n,
                      double mismatch_prob, uint32_t min_read_len, const hlo_score **out) {
    if (!ctx || !R || !quals || (!rec && n) || !out) return HLO_EINVAL;
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
    if (!ctx->prob_ready) {  // phred_to_prob (EdgeCalculator.cpp:59-63) for every byte value, the reference's own expression
        double tab[128];
        for (int b = 0; b < 128; ++b) tab[b] = pow(10, -(b - 33) / 10.0);
        CK(ctx->b_prob.ensure(sizeof tab + 2 * 128 * 128 * sizeof(double)));
        CK(cudaMemcpyAsync(ctx->b_prob.p, tab, sizeof tab, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        ctx->prob_ready = true;
    }
    CK(ctx->b_qual.ensure(d.n_bases + 64 + (size_t)d.n_reads + 16));  // 32 bytes of padding on both sides, then a flag per read
    CK(ctx->b_screc.ensure(n * sizeof(hlo_rec) + 32));
    CK(ctx->b_scclu.ensure(n * 4 + 4));
    CK(ctx->b_scout.ensure(n * sizeof(hlo_score) + 16));
    CK(ctx->h_score.ensure(n * sizeof(hlo_score) + 16));
    CK(const_cast<hlo_reads *>(R)->b_coff.ensure(((size_t)d.n_clusters + 1) * 4));
    reset_timing(ctx);
    CK(cudaEventRecord(ctx->ev[ST_BEGIN], s));
    CK(cudaMemcpyAsync(R->b_coff.p, R->h_cluster_off.data(), ((size_t)d.n_clusters + 1) * 4, cudaMemcpyHostToDevice, s));
    uint8_t *d_qual = ctx->b_qual.as<uint8_t>() + 32, *d_qflag = d_qual + ((d.n_bases + 32 + 15) & ~(uint64_t)15);
    CK(cudaMemsetAsync(ctx->b_qual.p, 0, 32, s));
    CK(cudaMemsetAsync(d_qual + d.n_bases, 0, 32, s));
    CK(cudaMemcpyAsync(d_qual, quals, d.n_bases, cudaMemcpyHostToDevice, s));
    if (n) CK(cudaMemcpyAsync(ctx->b_screc.p, rec, n * sizeof(hlo_rec), cudaMemcpyHostToDevice, s));
    if (n && rec_cluster) CK(cudaMemcpyAsync(ctx->b_scclu.p, rec_cluster, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ctx->ev[ST_H2D], s));
    launch_score_table(ctx->b_prob.as<double>(), mismatch_prob, ctx->b_prob.as<double>() + 128, s, &ctx->lc);
    launch_qual_flags(d_qual, R->b_off.as<uint64_t>(), d.n_reads, d_qflag, s, &ctx->lc);
    launch_score_records(d, d_qual, d_qflag, R->b_off.as<uint64_t>(), R->b_coff.as<uint32_t>(), ctx->b_screc.as<hlo_rec>(),
                         rec_cluster ? ctx->b_scclu.as<uint32_t>() : nullptr, n, ctx->b_prob.as<double>() + 128, min_read_len,
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

void *hlo_stream(const hlo_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int hlo_get_timing(const hlo_ctx *ctx, hlo_timing *t) {
    if (!ctx || !t) return HLO_EINVAL;
    *t = ctx->tm;
    return HLO_OK;
}

}  // extern "C"
