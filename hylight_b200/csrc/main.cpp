// hylight-ovlp — host binary of the B200-native short-read overlap stage.
//
// Replaces, per read cluster, the process chain of polyte.tune_params.py:488-529 (preprocessing_denovo)
// and :534-579 (run_sfo):
//     fastq2fasta.py ; minimap2 --sr -X -c -k 21 -w 11 -s 60 -m 30 -n 2 -r 0 -A 4 -B 2 --end-bonus 100 F F
//       | cut -f 1-12 | filter_ovlp_inline.py L 0.98 1 0.8 ; minimap22sfo.py ; sfo2overlaps.py
// and writes the same 13-column `.overlaps` file (and optionally the 8-column SFO file) that
// tools/HaploConduct/src/EdgeCalculator.cpp:561-666 parses — or, with --binary_out, the records themselves
// (hlo_write_binary16; reader patch in INTEGRATION.md §5).  All arithmetic runs in libhlovlp's CUDA kernels through
// the C ABI of include/hlovlp.h; this file only parses text, batches clusters (largest first), drives a
// few worker threads per GPU (clusters are independent: dynamic queue, host-side gather, no NCCL) and writes files.
// A worker owns one context and one set of page-locked input buffers for its whole life: per batch it parses into
// them, calls hlo_submit / hlo_wait (nothing is allocated, and the device is never waited for between stages), and
// writes the cluster files.  There is no CPU code path for the overlap computation.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#include "../../include/hlovlp.h"

namespace {

struct Cluster {
    std::string in1, in2, out, sfo;
    uint64_t bytes = 0;  // size of the input files: an upper bound of the bases, and the batching weight
    std::string err;
    uint64_t n_records = 0;
};

// one cluster's reads as parsed from its files
struct Parsed {
    std::string bases, quals;
    std::vector<uint32_t> lens, ids;
    void clear() { bases.clear(), quals.clear(), lens.clear(), ids.clear(); }
};

struct Options {
    hlo_params P;
    std::vector<int> devices;
    uint64_t batch_bases = 96ull << 20;
    int workers_per_device = 3;
    bool rename = false;
    bool scores = false;          // also write <out>.scored: 13 columns + HaploConduct's overlap score and mismatch rate
    bool binary_out = false;      // <out> holds the records (hlo_write_binary16) instead of text
    double mismatch_prob = 0.0;   // ViralQuasispecies --mismatch
    uint32_t min_read_len = 0;    // ViralQuasispecies --min_read_len
    std::string stats;
    std::vector<Cluster> clusters;
};

// page-locked host array that only grows (hlo_host_alloc)
template <class T>
struct PinVec {
    T *p = nullptr;
    size_t cap = 0;
    bool ensure(size_t n) {
        if (n <= cap) return true;
        if (p) hlo_host_free(p);
        cap = n + n / 4 + 1024;
        p = (T *)hlo_host_alloc(cap * sizeof(T));
        if (!p) cap = 0;
        return p != nullptr;
    }
};

bool read_file(const std::string &path, std::string &data, std::string &err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        err = "cannot open " + path;
        return false;
    }
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    const size_t old = data.size();
    if (sz > 0) {
        data.resize(old + (size_t)sz);
        const size_t got = fread(&data[old], 1, (size_t)sz, f);
        data.resize(old + got);
    }
    const bool ok = !ferror(f);
    fclose(f);
    if (!ok) err = "read error on " + path;
    return ok;
}

// decimal read name (the integers rename_fas.py:48-57 writes); FastqStorage.cpp:109-117 parses them with strtoul
bool parse_id(const char *s, const char *e, uint32_t &id) {
    if (s == e) return false;
    uint64_t v = 0;
    const char *p = s;
    for (; p < e && *p >= '0' && *p <= '9'; ++p) {
        v = v * 10 + (uint64_t)(*p - '0');
        if (v > 0xffffffffull) return false;
    }
    if (p == s) return false;
    if (p - s > 1 && *s == '0') return false;  // strtoul(.., 0) would read a leading 0 as octal
    if (p < e && *p != ' ' && *p != '\t' && *p != '\r') return false;
    id = (uint32_t)v;
    return true;
}

// FASTA (multi-line) or FASTQ (4-line records, fastq2fasta.py:32-40) -> appended to c
bool parse_reads(const std::string &path, Parsed &c, std::string &err, bool rename, bool want_quals, std::string &data) {
    data.clear();
    if (!read_file(path, data, err)) return false;
    const char *p = data.data(), *end = p + data.size();
    auto line_end = [&](const char *q) {
        const char *e = (const char *)memchr(q, '\n', (size_t)(end - q));
        return e ? e : end;
    };
    while (p < end && (*p == '\n' || *p == '\r')) ++p;
    if (p == end) return true;
    const bool fq = *p == '@';
    if (!fq && *p != '>') {
        err = path + ": neither FASTA nor FASTQ";
        return false;
    }
    if (!fq && want_quals) {
        err = path + ": --scores needs FASTQ input (phred qualities)";
        return false;
    }
    while (p < end) {
        if (*p == '\n' || *p == '\r') {
            ++p;
            continue;
        }
        const char *le = line_end(p);
        if (*p != (fq ? '@' : '>')) {
            err = path + ": malformed record header";
            return false;
        }
        uint32_t id = (uint32_t)c.ids.size();
        if (!rename && !parse_id(p + 1, le, id)) {
            err = path + ": read name '" + std::string(p + 1, le) + "' is not a plain decimal integer (use --rename)";
            return false;
        }
        c.ids.push_back(id);
        p = le < end ? le + 1 : end;
        const size_t b0 = c.bases.size();
        if (fq) {
            const char *se = line_end(p);
            const char *s_end = se;
            while (s_end > p && s_end[-1] == '\r') --s_end;
            c.bases.append(p, s_end);
            const size_t slen = (size_t)(s_end - p);
            p = se < end ? se + 1 : end;
            for (int k = 0; k < 2 && p < end; ++k) {  // '+' line and the qualities
                const char *e2 = line_end(p);
                if (k == 1 && want_quals) {
                    const char *q_end = e2;
                    while (q_end > p && q_end[-1] == '\r') --q_end;
                    if ((size_t)(q_end - p) != slen) {
                        err = path + ": quality string and sequence differ in length";
                        return false;
                    }
                    c.quals.append(p, q_end);
                }
                p = e2 < end ? e2 + 1 : end;
            }
        } else {
            while (p < end && *p != '>') {
                const char *se = line_end(p);
                const char *s_end = se;
                while (s_end > p && s_end[-1] == '\r') --s_end;
                c.bases.append(p, s_end);
                p = se < end ? se + 1 : end;
            }
        }
        c.lens.push_back((uint32_t)(c.bases.size() - b0));
    }
    return true;
}

void usage() {
    fprintf(stderr,
            "usage: hylight-ovlp --in reads.fa|fq --out overlaps.txt [--sfo sfoverlaps.out] [options]\n"
            "       hylight-ovlp --p1 X.1.fq --p2 X.2.fq --rename --out overlaps.txt\n"
            "       hylight-ovlp --clusters list.tsv [options]      (lines: in<TAB>out[<TAB>sfo]; in may be 'a.1.fq,a.2.fq')\n"
            "options (defaults = polyte.tune_params.py:500-510):\n"
            "  --min_ovlp_len 60  --min_identity 0.98  --max_overhang 1  --overhang_ratio 0.8\n"
            "  -k 21  -w 11  --min_seeds 2  --min_chain_score 30\n"
            "  --mid_occ 0         occurrence cut-off; 0 = minimap2's default, derived from every cluster's index\n"
            "  --stage_b           the thresholds of HyLight's contig stage (HyLight.py:309-311): --min_ovlp_len 90\n"
            "                      --min_identity 0.99 --max_overhang 2\n"
            "  --rename            name reads 0..N-1 in file order (what rename_fas.py does)\n"
            "  --devices 0,1,..|all   GPUs to use (default all)\n"
            "  --batch_bases N     input bytes per device batch (default 100663296)\n"
            "  --workers_per_device N  host workers (context + stream + pinned buffers each) per GPU; one's kernels run\n"
            "                      while the others parse, upload, download and write files (default 3)\n"
            "  --binary_out        write the records themselves (32-byte header + 16-byte records, include/hlovlp.h\n"
            "                      hlo_write_binary16) instead of text: the hand-off for a HaploConduct with the reader\n"
            "                      patch of INTEGRATION.md\n"
            "  --scores            also compute HaploConduct's phred-weighted overlap score and mismatch rate\n"
            "                      (EdgeCalculator.cpp:67-139) on the device: <out>.scored = 13 columns + the two (text), or\n"
            "                      appended to the binary file; FASTQ input; [--mismatch_prob P] [--min_read_len L]\n"
            "  --stats FILE        JSON-lines per batch\n");
}

bool parse_args(int argc, char **argv, Options &o, std::string &err) {
    hlo_params_default(&o.P);
    Cluster single;
    std::string list, devs = "all";
    bool have_single = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](const char *&dst) {
            if (i + 1 >= argc) return false;
            dst = argv[++i];
            return true;
        };
        const char *v = nullptr;
        if (a == "--rename") o.rename = true;
        else if (a == "--scores") o.scores = true;
        else if (a == "--binary_out") o.binary_out = true;
        else if (a == "--stage_b") o.P.min_ovlp_len = 90, o.P.min_identity = 0.99, o.P.max_overhang = 2;
        else if (a == "-h" || a == "--help") return false;
        else if (!val(v)) {
            err = "missing value for " + a;
            return false;
        } else if (a == "--in") single.in1 = v, have_single = true;
        else if (a == "--p1") single.in1 = v, have_single = true;
        else if (a == "--p2") single.in2 = v, have_single = true;
        else if (a == "--out") single.out = v;
        else if (a == "--sfo") single.sfo = v;
        else if (a == "--clusters") list = v;
        else if (a == "--min_ovlp_len") o.P.min_ovlp_len = atoi(v);
        else if (a == "--min_identity") o.P.min_identity = atof(v);
        else if (a == "--max_overhang") o.P.max_overhang = atoi(v);
        else if (a == "--overhang_ratio") o.P.overhang_ratio = atof(v);
        else if (a == "-k") o.P.k = atoi(v);
        else if (a == "-w") o.P.w = atoi(v);
        else if (a == "--min_seeds") o.P.min_seeds = atoi(v);
        else if (a == "--min_chain_score") o.P.min_chain_score = atoi(v);
        else if (a == "--mid_occ") o.P.mid_occ = atoi(v);
        else if (a == "--devices") devs = v;
        else if (a == "--batch_bases") o.batch_bases = strtoull(v, nullptr, 10);
        else if (a == "--workers_per_device") o.workers_per_device = std::max(1, atoi(v));
        else if (a == "--stats") o.stats = v;
        else if (a == "--mismatch_prob") o.mismatch_prob = atof(v);
        else if (a == "--min_read_len") o.min_read_len = (uint32_t)atoi(v);
        else {
            err = "unknown option " + a;
            return false;
        }
    }
    if (have_single) {
        if (single.out.empty()) {
            err = "--out is required";
            return false;
        }
        o.clusters.push_back(single);
    }
    if (!list.empty()) {
        std::string data;
        if (!read_file(list, data, err)) return false;
        size_t p = 0;
        while (p < data.size()) {
            size_t e = data.find('\n', p);
            if (e == std::string::npos) e = data.size();
            std::string line = data.substr(p, e - p);
            p = e + 1;
            if (line.empty() || line[0] == '#') continue;
            std::vector<std::string> f;
            size_t q = 0;
            while (true) {
                size_t t = line.find('\t', q);
                f.push_back(line.substr(q, t == std::string::npos ? std::string::npos : t - q));
                if (t == std::string::npos) break;
                q = t + 1;
            }
            if (f.size() < 2) {
                err = "cluster list line needs in<TAB>out: " + line;
                return false;
            }
            Cluster c;
            const size_t comma = f[0].find(',');
            c.in1 = f[0].substr(0, comma);
            if (comma != std::string::npos) c.in2 = f[0].substr(comma + 1);
            c.out = f[1];
            if (f.size() > 2) c.sfo = f[2];
            o.clusters.push_back(c);
        }
    }
    if (o.clusters.empty()) {
        err = "no input";
        return false;
    }
    const int n_dev = hlo_device_count();
    if (n_dev <= 0) {
        err = "no CUDA device visible; hylight-ovlp has no CPU fallback";
        return false;
    }
    if (devs == "all") {
        for (int d = 0; d < n_dev; ++d) o.devices.push_back(d);
    } else {
        size_t p = 0;
        while (p <= devs.size()) {
            size_t e = devs.find(',', p);
            if (e == std::string::npos) e = devs.size();
            const std::string tok = devs.substr(p, e - p);
            const size_t dash = tok.find('-');
            const int lo = atoi(tok.c_str()), hi = dash == std::string::npos ? lo : atoi(tok.c_str() + dash + 1);
            for (int d = lo; d <= hi; ++d) {
                if (d < 0 || d >= n_dev) {
                    err = "device " + std::to_string(d) + " not present";
                    return false;
                }
                o.devices.push_back(d);
            }
            p = e + 1;
        }
    }
    return true;
}

// fn(i) for i in [0, n) on up to `threads` host threads
template <class F>
void parallel_for(size_t n, size_t threads, F fn) {
    if (threads <= 1 || n <= 1) {
        for (size_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    const size_t nt = std::min(threads, n);
    for (size_t t = 0; t < nt; ++t)
        th.emplace_back([&] {
            for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto &t : th) t.join();
}

struct Shared {
    Options *o;
    size_t helpers = 1;  // host threads a worker may use to parse and to write the clusters of its batch
    std::vector<std::vector<size_t>> batches;  // cluster indices per batch, largest batch first
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    std::mutex log_mu;
    FILE *stats = nullptr;
};

double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void worker(Shared *S, int device) {
    Options &o = *S->o;
    // the context (CUDA start-up: a few hundred ms on a cold process) comes up while the first batch is parsed
    hlo_ctx *ctx = nullptr;
    int ctx_rc = 0;
    double ctx_ms = 0.0;
    std::thread ctx_thread([&] {
        const auto tc0 = std::chrono::steady_clock::now();
        ctx_rc = hlo_ctx_create(device, &o.P, &ctx);
        if (!ctx_rc) {
            hlo_set_option(ctx, "aux", 0);  // the files need the records and the per-cluster offsets only
            // text: 8-byte records (OHB and the overlap length follow from the read lengths parsed here);
            // binary hand-off: the self-contained 16-byte form
            hlo_set_option(ctx, "compact", o.binary_out ? 1 : 2);
            if (o.scores) {
                hlo_set_option(ctx, "scores", 1);
                hlo_set_score_params(ctx, o.mismatch_prob, o.min_read_len);
            }
        }
        ctx_ms = ms_since(tc0);
    });
    bool ctx_joined = false;
    PinVec<char> p_bases, p_quals;
    PinVec<uint64_t> p_off;
    PinVec<uint32_t> p_coff, p_ids;
    std::vector<Parsed> parsed;
    std::vector<std::string> file_buf;  // one per helper slot: the raw text of the file being parsed
    for (;;) {
        const size_t b = S->next.fetch_add(1);
        if (b >= S->batches.size() || S->failed) break;
        const auto t0 = std::chrono::steady_clock::now();
        const std::vector<size_t> &idx = S->batches[b];
        const size_t nc = idx.size();
        if (parsed.size() < nc) parsed.resize(nc);
        if (file_buf.size() < nc) file_buf.resize(nc);
        // parse the clusters of the batch (in parallel), then lay them out back to back in the pinned buffers
        std::atomic<bool> parse_ok{true};
        parallel_for(nc, S->helpers, [&](size_t j) {
            Cluster &c = o.clusters[idx[j]];
            Parsed &P = parsed[j];
            P.clear();
            if (!(parse_reads(c.in1, P, c.err, o.rename, o.scores, file_buf[j]) &&
                  (c.in2.empty() || parse_reads(c.in2, P, c.err, o.rename, o.scores, file_buf[j]))))
                parse_ok = false;
        });
        std::string err;
        bool ok = parse_ok;
        size_t nb = 0, nr = 0;
        std::vector<size_t> b_start(nc + 1, 0), r_start(nc + 1, 0);
        if (ok) {
            for (size_t j = 0; j < nc; ++j) {
                b_start[j] = nb, r_start[j] = nr;
                nb += parsed[j].bases.size(), nr += parsed[j].ids.size();
            }
            b_start[nc] = nb, r_start[nc] = nr;
            if (nr > 0xfffffff0ull) ok = false, err = "more than 2^32 reads in one batch: lower --batch_bases";
        } else {
            for (size_t ci : idx)
                if (!o.clusters[ci].err.empty()) err = o.clusters[ci].err;
        }
        if (ok && !(p_bases.ensure(nb + 16) && p_off.ensure(nr + 1) && p_coff.ensure(nc + 1) && p_ids.ensure(nr + 1) &&
                    (!o.scores || p_quals.ensure(nb + 16))))
            ok = false, err = "out of page-locked host memory";
        if (ok) {
            parallel_for(nc, S->helpers, [&](size_t j) {
                const Parsed &P = parsed[j];
                memcpy(p_bases.p + b_start[j], P.bases.data(), P.bases.size());
                if (o.scores) memcpy(p_quals.p + b_start[j], P.quals.data(), P.quals.size());
                uint64_t off = b_start[j];
                for (size_t r = 0; r < P.lens.size(); ++r) {
                    p_off.p[r_start[j] + r] = off;
                    p_ids.p[r_start[j] + r] = P.ids[r];
                    off += P.lens[r];
                }
                p_coff.p[j] = (uint32_t)r_start[j];
            });
            p_off.p[nr] = nb, p_coff.p[nc] = (uint32_t)nr;
        }
        const double parse_ms = ms_since(t0);
        if (!ctx_joined) {
            ctx_thread.join();
            ctx_joined = true;
            if (ctx_rc) {
                std::lock_guard<std::mutex> g(S->log_mu);
                fprintf(stderr, "hylight-ovlp: device %d: %s\n", device, hlo_strerror(ctx_rc));
                S->failed = 1;
                return;
            }
        }
        const auto t_call = std::chrono::steady_clock::now();
        hlo_recs out{};
        int rc = 0;
        if (ok) {
            hlo_batch B{};
            B.bases = p_bases.p, B.quals = o.scores ? p_quals.p : nullptr, B.base_off = p_off.p, B.n_reads = (uint32_t)nr;
            B.cluster_off = p_coff.p, B.n_clusters = (uint32_t)nc, B.ids = o.rename ? nullptr : p_ids.p;
            rc = hlo_submit(ctx, &B);
            if (!rc) rc = hlo_wait(ctx, &out);
            if (rc) err = std::string(hlo_strerror(rc)) + ": " + hlo_last_error(ctx);
        }
        const double call_ms = ms_since(t_call);
        const auto t_write = std::chrono::steady_clock::now();
        if (err.empty()) {  // per-cluster files, written in parallel
            std::mutex emu;
            parallel_for(nc, S->helpers, [&](size_t j) {
                Cluster &c = o.clusters[idx[j]];
                const uint64_t r0 = out.cluster_off[j], n = out.cluster_off[j + 1] - r0;
                c.n_records = n;
                const hlo_score *sc = out.score ? out.score + r0 : nullptr;
                int wrc = 0;
                std::vector<hlo_rec> full;   // the 32-byte form, only where a writer needs it
                std::vector<hlo_rec16> r16;  // the 16-byte form for the binary file when the library sent 8-byte records
                std::vector<uint32_t> len_by_id;
                if (out.rec8) {  // names -> lengths (a name is the read's position unless the file carried its own)
                    const Parsed &P = parsed[j];
                    uint32_t mx = 0;
                    for (uint32_t id : P.ids) mx = std::max(mx, id);
                    len_by_id.assign((size_t)mx + 1, 0u);
                    for (size_t r = 0; r < P.ids.size(); ++r) len_by_id[P.ids[r]] = P.lens[r];
                }
                auto need_full = [&]() {
                    if (!full.empty() || !n) return;
                    full.resize(n);
                    for (uint64_t i = 0; i < n; ++i) {
                        if (out.rec8) hlo_rec8_expand(&out.rec8[r0 + i], len_by_id[out.rec8[r0 + i].idA], len_by_id[out.rec8[r0 + i].idB], &full[i]);
                        else if (out.rec16) hlo_rec16_expand(&out.rec16[r0 + i], &full[i]);
                        else full[i] = out.rec[r0 + i];
                    }
                };
                if (o.binary_out) {
                    wrc = hlo_write_binary16(out.rec16 ? out.rec16 + r0 : nullptr, sc, n, c.out.c_str());
                } else if (out.rec8) {
                    wrc = hlo_write_overlaps8(out.rec8 + r0, n, len_by_id.data(), (uint32_t)len_by_id.size(), c.out.c_str());
                } else if (out.rec16) {
                    wrc = hlo_write_overlaps16(out.rec16 + r0, n, c.out.c_str());
                } else {
                    wrc = hlo_write_overlaps(out.rec + r0, n, c.out.c_str());
                }
                if (!wrc && !c.sfo.empty()) {
                    need_full();
                    wrc = hlo_write_sfo(full.data(), n, c.sfo.c_str());
                }
                if (!wrc && sc && !o.binary_out) {
                    need_full();
                    wrc = hlo_write_overlaps_scored(full.data(), sc, n, (c.out + ".scored").c_str());
                }
                if (wrc) {
                    std::lock_guard<std::mutex> g(emu);
                    err = std::string(hlo_strerror(wrc)) + " writing " + c.out;
                }
            });
        }
        const double write_ms = ms_since(t_write);
        hlo_timing tm{};
        hlo_get_timing(ctx, &tm);
        const double wall = ms_since(t0);
        std::lock_guard<std::mutex> g(S->log_mu);
        if (!err.empty()) {
            fprintf(stderr, "hylight-ovlp: batch %zu on device %d failed: %s\n", b, device, err.c_str());
            S->failed = 1;
            break;
        }
        if (S->stats)
            fprintf(S->stats,
                    "{\"batch\": %zu, \"device\": %d, \"clusters\": %zu, \"reads\": %llu, \"bases\": %llu, \"minimizers\": %llu, "
                    "\"seed_hits\": %llu, \"candidates\": %llu, \"accepted\": %llu, \"records\": %llu, \"deferred\": %llu, "
                    "\"overflow_reads\": %llu, \"launches\": %llu, \"reruns\": %llu, \"syncs\": %llu, \"mid_occ\": [%u, %u], "
                    "\"ms\": {\"h2d\": %.3f, \"pack\": %.3f, \"sketch\": %.3f, "
                    "\"sort\": %.3f, \"index\": %.3f, \"pairgen\": %.3f, \"verify\": %.3f, \"dedup\": %.3f, \"score\": %.3f, \"d2h\": %.3f, "
                    "\"device_total\": %.3f, \"wall\": %.3f, \"host_parse\": %.3f, \"host_call\": %.3f, \"host_write\": %.3f, "
                    "\"ctx_create\": %.3f}}\n",
                    b, device, nc, (unsigned long long)tm.n_reads, (unsigned long long)tm.n_bases,
                    (unsigned long long)tm.n_minimizers, (unsigned long long)tm.n_seed_hits, (unsigned long long)tm.n_candidates,
                    (unsigned long long)tm.n_accepted, (unsigned long long)tm.n_records, (unsigned long long)tm.n_deferred,
                    (unsigned long long)tm.n_overflow_reads, (unsigned long long)tm.total_launches, (unsigned long long)tm.n_reruns,
                    (unsigned long long)tm.n_syncs, tm.mid_occ_lo, tm.mid_occ_hi, tm.h2d_ms, tm.pack_ms,
                    tm.sketch_ms, tm.sort_ms, tm.index_ms, tm.pairgen_ms, tm.verify_ms, tm.dedup_ms, tm.score_ms, tm.d2h_ms, tm.total_ms, wall,
                    parse_ms, call_ms, write_ms, ctx_ms);
        ctx_ms = 0.0;  // (reported with the worker's first batch)
    }
    if (!ctx_joined) ctx_thread.join();
    // no hlo_ctx_destroy: the process is about to end, and releasing gigabytes of device and pinned memory buffer by
    // buffer (then the CUDA context) costs about a second that the operating system does not need
    (void)ctx;
}

uint64_t file_size(const std::string &p) {
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) return 0;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fclose(f);
    return n > 0 ? (uint64_t)n : 0;
}

}  // namespace

int main(int argc, char **argv) {
    const auto t_main = std::chrono::steady_clock::now();
    Options o;
    std::string err;
    if (!parse_args(argc, argv, o, err)) {
        if (!err.empty()) fprintf(stderr, "hylight-ovlp: %s\n", err.c_str());
        usage();
        return 2;
    }
    Shared S;
    S.o = &o;
    // batches: clusters largest first (the long jobs start early, the short ones fill the tail), packed
    // until ~batch_bases (file size is an upper bound of the base count); every cluster has its own output file, so
    // the order of the list does not matter
    {
        for (Cluster &c : o.clusters) c.bytes = file_size(c.in1) + (c.in2.empty() ? 0 : file_size(c.in2));
        std::vector<size_t> order(o.clusters.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return o.clusters[a].bytes > o.clusters[b].bytes; });
        std::vector<size_t> cur;
        uint64_t acc = 0;
        for (size_t i : order) {
            const uint64_t sz = o.clusters[i].bytes;
            if (!cur.empty() && acc + sz > o.batch_bases) {
                S.batches.push_back(cur);
                cur.clear(), acc = 0;
            }
            cur.push_back(i), acc += sz;
        }
        if (!cur.empty()) S.batches.push_back(cur);
    }
    if (!o.stats.empty()) {
        S.stats = o.stats == "-" ? stderr : fopen(o.stats.c_str(), "w");
        if (!S.stats) {
            fprintf(stderr, "hylight-ovlp: cannot open %s\n", o.stats.c_str());
            return 1;
        }
    }
    std::vector<std::thread> th;
    const size_t n_workers = std::min(o.devices.size() * (size_t)o.workers_per_device, S.batches.size());
    S.helpers = std::max<size_t>(1, std::thread::hardware_concurrency() / std::max<size_t>(1, n_workers));
    for (size_t i = 0; i < n_workers; ++i) th.emplace_back(worker, &S, o.devices[i % o.devices.size()]);
    for (auto &t : th) t.join();
    if (S.stats) {  // one closing line: the whole run as the process saw it (CUDA start-up included, process teardown not)
        uint64_t n_rec = 0;
        for (const Cluster &c : o.clusters) n_rec += c.n_records;
        fprintf(S.stats, "{\"summary\": true, \"clusters\": %zu, \"batches\": %zu, \"workers\": %zu, \"records\": %llu, \"wall_ms\": %.3f}\n",
                o.clusters.size(), S.batches.size(), n_workers, (unsigned long long)n_rec, ms_since(t_main));
    }
    if (S.stats && S.stats != stderr) fclose(S.stats);
    fflush(nullptr);
    _exit(S.failed ? 1 : 0);  // every output file is closed (and renamed into place) by now; skip the CUDA teardown
}
