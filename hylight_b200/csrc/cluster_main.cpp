// hylight-cluster — the size-capped, order-dependent union-find that turns the all-vs-all overlaps of the whole
// short-read set into read clusters, and the per-cluster FASTQ split (host half).
//
// One native binary for three reference scripts, same files in, same clusters out:
//   bin_pointer_limited_filechunks_shortpath2.py   (HyLight.py:218)  overlaps (PAF, `sort -k12 -nr` order) -> pointer forest
//   getclusters.py                                 (HyLight.py:219)  forest -> clusters of >= 20 reads
//   get_fq_cluster.py                              (HyLight.py:220)  interleaved FASTQ -> <outdir>/<cluster>/<cluster>.{1,2}.fq
//
// The reference's result depends on the ORDER in which overlaps are applied and on how it batches them, so the
// batching is restated, not "improved" (bin_pointer_limited_filechunks_shortpath2.py):
//   * the file is cut into chunks of >= 2 600 000 bytes ending at a line end (chunkify, :31-42); `threads` chunks form a round;
//   * at the start of a round every line of its chunks is FILTERED against the state at that moment: kept iff the two
//     reads are in different clusters and the sizes add up to LESS than the cap (getchunkfile, :96-109 — forked workers
//     see the parent's state as of the fork);
//   * the kept lines are then applied in file order (clusteralgorithm, :73-93): both heads are looked up again, and the
//     clusters merge iff they still differ and the sizes add up to AT MOST the cap; the head reached by the shorter
//     pointer path is hung under the other one (ties: the first read's head goes under the second's), the surviving
//     cluster number is that of the new head; no path compression (the reference has it commented out, :60-62);
//   * leftover Chunkfile_* of an earlier round that the reference re-reads in a short last round are lines it has
//     already applied: sizes only grow, so applying them again changes nothing — not restated.
// A read's name is column 1 / column 6 minus its last two characters ("/1", "/2": mates share a cluster).
//
//   hylight-cluster --paf shortr2.paf --readnames readnames.txt --size 15000 --threads 48
//                   [--json PREFIX]            PREFIX_clusters_grouped.json + PREFIX_clustersizes.json (getclusters.py's outputs)
//                   [--fastq reads.fq --outdir DIR]   the split FASTQ files (get_fq_cluster.py)
//                   [--min_cluster 20] [--chunk_bytes 2600000] [--keep_tail]
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

[[noreturn]] void die(const char *fmt, const std::string &a = "") {
    fprintf(stderr, "hylight-cluster: ");
    fprintf(stderr, fmt, a.c_str());
    fprintf(stderr, "\n");
    exit(1);
}

struct Mapped {
    const char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    explicit Mapped(const std::string &path) {
        fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) die("cannot open %s", path);
        struct stat st;
        if (fstat(fd, &st)) die("cannot stat %s", path);
        n = (size_t)st.st_size;
        if (n) {
            void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) die("cannot map %s", path);
            p = (const char *)m;
        }
    }
    ~Mapped() {
        if (p) munmap((void *)p, n);
        if (fd >= 0) close(fd);
    }
};

struct Forest {
    std::vector<std::string> names;
    std::unordered_map<std::string_view, uint32_t> index;
    std::vector<int32_t> parent;   // -1: a head
    std::vector<uint32_t> cid;     // cluster number of a head (1-based line number of the read that started it)
    std::vector<uint64_t> csize;   // size by cluster number (live numbers only)

    void load(const std::string &path) {
        std::ifstream f(path);
        if (!f) die("cannot open %s", path);
        std::string line;
        while (std::getline(f, line)) {
            while (!line.empty() && (line.back() == '\n' || line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
            names.push_back(line);
        }
        index.reserve(names.size() * 2);
        parent.assign(names.size(), -1);
        cid.resize(names.size());
        csize.assign(names.size() + 1, 1);
        // a name listed twice keeps its LAST number, like the dict the reference fills line by line (:115-120)
        for (uint32_t i = 0; i < names.size(); ++i) index[std::string_view(names[i])] = i, cid[i] = i + 1;
    }
    uint32_t find(std::string_view name) const {
        auto it = index.find(name);
        if (it == index.end()) die("read %s of the overlap file is not in the read-name list", std::string(name));
        return it->second;
    }
    uint32_t head(uint32_t r, uint32_t *pathlen = nullptr) const {
        uint32_t n = 1;
        while (parent[r] >= 0) r = (uint32_t)parent[r], ++n;
        if (pathlen) *pathlen = n;
        return r;
    }
};

// the two read names of a PAF line: columns 1 and 6 without their last two characters
bool line_reads(std::string_view line, std::string_view *a, std::string_view *b) {
    // rstrip() first, as the filter does
    while (!line.empty() && (line.back() == ' ' || line.back() == '\t' || line.back() == '\r' || line.back() == '\n')) line.remove_suffix(1);
    size_t pos = 0;
    std::string_view col[6];
    for (int c = 0; c < 6; ++c) {
        if (pos > line.size()) return false;
        size_t e = line.find('\t', pos);
        if (e == std::string_view::npos) e = line.size();
        col[c] = line.substr(pos, e - pos);
        pos = e + 1;
    }
    *a = col[0].substr(0, col[0].size() >= 2 ? col[0].size() - 2 : 0);
    *b = col[5].substr(0, col[5].size() >= 2 ? col[5].size() - 2 : 0);
    return true;
}

struct Chunk {
    size_t start, size;
};
// chunkify (:31-42): seek chunk_bytes ahead, finish the line, cut; the last chunk may reach past the end of the file
std::vector<Chunk> chunkify(const Mapped &m, size_t chunk_bytes) {
    std::vector<Chunk> out;
    size_t end = 0;
    for (;;) {
        const size_t start = end;
        size_t p = start + chunk_bytes;
        if (p < m.n) {
            const void *nl = memchr(m.p + p, '\n', m.n - p);
            p = nl ? (size_t)((const char *)nl - m.p) + 1 : m.n;
        }
        end = p;
        out.push_back({start, end - start});
        if (end > m.n) break;
    }
    return out;
}

template <class F>
void for_lines(const Mapped &m, const Chunk &c, F &&f) {
    size_t p = std::min(c.start, m.n);
    const size_t e = std::min(c.start + c.size, m.n);
    while (p < e) {
        const void *nl = memchr(m.p + p, '\n', e - p);
        const size_t q = nl ? (size_t)((const char *)nl - m.p) : e;
        f(std::string_view(m.p + p, q - p));
        p = q + 1;
    }
}

struct Pair {
    uint32_t a, b;
};

void json_string(std::string &o, const std::string &s) {
    o += '"';
    for (unsigned char ch : s) {
        if (ch == '"' || ch == '\\') o += '\\', o += (char)ch;
        else if (ch < 0x20) {
            char b[8];
            snprintf(b, sizeof b, "\\u%04x", ch);
            o += b;
        } else o += (char)ch;
    }
    o += '"';
}

}  // namespace

int main(int argc, char **argv) {
    std::string paf, readnames, size_arg, json_prefix, fastq, outdir;
    int threads = 1;
    bool keep_tail = false;
    uint64_t min_cluster = 20, chunk_bytes = 2600000;
    for (int i = 1; i < argc; ++i) {
        const std::string k = argv[i];
        auto val = [&]() -> std::string {
            if (i + 1 >= argc) die("missing value after %s", k);
            return argv[++i];
        };
        if (k == "--paf") paf = val();
        else if (k == "--readnames") readnames = val();
        else if (k == "--size") size_arg = val();
        else if (k == "--threads") threads = std::max(1, atoi(val().c_str()));
        else if (k == "--json") json_prefix = val();
        else if (k == "--fastq") fastq = val();
        else if (k == "--outdir") outdir = val();
        else if (k == "--min_cluster") min_cluster = strtoull(val().c_str(), nullptr, 10);
        else if (k == "--keep_tail") keep_tail = true;
        else if (k == "--chunk_bytes") chunk_bytes = strtoull(val().c_str(), nullptr, 10);
        else die("unknown option %s", k);
    }
    if (paf.empty() || readnames.empty() || size_arg.empty()) die("usage: hylight-cluster --paf F --readnames F --size N|inf [--threads T] [--json PREFIX] [--fastq F --outdir D]");
    if (fastq.empty() != outdir.empty()) die("--fastq and --outdir go together");
    if (chunk_bytes == 0) die("--chunk_bytes must be positive");
    const bool unlimited = size_arg == "inf";
    const uint64_t maxsize = unlimited ? UINT64_MAX : strtoull(size_arg.c_str(), nullptr, 10);

    Forest F;
    F.load(readnames);
    Mapped M(paf);
    const std::vector<Chunk> chunks = chunkify(M, (size_t)chunk_bytes);

    uint64_t lines_kept = 0, merges = 0;
    for (size_t s0 = 0; s0 < chunks.size(); s0 += (size_t)threads) {
        const size_t s1 = std::min(chunks.size(), s0 + (size_t)threads);
        // filter: every chunk of the round against the state as it is now (read-only: one thread per chunk)
        std::vector<std::vector<Pair>> kept(s1 - s0);
        std::vector<std::thread> pool;
        std::atomic<int> bad{0};
        for (size_t s = s0; s < s1; ++s)
            pool.emplace_back([&, s]() {
                std::vector<Pair> &out = kept[s - s0];
                for_lines(M, chunks[s], [&](std::string_view line) {
                    std::string_view a, b;
                    if (!line_reads(line, &a, &b)) {
                        bad = 1;
                        return;
                    }
                    const uint32_t ra = F.find(a), rb = F.find(b);
                    const uint32_t ca = F.cid[F.head(ra)], cb = F.cid[F.head(rb)];
                    if (ca != cb && (unlimited || F.csize[ca] + F.csize[cb] < maxsize)) out.push_back({ra, rb});
                });
            });
        for (auto &t : pool) t.join();
        if (bad) die("%s: a line with fewer than 6 tab-separated columns", paf);
        // apply, in file order
        for (const auto &v : kept)
            for (const Pair &p : v) {
                ++lines_kept;
                uint32_t l1, l2;
                const uint32_t h1 = F.head(p.a, &l1), h2 = F.head(p.b, &l2);
                const uint32_t c1 = F.cid[h1], c2 = F.cid[h2];
                if (c1 == c2) continue;
                const uint64_t sz = F.csize[c1] + F.csize[c2];
                if (!unlimited && sz > maxsize) continue;
                ++merges;
                if (l2 < l1) F.parent[h2] = (int32_t)h1, F.csize[c1] = sz, F.csize[c2] = 0;
                else F.parent[h1] = (int32_t)h2, F.csize[c2] = sz, F.csize[c1] = 0;
            }
    }

    // getclusters.py: cluster number of every read, clusters of at least min_cluster reads
    const uint32_t n = (uint32_t)F.names.size();
    std::vector<uint32_t> of(n);
    for (uint32_t i = 0; i < n; ++i) of[i] = F.cid[F.head(i)];
    std::vector<std::vector<uint32_t>> members(n + 1);
    std::vector<uint32_t> clustered;
    for (uint32_t i = 0; i < n; ++i)
        if (F.index[std::string_view(F.names[i])] == i && F.csize[of[i]] >= min_cluster) clustered.push_back(i);
    // getclusters.py:55-61 hands the reads to its workers in `threads` slices of len // threads reads each: the last
    // len % threads reads (in read-name order) fall off the end and appear in no cluster.  Same here unless --keep_tail.
    uint64_t n_tail_dropped = 0;
    if (!keep_tail) {
        const size_t keep = clustered.size() / (size_t)threads * (size_t)threads;
        n_tail_dropped = clustered.size() - keep;
        clustered.resize(keep);
    }
    for (uint32_t i : clustered) members[of[i]].push_back(i);
    std::vector<uint8_t> in_cluster(n, 0);
    for (uint32_t i : clustered) in_cluster[i] = 1;
    uint64_t n_clusters = 0, n_clustered = 0;
    for (uint32_t c = 1; c <= n; ++c)
        if (!members[c].empty()) ++n_clusters, n_clustered += members[c].size();

    if (!json_prefix.empty()) {
        std::string o = "{";
        bool first = true;
        for (uint32_t c = 1; c <= n; ++c) {
            if (members[c].empty()) continue;
            if (!first) o += ", ";
            first = false;
            o += '"' + std::to_string(c) + "\": [";
            for (size_t j = 0; j < members[c].size(); ++j) {
                if (j) o += ", ";
                json_string(o, F.names[members[c][j]]);
            }
            o += "]";
        }
        o += "}";
        std::ofstream(json_prefix + "_clusters_grouped.json") << o;
        std::string z = "{";
        first = true;
        for (uint32_t c = 1; c <= n; ++c) {
            if (F.csize[c] == 0) continue;
            if (!first) z += ", ";
            first = false;
            z += '"' + std::to_string(c) + "\": " + std::to_string(F.csize[c]);
        }
        z += "}";
        std::ofstream(json_prefix + "_clustersizes.json") << z;
    }

    // get_fq_cluster.py: 4-line records; the read is the token between the last two of '@' '/' in the header, mate 1 iff the
    // header ends in "/1"
    uint64_t n_records = 0;
    if (!fastq.empty()) {
        if (mkdir(outdir.c_str(), 0777) && errno != EEXIST) die("cannot create %s", outdir);
        std::vector<std::string> buf1(n + 1), buf2(n + 1);
        std::vector<uint8_t> made(n + 1, 0);  // bit 0 / 1: the .1.fq / .2.fq of the cluster exists already (append from then on)
        auto flush = [&](uint32_t c, int mate) {
            const std::string d = outdir + "/" + std::to_string(c);
            if (!made[c] && mkdir(d.c_str(), 0777) && errno != EEXIST) die("cannot create %s", d);
            std::string &b = mate == 1 ? buf1[c] : buf2[c];
            const std::string path = d + "/" + std::to_string(c) + (mate == 1 ? ".1.fq" : ".2.fq");
            std::ofstream f(path, std::ios::binary | ((made[c] >> (mate - 1)) & 1 ? std::ios::app : std::ios::trunc));
            if (!f) die("cannot write %s", path);
            f.write(b.data(), (std::streamsize)b.size());
            made[c] |= (uint8_t)(1 << (mate - 1));
            b.clear();
            b.shrink_to_fit();
        };
        Mapped Q(fastq);
        size_t p = 0;
        while (p < Q.n) {
            size_t e = p;
            int nl = 0;
            size_t h_end = p;
            while (e < Q.n && nl < 4) {
                const void *x = memchr(Q.p + e, '\n', Q.n - e);
                e = x ? (size_t)((const char *)x - Q.p) + 1 : Q.n;
                if (++nl == 1) h_end = e;
            }
            std::string_view header(Q.p + p, h_end - p);
            std::string_view h = header;
            if (!h.empty() && h.back() == '\n') h.remove_suffix(1);
            // re.split('[@/]', header)[-2]
            const size_t last = h.find_last_of("@/");
            if (last != std::string_view::npos) {
                const std::string_view before = h.substr(0, last);
                const size_t prev = before.find_last_of("@/");
                const std::string_view name = prev == std::string_view::npos ? before : before.substr(prev + 1);
                auto it = F.index.find(name);
                if (it != F.index.end()) {
                    const uint32_t c = of[it->second];
                    if (in_cluster[it->second]) {
                        const bool mate1 = h.size() >= 2 && h.substr(h.size() - 2) == "/1";
                        std::string &b = (mate1 ? buf1 : buf2)[c];
                        b.append(Q.p + p, e - p);
                        if (b.size() > (4u << 20)) flush(c, mate1 ? 1 : 2);
                        ++n_records;
                    }
                }
            }
            p = e;
        }
        for (uint32_t c = 1; c <= n; ++c) {
            if (members[c].empty()) continue;
            flush(c, 1), flush(c, 2);  // (both files exist for every cluster, empty or not, as in the reference)
        }
    }
    printf("{\"reads\": %u, \"chunks\": %zu, \"rounds\": %zu, \"lines_kept\": %llu, \"merges\": %llu, \"clusters\": %llu, \"clustered_reads\": %llu, "
           "\"fastq_records\": %llu, \"tail_reads_dropped\": %llu}\n",
           n, chunks.size(), (chunks.size() + (size_t)threads - 1) / (size_t)threads, (unsigned long long)lines_kept, (unsigned long long)merges,
           (unsigned long long)n_clusters, (unsigned long long)n_clustered, (unsigned long long)n_records, (unsigned long long)n_tail_dropped);
    return 0;
}
