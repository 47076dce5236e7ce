// emit.cpp — host-side text emitters of libhlovlp: byte-identical formatting to the reference scripts.
//   SFO line      : minimap22sfo.py:73          "idA idB ori OHA OHB OLA OLB K"
//   .overlaps line: sfo2overlaps.py:150-200     13 tab-separated fields, PERC1 = python3 round() (half-even)
// Records arrive sorted as sfo2overlaps.py:53 sorts its temporary file; the final `uniq`
// (sfo2overlaps.py:107) is applied while writing.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hlovlp.h"
#include "common.cuh"

namespace {

int format_overlap(const hlo_rec *r, char *buf, size_t cap) {
    const long long OHA = r->OHA, OHB = r->OHB, OLA = r->OLA, OLB = r->OLB;
    const long long ovlen = OLA < OLB ? OLA : OLB;
    long long readlenA, readlenB, pos1;
    uint32_t id1, id2;
    const char ori = r->ori ? '-' : '+';
    char ori1, ori2;
    if (OHA >= 0) {  // read A is first (sfo2overlaps.py:162-175)
        if (OHB >= 0) readlenA = OLA + OHA, readlenB = OLB + OHB;
        else readlenA = OLA + OHA - OHB, readlenB = OLB;
        id1 = r->idA, id2 = r->idB, pos1 = OHA, ori1 = '+', ori2 = ori;
    } else {  // read B is first (:176-189)
        if (OHB >= 0) readlenA = OLA, readlenB = -OHA + OLB + OHB;
        else readlenA = OLA - OHB, readlenB = -OHA + OLB;
        id1 = r->idB, id2 = r->idA, pos1 = -OHA, ori1 = ori, ori2 = '+';
    }
    const long long den = readlenA < readlenB ? readlenA : readlenB;
    if (den <= 0) return -1;  // sfo2overlaps.py:198 asserts minreadlen > 0
    // min(round(100*ovlen/minreadlen), 100): python3 round() is half-to-even on the (exact enough) quotient
    const long long num = 100 * ovlen;
    long long perc = num / den;
    const long long rem = num % den;
    if (2 * rem > den || (2 * rem == den && (perc & 1))) ++perc;
    if (perc > 100) perc = 100;
    return snprintf(buf, cap, "%u\t%u\t%lld\t-\t-\t%c\t%c\t%lld\t-\t%lld\t-\ts\ts", id1, id2, pos1, ori1, ori2, perc, ovlen);
}

// decimal digits of v at p; returns the end
inline char *put_u64(char *p, unsigned long long v) {
    char tmp[24];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

// the same 13 fields as format_overlap, without snprintf (the files hold millions of lines); returns the length or -1
int format_overlap_fast(const hlo_rec *r, char *buf) {
    const long long OHA = r->OHA, OHB = r->OHB, OLA = r->OLA, OLB = r->OLB;
    const long long ovlen = OLA < OLB ? OLA : OLB;
    long long readlenA, readlenB, pos1;
    uint32_t id1, id2;
    const char ori = r->ori ? '-' : '+';
    char ori1, ori2;
    if (OHA >= 0) {
        if (OHB >= 0) readlenA = OLA + OHA, readlenB = OLB + OHB;
        else readlenA = OLA + OHA - OHB, readlenB = OLB;
        id1 = r->idA, id2 = r->idB, pos1 = OHA, ori1 = '+', ori2 = ori;
    } else {
        if (OHB >= 0) readlenA = OLA, readlenB = -OHA + OLB + OHB;
        else readlenA = OLA - OHB, readlenB = -OHA + OLB;
        id1 = r->idB, id2 = r->idA, pos1 = -OHA, ori1 = ori, ori2 = '+';
    }
    const long long den = readlenA < readlenB ? readlenA : readlenB;
    if (den <= 0) return -1;
    const long long num = 100 * ovlen;
    long long perc = num / den;
    const long long rem = num % den;
    if (2 * rem > den || (2 * rem == den && (perc & 1))) ++perc;
    if (perc > 100) perc = 100;
    char *p = buf;
    p = put_u64(p, id1), *p++ = '\t';
    p = put_u64(p, id2), *p++ = '\t';
    p = put_u64(p, (unsigned long long)pos1);
    memcpy(p, "\t-\t-\t", 5), p += 5;
    *p++ = ori1, *p++ = '\t', *p++ = ori2, *p++ = '\t';
    p = put_u64(p, (unsigned long long)perc);
    memcpy(p, "\t-\t", 3), p += 3;
    p = put_u64(p, (unsigned long long)ovlen);
    memcpy(p, "\t-\ts\ts", 6), p += 6;
    *p = 0;
    return (int)(p - buf);
}

}  // namespace

extern "C" {

int hlo_format_overlap_line(const hlo_rec *rec, char *buf, size_t cap) {
    if (!rec || !buf || cap < 128) return HLO_EINVAL;
    return format_overlap(rec, buf, cap);
}

int hlo_write_sfo(const hlo_rec *rec, uint64_t n, const char *path) {
    if ((!rec && n) || !path) return HLO_EINVAL;
    std::string tmp = std::string(path) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "w");
    if (!f) return HLO_EIO;
    for (uint64_t i = 0; i < n; ++i) {
        const hlo_rec &r = rec[i];
        // undo the numeric flip where minimap22sfo.py's string order of the ids differs
        uint32_t idA = r.idA, idB = r.idB;
        long long OHA = r.OHA, OHB = r.OHB;
        uint32_t OLA = r.OLA, OLB = r.OLB;
        if (hlo::name_less(idB, idA)) {
            idA = r.idB, idB = r.idA;
            if (r.ori) OHA = r.OHB, OHB = r.OHA;
            else OHA = -(long long)r.OHA, OHB = -(long long)r.OHB;
            OLA = r.OLB, OLB = r.OLA;
        }
        if (fprintf(f, "%u\t%u\t%c\t%lld\t%lld\t%u\t%u\t%u\n", idA, idB, r.ori ? 'I' : 'N', OHA, OHB, OLA, OLB, r.K) < 0) {
            fclose(f);
            remove(tmp.c_str());
            return HLO_EIO;
        }
    }
    if (fclose(f) != 0 || rename(tmp.c_str(), path) != 0) {
        remove(tmp.c_str());
        return HLO_EIO;
    }
    return HLO_OK;
}

static int write_overlaps_impl(const hlo_rec *rec, const hlo_score *score, uint64_t n, const char *path);

// 16-byte records are expanded block-wise and written through the 32-byte code
static int write16_blocks(const hlo_rec16 *rec, uint64_t n, const char *path, bool sfo) {
    if ((!rec && n) || !path) return HLO_EINVAL;
    std::string all;
    std::vector<hlo_rec> tmp((size_t)n);
    for (uint64_t i = 0; i < n; ++i) hlo_rec16_expand(&rec[i], &tmp[(size_t)i]);
    return sfo ? hlo_write_sfo(tmp.data(), n, path) : write_overlaps_impl(tmp.data(), nullptr, n, path);
}
int hlo_write_overlaps16(const hlo_rec16 *rec, uint64_t n, const char *path) { return write16_blocks(rec, n, path, false); }
int hlo_write_sfo16(const hlo_rec16 *rec, uint64_t n, const char *path) { return write16_blocks(rec, n, path, true); }

int hlo_write_overlaps(const hlo_rec *rec, uint64_t n, const char *path) { return write_overlaps_impl(rec, nullptr, n, path); }

int hlo_write_overlaps_scored(const hlo_rec *rec, const hlo_score *score, uint64_t n, const char *path) {
    if (!score && n) return HLO_EINVAL;
    return write_overlaps_impl(rec, score, n, path);
}

static int write_overlaps_impl(const hlo_rec *rec, const hlo_score *score, uint64_t n, const char *path) {
    if ((!rec && n) || !path) return HLO_EINVAL;
    std::string tmp = std::string(path) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "w");
    if (!f) return HLO_EIO;
    static const size_t BUF = 4 << 20;
    std::string out;
    out.reserve(BUF + 256);
    char line[160], prev[160];
    prev[0] = 0;
    bool ok = true;
    for (uint64_t i = 0; i < n && ok; ++i) {
        const int len = format_overlap_fast(&rec[i], line);
        if (len < 0) {
            ok = false;
            break;
        }
        if (i && strcmp(line, prev) == 0) continue;  // `uniq`
        out.append(line, (size_t)len);
        if (score) {  // two more columns: overlap score and mismatch rate, round-trip precision
            char sc[64];
            const int sl = snprintf(sc, sizeof sc, "\t%.17g\t%.17g", score[i].score, score[i].mismatch_rate);
            out.append(sc, (size_t)sl);
        }
        out.push_back('\n');
        memcpy(prev, line, (size_t)len + 1);
        if (out.size() >= BUF) {
            ok = fwrite(out.data(), 1, out.size(), f) == out.size();
            out.clear();
        }
    }
    if (ok && !out.empty()) ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    if (fclose(f) != 0) ok = false;
    if (!ok || rename(tmp.c_str(), path) != 0) {  // never leave a partial file behind
        remove(tmp.c_str());
        return HLO_EIO;
    }
    return HLO_OK;
}

}  // extern "C"
