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

// decimal digits of v at p; returns the end (two digits per division: the files hold hundreds of millions of numbers)
static const char DIGIT_PAIRS[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960"
    "616263646566676869707172737475767778798081828384858687888990919293949596979899";
inline char *put_u64(char *p, unsigned long long v) {
    char tmp[24];
    int n = 24;
    while (v >= 100) {
        const unsigned r = (unsigned)(v % 100) * 2;
        v /= 100;
        tmp[--n] = DIGIT_PAIRS[r + 1], tmp[--n] = DIGIT_PAIRS[r];
    }
    if (v >= 10) tmp[--n] = DIGIT_PAIRS[v * 2 + 1], tmp[--n] = DIGIT_PAIRS[v * 2];
    else tmp[--n] = (char)('0' + v);
    memcpy(p, tmp + n, (size_t)(24 - n));
    return p + (24 - n);
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

// the 13-column file from any record form: get(i, r) yields record i (false: malformed input).  Lines are formatted
// into one buffer and leave with a single write; the file appears under its name only when it is complete.
template <class Get>
int write_overlaps_from(uint64_t n, const hlo_score *score, const char *path, Get get) {
    if (!path) return HLO_EINVAL;
    std::string tmp = std::string(path) + ".tmp";
    std::vector<char> out;
    out.resize((size_t)n * (score ? 120 : 72) + 64);  // ids < 2^32: a line is at most 10+10+5+5+3+5 digits + 25 fixed characters
    char *w = out.data();
    char prev[160];
    int prev_len = -1;
    hlo_rec r;
    for (uint64_t i = 0; i < n; ++i) {
        if (!get(i, r)) return HLO_EINVAL;
        char *line = w;
        const int len = format_overlap_fast(&r, line);
        if (len < 0) return HLO_EINVAL;
        if (len == prev_len && memcmp(line, prev, (size_t)len) == 0) continue;  // `uniq` (sfo2overlaps.py:107)
        memcpy(prev, line, (size_t)len), prev_len = len;
        w += len;
        if (score) w += snprintf(w, 64, "\t%.17g\t%.17g", score[i].score, score[i].mismatch_rate);  // round-trip precision
        *w++ = '\n';
    }
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return HLO_EIO;
    const size_t bytes = (size_t)(w - out.data());
    bool ok = bytes == 0 || fwrite(out.data(), 1, bytes, f) == bytes;
    if (fclose(f) != 0) ok = false;
    if (!ok || rename(tmp.c_str(), path) != 0) {  // never leave a partial file behind
        remove(tmp.c_str());
        return HLO_EIO;
    }
    return HLO_OK;
}

int write_overlaps_impl(const hlo_rec *rec, const hlo_score *score, uint64_t n, const char *path) {
    if (!rec && n) return HLO_EINVAL;
    return write_overlaps_from(n, score, path, [&](uint64_t i, hlo_rec &r) {
        r = rec[i];
        return true;
    });
}

}  // namespace

extern "C" {

int hlo_pack2b(const char *bases, const uint64_t *base_off, uint32_t n_reads, uint32_t *words, uint64_t *n_words, uint32_t *read_len,
               uint64_t *amb, uint64_t amb_cap, uint64_t *n_amb) {
    if (!bases || !base_off || !words || !read_len || !n_words || !n_amb) return HLO_EINVAL;
    static uint8_t lut[256];
    static bool init = false;
    if (!init) {  // (idempotent: concurrent first calls write the same bytes)
        memset(lut, 4, sizeof lut);
        lut['A'] = lut['a'] = 0, lut['C'] = lut['c'] = 1, lut['G'] = lut['g'] = 2, lut['T'] = lut['t'] = lut['U'] = lut['u'] = 3;
        init = true;
    }
    uint64_t w = 0, na = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        if (base_off[r + 1] < base_off[r]) return HLO_EINVAL;
        const uint64_t L = base_off[r + 1] - base_off[r];
        if (L > 0xffffffffull) return HLO_ERANGE;
        read_len[r] = (uint32_t)L;
        const uint8_t *s = reinterpret_cast<const uint8_t *>(bases) + base_off[r];
        for (uint64_t i0 = 0; i0 < L; i0 += 16) {
            const uint64_t n = L - i0 < 16 ? L - i0 : 16;
            uint32_t v = 0;
            for (uint64_t i = 0; i < n; ++i) {
                const uint32_t c = lut[s[i0 + i]];
                if (c > 3) {
                    if (amb && na < amb_cap) amb[na] = (uint64_t)r << 32 | (i0 + i);
                    ++na;
                } else {
                    v |= c << (2 * i);
                }
            }
            words[w++] = v;
        }
    }
    *n_words = w, *n_amb = na;
    return na > amb_cap ? HLO_ERANGE : HLO_OK;
}

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

int hlo_write_overlaps16(const hlo_rec16 *rec, uint64_t n, const char *path) {
    if (!rec && n) return HLO_EINVAL;
    return write_overlaps_from(n, nullptr, path, [&](uint64_t i, hlo_rec &r) {
        hlo_rec16_expand(&rec[i], &r);
        return true;
    });
}
int hlo_write_sfo16(const hlo_rec16 *rec, uint64_t n, const char *path) {
    if ((!rec && n) || !path) return HLO_EINVAL;
    std::vector<hlo_rec> tmp((size_t)n);
    for (uint64_t i = 0; i < n; ++i) hlo_rec16_expand(&rec[i], &tmp[(size_t)i]);
    return hlo_write_sfo(tmp.data(), n, path);
}

// 8-byte records carry (idA, idB, OHA, K, ori); the other fields follow from the two read lengths (hlo_rec8_expand)
int hlo_write_overlaps8(const hlo_rec8 *rec, uint64_t n, const uint32_t *len_by_id, uint32_t n_ids, const char *path) {
    if ((!rec && n) || (!len_by_id && n)) return HLO_EINVAL;
    return write_overlaps_from(n, nullptr, path, [&](uint64_t i, hlo_rec &r) {
        if (rec[i].idA >= n_ids || rec[i].idB >= n_ids) return false;
        hlo_rec8_expand(&rec[i], len_by_id[rec[i].idA], len_by_id[rec[i].idB], &r);
        return true;
    });
}
int hlo_write_sfo8(const hlo_rec8 *rec, uint64_t n, const uint32_t *len_by_id, uint32_t n_ids, const char *path) {
    if ((!rec && n) || !path || (!len_by_id && n)) return HLO_EINVAL;
    std::vector<hlo_rec> tmp((size_t)n);
    for (uint64_t i = 0; i < n; ++i) {
        if (rec[i].idA >= n_ids || rec[i].idB >= n_ids) return HLO_EINVAL;
        hlo_rec8_expand(&rec[i], len_by_id[rec[i].idA], len_by_id[rec[i].idB], &tmp[(size_t)i]);
    }
    return hlo_write_sfo(tmp.data(), n, path);
}

// binary hand-off (include/hlovlp.h; the reader patch for EdgeCalculator.cpp:569-604 is in INTEGRATION.md §5)
int hlo_write_binary16(const hlo_rec16 *rec, const hlo_score *score, uint64_t n, const char *path) {
    if ((!rec && n) || !path) return HLO_EINVAL;
    std::string tmp = std::string(path) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return HLO_EIO;
    struct {
        char magic[4];
        uint32_t version, rec_bytes, flags;
        uint64_t n, reserved;
    } h = {{'H', 'L', 'O', 'B'}, 1u, (uint32_t)sizeof(hlo_rec16), score ? 1u : 0u, n, 0ull};
    static_assert(sizeof h == 32, "header is 32 bytes");
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    if (ok && n) ok = fwrite(rec, sizeof(hlo_rec16), (size_t)n, f) == (size_t)n;
    if (ok && n && score) ok = fwrite(score, sizeof(hlo_score), (size_t)n, f) == (size_t)n;
    if (fclose(f) != 0) ok = false;
    if (!ok || rename(tmp.c_str(), path) != 0) {
        remove(tmp.c_str());
        return HLO_EIO;
    }
    return HLO_OK;
}

int hlo_write_overlaps(const hlo_rec *rec, uint64_t n, const char *path) { return write_overlaps_impl(rec, nullptr, n, path); }

int hlo_write_overlaps_scored(const hlo_rec *rec, const hlo_score *score, uint64_t n, const char *path) {
    if (!score && n) return HLO_EINVAL;
    return write_overlaps_impl(rec, score, n, path);
}

}  // extern "C"
