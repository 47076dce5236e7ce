/*
 * ec_ref_harness.cpp — TEST INFRASTRUCTURE ONLY (tests/ load the library this builds; the product never does).
 *
 * Compiles the reference's OWN EdgeCalculator::score / phred_to_prob / overlap_score
 * (/root/reference/tools/HaploConduct/src/EdgeCalculator.cpp:25-140) into oracle/_ref/libecref.so, so that the CPU
 * restatement (hlo_oracle.c:orc_overlap_score) and the CUDA score kernel are pinned to the reference's code and not
 * only to each other.  Those three member functions need nothing from Boost (the rest of the file does, and Boost is
 * absent from the image): oracle/Makefile extracts the line range into oracle/_ref/ec_extract.inc at build time —
 * no reference source is copied into the repository — and this file supplies the few declarations they use.
 */
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

struct ProgramSettings {  // the two fields the extracted functions read (Types.h: ProgramSettings)
    double mismatch;
    unsigned int min_read_len;
};

class EdgeCalculator {
   public:
    ProgramSettings program_settings;
    double score(char nt1, char nt2, double p1, double p2, int &mismatch_count);
    double phred_to_prob(const int phred);
    double overlap_score(std::string seq1, std::string seq2, std::string score1, std::string score2, const unsigned int pos,
                         double &mismatch_rate);
};

#include "_ref/ec_extract.inc"

extern "C" double ecref_overlap_score(const char *seq1, int l1, const char *seq2, int l2, const char *ph1, const char *ph2,
                                      unsigned int pos, double mismatch, unsigned int min_read_len, double *mismatch_rate) {
    EdgeCalculator ec;
    ec.program_settings.mismatch = mismatch;
    ec.program_settings.min_read_len = min_read_len;
    double mr = 1.0;
    const double s = ec.overlap_score(std::string(seq1, l1), std::string(seq2, l2), std::string(ph1, l1), std::string(ph2, l2), pos, mr);
    *mismatch_rate = mr;
    return s;
}

extern "C" double ecref_phred_to_prob(int phred) {
    EdgeCalculator ec;
    return ec.phred_to_prob(phred);
}
