// Host side of the sequencing profile: parse the reference's .profile text format, normalise and
// accumulate exactly as the reference does in FP64, then turn every CDF row into integer thresholds
// on the 32-bit draw so that the kernels sample with integer compares only.
//
// Reference: Profile::load /root/reference/lib/profile/Profile.cpp:930-1234, normParas :832-928,
// initCDFs :1363-1430, randIndx /root/reference/lib/mydefine/MyDefine.cpp:274-282.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace scs {

constexpr int kQualN = 94;   // Phred+33 .. 126
constexpr int kNB = 4;

// One sampler row in threshold form. A draw x in [0, 2^32) maps to index
//   min(#{k < eff : thr[k] <= x}, eff)          (thr non-decreasing)
// which equals the reference's  first k with  eps+(1-eps)*x/2^32 <= cdf[k], else n-1.
struct ThrRow {
    std::vector<uint32_t> thr;   // n entries (entries >= eff are padding)
    int eff = 0;
};

struct HostProfile {
    int readLength = 0, bins = 0, kmer = 3, kmerCount = 0;
    double insertRate = 0, delRate = 0, stdISize = 0, gcStd = 0;
    double gcMeans[101];
    bool paired = true, hasISize = false, hasSubs2 = false;
    int minInsert = 0, maxInsert = 0;

    // FP64 CDFs as the reference holds them
    std::vector<double> insCdf, delCdf, iSizeCdf;
    std::vector<double> subsCdf1, subsCdf2;   // [kmer][bin][4]
    std::vector<double> qualCdf;              // [pair 16][bin][94]

    // threshold form
    uint32_t thrInsert = 0;      // insertion  iff x <  thrInsert   (p <= insertRate)
    uint32_t thrDelete = 0;      // deletion   iff x <  thrDelete   (p <  delRate/(1-insertRate))
    bool thrInsertAll = false, thrDeleteAll = false;   // threshold == 2^32
    ThrRow insThr, delThr, iSizeThr;
    std::vector<uint32_t> subsThr1, subsThr2;   // [kmer][bin][4]: 3 thresholds + eff
    std::vector<uint32_t> qualThr;              // [pair][bin][94]
    std::vector<uint8_t> qualEff;               // [pair][bin]
    std::vector<uint8_t> qualLo;                // [pair][bin] number of leading zero thresholds

    std::string error;
    bool load(const std::string& path, bool pairedEnd, int isize);
};

// #{x in [0,2^32) : eps + (1-eps) * (x/2^32) <= c}  in [0, 2^32]
uint64_t count_draws_le(double c);
// #{x : x/2^32 <= c}  and  #{x : x/2^32 < c}
uint64_t count_unit_le(double c);
uint64_t count_unit_lt(double c);
ThrRow make_thr_row(const double* cdf, int n);

}  // namespace scs
