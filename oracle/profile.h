/* TEST INFRASTRUCTURE — CPU oracle (see draws.h header). Restatement of the load+predict half
 * of the reference's sequencing profile:
 *   load        /root/reference/lib/profile/Profile.cpp:930-1234
 *   init/kmers  Profile.cpp:69-123,171-214
 *   normParas   Profile.cpp:832-928   (Matrix::normalize lib/matrix/Matrix.h:483-503)
 *   initCDFs    Profile.cpp:1363-1430 (Matrix::cumsum Matrix.h:506-522)
 *   samplers    Profile.cpp:1482-1580, randIndx lib/mydefine/MyDefine.cpp:274-282
 *   predict     Profile.cpp:1582-1697
 * Everything is FP64 exactly as the reference computes it (the product turns the same
 * CDFs into integer thresholds; tests prove the two agree).
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "draws.h"

namespace orc {

static const double ZERO_FINAL = 2.2204e-16;   /* MyDefine.cpp:20 */

inline std::string trim(const std::string& s, const char* cl = " \t\r\n") {   /* MyDefine.cpp:295-307 */
    size_t a = s.find_first_not_of(cl);
    if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(cl);
    return s.substr(a, b - a + 1);
}
inline std::vector<std::string> split(const std::string& s, char d) {        /* lib/split/split.cpp:3-16 */
    std::vector<std::string> out; std::stringstream ss(s); std::string it;
    while (std::getline(ss, it, d)) out.push_back(it);
    return out;
}

struct Mat {   /* row-major dense double matrix: just what Profile uses of lib/matrix */
    int rows = 0, cols = 0; std::vector<double> v;
    void resize(int r, int c) { rows = r; cols = c; v.assign((size_t)r * c, 0.0); }
    double& at(int r, int c) { return v[(size_t)r * cols + c]; }
    const double* row(int r) const { return &v[(size_t)r * cols]; }
    void normalize_rows() {                   /* Matrix::normalize(0) */
        for (int i = 0; i < rows; i++) {
            double s = 0; for (int j = 0; j < cols; j++) s += at(i, j);
            for (int j = 0; j < cols; j++) at(i, j) /= (ZERO_FINAL + s);
        }
    }
    double rowsum(int i) { double s = 0; for (int j = 0; j < cols; j++) s += at(i, j); return s; }
    void cumsum_rows() {                      /* Matrix::cumsum */
        for (int i = 0; i < rows; i++) for (int j = 1; j < cols; j++) at(i, j) = at(i, j) + at(i, j - 1);
    }
};

inline int base_index(char c) {               /* getIndexOfBase, MyDefine.cpp:326-334, bases="ACGT" */
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}
inline char complement_base(char b) {         /* getComplementBase, MyDefine.cpp:352-367 */
    switch (b) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c';
        default: return 'N';
    }
}

struct Profile {
    std::string bases; int N = 4, kmer = 3, bins = 0, readLength = 0, kmerCount = 0;
    double insertRate = 0, delRate = 0, stdISize = 0, gcStd = 0;
    double gcMeans[101];
    Mat insCdf, delCdf;                  /* 1 x n */
    std::vector<Mat> subsCdf1, subsCdf2; /* kmerCount x (bins x 4); subsCdf2 empty => use 1 */
    std::vector<Mat> qualityCdf;         /* 16 x (bins x 94) */
    bool hasISize = false; int minInsert = 0; Mat iSizeCdf;
    bool paired = true;
    std::string err;

    static const int QN = 94;   /* maxBaseQuality-minBaseQuality+1 = 126-33+1 */

    /* kmer numbering of initKmers (Profile.cpp:69-123): j=k-1..0 leading X's, last char fastest */
    int kmer_index(const char* s) const {   /* s has `kmer` chars; -1 if not a known k-mer */
        int lead = 0; while (lead < kmer - 1 && s[lead] == 'X') lead++;
        int idx = 0, p = 1;
        for (int j = kmer - 1; j > lead; j--) { p *= N; idx += p; }   /* sum_{t=1}^{kmer-1-lead} N^t */
        int v = 0;
        for (int i = lead; i < kmer; i++) { int b = base_index(s[i]); if (b < 0) return -1; v = v * N + b; }
        return idx + v;
    }

    static bool next_line(std::ifstream& f, std::string& line) {   /* getNextLine, MyDefine.cpp:337-349 */
        line = "";
        while (std::getline(f, line)) { if (!line.empty() && line[0] != '#') break; }
        return !line.empty();
    }

    bool load(const std::string& path, bool pairedEnd, int isize) {
        paired = pairedEnd;
        std::ifstream f(path.c_str());
        if (!f.is_open()) { err = "can not open file " + path; return false; }
        std::string line; int binCount = -1; kmer = -1; readLength = -1; bases = "";
        while (next_line(f, line)) {
            std::vector<std::string> fs = split(line, ':');
            if (fs.size() != 2) { err = "malformed model file header"; return false; }
            std::string k = trim(fs[0]), val = trim(fs[1]);
            if (k == "bases") bases = val;
            else if (k == "binCount") binCount = atoi(val.c_str());
            else if (k == "kmer") kmer = atoi(val.c_str());
            else if (k == "readLength") readLength = atoi(val.c_str());
            else { err = "malformed model file header"; return false; }
            if (!bases.empty() && binCount > 0 && kmer > 0 && readLength > 0) break;
        }
        if (bases.empty() || binCount <= 0 || kmer <= 0 || readLength <= 0) { err = "malformed model file"; return false; }
        if (bases != "ACGT") { err = "oracle supports bases ACGT only"; return false; }
        N = 4;
        bins = readLength;                       /* Profile::init overwrites bins, Profile.cpp:183 */
        if (binCount != bins) { err = "binCount != readLength (the reference would write out of range)"; return false; }
        kmerCount = 0; { int p = 1; for (int i = 0; i < kmer; i++) { p *= N; kmerCount += p; } }
        std::vector<Mat> subs1(kmerCount), subs2(kmerCount), qual(N * N);
        for (auto& m : subs1) m.resize(bins, N);
        for (auto& m : subs2) m.resize(bins, N);
        for (auto& m : qual) m.resize(bins, QN);
        Mat insF, delF; insF.resize(1, 1); delF.resize(1, 1);
        for (int i = 0; i < 101; i++) gcMeans[i] = 0;
        int loaded = 0;
        while (next_line(f, line)) {
            if (line == "[Insert Rate]") { if (!next_line(f, line)) return bad(); insertRate = atof(trim(line).c_str()); loaded++; }
            else if (line == "[Insert Frequency]") {
                if (!next_line(f, line)) return bad();
                auto fs = split(line, '\t'); insF.resize(1, (int)fs.size());
                for (size_t j = 0; j < fs.size(); j++) insF.at(0, (int)j) = atof(trim(fs[j]).c_str());
                loaded++;
            }
            else if (line == "[Deletion Rate]") { if (!next_line(f, line)) return bad(); delRate = atof(trim(line).c_str()); loaded++; }
            else if (line == "[Deletion Frequency]") {
                if (!next_line(f, line)) return bad();
                auto fs = split(line, '\t'); delF.resize(1, (int)fs.size());
                for (size_t j = 0; j < fs.size(); j++) delF.at(0, (int)j) = atof(trim(fs[j]).c_str());
                loaded++;
            }
            else if (line == "[Substitution Probs]") {
                for (int i = 0; i < kmerCount; i++) {
                    if (!next_line(f, line)) return bad();
                    auto fs = split(line, ':');
                    if (fs.size() != 2 || trim(fs[0]) != "kmer") return bad();
                    std::string km = trim(fs[1]);
                    int ki = ((int)km.size() == kmer) ? kmer_index(km.c_str()) : -1;
                    if (ki < 0) return bad();
                    for (int j = 0; j < bins * 2; j++) {
                        if (!next_line(f, line)) return bad();
                        auto r = split(line, '\t');
                        if ((int)r.size() != N) return bad();
                        for (int k = 0; k < N; k++) {
                            double p = atof(trim(r[k]).c_str());
                            if (j < bins) subs1[ki].at(j, k) = p; else subs2[ki].at(j - bins, k) = p;
                        }
                    }
                }
                loaded++;
            }
            else if (line == "[Base Quality Distribution]") {
                for (int i = 0; i < N * N; i++) {
                    if (!next_line(f, line)) return bad();
                    auto fs = split(line, ':');
                    if (fs.size() != 2 || trim(fs[0]) != "basePairIndx") return bad();
                    int bp = atoi(trim(fs[1]).c_str());
                    if (bp < 0 || bp > N * N - 1) return bad();
                    for (int j = 0; j < bins; j++) {
                        if (!next_line(f, line)) return bad();
                        auto r = split(line, '\t');
                        if ((int)r.size() != QN) return bad();
                        for (int k = 0; k < QN; k++) qual[bp].at(j, k) = atof(trim(r[k]).c_str());
                    }
                }
                loaded++;
            }
            else if (line == "[Insert Size Standard Deviation]") { if (!next_line(f, line)) return bad(); stdISize = atof(trim(line).c_str()); loaded++; }
            else if (line == "[Log Ratio Mean Value]") {
                for (int j = 0; j < 101; j++) {
                    if (!next_line(f, line)) return bad();
                    auto fs = split(line, '\t');
                    if (fs.size() != 2) return bad();
                    int gc = atoi(fs[0].c_str());
                    if (gc < 0 || gc > 100) return bad();
                    gcMeans[gc] = atof(fs[1].c_str());
                }
                loaded++;
            }
            else if (line == "[Log Ratio Standard Deviation]") { if (!next_line(f, line)) return bad(); gcStd = atof(trim(line).c_str()); loaded++; }
        }
        if (loaded < 9) { err = "corrupted model file, failed to load some parameters"; return false; }

        /* normParas(true): Profile.cpp:840-862 */
        for (int i = 0; i < kmerCount; i++) {
            int last = i < N ? i : (i < N + N * N ? (i - N) % N : (i - N - N * N) % N);   /* index of the k-mer's last base */
            subs1[i].normalize_rows();
            for (int j = 0; j < bins; j++) if (subs1[i].rowsum(j) < ZERO_FINAL) subs1[i].at(j, last) = 1;
            subs2[i].normalize_rows();
            for (int j = 0; j < bins; j++) if (subs2[i].rowsum(j) < ZERO_FINAL) subs2[i].at(j, last) = 1;
        }
        for (auto& m : qual) m.normalize_rows();
        /* insert-size table, Profile.cpp:908-926 */
        Mat iSizeDist;
        hasISize = false;
        if (paired && stdISize > 0) {
            int mean = isize + 1;
            int intervalLen = (int)(6 * stdISize);
            int minI = std::max(mean - intervalLen / 2, readLength);
            int maxI = 2 * mean - minI;
            int cnt = maxI - minI + 1;
            if (cnt > 0) {
                hasISize = true; minInsert = minI;
                iSizeDist.resize(1, cnt);
                const double PI = 3.1415926;   /* normpdf, MyDefine.cpp:54-57 */
                for (int i = 0; i < cnt; i++) {
                    double x = minI + i;
                    iSizeDist.at(0, i) = exp(-pow(x - mean, 2) / (2 * pow(stdISize, 2))) / (sqrt(2 * PI) * stdISize);
                }
                iSizeDist.normalize_rows();
            }
        }
        /* initCDFs: Profile.cpp:1363-1430 */
        insCdf = insF; insCdf.cumsum_rows();
        delCdf = delF; delCdf.cumsum_rows();
        qualityCdf = qual;
        for (auto& m : qualityCdf) { m.normalize_rows(); m.cumsum_rows(); }   /* normalised a second time, :1393 */
        if (hasISize) { iSizeCdf = iSizeDist; iSizeCdf.cumsum_rows(); }
        subsCdf1 = subs1; for (auto& m : subsCdf1) m.cumsum_rows();
        subsCdf2.clear();
        if (paired && stdISize > 0) { subsCdf2 = subs2; for (auto& m : subsCdf2) m.cumsum_rows(); }
        return true;
    }
    bool bad() { err = "malformed profile file"; return false; }

    /* randIndx(double*, unsigned), MyDefine.cpp:274-282 */
    static unsigned rand_index(const double* cdf, unsigned ac, uint32_t x) {
        double r = uni_real(x, ZERO_FINAL, 1);
        for (unsigned k = 0; k < ac; k++) if (r <= cdf[k]) return k;
        return ac - 1;
    }

    int yield_insert_size(Draws& d) const {   /* Profile.cpp:1482-1489; caller checks hasISize */
        return minInsert + (int)rand_index(iSizeCdf.row(0), iSizeCdf.cols, d.next(E_REAL));
    }

    /* Free-running streams only (no tape): the indel stage of predict() draws the DISTANCE to the next indel event instead of
     * two draws per read position. Per position the reference decides "insertion" with probability pI = P(p <= insertRate),
     * else "deletion" with probability pD = P(p2 < delRate / (1 - insertRate)) — the same at every position, so the number of
     * event-free positions before the next event is geometric with q = (1 - pI)(1 - pD), and the event is an insertion with
     * probability pI / (1 - q). pI and pD are taken as the exact rationals T / 2^32 of the 32-bit draws that satisfy the
     * reference's comparisons, so that the CUDA path (integer thresholds) forms bit-identical constants. Same distribution as
     * Profile.cpp:1603-1630 at ~1 draw per read instead of 2 per base; replay keeps the reference's consumption. */
    struct IndelGeom { bool any = false; double logQ = 0; uint64_t thrInsType = 0; };
    static uint64_t count_unit_le(double c) { if (c < 0) return 0; double t = floor(ldexp(c, 32)) + 1; return t >= 4294967296.0 ? (1ull << 32) : (uint64_t)t; }
    static uint64_t count_unit_lt(double c) { if (c <= 0) return 0; double t = ceil(ldexp(c, 32)); return t >= 4294967296.0 ? (1ull << 32) : (uint64_t)t; }
    IndelGeom indel_geom() const {
        IndelGeom g;
        const uint64_t tI = count_unit_le(insertRate), tD = count_unit_lt(delRate / (1 - insertRate));
        const double pI = (double)tI / 4294967296.0, pD = (double)tD / 4294967296.0;
        const double q = (1.0 - pI) * (1.0 - pD);
        g.any = q < 1.0;
        g.logQ = det_log(q);
        g.thrInsType = tD == 0 ? (1ull << 32) : tI == 0 ? 0 : count_unit_lt(pI / (1.0 - q));
        return g;
    }

    /* Profile::predict(char*, int), Profile.cpp:1582-1697. Returns bases and qualities (equal length). */
    void predict(const std::string& ref, bool isRead1, Draws& d, std::string& outSeq, std::string& outQual) const {
        int n = (int)ref.size();
        std::vector<std::vector<int>> ins(n);
        std::vector<int> indelLens; indelLens.reserve(n + 8);
        int indelLength = 0;
        if (!d.is_tape()) {
            const IndelGeom G = indel_geom();
            for (int j = 0; G.any && j < n;) {
                const double u = ((double)d.next(E_REAL) + 0.5) / 4294967296.0;
                const double gd = floor(det_log(u) / G.logQ);
                if (!(gd < (double)(n - j))) break;
                for (int i = 0; i < (int)gd; i++) indelLens.push_back(0);
                j += (int)gd;
                int k; std::vector<int>& bi = ins[j];
                if ((uint64_t)d.next(E_REAL) < G.thrInsType) {
                    k = (int)rand_index(insCdf.row(0), insCdf.cols, d.next(E_REAL));
                    for (int i = 0; i < k; i++) bi.push_back((int)uni_int(d.next(E_INT), 0, N - 1));
                } else k = (int)rand_index(delCdf.row(0), delCdf.cols, d.next(E_REAL));
                if (bi.empty() && k > 0) {
                    k = std::min(n - j, k);
                    indelLength -= k;
                    indelLens.push_back(k);
                    for (int i = 1; i < k; i++) indelLens.push_back(0);
                    j += k;
                } else {
                    indelLength += k; j++; indelLens.push_back(k);
                }
            }
            indelLens.resize(n, 0);
        } else
        for (int j = 0; j < n;) {
            /* getIndelSeq, Profile.cpp:1552-1570 */
            int k = 0; std::vector<int>& bi = ins[j];
            double p = uni_real(d.next(E_REAL), 0, 1);
            if (p <= insertRate) {
                k = (int)rand_index(insCdf.row(0), insCdf.cols, d.next(E_REAL));
                for (int i = 0; i < k; i++) bi.push_back((int)uni_int(d.next(E_INT), 0, N - 1));   /* never T, :1560 */
            } else {
                p = uni_real(d.next(E_REAL), 0, 1);
                if (p < delRate / (1 - insertRate)) k = (int)rand_index(delCdf.row(0), delCdf.cols, d.next(E_REAL));
            }
            if (bi.empty() && k > 0) {
                k = std::min(n - j, k);
                indelLength -= k;
                indelLens.push_back(k);
                for (int i = 1; i < k; i++) indelLens.push_back(0);
                j += k;
            } else {
                indelLength += k; j++; indelLens.push_back(k);
            }
        }
        if (n + indelLength < 50) {
            indelLength = 0;
            for (auto& v : ins) v.clear();
            indelLens.assign(n, 0);
        }
        std::string src; src.reserve(n + indelLength);
        for (int j = 0; j < n;) {
            if (ins[j].empty() && indelLens[j] > 0) { j += indelLens[j]; continue; }
            else if (indelLens[j] == 0) { src.push_back(ref[j]); j++; }
            else { src.push_back(ref[j]); for (int b : ins[j]) src.push_back(bases[b]); j++; }
        }
        n += indelLength;
        std::string ctx(kmer - 1, 'X'); ctx += src;
        outSeq.assign(n, 'N'); outQual.assign(n, '!');
        for (int j = 0; j < n; j++) {
            int refIndx = base_index(src[j]);
            int bin = j * bins / n;
            int ki = kmer_index(&ctx[j]);
            int k;
            if (ki == -1) k = base_index(ctx[j + kmer - 1]);          /* getSubBaseIndx*, :1527-1529 */
            else {
                const Mat& m = (!isRead1 && !subsCdf2.empty()) ? subsCdf2[ki] : subsCdf1[ki];
                k = (int)rand_index(m.row(bin), N, d.next(E_REAL));
            }
            outSeq[j] = (k == -1) ? 'N' : bases[k];
            if (k == -1) outQual[j] = (char)uni_int(d.next(E_INT), 33, 53);           /* getRandBaseQuality */
            else outQual[j] = (char)(33 + rand_index(qualityCdf[refIndx * N + k].row(bin), QN, d.next(E_REAL)));
        }
    }
};

}  // namespace orc
