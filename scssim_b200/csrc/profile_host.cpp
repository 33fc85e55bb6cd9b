#include "profile_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace scs {

static const double kEps = 2.2204e-16;   // ZERO_FINAL, /root/reference/lib/mydefine/MyDefine.cpp:20

// ---- draw -> decision thresholds --------------------------------------------------------------
// The reference forms r = start + (end-start) * (x / 2^32) in FP64 from a 32-bit engine output x and
// compares it with a table entry. r is non-decreasing in x, so each comparison is a cut point on x,
// found here by bisection over x with the reference's own FP64 expression.
static inline double draw_r(uint64_t x) { return kEps + (1.0 - kEps) * ((double)x / 4294967296.0); }
static inline double draw_u(uint64_t x) { return 0.0 + (1.0 - 0.0) * ((double)x / 4294967296.0); }

template <class Pred> static uint64_t first_false(Pred holds) {   // holds(x) is true on a prefix of [0, 2^32)
    uint64_t lo = 0, hi = 1ull << 32;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (holds(mid)) lo = mid + 1; else hi = mid; }
    return lo;
}
uint64_t count_draws_le(double c) { return first_false([c](uint64_t x) { return draw_r(x) <= c; }); }
uint64_t count_unit_le(double c) { return first_false([c](uint64_t x) { return draw_u(x) <= c; }); }
uint64_t count_unit_lt(double c) { return first_false([c](uint64_t x) { return draw_u(x) < c; }); }

ThrRow make_thr_row(const double* cdf, int n) {
    ThrRow r; r.thr.assign(n, 0xFFFFFFFFu);
    int eff = n - 1;   // "else n-1" of randIndx: the last entry never needs a compare
    for (int k = 0; k < n - 1; k++) {
        uint64_t t = count_draws_le(cdf[k]);
        if (t >= (1ull << 32)) { eff = k; break; }   // every draw satisfies entry k
        r.thr[k] = (uint32_t)t;
    }
    r.eff = eff < 0 ? 0 : eff;
    return r;
}

// ---- text format --------------------------------------------------------------------------------
namespace {
struct Lines {   // getNextLine(): skips empty lines and '#' comments (MyDefine.cpp:337-349)
    std::vector<char> buf; size_t pos = 0; int lineNum = 0;
    bool open(const std::string& path) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        buf.resize((size_t)n + 1);
        size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
        fclose(f);
        buf[got] = 0; buf.resize(got + 1);
        return true;
    }
    // returns pointer to a NUL-terminated line (inside buf) or nullptr at end of file
    char* next() {
        size_t n = buf.size() - 1;
        while (pos < n) {
            char* s = &buf[pos];
            char* e = (char*)memchr(s, '\n', n - pos);
            size_t len = e ? (size_t)(e - s) : n - pos;
            pos += len + (e ? 1 : 0);
            s[len] = 0;
            lineNum++;
            if (len > 0 && s[0] != '#') return s;
        }
        return nullptr;
    }
};
inline char* trim_inplace(char* s) {   // trim(" \t\r\n")
    while (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\n') s++;
    size_t n = strlen(s);
    while (n > 0 && (s[n - 1] == ' ' || s[n - 1] == '\t' || s[n - 1] == '\r' || s[n - 1] == '\n')) s[--n] = 0;
    return s;
}
// split on one delimiter like std::getline does: a trailing empty field is not produced
inline int split_inplace(char* s, char d, std::vector<char*>& out) {
    out.clear();
    if (*s == 0) return 0;
    char* cur = s;
    for (;;) {
        char* e = strchr(cur, d);
        if (!e) { out.push_back(cur); break; }
        *e = 0; out.push_back(cur); cur = e + 1;
        if (*cur == 0) break;
    }
    return (int)out.size();
}
inline void normalize_row(double* r, int n) {   // Matrix::normalize(0), lib/matrix/Matrix.h:483-503
    double s = 0; for (int j = 0; j < n; j++) s += r[j];
    for (int j = 0; j < n; j++) r[j] /= (kEps + s);
}
inline void cumsum_row(double* r, int n) { for (int j = 1; j < n; j++) r[j] = r[j] + r[j - 1]; }
inline int base_code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }
}  // namespace

bool HostProfile::load(const std::string& path, bool pairedEnd, int isize) {
    paired = pairedEnd;
    Lines L;
    if (!L.open(path)) { error = "can not open file " + path; return false; }
    auto malformed = [&](const char* line) {
        error = "Error: malformed model file " + path + " @line " + std::to_string(L.lineNum) + "\n" + (line ? line : "");
        return false;
    };
    std::vector<char*> f;
    std::string bases; int binCount = -1; kmer = -1; readLength = -1;
    char* line;
    while ((line = L.next())) {
        if (split_inplace(line, ':', f) != 2) return malformed(line);
        char* k = trim_inplace(f[0]); char* v = trim_inplace(f[1]);
        if (!strcmp(k, "bases")) { bases = v; if (bases.empty()) return malformed(line); }
        else if (!strcmp(k, "binCount")) { binCount = atoi(v); if (binCount <= 0) return malformed(line); }
        else if (!strcmp(k, "kmer")) { kmer = atoi(v); if (kmer <= 0) return malformed(line); }
        else if (!strcmp(k, "readLength")) { readLength = atoi(v); if (readLength <= 0) return malformed(line); }
        else return malformed(line);
        if (!bases.empty() && binCount > 0 && kmer > 0 && readLength > 0) break;
    }
    if (bases.empty() || binCount <= 0 || kmer <= 0 || readLength <= 0) { error = "Error: malformed model file " + path; return false; }
    if (bases != "ACGT" || kmer != 3) { error = "Error: only profiles with bases ACGT and kmer 3 are supported (" + path + ")"; return false; }
    bins = readLength;   // Profile::init(), Profile.cpp:183
    if (binCount != bins) { error = "Error: binCount differs from readLength in " + path; return false; }
    if (readLength > 1024) { error = "Error: readLength above 1024 is not supported"; return false; }
    kmerCount = 4 + 16 + 64;
    std::vector<double> s1((size_t)kmerCount * bins * 4, 0.0), s2((size_t)kmerCount * bins * 4, 0.0);
    std::vector<double> q((size_t)16 * bins * kQualN, 0.0);
    std::vector<double> insF(1, 0.0), delF(1, 0.0);
    for (int i = 0; i < 101; i++) gcMeans[i] = 0;
    int loaded = 0;
    auto need = [&]() -> char* { char* l = L.next(); if (!l) error = "Error: malformed profile file " + path; return l; };
    while ((line = L.next())) {
        if (!strcmp(line, "[Insert Rate]")) { if (!(line = need())) return false; insertRate = atof(trim_inplace(line)); loaded++; }
        else if (!strcmp(line, "[Deletion Rate]")) { if (!(line = need())) return false; delRate = atof(trim_inplace(line)); loaded++; }
        else if (!strcmp(line, "[Insert Frequency]") || !strcmp(line, "[Deletion Frequency]")) {
            bool isIns = line[1] == 'I';
            if (!(line = need())) return false;
            int n = split_inplace(line, '\t', f);
            if (n < 1) return malformed(line);
            std::vector<double>& dst = isIns ? insF : delF;
            dst.assign(n, 0.0);
            for (int j = 0; j < n; j++) dst[j] = atof(trim_inplace(f[j]));
            loaded++;
        }
        else if (!strcmp(line, "[Substitution Probs]")) {
            for (int i = 0; i < kmerCount; i++) {
                if (!(line = need())) return false;
                if (split_inplace(line, ':', f) != 2 || strcmp(trim_inplace(f[0]), "kmer")) return malformed(line);
                char* km = trim_inplace(f[1]);
                // k-mer numbering of Profile::initKmers (Profile.cpp:69-123): XXb, Xbb, bbb blocks
                int ki = -1;
                if (strlen(km) == 3) {
                    int b0 = base_code(km[0]), b1 = base_code(km[1]), b2 = base_code(km[2]);
                    if (km[0] == 'X' && km[1] == 'X' && b2 >= 0) ki = b2;
                    else if (km[0] == 'X' && b1 >= 0 && b2 >= 0) ki = 4 + b1 * 4 + b2;
                    else if (b0 >= 0 && b1 >= 0 && b2 >= 0) ki = 20 + b0 * 16 + b1 * 4 + b2;
                }
                if (ki < 0) { error = "Error: unrecognized kmer @line " + std::to_string(L.lineNum) + " in profile file " + path; return false; }
                for (int j = 0; j < bins * 2; j++) {
                    if (!(line = need())) return false;
                    if (split_inplace(line, '\t', f) != 4) return malformed(line);
                    double* dst = (j < bins ? &s1[((size_t)ki * bins + j) * 4] : &s2[((size_t)ki * bins + (j - bins)) * 4]);
                    for (int k = 0; k < 4; k++) dst[k] = atof(trim_inplace(f[k]));
                }
            }
            loaded++;
        }
        else if (!strcmp(line, "[Base Quality Distribution]")) {
            for (int i = 0; i < 16; i++) {
                if (!(line = need())) return false;
                if (split_inplace(line, ':', f) != 2 || strcmp(trim_inplace(f[0]), "basePairIndx")) return malformed(line);
                int bp = atoi(trim_inplace(f[1]));
                if (bp < 0 || bp > 15) { error = "Error: unrecognized basePairIndx @line " + std::to_string(L.lineNum) + " in profile file " + path; return false; }
                for (int j = 0; j < bins; j++) {
                    if (!(line = need())) return false;
                    if (split_inplace(line, '\t', f) != kQualN) return malformed(line);
                    double* dst = &q[((size_t)bp * bins + j) * kQualN];
                    for (int k = 0; k < kQualN; k++) dst[k] = atof(trim_inplace(f[k]));
                }
            }
            loaded++;
        }
        else if (!strcmp(line, "[Insert Size Standard Deviation]")) { if (!(line = need())) return false; stdISize = atof(trim_inplace(line)); loaded++; }
        else if (!strcmp(line, "[Log Ratio Mean Value]")) {
            for (int j = 0; j < 101; j++) {
                if (!(line = need())) return false;
                if (split_inplace(line, '\t', f) != 2) return malformed(line);
                int gc = atoi(f[0]);
                if (gc < 0 || gc > 100) return malformed(line);
                gcMeans[gc] = atof(f[1]);
            }
            loaded++;
        }
        else if (!strcmp(line, "[Log Ratio Standard Deviation]")) { if (!(line = need())) return false; gcStd = atof(trim_inplace(line)); loaded++; }
    }
    if (loaded < 9) { error = "Error: corrupted model file " + path + ", failed to load some parameters!"; return false; }

    // normParas(true), Profile.cpp:840-862
    for (int ki = 0; ki < kmerCount; ki++) {
        int last = ki < 4 ? ki : (ki < 20 ? (ki - 4) & 3 : (ki - 20) & 3);
        for (int j = 0; j < bins; j++) {
            for (std::vector<double>* t : {&s1, &s2}) {
                double* r = &(*t)[((size_t)ki * bins + j) * 4];
                normalize_row(r, 4);
                if (r[0] + r[1] + r[2] + r[3] < kEps) r[last] = 1;   // all-zero row -> keep the base
            }
        }
    }
    // quality rows are normalised in normParas and once more in initCDFs (Profile.cpp:861, :1393)
    for (size_t r = 0; r < (size_t)16 * bins; r++) { normalize_row(&q[r * kQualN], kQualN); normalize_row(&q[r * kQualN], kQualN); cumsum_row(&q[r * kQualN], kQualN); }
    qualCdf.swap(q);
    for (size_t r = 0; r < (size_t)kmerCount * bins; r++) { cumsum_row(&s1[r * 4], 4); cumsum_row(&s2[r * 4], 4); }
    subsCdf1.swap(s1);
    hasSubs2 = paired && stdISize > 0;   // Profile.cpp:1416-1426
    if (hasSubs2) subsCdf2.swap(s2);
    insCdf = insF; cumsum_row(insCdf.data(), (int)insCdf.size());
    delCdf = delF; cumsum_row(delCdf.data(), (int)delCdf.size());
    // insert-size table, Profile.cpp:908-926 (normpdf with PI = 3.1415926, MyDefine.cpp:54-57)
    hasISize = false; iSizeCdf.clear();
    if (paired && stdISize > 0) {
        int mean = isize + 1;
        int interval = (int)(6 * stdISize);
        int lo = std::max(mean - interval / 2, readLength);
        int hi = 2 * mean - lo;
        if (hi >= lo) {
            hasISize = true; minInsert = lo; maxInsert = hi;
            iSizeCdf.resize(hi - lo + 1);
            const double PI = 3.1415926;
            for (int i = 0; i <= hi - lo; i++) {
                double x = lo + i;
                iSizeCdf[i] = exp(-pow(x - mean, 2) / (2 * pow(stdISize, 2))) / (sqrt(2 * PI) * stdISize);
            }
            normalize_row(iSizeCdf.data(), (int)iSizeCdf.size());
            cumsum_row(iSizeCdf.data(), (int)iSizeCdf.size());
        }
    }

    // ---- threshold form ----
    uint64_t ti = count_unit_le(insertRate);
    thrInsertAll = ti >= (1ull << 32); thrInsert = thrInsertAll ? 0xFFFFFFFFu : (uint32_t)ti;
    uint64_t td = count_unit_lt(delRate / (1 - insertRate));
    thrDeleteAll = td >= (1ull << 32); thrDelete = thrDeleteAll ? 0xFFFFFFFFu : (uint32_t)td;
    insThr = make_thr_row(insCdf.data(), (int)insCdf.size());
    delThr = make_thr_row(delCdf.data(), (int)delCdf.size());
    if (hasISize) iSizeThr = make_thr_row(iSizeCdf.data(), (int)iSizeCdf.size());
    auto subs_thr = [&](const std::vector<double>& cdf, std::vector<uint32_t>& out) {
        out.assign((size_t)kmerCount * bins * 4, 0);
        for (size_t r = 0; r < (size_t)kmerCount * bins; r++) {
            ThrRow t = make_thr_row(&cdf[r * 4], 4);
            out[r * 4 + 0] = t.thr[0]; out[r * 4 + 1] = t.thr[1]; out[r * 4 + 2] = t.thr[2]; out[r * 4 + 3] = (uint32_t)t.eff;
        }
    };
    subs_thr(subsCdf1, subsThr1);
    if (hasSubs2) subs_thr(subsCdf2, subsThr2);
    qualThr.assign((size_t)16 * bins * kQualN, 0xFFFFFFFFu);
    qualEff.assign((size_t)16 * bins, 0); qualLo.assign((size_t)16 * bins, 0);
    for (size_t r = 0; r < (size_t)16 * bins; r++) {
        ThrRow t = make_thr_row(&qualCdf[r * kQualN], kQualN);
        memcpy(&qualThr[r * kQualN], t.thr.data(), kQualN * sizeof(uint32_t));
        qualEff[r] = (uint8_t)t.eff;
        int lo = 0; while (lo < t.eff && t.thr[lo] == 0) lo++;
        qualLo[r] = (uint8_t)lo;
    }
    return true;
}

}  // namespace scs
