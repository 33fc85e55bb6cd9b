/* TEST INFRASTRUCTURE — C entry points of the CPU oracle for ctypes (tests/ only). */
#include <cstring>

#include "genreads.h"

using namespace orc;

extern "C" {

/* Philox4x32-10 block, for known-answer tests */
void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
uint32_t orc_philox_draw(uint64_t seed, int domain, int engine, uint64_t entity, uint64_t i) { return philox_draw(seed, domain, engine, entity, i); }
double orc_det_log(double x) { return det_log(x); }

void* orc_profile_load(const char* path, int paired, int isize, char* errbuf, int errlen) {
    Profile* p = new Profile();
    if (!p->load(path, paired != 0, isize)) {
        if (errbuf && errlen > 0) { strncpy(errbuf, p->err.c_str(), errlen - 1); errbuf[errlen - 1] = 0; }
        delete p; return nullptr;
    }
    return p;
}
void orc_profile_free(void* p) { delete (Profile*)p; }
int orc_profile_info(void* h, int* out) {   /* readLength, bins, kmerCount, hasISize, minInsert, iSizeCount, insN, delN, hasSubs2 */
    Profile* p = (Profile*)h;
    out[0] = p->readLength; out[1] = p->bins; out[2] = p->kmerCount; out[3] = p->hasISize; out[4] = p->minInsert;
    out[5] = p->hasISize ? p->iSizeCdf.cols : 0; out[6] = p->insCdf.cols; out[7] = p->delCdf.cols; out[8] = !p->subsCdf2.empty();
    return 0;
}
void orc_profile_scalars(void* h, double* out) {   /* insertRate, delRate, stdISize, gcStd, gcMeans[101] */
    Profile* p = (Profile*)h;
    out[0] = p->insertRate; out[1] = p->delRate; out[2] = p->stdISize; out[3] = p->gcStd;
    for (int i = 0; i < 101; i++) out[4 + i] = p->gcMeans[i];
}
/* which: 0 ins, 1 del, 2 isize, 3 subs1[idx], 4 subs2[idx], 5 quality[idx]; copies rows*cols doubles */
int orc_profile_cdf(void* h, int which, int idx, double* out) {
    Profile* p = (Profile*)h; const Mat* m = nullptr;
    switch (which) {
        case 0: m = &p->insCdf; break; case 1: m = &p->delCdf; break; case 2: m = &p->iSizeCdf; break;
        case 3: m = &p->subsCdf1[idx]; break; case 4: if (p->subsCdf2.empty()) return -1; m = &p->subsCdf2[idx]; break;
        case 5: m = &p->qualityCdf[idx]; break; default: return -1;
    }
    memcpy(out, m->v.data(), m->v.size() * sizeof(double));
    return (int)m->v.size();
}

/* predict() on one source window with an explicit draw tape: draws are consumed in order,
 * real[] for the "real" engine and ints[] for the "int" engine. Returns the output length, or
 * -1 if a tape ran dry. used[0..1] = draws consumed. */
struct ArrayDraws : Draws {
    const uint32_t* r; const uint32_t* q; uint64_t nr, nq, cr = 0, cq = 0; bool dry = false;
    void begin(int, uint64_t) override {}
    uint32_t next(int e) override {
        if (e == E_REAL) { if (cr >= nr) { dry = true; return 0; } return r[cr++]; }
        if (cq >= nq) { dry = true; return 0; } return q[cq++];
    }
    double gc_factor(double, double) override { return 0; }
    bool is_tape() const override { return true; }
};
int orc_predict(void* h, const char* ref, int n, int is_read1, const uint32_t* real, uint64_t nreal, const uint32_t* ints,
                uint64_t nint, char* out_seq, char* out_qual, uint64_t* used) {
    Profile* p = (Profile*)h; ArrayDraws d; d.r = real; d.q = ints; d.nr = nreal; d.nq = nint;
    std::string s, q;
    p->predict(std::string(ref, n), is_read1 != 0, d, s, q);
    if (d.dry) return -1;
    memcpy(out_seq, s.data(), s.size()); memcpy(out_qual, q.data(), q.size());
    if (used) { used[0] = d.cr; used[1] = d.cq; }
    return (int)s.size();
}

int orc_main(int argc, char** argv);
}
