/* TEST INFRASTRUCTURE — CPU oracle for the `scssim genreads` hot path. Not shipped,
 * never linked into or called from the product (scssim_b200/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * draws.h: where the oracle gets its random numbers.
 *
 * The reference (three RNG families, wall-clock seeded, SURVEY.md finding 4) is
 * restated as a *deterministic function of per-entity draw streams*:
 *   stream(domain, entity, engine)[i]  ->  u32
 * and every uniform the reference forms is  start + (end-start) * (x / 2^32)
 * (/root/reference/lib/threadpool/ThreadPool.cpp:203-212; libc rand() r/2^31 ==
 * (r<<1)/2^32, lib/mydefine/MyDefine.cpp:285-292).
 *
 * Two sources:
 *   TapeDraws   — the flat logs written by oracle/_ref/bin/scssim_replay -t 1, consumed in the
 *                 reference's own sequential order. While it runs it records, per entity, where
 *                 that entity's draws start on each tape; the CUDA path replays from those offsets.
 *   PhiloxDraws — Philox4x32-10, key = seed, counter = (entity, i/4, domain*2+engine): the
 *                 free-running source. The CUDA path uses the identical function, so oracle and
 *                 GPU FASTQ are byte-identical in this mode too.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace orc {

enum Domain : int {
    D_FRAG = 0,   /* Genome::splitToFrags          entity = sequence index            (tape mrand) */
    D_POIS = 1,   /* Malbac::setPrimers/poissRand  entity = round<<40 | template      (tape mrand) */
    D_AMPF = 2,   /* Fragment::amplify             entity = pass<<40 | fragment       (wreal,wint) */
    D_AMPS = 3,   /* Amplicon::amplify             entity = cycle<<40 | semi index    (wreal,wint) */
    D_GCF = 4,    /* Profile::getGCFactor          entity = full amplicon index       (tape gcf)   */
    D_MULTM = 5,  /* randIndx_hp leftover draws    entity = 0                         (tape mreal) */
    D_MULTC = 6,  /* batchSampling per chunk       entity = chunk index               (wreal)      */
    D_READ = 7,   /* Amplicon::yieldReads          entity = read (SE) / pair (PE) id  (wreal,wint) */
    D_COUNT = 8
};
enum Engine : int { E_REAL = 0, E_INT = 1 };

/* ---- Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants) ---- */
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

inline uint32_t philox_draw(uint64_t seed, int domain, int engine, uint64_t entity, uint64_t i) {
    uint32_t ctr[4] = {(uint32_t)entity, (uint32_t)(entity >> 32), (uint32_t)(i >> 2),
                       (uint32_t)(domain * 2 + engine)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    philox4x32_10(ctr, key, out);
    return out[i & 3];
}

/* ---- deterministic log: only IEEE +,-,*,/ and frexp, so CPU and GPU agree bit for bit ---- */
inline double det_log(double x) {
    if (x <= 0.0) return -INFINITY;
    int e;
    double m = frexp(x, &e);                 /* m in [0.5, 1) */
    if (m < 0.70710678118654752440) { m = m * 2.0; e -= 1; }
    double t = (m - 1.0) / (m + 1.0);
    double t2 = t * t;
    double s = 1.0 / 27.0;
    s = s * t2 + 1.0 / 25.0;
    s = s * t2 + 1.0 / 23.0;
    s = s * t2 + 1.0 / 21.0;
    s = s * t2 + 1.0 / 19.0;
    s = s * t2 + 1.0 / 17.0;
    s = s * t2 + 1.0 / 15.0;
    s = s * t2 + 1.0 / 13.0;
    s = s * t2 + 1.0 / 11.0;
    s = s * t2 + 1.0 / 9.0;
    s = s * t2 + 1.0 / 7.0;
    s = s * t2 + 1.0 / 5.0;
    s = s * t2 + 1.0 / 3.0;
    s = s * t2 + 1.0;
    double lm = 2.0 * t * s;
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    return (double)e * LN2_HI + (lm + (double)e * LN2_LO);
}

struct EntityMark { uint64_t entity; uint64_t off_real; uint64_t off_int; };

struct Draws {
    virtual ~Draws() {}
    virtual void begin(int domain, uint64_t entity) = 0;
    virtual uint32_t next(int engine) = 0;
    /* accepted GC factor for `gc` percent: tape value, or Philox polar-normal */
    virtual double gc_factor(double mean, double sd) = 0;
    virtual bool is_tape() const = 0;
};

struct PhiloxDraws : Draws {
    uint64_t seed; int dom = 0; uint64_t ent = 0; uint64_t cur[2] = {0, 0};
    explicit PhiloxDraws(uint64_t s) : seed(s) {}
    void begin(int domain, uint64_t entity) override { dom = domain; ent = entity; cur[0] = cur[1] = 0; }
    uint32_t next(int engine) override { return philox_draw(seed, dom, engine, ent, cur[engine]++); }
    double gc_factor(double mean, double sd) override {
        /* Marsaglia polar method on two u32 draws per attempt; redraw until v >= 0
         * (Profile.cpp:1508-1511). Distributional stand-in for libstdc++'s
         * normal_distribution; uses det_log so the GPU reproduces it exactly. */
        for (;;) {
            double u1 = ((double)next(E_REAL) + 0.5) / 4294967296.0;
            double u2 = ((double)next(E_REAL) + 0.5) / 4294967296.0;
            double v1 = 2.0 * u1 - 1.0, v2 = 2.0 * u2 - 1.0;
            double s = v1 * v1 + v2 * v2;
            if (s >= 1.0 || s == 0.0) continue;
            double f = sqrt(-2.0 * det_log(s) / s);
            double v = mean + sd * (v1 * f);
            if (v >= 0) return v;
        }
    }
    bool is_tape() const override { return false; }
};

struct Tape {
    std::vector<uint32_t> v; uint64_t cur = 0; std::string name;
    void load(const std::string& path) {
        name = path;
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) { fprintf(stderr, "oracle: cannot open tape %s\n", path.c_str()); exit(2); }
        fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        v.resize(n / 4);
        if (n && fread(v.data(), 4, v.size(), f) != v.size()) { fprintf(stderr, "oracle: short tape\n"); exit(2); }
        fclose(f);
    }
    uint32_t next() {
        if (cur >= v.size()) { fprintf(stderr, "oracle: tape %s exhausted at %llu\n", name.c_str(), (unsigned long long)cur); exit(3); }
        return v[cur++];
    }
};

struct TapeDraws : Draws {
    Tape wreal, wint, mrand, mreal, mint;
    std::vector<double> gcf; uint64_t gcf_cur = 0;
    int dom = 0;
    std::vector<EntityMark> marks[D_COUNT];
    explicit TapeDraws(const std::string& prefix) {
        wreal.load(prefix + ".wreal.bin"); wint.load(prefix + ".wint.bin");
        mrand.load(prefix + ".mrand.bin"); mreal.load(prefix + ".mreal.bin"); mint.load(prefix + ".mint.bin");
        FILE* f = fopen((prefix + ".gcf.bin").c_str(), "rb");
        if (!f) { fprintf(stderr, "oracle: cannot open gcf tape\n"); exit(2); }
        fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        gcf.resize(n / 8);
        if (n && fread(gcf.data(), 8, gcf.size(), f) != gcf.size()) exit(2);
        fclose(f);
    }
    Tape& tape(int engine) {
        switch (dom) {
            case D_FRAG: case D_POIS: return mrand;
            case D_MULTM: return engine == E_REAL ? mreal : mint;
            default: return engine == E_REAL ? wreal : wint;
        }
    }
    void begin(int domain, uint64_t entity) override {
        dom = domain;
        if (domain == D_GCF) { marks[domain].push_back({entity, gcf_cur, 0}); return; }
        marks[domain].push_back({entity, tape(E_REAL).cur, tape(E_INT).cur});
    }
    uint32_t next(int engine) override { return tape(engine).next(); }
    double gc_factor(double, double) override {
        if (gcf_cur >= gcf.size()) { fprintf(stderr, "oracle: gcf tape exhausted\n"); exit(3); }
        return gcf[gcf_cur++];
    }
    bool is_tape() const override { return true; }
    bool fully_consumed() const {
        return wreal.cur == wreal.v.size() && wint.cur == wint.v.size() && mrand.cur == mrand.v.size() &&
               mreal.cur == mreal.v.size() && mint.cur == mint.v.size() && gcf_cur == gcf.size();
    }
};

/* the reference's two uniform mappings, on a u32 draw x */
inline double uni_real(uint32_t x, double start, double end) {
    return start + (end - start) * ((double)x / 4294967296.0);
}
inline long uni_int(uint32_t x, long start, long end) {
    return (long)((double)start + (double)(end - start) * ((double)x / 4294967296.0));
}

}  // namespace orc
