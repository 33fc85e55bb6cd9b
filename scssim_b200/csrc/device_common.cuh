// Device-side building blocks shared by every kernel of the genreads path:
//  * Philox4x32-10 counter-based draws keyed by (seed, domain, engine, entity) — or, in replay
//    mode, the same (entity, index) addressing into the recorded tapes of the reference;
//  * the reference's uniform mappings reduced to exact integer arithmetic
//    (/root/reference/lib/threadpool/ThreadPool.cpp:203-212);
//  * det_log: a logarithm made only of IEEE +,-,*,/ so host and device agree bit for bit;
//  * 2-bit packed genome access.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Checked build (make variant NAME=checked DEFS=-DSCS_CHECKED=1 ALL=1): every index the kernels form into shared-memory scratch,
// staging cells, slabs, pools and the packed genome is asserted. compute-sanitizer is closed on this GPU pool; the GPU suite is run
// against this build instead (profiles/NOTES_r02.md).
#ifdef SCS_CHECKED
#include <cassert>
#define SCS_CHECK(cond) assert(cond)
#else
#define SCS_CHECK(cond) ((void)0)
#endif

namespace scs {

enum Domain : int { D_FRAG = 0, D_POIS = 1, D_AMPF = 2, D_AMPS = 3, D_GCF = 4, D_MULTM = 5, D_MULTC = 6, D_READ = 7 };
enum Engine : int { E_REAL = 0, E_INT = 1 };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Out-of-line copy for the scattered single-draw call sites: every inlined Philox is ~100 instructions (1.6 KB of SASS), and the
// read kernels must stay inside the 32 KB instruction cache (ncu: 17 % no_instruction stalls with everything inlined).
__device__ __noinline__ void philox4x32_10_call(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out);

// Where an entity's draws come from. tape == nullptr -> Philox.
struct DrawSrc {
    uint64_t seed;
    const uint32_t* tape[2];   // replay: flat u32 tapes for E_REAL / E_INT of this domain
    const uint64_t* marks;     // replay: triples (entity, off_real, off_int) in entity order of the domain
};

// Per-entity stream handle: (engine, i) -> u32
struct Stream {
    uint32_t k0, k1, e0, e1, dom2;       // Philox
    const uint32_t* t[2];                // replay (already offset to the entity's first draw), or nullptr
    __device__ __forceinline__ void init(const DrawSrc& s, int domain, uint64_t entity, uint64_t mark_index) {
        k0 = (uint32_t)s.seed; k1 = (uint32_t)(s.seed >> 32);
        e0 = (uint32_t)entity; e1 = (uint32_t)(entity >> 32); dom2 = (uint32_t)(domain * 2);
        if (s.tape[0] != nullptr) {
            t[0] = s.tape[0] + s.marks[3 * mark_index + 1];
            t[1] = s.tape[1] + s.marks[3 * mark_index + 2];
        } else { t[0] = nullptr; t[1] = nullptr; }
    }
    __device__ __forceinline__ bool replay() const { return t[0] != nullptr; }
    // 4 consecutive draws i = 4*b .. 4*b+3
    __device__ __forceinline__ void block(int engine, uint32_t b, uint32_t out[4]) const {
        if (t[0] != nullptr) {
            const uint32_t* p = t[engine] + 4ull * b;
            out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3];
        } else {
            philox4x32_10(e0, e1, b, dom2 + (uint32_t)engine, k0, k1, out);
        }
    }
    __device__ __forceinline__ uint32_t at(int engine, uint32_t i) const {
        if (t[0] != nullptr) return t[engine][i];
        uint32_t o[4];
#ifdef SCS_PHILOX_OUTLINE
        philox4x32_10_call(e0, e1, i >> 2, dom2 + (uint32_t)engine, k0, k1, o);
#else
        philox4x32_10(e0, e1, i >> 2, dom2 + (uint32_t)engine, k0, k1, o);
#endif
        return o[i & 3];
    }
};

// start + (end-start) * (x / 2^32) truncated, for integer start/end with (end-start) < 2^21: exact
// in FP64 (no rounding ever happens), hence this integer form is identical to the reference's.
__host__ __device__ __forceinline__ uint32_t uni_trunc(uint32_t x, uint32_t start, uint32_t span) {
    return start + (uint32_t)(((uint64_t)span * x) >> 32);
}

// number of leading thresholds <= x in a non-decreasing row
__device__ __forceinline__ int count_le(const uint32_t* __restrict__ row, int n, uint32_t x) {
    int lo = 0, hi = n;   // invariant: row[k] <= x for k < lo ; row[k] > x for k >= hi
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (row[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__host__ __device__ inline double det_log(double x) {
    if (x <= 0.0) return -INFINITY;
    int e;
    double m = frexp(x, &e);
    if (m < 0.70710678118654752440) { m = m * 2.0; e -= 1; }
#ifdef __CUDA_ARCH__
    double t = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0));
    double t2 = __dmul_rn(t, t);
    double s = 1.0 / 27.0;
#define SCS_H(c) s = __dadd_rn(__dmul_rn(s, t2), (c))
#else
    double t = (m - 1.0) / (m + 1.0);
    double t2 = t * t;
    double s = 1.0 / 27.0;
#define SCS_H(c) s = s * t2 + (c)
#endif
    SCS_H(1.0 / 25.0); SCS_H(1.0 / 23.0); SCS_H(1.0 / 21.0); SCS_H(1.0 / 19.0); SCS_H(1.0 / 17.0);
    SCS_H(1.0 / 15.0); SCS_H(1.0 / 13.0); SCS_H(1.0 / 11.0); SCS_H(1.0 / 9.0); SCS_H(1.0 / 7.0);
    SCS_H(1.0 / 5.0); SCS_H(1.0 / 3.0); SCS_H(1.0);
#undef SCS_H
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
#ifdef __CUDA_ARCH__
    double lm = __dmul_rn(__dmul_rn(2.0, t), s);
    return __dadd_rn(__dmul_rn((double)e, LN2_HI), __dadd_rn(lm, __dmul_rn((double)e, LN2_LO)));
#else
    double lm = 2.0 * t * s;
    return (double)e * LN2_HI + (lm + (double)e * LN2_LO);
#endif
}

// ---- packed genome: 2 bits per base, 32 bases per u64 word, A=0 C=1 G=2 T=3 ----
// Bases other than A/C/G/T (N, IUPAC codes; the reference's complement maps them all to 'N',
// lib/mydefine/MyDefine.cpp:352-367) are code 4: stored as 0 in the packed words plus one bit in `nmask`
// (32 bases per u32 word). has_n == 0 (no such base in the genome) lets every kernel skip the mask.
struct Genome {
    const uint64_t* __restrict__ words;
    const uint32_t* __restrict__ nmask;
    uint64_t n_bases;
    int has_n;
};
__device__ __forceinline__ uint32_t genome_base(const Genome& g, uint64_t pos) {
    SCS_CHECK(pos < g.n_bases);
    uint32_t b = (uint32_t)(__ldg(g.words + (pos >> 5)) >> ((pos & 31) * 2)) & 3u;
    if (g.has_n && ((__ldg(g.nmask + (pos >> 5)) >> (pos & 31)) & 1u)) b = 4u;
    return b;
}
__device__ __forceinline__ uint32_t comp_code(uint32_t b) { return b < 4u ? 3u - b : 4u; }
// base i of an oriented window: rc ? complement(G[gstart - i]) : G[gstart + i]
__device__ __forceinline__ uint32_t window_base(const Genome& g, uint64_t gstart, int rc, uint32_t i) {
    return rc ? comp_code(genome_base(g, gstart - i)) : genome_base(g, gstart + i);
}

// Template descriptor shared by fragments, semi and full amplicons: an oriented genome window plus
// a sparse overlay of substitutions. d0 = gstart(40) | rc(1) | len(17) ; errors = u32 pos(17)|base(2)<<17
struct Tmpl {
    uint64_t gstart; uint32_t len; uint32_t rc;
};
__host__ __device__ __forceinline__ uint64_t pack_desc(uint64_t gstart, uint32_t rc, uint32_t len) {
    return gstart | ((uint64_t)rc << 40) | ((uint64_t)len << 41);
}
__host__ __device__ __forceinline__ Tmpl unpack_desc(uint64_t d) {
    Tmpl t; t.gstart = d & ((1ull << 40) - 1); t.rc = (uint32_t)(d >> 40) & 1u; t.len = (uint32_t)(d >> 41) & 0x1FFFFu; return t;
}
__host__ __device__ __forceinline__ uint32_t pack_err(uint32_t pos, uint32_t base) { return pos | (base << 17); }
__host__ __device__ __forceinline__ uint32_t err_pos(uint32_t e) { return e & 0x1FFFFu; }
__host__ __device__ __forceinline__ uint32_t err_base(uint32_t e) { return (e >> 17) & 3u; }

// 16 bytes starting at an arbitrary address: two aligned 16-byte loads (the second one is the neighbour thread's first,
// an L1 hit) and a word/byte shifter. The caller guarantees 31 readable bytes past p (buffers are padded).
__device__ __forceinline__ void load16(const uint8_t* p, uint32_t X[4]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint4* q = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    const uint4 v0 = __ldg(q), v1 = __ldg(q + 1);
    const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const uint32_t sh = (uint32_t)(a & 15), s8 = (sh & 3) * 8;
    uint32_t t[6], u[5];
#pragma unroll
    for (int i = 0; i < 6; i++) t[i] = (sh & 8) ? w[i + 2] : w[i];
#pragma unroll
    for (int i = 0; i < 5; i++) u[i] = (sh & 4) ? t[i + 1] : t[i];
#pragma unroll
    for (int m = 0; m < 4; m++) X[m] = __funnelshift_r(u[m], u[m + 1], s8);
}
// word m of a 16-byte vector: mask of the bytes whose index in the vector is below k (k may be <= 0 or >= 16)
__device__ __forceinline__ uint32_t below_mask(int k, int m) {
    const int km = k - 4 * m;
    return km <= 0 ? 0u : (km >= 4 ? 0xffffffffu : (1u << (8 * km)) - 1u);
}


}  // namespace scs
