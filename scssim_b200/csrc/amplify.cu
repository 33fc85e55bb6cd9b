// K1 assign_primers + K2 amplify: MALBAC whole-genome amplification on the device.
//
// Replaces Malbac::amplify / setPrimers / amplifyFrags / amplifySemiAmplicons
// (/root/reference/lib/malbac/Malbac.cpp:173-201,236-283,318-368), poissRand
// (lib/mydefine/MyDefine.cpp:69-80), Fragment::amplify (lib/fragment/Fragment.cpp:52-137),
// Amplicon::amplify (lib/amplicon/Amplicon.cpp:156-240) and the linked-list splicing
// (Amplicon.cpp:574-585, Malbac.cpp:105-141).
//
// Design: a template (fragment or semi amplicon) is an oriented genome window + a sparse overlay of
// substitutions; no sequence is ever materialised. One warp owns one template and walks its primers
// in order (they are sequentially dependent through the attached-site bitmap and the draw cursors);
// the 50 primer-site tries (and, in replay, the ~1500 per-base error draws) of each primer are evaluated
// lane-parallel with ballots picking the first event in the reference's order; the 8-mer under a site is
// two word loads of the packed genome and the GC content of a product four loads of a per-word GC index. Products go to
// per-template slots (exclusive scan of the primer counts) and a second small kernel compacts them
// into the reference's list order (reverse creation order inside a batch, batches appended).
#include <algorithm>
#include <cmath>

#include <chrono>
#include <functional>

#include "ctx.h"

namespace scs {

constexpr int kAmpMin = 1000, kAmpMax = 2000;   // Config.cpp:39-40
constexpr int kMaxErrPerAmp = 96;                // per-amplicon staging (own + inherited substitutions)
constexpr int kSiteCap = 64;                     // attached primer sites of a fragment kept as a list in shared memory; beyond
                                                 // that (large gamma) the warp switches to a bitmap in global scratch
constexpr int kFragBitmapWords = (100000 + 32) / 32 + 1;   // fragments are <= 100 kb (Fragment.cpp:15)

struct AmpParams {
    uint32_t thr_ber;        // error iff x < thr_ber   (p < ber, ber = 3.4e-4, Config.cpp:46)
    double log1m_ber;        // det_log(1 - ber): free-running streams draw the gap to the next error, floor(log(u) / log(1 - ber))
    uint64_t round_tag;      // round << 40
    uint64_t base0;          // global index of this rank's template 0 (fragments: contiguous per rank)
    uint64_t mark_base;      // replay: index of template 0's mark inside the domain's mark array
    int use_geom;            // semi amplicons: global index through the list geometry
    ListGeom geom;
};
__device__ __forceinline__ uint64_t tmpl_entity(const AmpParams& ap, uint64_t t) {
    return ap.round_tag | (ap.base0 + (ap.use_geom ? global_index(ap.geom, t) : t));
}

// sum of the lengths of the products sitting in the per-template slots
__global__ void __launch_bounds__(256) sum_len_slots_kernel(uint64_t n_tmpl, const uint32_t* __restrict__ created, const uint64_t* __restrict__ slot_off,
                                                            const uint64_t* __restrict__ tdesc, unsigned long long* __restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (t < n_tmpl) { const uint32_t m = created[t]; const uint64_t s0 = slot_off[t]; for (uint32_t i = 0; i < m; i++) v += unpack_desc(tdesc[s0 + i]).len; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

// ---------------------------------------------------------------------------------- K1
// k ~ Poisson(lambda) by Knuth's product method in log space, exactly the reference's loop.
__global__ void __launch_bounds__(256) assign_primers_kernel(DrawSrc src, AmpParams ap, const uint64_t* __restrict__ desc,
                                                             uint64_t n, double expected, double totalLen, uint32_t mask, uint32_t* __restrict__ primers,
                                                             unsigned long long* __restrict__ count) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long k = 0;
    if (t < n) {
        Tmpl T = unpack_desc(desc[t]);
        double lambda = __dmul_rn(expected, __ddiv_rn(__dmul_rn(1.0, (double)T.len), totalLen));
        Stream s; s.init(src, D_POIS, tmpl_entity(ap, t), ap.mark_base + t);
        double log1 = 0.0, log2 = -lambda; long long x = -1; uint32_t i = 0;
        do {
            uint32_t d = s.at(E_REAL, i++);
            double u = (double)d / 4294967296.0;
            log1 = __dadd_rn(log1, det_log(u));
            x++;
        } while (log1 >= log2);
        k = (unsigned long long)x;
        primers[t] = (uint32_t)k & mask;
    }
    // block reduction of k
    __shared__ unsigned long long ws[8];
    unsigned long long v = k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long s = 0; for (int w = 0; w < 8; w++) s += ws[w]; if (s) atomicAdd(count, s); }
}

// ---------------------------------------------------------------------------------- K2
__device__ __forceinline__ uint32_t tmpl_base(const Genome& g, const Tmpl& T, const uint32_t* __restrict__ errs, uint32_t nerr, uint32_t i) {
    SCS_CHECK(i < T.len);
    uint32_t b = window_base(g, T.gstart, T.rc, i);
    for (uint32_t e = 0; e < nerr; e++) { uint32_t v = errs[e]; if (err_pos(v) == i) b = err_base(v); }
    return b;
}

// Stream with the replay switch resolved at compile time: the free-running kernels carry no tape branches and the replay kernels
// no Philox (the kernel is instruction-fetch bound otherwise: ncu showed 5 no_instruction stalls per issue with both inlined).
template <bool REPLAY> struct AStream : Stream {
    __device__ __forceinline__ bool replay() const { return REPLAY; }
    __device__ __forceinline__ void block(int engine, uint32_t b, uint32_t out[4]) const {
        if (REPLAY) { const uint32_t* p = t[engine] + 4ull * b; out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3]; }
        else philox4x32_10(e0, e1, b, dom2 + (uint32_t)engine, k0, k1, out);
    }
    __device__ __forceinline__ uint32_t at(int engine, uint32_t i) const {
        if (REPLAY) return t[engine][i];
        uint32_t o[4];
        philox4x32_10(e0, e1, i >> 2, dom2 + (uint32_t)engine, k0, k1, o);
        return (i & 2u) ? ((i & 1u) ? o[3] : o[2]) : ((i & 1u) ? o[1] : o[0]);   // selects, not a dynamically indexed (local-memory) array
    }
};

// Draws ci+lane of the int engine and cr+lane of the real engine, for every lane: 32 consecutive draws of an engine lie in 9
// Philox blocks, so ONE block per lane serves both engines (lanes 0-8 the int blocks, lanes 16-24 the real blocks) and four
// shuffles per engine hand every lane its draw — instead of two whole Philox evaluations per lane of which one word each is used.
template <class SX>
__device__ __forceinline__ void warp_try_draws(const SX& S, uint32_t ci, uint32_t cr, int lane, bool live, uint32_t* xi, uint32_t* xr) {
    if (S.replay()) { *xi = live ? S.t[E_INT][ci + lane] : 0u; *xr = live ? S.t[E_REAL][cr + lane] : 0u; return; }
    const bool real_half = lane >= 16;
    uint32_t o[4];
    S.block(real_half ? E_REAL : E_INT, ((real_half ? cr : ci) >> 2) + (uint32_t)(lane & 15), o);
    const uint32_t pi = (uint32_t)lane + (ci & 3u), pr = (uint32_t)lane + (cr & 3u);   // position inside the engine's 9 blocks
    uint32_t a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { a[q] = __shfl_sync(0xffffffffu, o[q], (int)(pi >> 2)); b[q] = __shfl_sync(0xffffffffu, o[q], 16 + (int)(pr >> 2)); }
    *xi = (pi & 2u) ? ((pi & 1u) ? a[3] : a[2]) : ((pi & 1u) ? a[1] : a[0]);
    *xr = (pr & 2u) ? ((pr & 1u) ? b[3] : b[2]) : ((pr & 1u) ? b[1] : b[0]);
}

// The 8 bases [sp, sp+8) of a template as a primer-type index (first base most significant, Malbac.cpp:91-98), after the
// template's substitution overlay; false when one of them is N (such an 8-mer never binds). Two word loads instead of 8 base
// fetches (this was 27 % of the kernel's instructions).
__device__ __forceinline__ bool window_kmer8(const Genome& g, const Tmpl& T, const uint32_t* __restrict__ errs, uint32_t nerr, uint32_t sp, uint32_t* pidx) {
    SCS_CHECK(sp + 8 <= T.len);
    const uint64_t lo = T.rc ? (T.gstart - sp - 7) : (T.gstart + sp);   // genome interval [lo, lo+8)
    const uint64_t w = lo >> 5; const uint32_t k = (uint32_t)lo & 31u;
    SCS_CHECK(lo + 8 <= g.n_bases);
    uint64_t x = __ldg(g.words + w) >> (2 * k);
    if (k > 24) x |= __ldg(g.words + w + 1) << (64 - 2 * k);
    const uint32_t v = (uint32_t)x & 0xFFFFu;                            // base j of the interval at bits 2j
    uint32_t nm = 0;                                                     // bit j: base j of the interval is N
    if (g.has_n) { uint32_t m = __ldg(g.nmask + w) >> k; if (k > 24) m |= __ldg(g.nmask + w + 1) << (32 - k); nm = m & 0xFFu; }
    // window base q is interval base q (forward) or the complement of interval base 7-q (reverse); index = sum base_q 4^(7-q)
    uint32_t idx;
    if (T.rc) idx = ~v & 0xFFFFu;
    else { const uint32_t r = __brev(v) >> 16; idx = ((r & 0x5555u) << 1) | ((r >> 1) & 0x5555u); }
    for (uint32_t e = 0; e < nerr; e++) {
        const uint32_t ev = errs[e], q = err_pos(ev) - sp;               // unsigned wrap: q < 8 iff the substitution lies in the 8-mer
        if (q < 8u) { const uint32_t sh = 2u * (7u - q); idx = (idx & ~(3u << sh)) | (err_base(ev) << sh); nm &= ~(1u << (T.rc ? 7u - q : q)); }
    }
    *pidx = idx;
    return nm == 0;
}

// per-word prefix counts of the packed genome (build_gc_index, genome.cu)
struct GcIndex { const uint32_t* __restrict__ gc; const uint32_t* __restrict__ n; };

// GC count and N count of template window [s, s+l) from the GC index: the whole words between the window's first and last
// word come from two prefix entries, the two partial words from popcounts (+ overlay fix-up by the caller). Every lane
// evaluates the same four (six with N) loads. Packed N bases are stored as code 0 (= A), so they never count as GC.
__device__ __forceinline__ uint32_t window_gc_raw(const Genome& g, const GcIndex& ix, const Tmpl& T, uint32_t s, uint32_t l, uint32_t* n_count) {
    const uint64_t lo = T.rc ? (T.gstart - (s + l - 1)) : (T.gstart + s);   // genome interval [lo, lo+l)
    const uint64_t hi = lo + l;
    SCS_CHECK(l > 0 && hi <= g.n_bases);
    const uint64_t w0 = lo >> 5, w1 = (hi - 1) >> 5;
    const uint32_t k0 = (uint32_t)lo & 31u, k1 = (uint32_t)(hi - (w1 << 5));   // bases of word w0 before the window (0..31), of w1 inside it (1..32)
    const uint64_t x0 = __ldg(g.words + w0), x1 = __ldg(g.words + w1);
    const uint64_t m0 = ((x0 ^ (x0 >> 1)) & 0x5555555555555555ull) & ((1ull << (2 * k0)) - 1ull);
    const uint64_t m1 = ((x1 ^ (x1 >> 1)) & 0x5555555555555555ull) & (~0ull >> (64 - 2 * k1));
    const uint32_t cnt = __ldg(ix.gc + w1) - __ldg(ix.gc + w0) + (uint32_t)__popcll(m1) - (uint32_t)__popcll(m0);
    uint32_t nn = 0;
    if (g.has_n) {
        const uint32_t n0 = __ldg(g.nmask + w0) & ((1u << k0) - 1u), n1 = __ldg(g.nmask + w1) & (~0u >> (32 - k1));
        nn = __ldg(ix.n + w1) - __ldg(ix.n + w0) + (uint32_t)__popc(n1) - (uint32_t)__popc(n0);
    }
    *n_count = nn;
    return cnt;
}

template <bool FROM_FRAG, bool REPLAY, int BITMAP_WORDS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) amplify_kernel(Genome g, GcIndex gcx, DrawSrc src, AmpParams ap, uint64_t n_tmpl, const uint64_t* __restrict__ desc,
                                                             const uint32_t* __restrict__ primers, const uint64_t* __restrict__ errref,
                                                             const uint64_t* __restrict__ slot_off, uint64_t* __restrict__ out_desc,
                                                             uint32_t* __restrict__ out_gc, uint64_t* __restrict__ out_errref,
                                                             uint32_t* __restrict__ created, uint32_t* err_pool, unsigned long long* err_top,
                                                             uint64_t err_cap, int* __restrict__ flags, long long* primer_counts,
                                                             unsigned long long* __restrict__ ticket, uint32_t* __restrict__ gbitmaps, uint32_t lane_cap) {
    // lane_cap != 0: templates with at most that many primers were amplified by amplify_semis_lanes_kernel — skip them
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // attached-site set (posAttached[], Fragment.cpp:70): semi templates (<= 2000 bp) use a bitmap in shared memory; fragments
    // keep a short list of sites in shared memory and fall back to a per-warp bitmap in global memory for many primers
    uint32_t* bitmap = smem + (size_t)warp * BITMAP_WORDS;
    uint32_t* gbitmap = FROM_FRAG ? gbitmaps + ((size_t)blockIdx.x * WARPS + warp) * kFragBitmapWords : nullptr;
    uint32_t* errbuf = smem + (size_t)WARPS * BITMAP_WORDS + (size_t)warp * kMaxErrPerAmp;
    for (;;) {
        // dynamic work: fragments (every one has primers, each worth thousands of draws) are handed out one per warp;
        // semi amplicons (mostly zero primers) in chunks of 32 consecutive templates filtered by a ballot
        constexpr uint64_t kChunk = FROM_FRAG ? 1 : 32;
        unsigned long long chunk = 0;
        if (lane == 0) chunk = atomicAdd(ticket, 1ull);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        uint64_t t0 = chunk * kChunk;
        if (t0 >= n_tmpl) break;
        uint64_t tl = t0 + lane;
        uint32_t myp = 0; uint64_t myd = 0;
        if (lane < (int)kChunk && tl < n_tmpl) { myp = primers[tl]; myd = desc[tl]; if (unpack_desc(myd).len < (uint32_t)(kAmpMin + 27)) myp = 0; }
        if (lane_cap != 0 && myp <= lane_cap) myp = 0;
        else if (lane < (int)kChunk && tl < n_tmpl && myp == 0) created[tl] = 0;
        uint32_t work = __ballot_sync(0xffffffffu, myp != 0);
        while (work) {
            int src_lane = __ffs(work) - 1; work &= work - 1;
            const uint64_t t = t0 + src_lane;
            const uint32_t primerNum = __shfl_sync(0xffffffffu, myp, src_lane);
            const Tmpl T = unpack_desc(__shfl_sync(0xffffffffu, myd, src_lane));
            const uint32_t* terr = nullptr; uint32_t tnerr = 0;
            if (!FROM_FRAG) { uint64_t er = errref[t]; tnerr = (uint32_t)(er & 0xFFFF); terr = err_pool + (er >> 16); }
            const uint64_t slot0 = slot_off[t];
            AStream<REPLAY> S; S.init(src, FROM_FRAG ? D_AMPF : D_AMPS, tmpl_entity(ap, t), ap.mark_base + t);
            const bool use_list = FROM_FRAG && primerNum <= (uint32_t)kSiteCap;
            if (!use_list) {
                uint32_t* bm = FROM_FRAG ? gbitmap : bitmap;
                const uint32_t bw = (T.len + 31) >> 5;
                for (uint32_t w = lane; w < bw; w += 32) bm[w] = 0;
            }
            __syncwarp();
            uint32_t ci = 0, cr = 0, made = 0;
            for (uint32_t pi = 0; pi < primerNum; pi++) {
                // ---- primer site: up to 50 tries (Fragment.cpp:73-95); try k uses int draw ci+k-1 and real draw cr+k-1
                uint32_t spos = 0, alen = 0; int acc = 0;
                for (uint32_t tb = 0; tb < 50 && !acc; tb += 32) {
                    uint32_t k = tb + lane;   // 0-based try
                    bool ok = false; uint32_t sp = 0, al = 0; uint32_t pidx = 0;
                    uint32_t xi, xr;
                    warp_try_draws(S, ci + tb, cr + tb, lane, k < 50, &xi, &xr);
                    if (k < 50) {
                        sp = uni_trunc(xi, 27, T.len - 27);
                        al = uni_trunc(xr, kAmpMin, kAmpMax + 1 - kAmpMin);
                        ok = (sp + al <= T.len);
                        if (ok) {
                            if (use_list) { for (uint32_t q = 0; q < made; q++) ok &= (bitmap[q] != sp); }
                            else { const uint32_t* bm = FROM_FRAG ? gbitmap : bitmap; ok = !((bm[sp >> 5] >> (sp & 31)) & 1u); }
                        }
                        if (ok) ok = window_kmer8(g, T, terr, tnerr, sp, &pidx);   // 8-mers holding an N never bind
                        if (ok) ok = primer_counts[pidx] > 0;
                    }
                    uint32_t cand = __ballot_sync(0xffffffffu, ok);
                    while (cand && !acc) {
                        int wl = __ffs(cand) - 1; cand &= cand - 1;
                        int got = 0;
                        if (lane == wl) {   // updatePrimerCount(s, -1), Malbac.cpp:91-103
                            long long old = atomicAdd((unsigned long long*)&primer_counts[pidx], (unsigned long long)-1ll);
                            if (old > 0) got = 1; else atomicAdd((unsigned long long*)&primer_counts[pidx], 1ull);
                        }
                        got = __shfl_sync(0xffffffffu, got, wl);
                        if (got) { acc = 1; spos = __shfl_sync(0xffffffffu, sp, wl); alen = __shfl_sync(0xffffffffu, al, wl); ci += tb + wl + 1; cr += tb + wl + 1; }
                    }
                }
                if (!acc) { ci += 51; cr += 51; break; }   // 51st try draws, then the template is abandoned
                if (lane == 0) {
                    SCS_CHECK(!use_list || made < (uint32_t)BITMAP_WORDS);
                    if (use_list) bitmap[made] = spos;   // every accepted primer yields exactly one product: `made` indexes the list
                    else { uint32_t* bm = FROM_FRAG ? gbitmap : bitmap; bm[spos >> 5] |= 1u << (spos & 31); }
                }
                // ---- GC content of the window (countGC, MyDefine.cpp:434-452)
                uint32_t nN = 0;
                int gc = (int)window_gc_raw(g, gcx, T, spos, alen, &nN);
                if (!FROM_FRAG) for (uint32_t e = 0; e < tnerr; e++) {
                    uint32_t v = terr[e], p = err_pos(v);
                    if (p >= spos && p < spos + alen) {
                        uint32_t raw = window_base(g, T.gstart, T.rc, p), nb = err_base(v);
                        gc += (int)((nb == 1u) | (nb == 2u)) - (int)((raw == 1u) | (raw == 2u));
                        if (raw == 4u) nN--;   // the substitution replaced an N
                    }
                }
                if (nN) gc = 0;   // countGC() gives 0 as soon as the window holds an N (MyDefine.cpp:448-450)
                // ---- polymerase errors at positions j = 8 .. alen-1, each with probability ber (Fragment.cpp:105-123)
                uint32_t nown = 0;
                if (!S.replay()) {
                    // free-running streams: the distance to the next error is drawn directly, P(gap = g) = (1-ber)^g * ber — the same
                    // distribution as the reference's one Bernoulli draw per base, at one draw per error (+1) instead of ~1500 per
                    // product. Every lane evaluates the same draws (det_log: bit-identical to the CPU oracle).
                    uint64_t j = 7;
                    for (;;) {
                        const double u = __ddiv_rn(__dadd_rn((double)S.at(E_REAL, cr++), 0.5), 4294967296.0);
                        const double gd = floor(__ddiv_rn(det_log(u), ap.log1m_ber));
                        if (!(gd < (double)alen)) break;
                        j += 1 + (uint64_t)gd;
                        if (j >= alen) break;
                        const uint32_t base = tmpl_base(g, T, terr, tnerr, spos + (uint32_t)j);
                        uint32_t nb;
                        if (FROM_FRAG) { do { nb = S.at(E_INT, ci++) >> 30; } while (nb == base); }     // Fragment.cpp:110
                        else { do { nb = S.at(E_REAL, cr++) >> 30; } while (nb == base); }               // Amplicon.cpp:213
                        gc += (int)((nb == 1u) | (nb == 2u)) - (int)((base == 1u) | (base == 2u));
                        if (nown < kMaxErrPerAmp) { if (lane == 0) errbuf[nown] = pack_err((uint32_t)j, nb); } else if (lane == 0) atomicOr(flags, 1);
                        nown++;
                    }
                } else {
                // replay: the reference's own consumption, one draw per base. Draw d on the real stream belongs to position j = 8 + (d - dbase)
                uint32_t dbase = cr, dcur = cr; const uint32_t dend_pos = alen;   // position limit
                for (;;) {
                    uint32_t jcur = 8 + (dcur - dbase);
                    if (jcur >= dend_pos) break;
                    uint32_t blk = (dcur >> 2) + lane; uint32_t o[4];
                    S.block(E_REAL, blk, o);
                    uint32_t hit = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        uint32_t d = blk * 4 + q;
                        if (d >= dcur && (8 + (d - dbase)) < dend_pos && o[q] < ap.thr_ber) hit |= 1u << q;
                    }
                    uint32_t any = __ballot_sync(0xffffffffu, hit != 0);
                    if (!any) { dcur = ((dcur >> 2) + 32) << 2; continue; }
                    int hl = __ffs(any) - 1;
                    uint32_t hq = __shfl_sync(0xffffffffu, hit, hl);
                    uint32_t dh = ((dcur >> 2) + hl) * 4 + (__ffs(hq) - 1);
                    uint32_t j = 8 + (dh - dbase);
                    uint32_t base = tmpl_base(g, T, terr, tnerr, spos + j);
                    uint32_t nb, extra = 0;
                    if (FROM_FRAG) { do { nb = S.at(E_INT, ci++) >> 30; } while (nb == base); }               // Fragment.cpp:110
                    else { do { nb = S.at(E_REAL, dh + 1 + extra) >> 30; extra++; } while (nb == base); }       // Amplicon.cpp:213
                    gc += (int)((nb == 1u) | (nb == 2u)) - (int)((base == 1u) | (base == 2u));
                    if (nown < kMaxErrPerAmp) { if (lane == 0) errbuf[nown] = pack_err(j, nb); } else if (lane == 0) atomicOr(flags, 1);
                    nown++;
                    dcur = dh + 1 + extra; dbase += extra;
                }
                cr = dbase + (alen - 8);
                }
                if (nown > kMaxErrPerAmp) nown = kMaxErrPerAmp;
                __syncwarp();
                // ---- emit the product into slot (slot0 + made)
                uint64_t ngstart; uint32_t nrc;
                if (FROM_FRAG) {   // semi-amplicon template U = reverse complement of the copied window
                    nrc = T.rc ^ 1u; ngstart = T.rc ? (T.gstart - spos - alen + 1) : (T.gstart + spos + alen - 1);
                } else {           // full amplicon = window of U
                    nrc = T.rc; ngstart = T.rc ? (T.gstart - spos) : (T.gstart + spos);
                }
                // error overlay of the product: inherited (semi errors inside the window) then own
                uint32_t ninh = 0;
                if (!FROM_FRAG) {
                    for (uint32_t e = 0; e < tnerr; e++) {
                        uint32_t v = terr[e], p = err_pos(v);
                        if (p >= spos && p < spos + alen) {
                            bool over = false;
                            for (uint32_t q = 0; q < nown; q++) over |= (err_pos(errbuf[q]) == p - spos);
                            if (!over) { if (nown + ninh < kMaxErrPerAmp) { if (lane == 0) errbuf[nown + ninh] = pack_err(p - spos, err_base(v)); ninh++; } else if (lane == 0) atomicOr(flags, 1); }
                        }
                    }
                    __syncwarp();
                }
                uint32_t ntot = nown + ninh;
                unsigned long long eoff = 0;
                if (ntot) {
                    if (lane == 0) eoff = atomicAdd(err_top, (unsigned long long)ntot);
                    eoff = __shfl_sync(0xffffffffu, eoff, 0);
                    if (eoff + ntot > err_cap) { if (lane == 0) atomicOr(flags, 2); ntot = 0; eoff = 0; }
                    for (uint32_t q = lane; q < ntot; q += 32) {
                        SCS_CHECK(q < (uint32_t)kMaxErrPerAmp && eoff + q < err_cap);
                        uint32_t v = errbuf[q];
                        // own errors of a semi are recorded in copied-window coordinates; its template is the reverse complement
                        if (FROM_FRAG) v = pack_err(alen - 1 - err_pos(v), 3u - err_base(v));
                        err_pool[eoff + q] = v;
                    }
                }
                if (lane == 0) {
                    uint64_t slot = slot0 + made;
                    SCS_CHECK(made < primerNum && slot < slot_off[t] + primers[t] && spos + alen <= T.len && spos >= 27 && alen >= (uint32_t)kAmpMin && alen <= (uint32_t)kAmpMax);
                    out_desc[slot] = pack_desc(ngstart, nrc, alen);
                    out_gc[slot] = (uint32_t)max(0, gc);
                    out_errref[slot] = ((uint64_t)eoff << 16) | ntot;
                }
                made++;
                __syncwarp();
            }
            if (lane == 0) created[t] = made;
        }
    }
}

// ---------------------------------------------------------------------------------- K2b
// Free-running amplification of semi amplicons with one LANE per template (the warp-per-template kernel above spends ~500 warp
// instructions per primer, most of them 32-fold redundant: a semi amplicon has ~6 primers, each a short search and one product).
// Every lane walks its own template through the same sequence of draws as the sequential code; a warp alternates between
//   * search steps — every searching lane evaluates its next primer-site try (Fragment.cpp:73-95), and
//   * product steps — run once half of the warp's lanes hold an accepted site: GC content, polymerase errors, inherited
//     substitutions and the product record, all lane-parallel over different templates,
// and refills finished lanes from a per-warp queue of templates (chunks of 32 consecutive templates, those with primers only).
// The attached sites of a template are a list in shared memory (one product per accepted site); templates with more primers than
// the list holds are left to the warp kernel (flag 4 tells the host that there are any). Results are identical to the warp
// kernel's: the draws are addressed by (template, engine, index), not by who evaluates them.
constexpr int kLaneSites = 24;
#ifdef SCS_LANE_CTAS        // A/B builds: resident CTAs per SM the lane kernel is compiled and launched for (profiles/NOTES_r02.md)
#define SCS_LANE_BOUNDS __launch_bounds__(kLaneWarps * 32, SCS_LANE_CTAS)
#else                       // shipped: 64 registers without a cap, 4 CTAs per SM
#define SCS_LANE_CTAS 4
#define SCS_LANE_BOUNDS __launch_bounds__(kLaneWarps * 32)
#endif
#ifndef SCS_LANE_PEND_NUM   // a product step runs once NUM/DEN of the lanes in work hold an accepted site (A/B: profiles/NOTES_r02.md)
#define SCS_LANE_PEND_NUM 1
#define SCS_LANE_PEND_DEN 2
#endif
constexpr int kLaneWarps = 8;

__global__ void SCS_LANE_BOUNDS amplify_semis_lanes_kernel(Genome g, GcIndex gcx, DrawSrc src, AmpParams ap, uint64_t n_tmpl, const uint64_t* __restrict__ desc,
                                                                              const uint32_t* __restrict__ primers, const uint64_t* __restrict__ errref,
                                                                              const uint64_t* __restrict__ slot_off, uint64_t* __restrict__ out_desc,
                                                                              uint32_t* __restrict__ out_gc, uint64_t* __restrict__ out_errref,
                                                                              uint32_t* __restrict__ created, uint32_t* err_pool, unsigned long long* err_top,
                                                                              uint64_t err_cap, int* __restrict__ flags, long long* primer_counts,
                                                                              unsigned long long* __restrict__ ticket) {
    __shared__ uint16_t sites[kLaneSites][kLaneWarps * 32];   // [site][thread]: a warp's accesses to one site index are conflict-free
    __shared__ uint32_t queue[kLaneWarps][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t ebuf[kMaxErrPerAmp];   // substitutions of the product in work (local memory; ~1 entry used per product)
    // lane state
    bool active = false, pending = false;
    uint64_t t = 0, slot0 = 0, er = 0; Tmpl T{0, 0, 0}; uint32_t primerNum = 0, pi = 0, k = 0, ci = 0, cr = 0, made = 0, spos = 0, alen = 0;
    AStream<false> S; S.init(src, D_AMPS, 0, 0);
    // warp state
    uint32_t qhead = 0, qcount = 0; bool drained = false;
    for (;;) {
        // ---- refill idle lanes from the queue; top the queue up from the template list
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle) {
            const uint32_t want = (uint32_t)__popc(idle);
            while (!drained && qcount < want) {
                unsigned long long chunk = 0;
                if (lane == 0) chunk = atomicAdd(ticket, 1ull);
                chunk = __shfl_sync(0xffffffffu, chunk, 0);
                const uint64_t tl = chunk * 32ull + (uint64_t)lane;
                if (chunk * 32ull >= n_tmpl) { drained = true; break; }
                uint32_t p = 0;
                if (tl < n_tmpl) {
                    p = primers[tl];
                    if (unpack_desc(desc[tl]).len < (uint32_t)(kAmpMin + 27)) p = 0;
                    if (p == 0) created[tl] = 0;
                    else if (p > (uint32_t)kLaneSites) { p = 0; atomicOr(flags, 4); }   // the warp kernel's
                }
                const uint32_t has = __ballot_sync(0xffffffffu, p != 0);
                if (p != 0) queue[warp][(qhead + qcount + (uint32_t)__popc(has & lt_mask)) & 63u] = (uint32_t)tl;
                qcount += (uint32_t)__popc(has);
            }
            __syncwarp();
            const uint32_t r = (uint32_t)__popc(idle & lt_mask);
            if (!active && r < qcount) {
                t = queue[warp][(qhead + r) & 63u];
                T = unpack_desc(desc[t]); primerNum = primers[t]; er = errref[t]; slot0 = slot_off[t];
                const uint64_t ent = tmpl_entity(ap, t);
                S.e0 = (uint32_t)ent; S.e1 = (uint32_t)(ent >> 32);
                pi = 0; k = 0; ci = 0; cr = 0; made = 0; active = true; pending = false;
            }
            const uint32_t taken = min(qcount, want);
            qhead += taken; qcount -= taken;
            __syncwarp();
        }
        const uint32_t act = __ballot_sync(0xffffffffu, active);
        if (!act) break;   // nothing in work, nothing queued (the refill loop ran until the list was drained)
        const uint32_t* terr = err_pool + (er >> 16); const uint32_t tnerr = (uint32_t)(er & 0xFFFF);
        // ---- search step: try k of the current primer uses int draw ci+k and real draw cr+k
        if (active && !pending) {
            const uint32_t sp = uni_trunc(S.at(E_INT, ci + k), 27, T.len - 27);
            const uint32_t al = uni_trunc(S.at(E_REAL, cr + k), kAmpMin, kAmpMax + 1 - kAmpMin);
            bool ok = (sp + al <= T.len);
            if (ok) for (uint32_t q = 0; q < made; q++) ok &= ((uint32_t)sites[q][tid] != sp);
            uint32_t pidx = 0;
            if (ok) ok = window_kmer8(g, T, terr, tnerr, sp, &pidx);
            if (ok) ok = primer_counts[pidx] > 0;
            if (ok) {   // updatePrimerCount(s, -1), Malbac.cpp:91-103
                const long long old = atomicAdd((unsigned long long*)&primer_counts[pidx], (unsigned long long)-1ll);
                if (old <= 0) { atomicAdd((unsigned long long*)&primer_counts[pidx], 1ull); ok = false; }
            }
            if (ok) {
                SCS_CHECK(made < (uint32_t)kLaneSites);
                sites[made][tid] = (uint16_t)sp; spos = sp; alen = al; ci += k + 1; cr += k + 1; pending = true;
            } else if (++k == 50) {   // the 51st try draws, then the template is abandoned (Fragment.cpp:88-95)
                created[t] = made; active = false;
            }
        }
        // ---- product step, once half of the lanes in work hold an accepted site
        const uint32_t pend = __ballot_sync(0xffffffffu, pending);
        const uint32_t act2 = __ballot_sync(0xffffffffu, active);
        if (pend == 0 || SCS_LANE_PEND_DEN * __popc(pend) < SCS_LANE_PEND_NUM * __popc(act2)) continue;
        uint32_t ntot = 0; int gc = 0;
        if (pending) {
            // GC content of the window (countGC, MyDefine.cpp:434-452)
            uint32_t nN = 0;
            gc = (int)window_gc_raw(g, gcx, T, spos, alen, &nN);
            for (uint32_t e = 0; e < tnerr; e++) {
                const uint32_t v = terr[e], p = err_pos(v);
                if (p >= spos && p < spos + alen) {
                    const uint32_t raw = window_base(g, T.gstart, T.rc, p), nb = err_base(v);
                    gc += (int)((nb == 1u) | (nb == 2u)) - (int)((raw == 1u) | (raw == 2u));
                    if (raw == 4u) nN--;
                }
            }
            if (nN) gc = 0;
            // polymerase errors: gaps between errors drawn directly (see amplify_kernel), substitutes from the real engine (Amplicon.cpp:213)
            uint32_t nown = 0; uint64_t j = 7;
            for (;;) {
                const double u = __ddiv_rn(__dadd_rn((double)S.at(E_REAL, cr++), 0.5), 4294967296.0);
                const double gd = floor(__ddiv_rn(det_log(u), ap.log1m_ber));
                if (!(gd < (double)alen)) break;
                j += 1 + (uint64_t)gd;
                if (j >= alen) break;
                const uint32_t base = tmpl_base(g, T, terr, tnerr, spos + (uint32_t)j);
                uint32_t nb;
                do { nb = S.at(E_REAL, cr++) >> 30; } while (nb == base);
                gc += (int)((nb == 1u) | (nb == 2u)) - (int)((base == 1u) | (base == 2u));
                if (nown < (uint32_t)kMaxErrPerAmp) ebuf[nown] = pack_err((uint32_t)j, nb); else atomicOr(flags, 1);
                nown++;
            }
            if (nown > (uint32_t)kMaxErrPerAmp) nown = kMaxErrPerAmp;
            // inherited substitutions of the template inside the window, unless overwritten by an own one
            uint32_t ninh = 0;
            for (uint32_t e = 0; e < tnerr; e++) {
                const uint32_t v = terr[e], p = err_pos(v);
                if (p >= spos && p < spos + alen) {
                    bool over = false;
                    for (uint32_t q = 0; q < nown; q++) over |= (err_pos(ebuf[q]) == p - spos);
                    if (!over) { if (nown + ninh < (uint32_t)kMaxErrPerAmp) { ebuf[nown + ninh] = pack_err(p - spos, err_base(v)); ninh++; } else atomicOr(flags, 1); }
                }
            }
            ntot = nown + ninh;
        }
        // error pool space for the warp's products: one atomic for all of them
        uint32_t incl = ntot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const uint32_t wtot = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long wbase = 0;
        if (wtot) { if (lane == 31) wbase = atomicAdd(err_top, (unsigned long long)wtot); wbase = __shfl_sync(0xffffffffu, wbase, 31); }
        if (pending) {
            unsigned long long eoff = wbase + (incl - ntot);
            if (ntot && eoff + ntot > err_cap) { atomicOr(flags, 2); ntot = 0; }
            if (!ntot) eoff = 0;
            for (uint32_t q = 0; q < ntot; q++) { SCS_CHECK(eoff + q < err_cap); err_pool[eoff + q] = ebuf[q]; }
            const uint64_t slot = slot0 + made;
            SCS_CHECK(made < primerNum && spos + alen <= T.len && spos >= 27 && alen >= (uint32_t)kAmpMin && alen <= (uint32_t)kAmpMax);
            out_desc[slot] = pack_desc(T.rc ? (T.gstart - spos) : (T.gstart + spos), T.rc, alen);   // full amplicon = window of the semi
            out_gc[slot] = (uint32_t)max(0, gc);
            out_errref[slot] = ((uint64_t)eoff << 16) | ntot;
            made++; pi++; k = 0; pending = false;
            if (pi == primerNum) { created[t] = made; active = false; }
        }
    }
}

// move products from per-template slots into list order: batch position = total-1-(creation rank)
__global__ void __launch_bounds__(256) compact_products_kernel(uint64_t n_tmpl, const uint32_t* __restrict__ created, const uint64_t* __restrict__ cprefix,
                                                               const uint64_t* __restrict__ slot_off, const uint64_t* __restrict__ tdesc,
                                                               const uint32_t* __restrict__ tgc, const uint64_t* __restrict__ terr, uint64_t local_total,
                                                               uint64_t* __restrict__ ldesc, uint32_t* __restrict__ lgc, uint64_t* __restrict__ lerr,
                                                               uint32_t* __restrict__ lprimers, uint64_t list_base) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tmpl) return;
    uint32_t m = created[t];
    if (!m) return;
    uint64_t s0 = slot_off[t], r0 = cprefix[t];
    for (uint32_t i = 0; i < m; i++) {
        uint64_t dst = list_base + (local_total - 1 - (r0 + i));
        ldesc[dst] = tdesc[s0 + i]; lgc[dst] = tgc[s0 + i]; lerr[dst] = terr[s0 + i];
        if (lprimers) lprimers[dst] = 0;
    }
}

// ---------------------------------------------------------------------------------- host driver
namespace {

struct Round {
    scs_ctx* c; Genome g; uint32_t thr_ber;
    uint64_t pending_count = 0, semi_len_sum_global = 0;
    // scratch lives in the context: reused by every pass and every run
    DevBuf<unsigned long long>&dcount, &ticket; DevBuf<int>& flags;
    DevBuf<uint64_t>&slot_off, &cprefix, &tdesc, &terr; DevBuf<uint32_t>&tgc, &created, &gbitmaps;
    explicit Round(scs_ctx* ctx) : c(ctx), dcount(ctx->ascratch.dcount), ticket(ctx->ascratch.ticket), flags(ctx->ascratch.flags), slot_off(ctx->ascratch.slot_off),
                                   cprefix(ctx->ascratch.cprefix), tdesc(ctx->ascratch.tdesc), terr(ctx->ascratch.terr), tgc(ctx->ascratch.tgc),
                                   created(ctx->ascratch.created), gbitmaps(ctx->ascratch.gbitmaps) {}

    int allreduce_u64(uint64_t* v, size_t n) { return scs::allreduce_u64(c, v, n); }

    AmpParams params(int round, bool semis, int domain) {
        AmpParams ap{}; ap.thr_ber = thr_ber; ap.log1m_ber = det_log(1.0 - 3.4e-4); ap.round_tag = (uint64_t)round << 40;
        if (semis) { ap.use_geom = 1; ap.base0 = 0; ap.geom = c->semi_geom; ap.mark_base = mark_base(domain, round); }
        else { ap.use_geom = 0; ap.base0 = c->frag_global0; ap.mark_base = mark_base(domain, round) + c->frag_global0; }
        return ap;
    }

    // Malbac::setPrimers (Malbac.cpp:236-283)
    int set_primers(bool onlyFrags, int round) {
        uint64_t nF = c->frags.size(), nS = onlyFrags ? 0 : c->semis.n;
        // template count and total length over all ranks (lengths are integers: the FP64 sum is exact in any order); the
        // cell-wide semi-amplicon count and length are kept up to date by pass() — no collective here
        uint64_t nS_global = 0; for (int b = 0; b < c->semi_geom.nb; b++) nS_global += c->semi_geom.gtot[b];
        uint64_t templateNum = c->n_frags_global + (onlyFrags ? 0 : nS_global);
        double totalLen = (double)(c->frag_len_sum_global + (onlyFrags ? 0 : semi_len_sum_global));
        uint64_t expected = (uint64_t)((double)c->total_primers * c->P.gamma * (double)templateNum);
        SCS_CUDA(c, cudaMemsetAsync(dcount.p, 0, 8, c->st));
        // global template index: fragments first, then semis in (global) list order
        if (nF) {
            AmpParams ap = params(round, false, D_POIS);
            assign_primers_kernel<<<(unsigned)((nF + 255) / 256), 256, 0, c->st>>>(draw_src(c, D_POIS), ap, c->frag_desc.p, nF, (double)expected, totalLen,
                                                                                  0xFFFFFFFFu, c->frag_primers.p, dcount.p);
            SCS_LAUNCHED(c);
        }
        if (nS) {
            AmpParams ap = params(round, true, D_POIS);
            ap.base0 = c->n_frags_global; ap.mark_base += c->n_frags_global;
            assign_primers_kernel<<<(unsigned)((nS + 255) / 256), 256, 0, c->st>>>(draw_src(c, D_POIS), ap, c->semis.desc.p, nS, (double)expected, totalLen, 0xFFFu,
                                                                                  c->semis.primers.p, dcount.p);
            SCS_LAUNCHED(c);
        }
        uint64_t count = 0;
        SCS_CUDA(c, cudaMemcpyAsync(&count, dcount.p, 8, cudaMemcpyDeviceToHost, c->st));
        SCS_CUDA(c, cudaStreamSynchronize(c->st));
        pending_count += count;   // summed over ranks and charged to the primer budget by the next pass()'s exchange
        return SCS_OK;
    }

    // index of the first mark of `round` inside the replay mark array of a domain
    uint64_t mark_base(int domain, int round) {
        if (!c->replay.on) return 0;
        const std::vector<uint64_t>& m = c->replay.hmarks[domain];
        uint64_t n = m.size() / 3, lo = 0, hi = n;   // marks are sorted by entity = round<<40 | index
        uint64_t key = (uint64_t)round << 40;
        while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (m[3 * mid] < key) lo = mid + 1; else hi = mid; }
        return lo;
    }

    // one amplification pass over `n` templates; products appended to `dst`
    template <bool FROM_FRAG>
    int pass(int round, uint64_t n, const uint64_t* desc, const uint32_t* primers, const uint64_t* errref, AmpList& dst, ListGeom& geom) {
        const bool trace = getenv("SCS_TRACE") != nullptr;
        auto now_ms = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double tr0 = now_ms(); double tr_alloc = 0, tr_kernel = 0, tr_scan = 0;
        uint64_t total_slots = 0;
        SCS_CUDA(c, slot_off.reserve(n + 1)); SCS_CUDA(c, cprefix.reserve(n + 1)); SCS_CUDA(c, created.reserve(n + 1));
        uint64_t made_total = 0;
        if (n) {
            if (int rc = exclusive_scan_u32(c, primers, slot_off.p, n, &total_slots)) return rc;
            SCS_CUDA(c, tdesc.reserve(total_slots + 1)); SCS_CUDA(c, terr.reserve(total_slots + 1)); SCS_CUDA(c, tgc.reserve(total_slots + 1));
            // error pool: ~0.5 own + <= 0.5 inherited substitutions per product expected (ber x mean length 1500); 2 slots per
            // product + 64 k reserved (a pass of millions of products concentrates tightly around its mean), overflow is reported
            uint64_t etop = 0;
            SCS_CUDA(c, cudaMemcpyAsync(&etop, c->err_top.p, 8, cudaMemcpyDeviceToHost, c->st));
            SCS_CUDA(c, cudaStreamSynchronize(c->st));
            uint64_t need = etop + total_slots * 2 + 65536;
            SCS_CUDA(c, c->err_pool.reserve(need, etop, c->st));
            SCS_CUDA(c, cudaMemsetAsync(flags.p, 0, 4, c->st)); SCS_CUDA(c, cudaMemsetAsync(ticket.p, 0, 8, c->st));
            AmpParams ap = params(round, !FROM_FRAG, FROM_FRAG ? D_AMPF : D_AMPS);
            if (trace) { cudaStreamSynchronize(c->st); tr_alloc = now_ms(); }
            int dev = 0; cudaGetDevice(&dev); int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const GcIndex gcx{c->gc_pref.p, c->genome_has_n ? c->n_pref.p : nullptr};
            const bool replay = c->replay.on;
            // free-running semi-amplicon passes: one lane per template (amplify_semis_lanes_kernel), then the warp kernel for the
            // templates with more primers than a lane's site list holds, if there are any
            const bool lanes = !FROM_FRAG && !replay && n < (1ull << 32);
            std::function<void()> semis_warp_pass;
            if (FROM_FRAG) {
                constexpr int W = 8, BW = kSiteCap, CTAS = 5;
                size_t sm = (size_t)W * (BW + kMaxErrPerAmp) * 4;
                SCS_CUDA(c, gbitmaps.reserve((size_t)sms * CTAS * W * kFragBitmapWords));
                auto kern = replay ? amplify_kernel<true, true, BW, W> : amplify_kernel<true, false, BW, W>;
                kern<<<sms * CTAS, W * 32, sm, c->st>>>(g, gcx, draw_src(c, D_AMPF), ap, n, desc, primers, errref, slot_off.p, tdesc.p, tgc.p, terr.p, created.p,
                                                        c->err_pool.p, c->err_top.p, c->err_pool.cap, flags.p, c->primer_counts.p, ticket.p, gbitmaps.p, 0u);
            } else {
                constexpr int W = 8, BW = (kAmpMax + 32) / 32 + 1;
                size_t sm = (size_t)W * (BW + kMaxErrPerAmp) * 4;
                auto kern = replay ? amplify_kernel<false, true, BW, W> : amplify_kernel<false, false, BW, W>;
                semis_warp_pass = [=]() {
                    kern<<<sms * 8, W * 32, sm, c->st>>>(g, gcx, draw_src(c, D_AMPS), ap, n, desc, primers, errref, slot_off.p, tdesc.p, tgc.p, terr.p, created.p,
                                                         c->err_pool.p, c->err_top.p, c->err_pool.cap, flags.p, c->primer_counts.p, ticket.p, nullptr, lanes ? (uint32_t)kLaneSites : 0u);
                };
                if (lanes) amplify_semis_lanes_kernel<<<sms * SCS_LANE_CTAS, kLaneWarps * 32, 0, c->st>>>(g, gcx, draw_src(c, D_AMPS), ap, n, desc, primers, errref, slot_off.p, tdesc.p, tgc.p,
                                                                                             terr.p, created.p, c->err_pool.p, c->err_top.p, c->err_pool.cap, flags.p,
                                                                                             c->primer_counts.p, ticket.p);
                else semis_warp_pass();
            }
            SCS_LAUNCHED(c);
            // while the kernel runs: grow the destination list for the most it can produce (one product per primer). Mapping
            // fresh device memory costs 1-100 ms per GB on this platform (erratically); the list grows in place (VMM), so this is
            // safe under a running kernel and leaves nothing to allocate after it.
            {
                const uint64_t bound = dst.n + total_slots + 1;
                SCS_CUDA(c, dst.desc.reserve(bound, dst.n, c->st)); SCS_CUDA(c, dst.gc.reserve(bound, dst.n, c->st)); SCS_CUDA(c, dst.errref.reserve(bound, dst.n, c->st));
                if (FROM_FRAG) SCS_CUDA(c, dst.primers.reserve(bound, dst.n, c->st));
            }
            int hflags = 0;
            SCS_CUDA(c, cudaMemcpyAsync(&hflags, flags.p, 4, cudaMemcpyDeviceToHost, c->st));
            SCS_CUDA(c, cudaStreamSynchronize(c->st));
            if (lanes && (hflags & 4)) {
                SCS_CUDA(c, cudaMemsetAsync(ticket.p, 0, 8, c->st));
                semis_warp_pass(); SCS_LAUNCHED(c);
                SCS_CUDA(c, cudaMemcpyAsync(&hflags, flags.p, 4, cudaMemcpyDeviceToHost, c->st));
                SCS_CUDA(c, cudaStreamSynchronize(c->st));
            }
            tr_kernel = now_ms();
            if (hflags & 2) return c->fail(SCS_E_NOMEM, "amplify: error pool overflow");
            if (hflags & 1) return c->fail(SCS_E_UNSUPPORTED, "amplify: more than 96 substitutions on one amplicon");
            if (int rc = exclusive_scan_u32(c, created.p, cprefix.p, n, &made_total)) return rc;
            tr_scan = now_ms();
        }
        // list geometry across ranks (see ListGeom): count this rank's products per sub-batch, exchange, derive offsets
        const int W = std::max(1, c->P.world), R = c->P.rank;
        const int nsub = FROM_FRAG ? 1 : std::max(1, c->semi_geom.nb);
        std::vector<uint64_t> M((size_t)W * nsub + 2, 0);   // M[r][sb], then: length of the new semi amplicons, primers used
        if (FROM_FRAG || made_total == 0 || nsub == 1) M[(size_t)R * nsub] = made_total;
        else {
            // products per batch of the semi list = differences of the creation prefix at the batch boundaries
            uint64_t prev = 0;
            for (int sb = 0; sb < nsub; sb++) {
                uint64_t end = c->semi_geom.lend[sb], cp = made_total;
                if (end < n) { SCS_CUDA(c, cudaMemcpyAsync(&cp, cprefix.p + end, 8, cudaMemcpyDeviceToHost, c->st)); SCS_CUDA(c, cudaStreamSynchronize(c->st)); }
                M[(size_t)R * nsub + sb] = cp - prev; prev = cp;
            }
        }
        if (FROM_FRAG && made_total) {   // total length of this rank's new semi amplicons (sits in the per-template slots)
            SCS_CUDA(c, cudaMemsetAsync(dcount.p, 0, 8, c->st));
            sum_len_slots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>(n, created.p, slot_off.p, tdesc.p, dcount.p); SCS_LAUNCHED(c);
            SCS_CUDA(c, cudaMemcpyAsync(&M[(size_t)W * nsub], dcount.p, 8, cudaMemcpyDeviceToHost, c->st));
            SCS_CUDA(c, cudaStreamSynchronize(c->st));
        }
        M[(size_t)W * nsub + 1] = pending_count; pending_count = 0;
        if (int rc = allreduce_u64(M.data(), M.size())) return rc;
        semi_len_sum_global += M[(size_t)W * nsub];
        c->total_primers -= M[(size_t)W * nsub + 1];   // unsigned wrap as in the reference (unsigned long, Malbac.h:29)
        if (geom.nb >= 6) return c->fail(SCS_E_STATE, "amplify: too many batches");
        const int b = geom.nb++;
        uint64_t gtot = 0, l0 = 0, g0 = 0;
        geom.nsub[b] = nsub;
        for (int sb = 0; sb < nsub; sb++) {
            uint64_t subtot = 0, ahead = 0;   // ahead = products created before this rank's inside the sub-batch
            for (int r = 0; r < W; r++) { subtot += M[(size_t)r * nsub + sb]; if (FROM_FRAG ? (r < R) : (r > R)) ahead += M[(size_t)r * nsub + sb]; }
            geom.sub_l0[b][sb] = l0; geom.sub_g0[b][sb] = g0 + ahead;
            l0 += M[(size_t)R * nsub + sb]; g0 += subtot; gtot += subtot;
        }
        geom.ltot[b] = made_total; geom.gtot[b] = gtot;
        geom.gbase[b] = b ? geom.gbase[b - 1] + geom.gtot[b - 1] : 0;
        geom.lend[b] = (b ? geom.lend[b - 1] : 0) + made_total;
        uint64_t old = dst.n;
        SCS_CUDA(c, dst.desc.reserve(old + made_total + 1, old, c->st)); SCS_CUDA(c, dst.gc.reserve(old + made_total + 1, old, c->st));
        SCS_CUDA(c, dst.errref.reserve(old + made_total + 1, old, c->st));
        if (FROM_FRAG) SCS_CUDA(c, dst.primers.reserve(old + made_total + 1, old, c->st));
        if (made_total) {
            compact_products_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>(n, created.p, cprefix.p, slot_off.p, tdesc.p, tgc.p, terr.p, made_total,
                                                                                   dst.desc.p, dst.gc.p, dst.errref.p, FROM_FRAG ? dst.primers.p : nullptr, old);
            SCS_LAUNCHED(c);
            SCS_CUDA(c, cudaStreamSynchronize(c->st));
        }
        dst.n = old + made_total; dst.batch_end.push_back(dst.n);
        if (trace) fprintf(stderr, "[scs trace] amplify pass round %d from %s: %llu templates, %llu primers, %llu products | scan+alloc %.1f kernel %.1f scan %.1f geometry+append %.1f ms\n",
                           round, FROM_FRAG ? "fragments" : "semis", (unsigned long long)n, (unsigned long long)total_slots, (unsigned long long)made_total,
                           tr_alloc - tr0, tr_kernel - tr_alloc, tr_scan - tr_kernel, now_ms() - tr_scan);
        return SCS_OK;
    }
};

}  // namespace

int amplify(scs_ctx* c) {   // Malbac::amplify, Malbac.cpp:173-201
    if (!c->have_frags) return c->fail(SCS_E_STATE, "scs_amplify: call scs_create_frags first");
    if (c->P.world > 1 && c->replay.on) return c->fail(SCS_E_UNSUPPORTED, "replay runs on one rank only (the reference's logs are sequential)");
    StageTimer timer(c);
    Round R(c); R.g = c->dev_genome();
    if (int rc0 = build_gc_index(c)) return rc0;
    R.thr_ber = (uint32_t)std::min<uint64_t>(count_unit_lt(3.4e-4), 0xFFFFFFFFull);
    SCS_CUDA(c, R.dcount.reserve(1)); SCS_CUDA(c, R.ticket.reserve(1)); SCS_CUDA(c, R.flags.reserve(1));
    // createPrimers (Malbac.cpp:36-81): 4^8 primer types, -p copies each
    std::vector<long long> pc(65536, (long long)c->P.primers);
    SCS_CUDA(c, c->primer_counts.reserve(65536));
    SCS_CUDA(c, cudaMemcpyAsync(c->primer_counts.p, pc.data(), 65536 * 8, cudaMemcpyHostToDevice, c->st));
    c->total_primers = 65536ull * (uint64_t)c->P.primers;
    SCS_CUDA(c, c->err_top.reserve(1)); SCS_CUDA(c, cudaMemsetAsync(c->err_top.p, 0, 8, c->st));
    SCS_CUDA(c, c->err_pool.reserve(4096));
    c->semis.clear(); c->fulls.clear();
    c->semi_geom = ListGeom{}; c->full_geom = ListGeom{};
    const uint64_t nF = c->frags.size();
    int rc;
    if ((rc = R.set_primers(true, 0))) return rc;
    if ((rc = R.pass<true>(0, nF, c->frag_desc.p, c->frag_primers.p, nullptr, c->semis, c->semi_geom))) return rc;
    for (int i = 0; i < 5; i++) {
        if (c->total_primers == 0) break;
        if ((rc = R.set_primers(false, i + 1))) return rc;
        if ((rc = R.pass<false>(i + 1, c->semis.n, c->semis.desc.p, c->semis.primers.p, c->semis.errref.p, c->fulls, c->full_geom))) return rc;
        if (i < 4) if ((rc = R.pass<true>(i + 1, nF, c->frag_desc.p, c->frag_primers.p, nullptr, c->semis, c->semi_geom))) return rc;
    }
    SCS_CUDA(c, timer.stop(&c->stats.ms_amplify));
    c->stats.n_semis = c->semis.n; c->stats.n_fulls = c->fulls.n;
    c->stats.n_semis_global = 0; for (int b = 0; b < c->semi_geom.nb; b++) c->stats.n_semis_global += c->semi_geom.gtot[b];
    c->stats.n_fulls_global = 0; for (int b = 0; b < c->full_geom.nb; b++) c->stats.n_fulls_global += c->full_geom.gtot[b];
    c->stats.total_primers_left = c->total_primers;
    c->amplified = true; c->have_counts = false;
    return SCS_OK;
}

}  // namespace scs
