// K4 synth_reads + K5 pack_fastq: per-read synthesis (indels, k-mer-context substitutions, Phred
// qualities) and FASTQ packing, one warp per read slot (SE read / PE pair).
//
// Replaces Amplicon::yieldReads (/root/reference/lib/amplicon/Amplicon.cpp:402-565), the sequence
// rebuild Amplicon::getSequence (Amplicon.cpp:255-382), Profile::predict and its samplers
// (lib/profile/Profile.cpp:1482-1697), the sprintf/strncpy record formatting (Amplicon.cpp:459-541)
// and SeqWriter::write (lib/seqwriter/SeqWriter.cpp:41-54).
//
// Per slab of consecutive slots:  plan (lengths)  ->  exclusive scan (byte offsets)  ->  emit.
// Both kernels walk the slot's draw stream with the same cursors the reference's sequential code
// would have, but evaluate 64 positions per step (2 per lane, one Philox4x32 block per lane) and
// resolve the rare indel events with a ballot, so a read costs ~RL/32 steps instead of RL.
// All sampling is integer compares against the threshold tables built on the host. Records are
// assembled in shared memory at the destination's 16-byte phase and stored with 128-bit stores.
#include <algorithm>
#include <cstring>

#include "ctx.h"

namespace scs {

constexpr int kReadWarps = 8;         // warps per CTA
constexpr int kRLCap = 256;           // max profile read length handled by the kernels
constexpr int kSrcCap = 384;          // max read length after insertions (overflow -> error flag)
constexpr int kMaxEvents = 32;        // indel events per read kept in shared memory
constexpr int kRecCap = 16 + 40 + 2 * kSrcCap + 8;

struct ReadTables {
    const uint32_t* subs1; const uint32_t* subs2;   // [84][bins][4]  (3 thresholds + eff)
    const uint32_t* qual; const uint8_t* qualEff;    // [16][bins][94], [16][bins]
    const uint32_t* ins; const uint32_t* del; const uint32_t* isize;
    int insEff, delEff, isizeEff, minInsert, maxInsert;
    uint64_t thrIns, thrDel;                         // insertion iff x < thrIns ; else deletion iff x2 < thrDel
    int RL, paired;
};

struct WarpScratch {
    uint8_t ref[kRLCap];
    uint8_t src[kSrcCap];
    char rec[kRecCap];
    int16_t ev_pos[kMaxEvents]; int16_t ev_len[kMaxEvents]; uint32_t ev_ci[kMaxEvents];
};

// draws base+4*lane .. base+4*lane+3 of one engine, one Philox block per lane (+ a neighbour shuffle)
__device__ __forceinline__ void warp_draws4(const Stream& S, int eng, uint32_t base, int lane, uint32_t x[4]) {
    if (S.replay()) {
        const uint32_t* p = S.t[eng] + base + 4u * lane;
        x[0] = p[0]; x[1] = p[1]; x[2] = p[2]; x[3] = p[3];
        return;
    }
    uint32_t o[4], nx[4];
    const uint32_t B = (base >> 2) + lane, sh = base & 3u;
    S.block(eng, B, o);
    if (sh == 0) { x[0] = o[0]; x[1] = o[1]; x[2] = o[2]; x[3] = o[3]; return; }
#pragma unroll
    for (int q = 0; q < 4; q++) nx[q] = __shfl_down_sync(0xffffffffu, o[q], 1);
    if (lane == 31) S.block(eng, B + 1, nx);
    if (sh == 1) { x[0] = o[1]; x[1] = o[2]; x[2] = o[3]; x[3] = nx[0]; }
    else if (sh == 2) { x[0] = o[2]; x[1] = o[3]; x[2] = nx[0]; x[3] = nx[1]; }
    else { x[0] = o[3]; x[1] = nx[0]; x[2] = nx[1]; x[3] = nx[2]; }
}

// Indel pass of Profile::predict (Profile.cpp:1603-1630) over n source positions. Advances the real /
// int cursors exactly as the sequential code does; returns the output length n'. Events (position,
// +inserted / -deleted, first int draw) go to the warp scratch; *nev = 0 if there are none or if the
// "< 50 bases" guard discarded them.
__device__ __forceinline__ int indel_pass(const Stream& S, const ReadTables& T, int n, uint32_t& cr, uint32_t& ci, int lane, WarpScratch* ws,
                                          int* nev_out, int* flags) {
    int j = 0, delta = 0, nev = 0;
    while (j < n) {
        uint32_t x[4];
        warp_draws4(S, E_REAL, cr, lane, x);
        const int pA = j + 2 * lane;
        int ev = 0;
        if (pA < n) { if ((uint64_t)x[0] < T.thrIns) ev = 1; else if ((uint64_t)x[1] < T.thrDel) ev = 2; }
        if (!ev && pA + 1 < n) { if ((uint64_t)x[2] < T.thrIns) ev = 3; else if ((uint64_t)x[3] < T.thrDel) ev = 4; }
        const uint32_t any = __ballot_sync(0xffffffffu, ev != 0);
        if (!any) { int used = min(64, n - j); j += used; cr += 2u * used; continue; }
        const int l0 = __ffs(any) - 1;
        const int e = __shfl_sync(0xffffffffu, ev, l0);
        const int p = j + 2 * l0 + (e >= 3);
        const uint32_t crp = cr + 2u * (uint32_t)(p - j);
        if (e == 1 || e == 3) {   // insertion after base p: length draw, then k base draws on the int engine
            uint32_t xl = S.at(E_REAL, crp + 1);
            int k = min(count_le(T.ins, T.insEff, xl), T.insEff);
            cr = crp + 2;
            if (k > 0) {
                if (nev < kMaxEvents) { if (lane == 0) { ws->ev_pos[nev] = (int16_t)p; ws->ev_len[nev] = (int16_t)k; ws->ev_ci[nev] = ci; } }
                else if (lane == 0) atomicOr(flags, 4);
                nev++; delta += k; ci += (uint32_t)k;
            }
            j = p + 1;
        } else {                  // deletion of k bases starting at p
            uint32_t xl = S.at(E_REAL, crp + 2);
            int k = min(count_le(T.del, T.delEff, xl), T.delEff);
            k = min(n - p, k);
            cr = crp + 3;
            if (k > 0) {
                if (nev < kMaxEvents) { if (lane == 0) { ws->ev_pos[nev] = (int16_t)p; ws->ev_len[nev] = (int16_t)(-k); ws->ev_ci[nev] = 0; } }
                else if (lane == 0) atomicOr(flags, 4);
                nev++; delta -= k; j = p + k;
            } else j = p + 1;
        }
    }
    if (nev > kMaxEvents) nev = kMaxEvents;
    if (n + delta < 50) { delta = 0; nev = 0; }   // Profile.cpp:1623-1630
    *nev_out = nev;
    return n + delta;
}

// source sequence after indels (Profile.cpp:1632-1654); lane 0 walks the (rare) events
__device__ __forceinline__ const uint8_t* build_source(const Stream& S, int n, int nev, int lane, WarpScratch* ws) {
    if (nev == 0) return ws->ref;
    if (lane == 0) {
        int m = 0, e = 0;
        for (int j = 0; j < n;) {
            if (e < nev && ws->ev_pos[e] == j) {
                int len = ws->ev_len[e];
                if (len < 0) { j += -len; e++; continue; }
                if (m < kSrcCap) ws->src[m] = ws->ref[j]; m++;
                uint32_t c0 = ws->ev_ci[e];
                for (int i = 0; i < len; i++) { uint32_t b = uni_trunc(S.at(E_INT, c0 + i), 0, 3); if (m < kSrcCap) ws->src[m] = (uint8_t)b; m++; }   // A/C/G only, Profile.cpp:1560
                j++; e++;
            } else { if (m < kSrcCap) ws->src[m] = ws->ref[j]; m++; j++; }
        }
    }
    __syncwarp();
    return ws->src;
}

// substitution + quality loop (Profile.cpp:1656-1694): two output positions per lane per step
__device__ __forceinline__ void subst_quality_pass(const Stream& S, const ReadTables& T, const uint8_t* __restrict__ src, int np, int isRead1, uint32_t& cr,
                                                   int lane, char* __restrict__ oseq, char* __restrict__ oqual) {
    const uint32_t* __restrict__ subs = (!isRead1 && T.subs2) ? T.subs2 : T.subs1;
    const int bins = T.RL;
    for (int m0 = 0; m0 < np; m0 += 64) {
        uint32_t x[4];
        warp_draws4(S, E_REAL, cr + 2u * m0, lane, x);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int m = m0 + 2 * lane + h;
            if (m < np) {
                const uint32_t b0 = src[m];
                uint32_t ki;
                if (m == 0) ki = b0;
                else if (m == 1) ki = 4u + 4u * src[0] + b0;
                else ki = 20u + 16u * src[m - 2] + 4u * src[m - 1] + b0;
                const int bin = m * bins / np;
                const uint4 th = __ldg(reinterpret_cast<const uint4*>(subs) + ((size_t)ki * bins + bin));
                const uint32_t xs = x[2 * h], xq = x[2 * h + 1];
                uint32_t k = (uint32_t)(th.w > 0 && th.x <= xs) + (uint32_t)(th.w > 1 && th.y <= xs) + (uint32_t)(th.w > 2 && th.z <= xs);
                const size_t row = (size_t)(b0 * 4u + k) * bins + bin;
                const int eff = T.qualEff[row];
                const int q = count_le(T.qual + row * kQualN, eff, xq);
                oseq[m] = (char)((0x54474341u >> (8u * k)) & 0xFFu);   // "ACGT"[k]
                oqual[m] = (char)(33 + q);
            }
        }
    }
    cr += 2u * (uint32_t)np;
}

__device__ __forceinline__ int dec_digits(uint32_t v) {
    return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6 : v < 10000000u ? 7 : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}
// "@%d#%d" (+"/1" | "/2") + "\n": the %d of a negative amplicon index never occurs (index < 2^31)
__device__ __forceinline__ int header_len(uint32_t amp, uint32_t frag, int paired) { return 1 + dec_digits(amp) + 1 + dec_digits(frag) + (paired ? 2 : 0) + 1; }
__device__ __forceinline__ int write_header(char* p, uint32_t amp, uint32_t frag, int mate) {
    int n = 0; p[n++] = '@';
    int d = dec_digits(amp); for (int i = d - 1; i >= 0; i--) { p[n + i] = (char)('0' + amp % 10u); amp /= 10u; } n += d;
    p[n++] = '#';
    d = dec_digits(frag); for (int i = d - 1; i >= 0; i--) { p[n + i] = (char)('0' + frag % 10u); frag /= 10u; } n += d;
    if (mate) { p[n++] = '/'; p[n++] = (char)('0' + mate); }
    p[n++] = '\n';
    return n;
}

// copy n bytes smem -> global where (smem address & 15) == (global address & 15)
__device__ __forceinline__ void copy_out(char* __restrict__ dst, const char* __restrict__ srcp, int n, int lane) {
    const int head = min(n, (int)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
    if (lane < head) dst[lane] = srcp[lane];
    const int nvec = (n - head) >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(srcp + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (int i = lane; i < nvec; i += 32) d4[i] = s4[i];
    const int done = head + (nvec << 4);
    if (done + lane < n) dst[done + lane] = srcp[done + lane];
}

struct SlabArgs {
    uint64_t slot0, nslots;            // this launch covers local slots [slot0, slot0 + nslots)
    uint64_t slot_global0;             // global id of local slot 0 (Philox entity / replay mark index)
    uint64_t amp_global0;              // global index of local amplicon 0 (FASTQ header)
    uint64_t n_amp;
    const uint64_t* slot_base;         // [n_amp + 1] exclusive prefix of slots per amplicon
    const uint64_t* desc; const uint64_t* errref; const uint32_t* err_pool;
    const uint32_t* hdr_no;            // slow path: fragCount per slot (0 = dropped); nullptr -> slot index + 1
    const uint32_t* nfail;             // slow path: failed insert-size attempts before the slot's success
};

// Shared body of the plan and emit kernels for one slot.
template <bool EMIT>
__device__ __forceinline__ void do_slot(const Genome& g, const DrawSrc& dsrc, const ReadTables& T, const SlabArgs& A, uint64_t ls, int lane, WarpScratch* ws,
                                        uint32_t* __restrict__ plan, const uint64_t* __restrict__ off1, const uint64_t* __restrict__ off2,
                                        char* __restrict__ out1, char* __restrict__ out2, int* flags) {
    const uint64_t slot = A.slot0 + ls;
    // amplicon of this slot: last a with slot_base[a] <= slot
    uint64_t lo = 0, hi = A.n_amp;
    while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (__ldg(A.slot_base + mid) <= slot) lo = mid; else hi = mid; }
    const uint64_t a = lo;
    const Tmpl F = unpack_desc(__ldg(A.desc + a));
    const uint64_t er = __ldg(A.errref + a);
    const uint32_t nerr = (uint32_t)(er & 0xFFFF); const uint32_t* __restrict__ errs = A.err_pool + (er >> 16);
    const int RL = T.RL; const int ampLen = (int)F.len;
    uint32_t fragNo = A.hdr_no ? A.hdr_no[slot] : (uint32_t)(slot - __ldg(A.slot_base + a)) + 1u;
    if (fragNo == 0 || ampLen < RL) { if (!EMIT && lane == 0) plan[ls] = 0; return; }   // dropped slot / Amplicon.cpp:442
    Stream S; S.init(dsrc, D_READ, A.slot_global0 + slot, A.slot_global0 + slot);
    uint32_t cr = 0, ci = 0;
    int pos, isz = RL;
    if (T.paired) {
        if (A.nfail) cr = A.nfail[slot];   // failed attempts each consumed one real draw (Amplicon.cpp:483-490)
        isz = T.minInsert + min(count_le(T.isize, T.isizeEff, S.at(E_REAL, cr)), T.isizeEff);
        cr += 1;
        pos = (int)uni_trunc(S.at(E_INT, ci), 0, (uint32_t)(ampLen - isz + 1)); ci += 1;
    } else {
        pos = (int)uni_trunc(S.at(E_INT, ci), 0, (uint32_t)(ampLen - RL + 1)); ci += 1;
    }
    const uint32_t ampIdx = (uint32_t)(A.amp_global0 + a);
    uint32_t lens = 0;
    for (int mate = 1; mate <= (T.paired ? 2 : 1); mate++) {
        // source window (read 2 = reverse complement of the insert's far end, Amplicon.cpp:508-512)
        __syncwarp();
        for (int i = lane; i < RL; i += 32) {
            uint32_t fi = (mate == 1) ? (uint32_t)(pos + i) : (uint32_t)(pos + isz - 1 - i);
            uint32_t b = window_base(g, F.gstart, F.rc, fi);
            for (uint32_t e = 0; e < nerr; e++) { uint32_t v = errs[e]; if (err_pos(v) == fi) b = err_base(v); }
            ws->ref[i] = (uint8_t)(mate == 1 ? b : 3u - b);
        }
        __syncwarp();
        int nev = 0;
        const int np = indel_pass(S, T, RL, cr, ci, lane, ws, &nev, flags);
        if (np > kSrcCap) { if (lane == 0) { atomicOr(flags, 8); if (!EMIT) plan[ls] = 0; } return; }
        __syncwarp();
        if (!EMIT) {
            cr += 2u * (uint32_t)np;   // the substitution/quality pass draws twice per output base
            lens |= (uint32_t)np << (mate == 1 ? 0 : 16);
        } else {
            const uint8_t* src = build_source(S, RL, nev, lane, ws);
            const uint64_t o = (mate == 1) ? off1[ls] : off2[ls];
            char* dst = ((mate == 1) ? out1 : out2) + o;
            char* rec = ws->rec + (reinterpret_cast<uintptr_t>(dst) & 15);
            const int hl = header_len(ampIdx, fragNo, T.paired);
            if (lane == 0) {
                write_header(rec, ampIdx, fragNo, T.paired ? mate : 0);
                rec[hl + np] = '\n'; rec[hl + np + 1] = '+'; rec[hl + np + 2] = '\n'; rec[hl + 2 * np + 3] = '\n';
            }
            subst_quality_pass(S, T, src, np, mate == 1, cr, lane, rec + hl, rec + hl + np + 3);
            __syncwarp();
            copy_out(dst, rec, hl + 2 * np + 4, lane);
        }
    }
    if (!EMIT && lane == 0) plan[ls] = lens;
}

// plan: output lengths of both mates -> record sizes for the scan
__global__ void __launch_bounds__(kReadWarps * 32) plan_kernel(Genome g, DrawSrc dsrc, ReadTables T, SlabArgs A, uint32_t* __restrict__ plan,
                                                               uint32_t* __restrict__ size1, uint32_t* __restrict__ size2, int* flags,
                                                               unsigned long long* __restrict__ records) {
    __shared__ WarpScratch scratch[kReadWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t ls = (uint64_t)blockIdx.x * kReadWarps + warp;
    if (ls >= A.nslots) return;
    do_slot<false>(g, dsrc, T, A, ls, lane, &scratch[warp], plan, nullptr, nullptr, nullptr, nullptr, flags);
    __syncwarp();
    if (lane == 0) {
        const uint32_t p = plan[ls];
        uint32_t s1 = 0, s2 = 0;
        if (p) {
            const uint64_t slot = A.slot0 + ls;
            uint64_t lo = 0, hi = A.n_amp;
            while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (A.slot_base[mid] <= slot) lo = mid; else hi = mid; }
            const uint32_t fragNo = A.hdr_no ? A.hdr_no[slot] : (uint32_t)(slot - A.slot_base[lo]) + 1u;
            const int hl = header_len((uint32_t)(A.amp_global0 + lo), fragNo, T.paired);
            s1 = hl + 2 * (p & 0xFFFF) + 4;
            if (T.paired) s2 = hl + 2 * (p >> 16) + 4;
            atomicAdd(records, T.paired ? 2ull : 1ull);
        }
        size1[ls] = s1; size2[ls] = s2;
    }
}

__global__ void __launch_bounds__(kReadWarps * 32) emit_kernel(Genome g, DrawSrc dsrc, ReadTables T, SlabArgs A, const uint32_t* __restrict__ plan,
                                                               const uint64_t* __restrict__ off1, const uint64_t* __restrict__ off2,
                                                               char* __restrict__ out1, char* __restrict__ out2, int* flags) {
    __shared__ __align__(16) WarpScratch scratch[kReadWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t ls = (uint64_t)blockIdx.x * kReadWarps + warp;
    if (ls >= A.nslots) return;
    if (plan[ls] == 0) return;
    do_slot<true>(g, dsrc, T, A, ls, lane, &scratch[warp], nullptr, off1, off2, out1, out2, flags);
}

// ---- slow path (insert sizes that can exceed an amplicon, -s large): failed attempts per slot -----
__global__ void __launch_bounds__(256) fail_count_kernel(DrawSrc dsrc, ReadTables T, SlabArgs A, uint32_t* __restrict__ nfail) {
    uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= A.nslots) return;
    uint64_t lo = 0, hi = A.n_amp;
    while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (A.slot_base[mid] <= slot) lo = mid; else hi = mid; }
    const int ampLen = (int)unpack_desc(A.desc[lo]).len;
    Stream S; S.init(dsrc, D_READ, A.slot_global0 + slot, A.slot_global0 + slot);
    uint32_t f = 0;
    for (; f <= 1001u; f++) {
        int isz = T.minInsert + min(count_le(T.isize, T.isizeEff, S.at(E_REAL, f)), T.isizeEff);
        if (!(isz < T.RL || isz > ampLen)) break;
    }
    nfail[slot] = f;
}
// per amplicon: fragCount numbering and the ">1000 failures -> give up" rule (Amplicon.cpp:448-490)
__global__ void __launch_bounds__(256) fail_scan_kernel(SlabArgs A, const uint32_t* __restrict__ nfail, uint32_t* __restrict__ hdr_no) {
    uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A.n_amp) return;
    uint64_t s0 = A.slot_base[a], s1 = A.slot_base[a + 1];
    uint32_t cum = 0; bool dead = false;
    for (uint64_t s = s0; s < s1; s++) {
        if (!dead) { cum += nfail[s]; if (cum > 1000u) dead = true; }
        hdr_no[s] = dead ? 0u : (uint32_t)(s - s0) + 1u + cum;
    }
}

// ---------------------------------------------------------------------------------- profile upload
int upload_profile(scs_ctx* c) {
    if (!c->have_device) return SCS_OK;
    const HostProfile& P = c->prof; DevProfile& D = c->dprof;
    auto up32 = [&](DevBuf<uint32_t>& b, const std::vector<uint32_t>& v) -> cudaError_t {
        cudaError_t e = b.reserve(v.size() + 4); if (e != cudaSuccess) return e;
        return v.empty() ? cudaSuccess : cudaMemcpy(b.p, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
    };
    SCS_CUDA(c, up32(D.subs1, P.subsThr1));
    if (P.hasSubs2) SCS_CUDA(c, up32(D.subs2, P.subsThr2));
    SCS_CUDA(c, up32(D.qual, P.qualThr));
    SCS_CUDA(c, D.qualEff.reserve(P.qualEff.size() + 4));
    SCS_CUDA(c, cudaMemcpy(D.qualEff.p, P.qualEff.data(), P.qualEff.size(), cudaMemcpyHostToDevice));
    SCS_CUDA(c, up32(D.ins, P.insThr.thr)); SCS_CUDA(c, up32(D.del, P.delThr.thr));
    if (P.hasISize) SCS_CUDA(c, up32(D.isize, P.iSizeThr.thr));
    return SCS_OK;
}

static ReadTables make_tables(const scs_ctx* c) {
    const HostProfile& P = c->prof; const DevProfile& D = c->dprof;
    ReadTables T;
    T.subs1 = D.subs1.p; T.subs2 = P.hasSubs2 ? D.subs2.p : nullptr; T.qual = D.qual.p; T.qualEff = D.qualEff.p;
    T.ins = D.ins.p; T.del = D.del.p; T.isize = D.isize.p;
    T.insEff = P.insThr.eff; T.delEff = P.delThr.eff; T.isizeEff = P.hasISize ? P.iSizeThr.eff : 0;
    T.minInsert = P.minInsert; T.maxInsert = P.maxInsert;
    T.thrIns = P.thrInsertAll ? (1ull << 32) : P.thrInsert; T.thrDel = P.thrDeleteAll ? (1ull << 32) : P.thrDelete;
    T.RL = P.readLength; T.paired = c->P.paired;
    return T;
}

// ---------------------------------------------------------------------------------- slab pipeline
int yield_reads(scs_ctx* c, scs_sink_fn sink, void* user) {
    if (!c->have_profile) return c->fail(SCS_E_STATE, "scs_yield_reads: no profile loaded");
    if (!c->have_counts) { if (int rc = set_read_counts(c)) return rc; }
    const HostProfile& P = c->prof;
    if (P.readLength > kRLCap) return c->fail(SCS_E_UNSUPPORTED, "read length above 256 is not supported by the kernels");
    if (c->P.paired && !P.hasISize) return c->fail(SCS_E_ARG, "Error: unrecognized parameter name \"insertSize\"");   // Profile.cpp:1484 -> Config.cpp:71-78
    c->stats.records = 0; c->stats.fastq_bytes[0] = c->stats.fastq_bytes[1] = 0;
    c->stats.ms_reads = c->stats.ms_reads_kernels = c->stats.ms_emit_kernel = 0; c->stats.emit_launches = 0; c->stats.genome_window_bytes = 0;
    const uint64_t nslots = c->n_slots;
    if (nslots == 0) return SCS_OK;
    const int nfiles = c->P.paired ? 2 : 1;
    const uint64_t slab = c->P.slab_bytes ? c->P.slab_bytes : (128ull << 20);
    const uint64_t worst = 40 + 2ull * (P.readLength + 128) + 8;   // bytes per record bound used for batching (checked below)
    uint64_t batch = std::max<uint64_t>(1024, slab / worst);
    // device + pinned slabs, double buffered
    if (c->slab_cap < slab) {
        for (int b = 0; b < 2; b++) for (int f = 0; f < 2; f++) {
            c->slab_dev[b][f].release();
            if (c->slab_host[b][f]) { cudaFreeHost(c->slab_host[b][f]); c->slab_host[b][f] = nullptr; }
        }
        for (int b = 0; b < 2; b++) for (int f = 0; f < nfiles; f++) {
            SCS_CUDA(c, c->slab_dev[b][f].reserve(slab + 64));
            SCS_CUDA(c, cudaMallocHost((void**)&c->slab_host[b][f], slab + 64));
        }
        c->slab_cap = slab;
    } else for (int b = 0; b < 2; b++) for (int f = 0; f < nfiles; f++) if (!c->slab_host[b][f]) {
        SCS_CUDA(c, c->slab_dev[b][f].reserve(slab + 64));
        SCS_CUDA(c, cudaMallocHost((void**)&c->slab_host[b][f], slab + 64));
    }
    ReadTables T = make_tables(c);
    Genome g; g.words = c->genome_words.p; g.n_bases = c->genome_bases;
    DrawSrc dsrc = draw_src(c, D_READ);
    SlabArgs A; A.slot_global0 = 0; A.amp_global0 = 0; A.n_amp = c->fulls.n; A.slot_base = c->slot_base.p;
    A.desc = c->fulls.desc.p; A.errref = c->fulls.errref.p; A.err_pool = c->err_pool.p; A.hdr_no = nullptr; A.nfail = nullptr;
    DevBuf<int> flags; SCS_CUDA(c, flags.reserve(1)); SCS_CUDA(c, cudaMemsetAsync(flags.p, 0, 4, c->st));
    DevBuf<unsigned long long> drec; SCS_CUDA(c, drec.reserve(1)); SCS_CUDA(c, cudaMemsetAsync(drec.p, 0, 8, c->st));
    cudaEvent_t e0, e1, ek0, ek1, ee0, ee1, ecopy[2], ekern[2];
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&ek0); cudaEventCreate(&ek1); cudaEventCreate(&ee0); cudaEventCreate(&ee1);
    for (int b = 0; b < 2; b++) { cudaEventCreateWithFlags(&ecopy[b], cudaEventDisableTiming); cudaEventCreateWithFlags(&ekern[b], cudaEventDisableTiming); }
    cudaEventRecord(e0, c->st);
    // slow path: insert sizes that can fail (isize > amplicon length; amplicons are 1000..2000 long)
    DevBuf<uint32_t> nfail, hdrno;
    if (c->P.paired && P.maxInsert > 1000) {
        SCS_CUDA(c, nfail.reserve(nslots + 1)); SCS_CUDA(c, hdrno.reserve(nslots + 1));
        SlabArgs F = A; F.slot0 = 0; F.nslots = nslots;
        fail_count_kernel<<<(unsigned)((nslots + 255) / 256), 256, 0, c->st>>>(dsrc, T, F, nfail.p); SCS_LAUNCHED(c);
        fail_scan_kernel<<<(unsigned)((A.n_amp + 255) / 256), 256, 0, c->st>>>(F, nfail.p, hdrno.p); SCS_LAUNCHED(c);
        A.hdr_no = hdrno.p; A.nfail = nfail.p;
    }
    DevBuf<uint32_t> plan, size1, size2; DevBuf<uint64_t> off1, off2;
    SCS_CUDA(c, plan.reserve(batch + 1)); SCS_CUDA(c, size1.reserve(batch + 1)); SCS_CUDA(c, size2.reserve(batch + 1));
    SCS_CUDA(c, off1.reserve(batch + 1)); SCS_CUDA(c, off2.reserve(batch + 1));
    struct Pending { bool live = false; uint64_t bytes[2] = {0, 0}; } pend[2];
    auto drain = [&](int b) -> int {
        if (!pend[b].live) return SCS_OK;
        cudaError_t e = cudaEventSynchronize(ecopy[b]);
        if (e != cudaSuccess) return c->fail(SCS_E_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
        pend[b].live = false;
        if (sink) for (int f = 0; f < nfiles; f++) if (pend[b].bytes[f]) if (sink(user, f, c->slab_host[b][f], pend[b].bytes[f])) return c->fail(SCS_E_IO, "FASTQ sink failed");
        return SCS_OK;
    };
    int bi = 0; double msk = 0, mse = 0;
    for (uint64_t s0 = 0; s0 < nslots; s0 += batch, bi ^= 1) {
        const uint64_t m = std::min(batch, nslots - s0);
        if (int rc = drain(bi)) return rc;   // buffer bi is free again once its previous copy has been consumed
        A.slot0 = s0; A.nslots = m;
        const unsigned nb = (unsigned)((m + kReadWarps - 1) / kReadWarps);
        cudaEventRecord(ek0, c->st);
        plan_kernel<<<nb, kReadWarps * 32, 0, c->st>>>(g, dsrc, T, A, plan.p, size1.p, size2.p, flags.p, drec.p); SCS_LAUNCHED(c);
        uint64_t tot[2] = {0, 0};
        if (int rc = exclusive_scan_u32(c, size1.p, off1.p, m, &tot[0])) return rc;
        if (nfiles == 2) if (int rc = exclusive_scan_u32(c, size2.p, off2.p, m, &tot[1])) return rc;
        if (tot[0] > slab || tot[1] > slab) return c->fail(SCS_E_NOMEM, "FASTQ slab too small for one batch (raise slab_bytes)");
        cudaEventRecord(ee0, c->st);
        emit_kernel<<<nb, kReadWarps * 32, 0, c->st>>>(g, dsrc, T, A, plan.p, off1.p, off2.p, c->slab_dev[bi][0].p, nfiles == 2 ? c->slab_dev[bi][1].p : nullptr, flags.p);
        SCS_LAUNCHED(c); c->stats.emit_launches++;
        cudaEventRecord(ee1, c->st); cudaEventRecord(ek1, c->st);
        cudaEventRecord(ekern[bi], c->st);
        SCS_CUDA(c, cudaStreamWaitEvent(c->st_copy, ekern[bi], 0));
        for (int f = 0; f < nfiles; f++) if (tot[f]) SCS_CUDA(c, cudaMemcpyAsync(c->slab_host[bi][f], c->slab_dev[bi][f].p, tot[f], cudaMemcpyDeviceToHost, c->st_copy));
        cudaEventRecord(ecopy[bi], c->st_copy);
        pend[bi].live = true; pend[bi].bytes[0] = tot[0]; pend[bi].bytes[1] = tot[1];
        c->stats.fastq_bytes[0] += tot[0]; c->stats.fastq_bytes[1] += tot[1];
        SCS_CUDA(c, cudaEventSynchronize(ek1));
        float a = 0, b2 = 0; cudaEventElapsedTime(&a, ek0, ek1); cudaEventElapsedTime(&b2, ee0, ee1); msk += a; mse += b2;
        // while this slab's copy runs, hand the previous slab to the sink
        if (int rc = drain(bi ^ 1)) return rc;
    }
    if (int rc = drain(0)) return rc;
    if (int rc = drain(1)) return rc;
    SCS_CUDA(c, cudaStreamWaitEvent(c->st, ecopy[0], 0)); SCS_CUDA(c, cudaStreamWaitEvent(c->st, ecopy[1], 0));
    cudaEventRecord(e1, c->st); SCS_CUDA(c, cudaStreamSynchronize(c->st));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    c->stats.ms_reads = ms; c->stats.ms_reads_kernels = msk; c->stats.ms_emit_kernel = mse;
    int hflags = 0; SCS_CUDA(c, cudaMemcpy(&hflags, flags.p, 4, cudaMemcpyDeviceToHost));
    unsigned long long hrec = 0; SCS_CUDA(c, cudaMemcpy(&hrec, drec.p, 8, cudaMemcpyDeviceToHost)); c->stats.records = hrec;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(ek0); cudaEventDestroy(ek1); cudaEventDestroy(ee0); cudaEventDestroy(ee1);
    for (int b = 0; b < 2; b++) { cudaEventDestroy(ecopy[b]); cudaEventDestroy(ekern[b]); }
    if (hflags & 4) return c->fail(SCS_E_UNSUPPORTED, "more than 32 indel events in one read");
    if (hflags & 8) return c->fail(SCS_E_UNSUPPORTED, "read grew beyond 384 bases through insertions");
    return SCS_OK;
}

// ---------------------------------------------------------------------------------- test hooks
__global__ void __launch_bounds__(kReadWarps * 32) test_predict_kernel(ReadTables T, const char* __restrict__ srcAscii, int n_reads, int isRead1,
                                                                       const uint32_t* real, uint64_t stride_real, const uint32_t* ints, uint64_t stride_int,
                                                                       char* out_seq, char* out_qual, int out_stride, int* out_len, int* flags) {
    __shared__ __align__(16) WarpScratch scratch[kReadWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * kReadWarps + warp;
    if (r >= n_reads) return;
    WarpScratch* ws = &scratch[warp];
    Stream S; S.k0 = S.k1 = S.e0 = S.e1 = S.dom2 = 0; S.t[0] = real + (uint64_t)r * stride_real; S.t[1] = ints + (uint64_t)r * stride_int;
    const int RL = T.RL;
    for (int i = lane; i < RL; i += 32) {
        char ch = srcAscii[(size_t)r * RL + i];
        ws->ref[i] = (uint8_t)(ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3);
    }
    __syncwarp();
    uint32_t cr = 0, ci = 0; int nev = 0;
    const int np = indel_pass(S, T, RL, cr, ci, lane, ws, &nev, flags + r);
    __syncwarp();
    if (flags[r] != 0) { if (lane == 0) out_len[r] = -2; return; }   // more than kMaxEvents indel events
    if (np > kSrcCap || np > out_stride) { if (lane == 0) out_len[r] = -1; return; }
    __syncwarp();
    const uint8_t* src = build_source(S, RL, nev, lane, ws);
    subst_quality_pass(S, T, src, np, isRead1, cr, lane, out_seq + (size_t)r * out_stride, out_qual + (size_t)r * out_stride);
    if (lane == 0) out_len[r] = np;
}

int test_predict(scs_ctx* c, const char* src, int n_reads, int is_read1, const uint32_t* real, uint64_t stride_real, const uint32_t* ints,
                 uint64_t stride_int, char* out_seq, char* out_qual, int out_stride, int32_t* out_len) {
    if (!c->have_device) return c->fail(SCS_E_CUDA, "no CUDA device");
    if (!c->have_profile) return c->fail(SCS_E_STATE, "scs_test_predict: no profile loaded");
    const int RL = c->prof.readLength;
    if (RL > kRLCap) return c->fail(SCS_E_UNSUPPORTED, "read length above 256");
    DevBuf<char> dsrc, dseq, dqual; DevBuf<uint32_t> dreal, dint; DevBuf<int> dlen, flags;
    SCS_CUDA(c, dsrc.reserve((size_t)n_reads * RL + 16)); SCS_CUDA(c, dseq.reserve((size_t)n_reads * out_stride + 16)); SCS_CUDA(c, dqual.reserve((size_t)n_reads * out_stride + 16));
    SCS_CUDA(c, dreal.reserve((size_t)n_reads * stride_real + 4096)); SCS_CUDA(c, dint.reserve((size_t)n_reads * stride_int + 4096));
    SCS_CUDA(c, dlen.reserve(n_reads + 1)); SCS_CUDA(c, flags.reserve(n_reads + 1));
    SCS_CUDA(c, cudaMemset(dreal.p, 0, dreal.cap * 4)); SCS_CUDA(c, cudaMemset(dint.p, 0, dint.cap * 4)); SCS_CUDA(c, cudaMemset(flags.p, 0, flags.cap * 4));
    SCS_CUDA(c, cudaMemset(dseq.p, 0, dseq.cap)); SCS_CUDA(c, cudaMemset(dqual.p, 0, dqual.cap));
    SCS_CUDA(c, cudaMemcpy(dsrc.p, src, (size_t)n_reads * RL, cudaMemcpyHostToDevice));
    SCS_CUDA(c, cudaMemcpy(dreal.p, real, (size_t)n_reads * stride_real * 4, cudaMemcpyHostToDevice));
    SCS_CUDA(c, cudaMemcpy(dint.p, ints, (size_t)n_reads * stride_int * 4, cudaMemcpyHostToDevice));
    ReadTables T = make_tables(c);
    test_predict_kernel<<<(n_reads + kReadWarps - 1) / kReadWarps, kReadWarps * 32, 0, c->st>>>(T, dsrc.p, n_reads, is_read1, dreal.p, stride_real, dint.p, stride_int,
                                                                                              dseq.p, dqual.p, out_stride, dlen.p, flags.p);
    SCS_LAUNCHED(c);
    SCS_CUDA(c, cudaStreamSynchronize(c->st));
    SCS_CUDA(c, cudaMemcpy(out_seq, dseq.p, (size_t)n_reads * out_stride, cudaMemcpyDeviceToHost));
    SCS_CUDA(c, cudaMemcpy(out_qual, dqual.p, (size_t)n_reads * out_stride, cudaMemcpyDeviceToHost));
    SCS_CUDA(c, cudaMemcpy(out_len, dlen.p, (size_t)n_reads * 4, cudaMemcpyDeviceToHost));
    return SCS_OK;
}

// full amplicon sequences as text (parity check of descriptors + error overlays)
__global__ void full_seq_kernel(Genome g, const uint64_t* __restrict__ desc, const uint64_t* __restrict__ errref, const uint32_t* __restrict__ err_pool,
                                const uint64_t* __restrict__ offs, uint64_t n, char* __restrict__ out) {
    uint64_t a = blockIdx.x;
    if (a >= n) return;
    Tmpl F = unpack_desc(desc[a]); uint64_t er = errref[a]; uint32_t nerr = (uint32_t)(er & 0xFFFF); const uint32_t* errs = err_pool + (er >> 16);
    char* o = out + offs[a];
    for (uint32_t i = threadIdx.x; i < F.len; i += blockDim.x) {
        uint32_t b = window_base(g, F.gstart, F.rc, i);
        for (uint32_t e = 0; e < nerr; e++) if (err_pos(errs[e]) == i) b = err_base(errs[e]);
        o[i] = "ACGT"[b];
    }
    if (threadIdx.x == 0) o[F.len] = '\n';
}

int dump_full_seqs(scs_ctx* c, char* buf, uint64_t cap, int64_t* written) {
    const uint64_t n = c->fulls.n;
    std::vector<uint64_t> d(n);
    if (n) SCS_CUDA(c, cudaMemcpy(d.data(), c->fulls.desc.p, n * 8, cudaMemcpyDeviceToHost));
    std::vector<uint64_t> offs(n + 1, 0);
    for (uint64_t i = 0; i < n; i++) offs[i + 1] = offs[i] + unpack_desc(d[i]).len + 1;
    *written = (int64_t)offs[n];
    if (!buf) return SCS_OK;
    if (cap < offs[n]) return c->fail(SCS_E_ARG, "scs_dump: buffer too small");
    if (n == 0) return SCS_OK;
    DevBuf<uint64_t> doffs; DevBuf<char> dout;
    SCS_CUDA(c, doffs.reserve(n + 1)); SCS_CUDA(c, dout.reserve(offs[n] + 16));
    SCS_CUDA(c, cudaMemcpy(doffs.p, offs.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
    Genome g; g.words = c->genome_words.p; g.n_bases = c->genome_bases;
    full_seq_kernel<<<(unsigned)n, 128, 0, c->st>>>(g, c->fulls.desc.p, c->fulls.errref.p, c->err_pool.p, doffs.p, n, dout.p); SCS_LAUNCHED(c);
    SCS_CUDA(c, cudaStreamSynchronize(c->st));
    SCS_CUDA(c, cudaMemcpy(buf, dout.p, offs[n], cudaMemcpyDeviceToHost));
    return SCS_OK;
}

}  // namespace scs
