// K3 amplicon_weights + alloc_reads: GC-bias-weighted allocation of reads to full amplicons.
//
// Replaces Amplicon::getWeightedLength (/root/reference/lib/amplicon/Amplicon.cpp:396-400),
// Profile::getGCFactor (lib/profile/Profile.cpp:1503-1513), Malbac::setReadCounts
// (lib/malbac/Malbac.cpp:370-408) and randIndx_hp / batchSampling (lib/mydefine/MyDefine.cpp:191-272).
//
// The reference's multinomial for the remainder works on chunks of 1000 consecutive amplicons with
// a serial FP64 running sum inside each chunk; that structure is kept (one thread builds a chunk's
// CDF in the reference's summation order, then all of the chunk's samples are drawn in parallel by
// binary search on it), so the decisions are the reference's for the same draws.
#include <algorithm>
#include <cmath>

#include "ctx.h"

namespace scs {

constexpr uint32_t kChunk = 1000;   // min(1000, ac/threads) with threads = 1, MyDefine.cpp:206
static const double kEpsH = 2.2204e-16;

__device__ __forceinline__ double draw_r(uint32_t x) {   // randomDouble(ZERO_FINAL, 1)
    return __dadd_rn(2.2204e-16, __dmul_rn(__dadd_rn(1.0, -2.2204e-16), (double)x / 4294967296.0));
}

__global__ void __launch_bounds__(256) full_gidx_kernel(ListGeom G, uint64_t n, uint64_t* __restrict__ gidx) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gidx[i] = global_index(G, i);
}
__global__ void __launch_bounds__(256) scatter_f64_kernel(const double* __restrict__ src, const uint64_t* __restrict__ gidx, uint64_t n, double* __restrict__ dst) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[gidx[i]] = src[i];
}
__global__ void __launch_bounds__(256) gather_counts_kernel(const double* __restrict__ wg, const uint32_t* __restrict__ cg, const uint64_t* __restrict__ sbase_g,
                                                            const uint64_t* __restrict__ gidx, uint64_t n, int paired, double* __restrict__ w,
                                                            uint32_t* __restrict__ counts, uint64_t* __restrict__ slot_gbase, uint32_t* __restrict__ slots) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t g = gidx[i];
    const uint32_t v = cg[g];
    w[i] = wg[g]; counts[i] = v; slot_gbase[i] = sbase_g[g]; slots[i] = paired ? (v >> 1) : v;
}

// balance = 1: one rank's block of the gathered (desc, errref, global index) triples scattered into the cell-wide table, rebased
// to that rank's place in the cell-wide genome / error pool
__global__ void __launch_bounds__(256) scatter_amplicons_kernel(const uint64_t* __restrict__ triples, uint64_t stride, uint64_t n, uint64_t base_bases,
                                                                uint64_t err_base, uint64_t* __restrict__ gdesc, uint64_t* __restrict__ gerrref) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t g = triples[2 * stride + i];
    Tmpl t = unpack_desc(triples[i]);
    gdesc[g] = pack_desc(t.gstart + base_bases, t.rc, t.len);
    const uint64_t er = triples[stride + i];
    gerrref[g] = (er & 0xFFFFull) ? ((((er >> 16) + err_base) << 16) | (er & 0xFFFFull)) : 0ull;
}

__global__ void __launch_bounds__(256) weights_kernel(DrawSrc src, const double* __restrict__ gcf_tape, const uint64_t* __restrict__ gidx, uint64_t n,
                                                      const uint64_t* __restrict__ desc, const uint32_t* __restrict__ gc,
                                                      const double* __restrict__ gcMeans, double gcStd, double* __restrict__ w) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t len = unpack_desc(desc[i]).len;
    uint32_t pct = 100u * gc[i] / len;   // Amplicon.cpp:398
    double f = 0.0;
    if (pct <= 100u) {
        if (gcf_tape) f = gcf_tape[gidx[i]];
        else {
            // Marsaglia polar normal on Philox draws, redrawn until >= 0 (Profile.cpp:1508-1511)
            Stream s; s.init(src, D_GCF, gidx[i], 0);
            double mean = gcMeans[pct]; uint32_t k = 0;
            for (;;) {
                double u1 = __dadd_rn((double)s.at(E_REAL, k), 0.5) / 4294967296.0;
                double u2 = __dadd_rn((double)s.at(E_REAL, k + 1), 0.5) / 4294967296.0;
                k += 2;
                double v1 = __dadd_rn(__dmul_rn(2.0, u1), -1.0), v2 = __dadd_rn(__dmul_rn(2.0, u2), -1.0);
                double q = __dadd_rn(__dmul_rn(v1, v1), __dmul_rn(v2, v2));
                if (q >= 1.0 || q == 0.0) continue;
                double fac = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, det_log(q)), q));
                double v = __dadd_rn(mean, __dmul_rn(gcStd, __dmul_rn(v1, fac)));
                if (v >= 0) { f = v; break; }
            }
        }
    }
    w[i] = __ddiv_rn(__dmul_rn(f, (double)len), 1000000.0);   // fragSize^2, Config.cpp:41
}

// serial sum of each chunk of kChunk weights (reference summation order inside a chunk)
__global__ void __launch_bounds__(128) chunk_sum_kernel(const double* __restrict__ w, uint64_t n, double* __restrict__ sums) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = c * kChunk;
    if (s >= n) return;
    uint64_t e = min(n, s + kChunk);
    double t = 0.0;
    for (uint64_t i = s; i < e; i++) t = __dadd_rn(t, w[i]);
    sums[c] = t;
}

__global__ void __launch_bounds__(256) normalize_floor_kernel(double* __restrict__ w, uint64_t n, double denom, double reads, uint32_t* __restrict__ counts,
                                                              unsigned long long* __restrict__ total) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long rc = 0;
    if (i < n) {
        double v = __ddiv_rn(w[i], denom);
        w[i] = v;
        rc = (unsigned int)__dmul_rn(v, reads);   // Malbac.cpp:390
        counts[i] = (uint32_t)rc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rc += __shfl_down_sync(0xffffffffu, rc, o);
    if ((threadIdx.x & 31) == 0 && rc) atomicAdd(total, rc);
}

// per chunk: cdf[k] = cdf[k-1] + w[k]/total (serial, MyDefine.cpp:224-226)
__global__ void __launch_bounds__(128) chunk_cdf_kernel(const double* __restrict__ w, uint64_t n, const double* __restrict__ totals, double* __restrict__ cdf) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = c * kChunk;
    if (s >= n) return;
    uint64_t e = min(n, s + kChunk);
    double tot = totals[c], prev = 0.0;
    for (uint64_t i = s; i < e; i++) { prev = __dadd_rn(prev, __ddiv_rn(w[i], tot)); cdf[i] = prev; }
}

// one thread per sample: chunk by binary search on the sample prefix, index by binary search on the chunk CDF
__global__ void __launch_bounds__(256) chunk_sample_kernel(DrawSrc src, uint64_t chunk_global0, uint64_t n_amp, uint64_t n_chunks,
                                                           const uint64_t* __restrict__ sample_prefix, uint64_t n_samples,
                                                           const double* __restrict__ cdf, uint32_t* __restrict__ counts) {
    uint64_t sidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= n_samples) return;
    uint64_t lo = 0, hi = n_chunks;   // last chunk with prefix <= sidx
    while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (sample_prefix[mid] <= sidx) lo = mid; else hi = mid; }
    uint64_t c = lo, i = sidx - sample_prefix[c];
    Stream s; s.init(src, D_MULTC, chunk_global0 + c, chunk_global0 + c);
    double r = draw_r(s.at(E_REAL, (uint32_t)i));
    uint64_t a0 = c * kChunk; uint32_t m = (uint32_t)min((uint64_t)kChunk, n_amp - a0);
    const double* row = cdf + a0;
    uint32_t l = 0, h = m;   // first k with r <= row[k], else m-1 (randIndx, MyDefine.cpp:274-282)
    while (l < h) { uint32_t mid = (l + h) >> 1; if (r <= row[mid]) h = mid; else l = mid + 1; }
    if (l >= m) l = m - 1;
    atomicAdd(&counts[a0 + l], 1u);
}

__global__ void __launch_bounds__(256) odd_flags_kernel(const uint32_t* __restrict__ counts, uint64_t n, uint32_t* __restrict__ flags) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = counts[i] & 1u;
}
// PE: odd counts alternately +1 / -1 in list order (Malbac.cpp:398-407); then slots = pairs (PE) or reads (SE)
__global__ void __launch_bounds__(256) parity_slots_kernel(uint32_t* __restrict__ counts, uint64_t n, const uint64_t* __restrict__ odd_prefix, uint64_t odd_before,
                                                           int paired, uint32_t* __restrict__ slots) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t v = counts[i];
    if (paired) {
        if (v & 1u) { v = (((odd_before + odd_prefix[i]) & 1ull) == 0) ? v + 1 : v - 1; counts[i] = v; }
        slots[i] = v >> 1;
    } else slots[i] = v;
}

int set_read_counts(scs_ctx* c) {
    if (!c->amplified) return c->fail(SCS_E_STATE, "scs_set_read_counts: call scs_amplify first");
    if (!c->have_profile) return c->fail(SCS_E_STATE, "scs_set_read_counts: no profile loaded");
    StageTimer timer(c);
    const uint64_t n = c->fulls.n;   // local
    const ListGeom G = c->full_geom;
    uint64_t N = 0; for (int b = 0; b < G.nb; b++) N += G.gtot[b];   // all ranks
    const bool multi = c->P.world > 1;
    // Malbac::yieldReads: reads = refLen * coverage / readLength (Malbac.cpp:420), individual reads, whole cell
    c->reads_requested = (uint64_t)((double)c->ref_len_half * c->P.coverage / (double)c->prof.readLength);
    c->stats.reads_requested = c->reads_requested;
    SCS_CUDA(c, c->weights.reserve(n + 1)); SCS_CUDA(c, c->counts.reserve(n + 1)); SCS_CUDA(c, c->slot_base.reserve(n + 2));
    SCS_CUDA(c, c->full_gidx.reserve(n + 1)); SCS_CUDA(c, c->slot_gbase.reserve(n + 1));
    c->n_slots = 0;
    if (N == 0) { c->have_counts = true; return SCS_OK; }   // the reference crashes on an empty amplicon list; we emit nothing
    AllocScratch& LS = c->lscratch;
    DevBuf<double>& gcm = LS.gcm; SCS_CUDA(c, gcm.reserve(101));
    SCS_CUDA(c, cudaMemcpyAsync(gcm.p, c->prof.gcMeans, 101 * 8, cudaMemcpyHostToDevice, c->st));
    const unsigned nbl = (unsigned)((n + 255) / 256), nb = (unsigned)((N + 255) / 256);
    if (n) {
        full_gidx_kernel<<<nbl, 256, 0, c->st>>>(G, n, c->full_gidx.p); SCS_LAUNCHED(c);
        weights_kernel<<<nbl, 256, 0, c->st>>>(draw_src(c, D_GCF), c->replay.on ? c->replay.gcf.p : nullptr, c->full_gidx.p, n, c->fulls.desc.p, c->fulls.gc.p,
                                               gcm.p, c->prof.gcStd, c->weights.p); SCS_LAUNCHED(c);
    }
    // The allocation itself is defined on the whole cell's list. With several ranks every rank gets the full weight vector
    // (one all-reduce of a scattered copy) and runs the identical allocation, so the result does not depend on the rank count.
    DevBuf<double>& wg_buf = LS.wg_buf; DevBuf<uint32_t>& cg_buf = LS.cg_buf;
    double* wg = c->weights.p; uint32_t* cg = c->counts.p;
    if (multi) {
        SCS_CUDA(c, wg_buf.reserve(N + 1)); SCS_CUDA(c, cg_buf.reserve(N + 1));
        wg = wg_buf.p; cg = cg_buf.p;
        SCS_CUDA(c, cudaMemsetAsync(wg, 0, N * 8, c->st));
        if (n) { scatter_f64_kernel<<<nbl, 256, 0, c->st>>>(c->weights.p, c->full_gidx.p, n, wg); SCS_LAUNCHED(c); }
        if (int rc = allreduce_dev_f64(c, wg, N)) return rc;   // NCCL over NVLink (or the caller's hook)
    }
    const uint64_t nch = (N + kChunk - 1) / kChunk;
    DevBuf<double>& dsums = LS.dsums; SCS_CUDA(c, dsums.reserve(nch + 1));
    std::vector<double> hs(nch);
    chunk_sum_kernel<<<(unsigned)((nch + 127) / 128), 128, 0, c->st>>>(wg, N, dsums.p); SCS_LAUNCHED(c);
    SCS_CUDA(c, cudaMemcpyAsync(hs.data(), dsums.p, nch * 8, cudaMemcpyDeviceToHost, c->st));
    SCS_CUDA(c, cudaStreamSynchronize(c->st));
    double S = 0; for (uint64_t k = 0; k < nch; k++) S += hs[k];
    DevBuf<unsigned long long>& dtot = LS.dtot; SCS_CUDA(c, dtot.reserve(1)); SCS_CUDA(c, cudaMemsetAsync(dtot.p, 0, 8, c->st));
    normalize_floor_kernel<<<nb, 256, 0, c->st>>>(wg, N, kEpsH + S, (double)(long)c->reads_requested, cg, dtot.p); SCS_LAUNCHED(c);
    chunk_sum_kernel<<<(unsigned)((nch + 127) / 128), 128, 0, c->st>>>(wg, N, dsums.p); SCS_LAUNCHED(c);
    unsigned long long floorSum = 0;
    SCS_CUDA(c, cudaMemcpyAsync(&floorSum, dtot.p, 8, cudaMemcpyDeviceToHost, c->st));
    SCS_CUDA(c, cudaMemcpyAsync(hs.data(), dsums.p, nch * 8, cudaMemcpyDeviceToHost, c->st));
    SCS_CUDA(c, cudaStreamSynchronize(c->st));
    // randIndx_hp (MyDefine.cpp:203-247): whole samples per chunk, then the leftover one by one over the chunk CDF
    unsigned long rem = (unsigned long)((long)c->reads_requested - (long)floorSum);
    std::vector<uint64_t> ns(nch + 1, 0); unsigned long cnt = 0;
    for (uint64_t k = 0; k < nch; k++) { ns[k] = (unsigned int)(hs[k] * (double)rem); cnt += ns[k]; }
    rem -= cnt;
    if (rem > 0) {
        std::vector<double> probs(nch);
        probs[0] = hs[0]; for (uint64_t k = 1; k < nch; k++) probs[k] = probs[k - 1] + hs[k];
        uint64_t i = 0;
        while (rem-- > 0) {
            uint32_t x = host_draw(c, D_MULTM, E_REAL, 0, 0, i++);
            double r = kEpsH + (1.0 - kEpsH) * ((double)x / 4294967296.0);
            uint64_t k = std::lower_bound(probs.begin(), probs.end(), r) - probs.begin();   // first k with r <= probs[k]
            if (k >= nch) k = nch - 1;
            ns[k] += 1;
        }
    }
    uint64_t nsamp = 0;
    std::vector<uint64_t> pref(nch + 1);
    for (uint64_t k = 0; k < nch; k++) { pref[k] = nsamp; nsamp += ns[k]; }
    pref[nch] = nsamp;
    if (nsamp) {
        DevBuf<double>& cdf = LS.cdf; SCS_CUDA(c, cdf.reserve(N + 1));
        DevBuf<uint64_t>& dpref = LS.dpref; SCS_CUDA(c, dpref.reserve(nch + 1));
        SCS_CUDA(c, cudaMemcpyAsync(dpref.p, pref.data(), (nch + 1) * 8, cudaMemcpyHostToDevice, c->st));
        chunk_cdf_kernel<<<(unsigned)((nch + 127) / 128), 128, 0, c->st>>>(wg, N, dsums.p, cdf.p); SCS_LAUNCHED(c);
        chunk_sample_kernel<<<(unsigned)((nsamp + 255) / 256), 256, 0, c->st>>>(draw_src(c, D_MULTC), 0, N, nch, dpref.p, nsamp, cdf.p, cg); SCS_LAUNCHED(c);
        SCS_CUDA(c, cudaStreamSynchronize(c->st));   // pref (host) is read by the copy above
    }
    DevBuf<uint32_t>& tmp = LS.tmp; SCS_CUDA(c, tmp.reserve(N + 1));
    DevBuf<uint64_t>&oddp = LS.oddp, &sbase_g = LS.sbase_g;
    if (c->P.paired) {
        SCS_CUDA(c, oddp.reserve(N + 1));
        odd_flags_kernel<<<nb, 256, 0, c->st>>>(cg, N, tmp.p); SCS_LAUNCHED(c);
        if (int rc = exclusive_scan_u32(c, tmp.p, oddp.p, N, nullptr)) return rc;
    }
    parity_slots_kernel<<<nb, 256, 0, c->st>>>(cg, N, oddp.p, 0, c->P.paired, tmp.p); SCS_LAUNCHED(c);
    if (!multi) {
        if (int rc = exclusive_scan_u32(c, tmp.p, c->slot_base.p, N, &c->n_slots)) return rc;
        SCS_CUDA(c, cudaMemcpyAsync(c->slot_gbase.p, c->slot_base.p, N * 8, cudaMemcpyDeviceToDevice, c->st));
    } else {
        SCS_CUDA(c, sbase_g.reserve(N + 1));
        if (int rc = exclusive_scan_u32(c, tmp.p, sbase_g.p, N, nullptr)) return rc;
        DevBuf<uint32_t>& lslots = LS.lslots; SCS_CUDA(c, lslots.reserve(n + 1));
        if (n) {
            gather_counts_kernel<<<nbl, 256, 0, c->st>>>(wg, cg, sbase_g.p, c->full_gidx.p, n, c->P.paired, c->weights.p, c->counts.p, c->slot_gbase.p, lslots.p);
            SCS_LAUNCHED(c);
            if (int rc = exclusive_scan_u32(c, lslots.p, c->slot_base.p, n, &c->n_slots)) return rc;
        }
    }
    SCS_CUDA(c, cudaMemcpyAsync(c->slot_base.p + n, &c->n_slots, 8, cudaMemcpyHostToDevice, c->st));
    c->global_view = false;
    if (multi && c->P.balance) {
        // ---- replicate genome + amplicon table over the ranks (all-gathers over NVLink) and cut the cell's slots by shard weight.
        // ---- Every rank's block sits at rank * stride of the cell-wide arrays (stride = the largest rank's size, padded), so the
        // ---- gathers are in place and need no displacements.
        const int W = c->P.world, R = c->P.rank;
        unsigned long long etop = 0;
        SCS_CUDA(c, memcpy_sync(c, &etop, c->err_top.p, 8, cudaMemcpyDeviceToHost));
        std::vector<uint64_t> v(5 * (size_t)W, 0);
        v[R] = c->genome_bases / 32; v[W + R] = etop; v[2 * W + R] = (uint64_t)c->genome_has_n; v[3 * W + R] = n; v[4 * W + R] = c->genome_version;
        if (int rc = allreduce_u64(c, v.data(), v.size())) return rc;
        uint64_t word_stride = 0, err_stride = 0, amp_stride = 0, version_sum = 0; int any_n = 0;
        for (int r = 0; r < W; r++) {
            word_stride = std::max(word_stride, v[r]); err_stride = std::max(err_stride, v[W + r]); amp_stride = std::max(amp_stride, v[3 * W + r]);
            any_n |= (int)v[2 * W + r]; version_sum += v[4 * W + r] * (uint64_t)(r + 1);
        }
        word_stride = (word_stride + 3) & ~3ull;   // 32-byte blocks: genome words (8 B) and mask words (4 B) both stay 16-byte aligned per rank
        err_stride = (err_stride + 1) & ~1ull; amp_stride = std::max<uint64_t>(amp_stride, 1);
        const uint64_t tot_words = word_stride * (uint64_t)W, tot_errs = err_stride * (uint64_t)W;
        SCS_CUDA(c, c->g_words.reserve(tot_words + 16)); SCS_CUDA(c, c->g_nmask.reserve(tot_words + 16));
        SCS_CUDA(c, c->g_desc.reserve(N + 1)); SCS_CUDA(c, c->g_errref.reserve(N + 1)); SCS_CUDA(c, c->g_errs.reserve(tot_errs + 2));
        SCS_CUDA(c, c->g_slot_base.reserve(N + 2)); SCS_CUDA(c, c->g_gather.reserve(3 * amp_stride * (uint64_t)W + 8));
        // the packed genome (and its N mask) only when some rank loaded a new one since the last replication
        if (c->g_genome_version != version_sum || c->g_genome_stride != word_stride) {
            if (v[R]) SCS_CUDA(c, cudaMemcpyAsync(c->g_words.p + (uint64_t)R * word_stride, c->genome_words.p, v[R] * 8, cudaMemcpyDeviceToDevice, c->st));
            if (int rc = allgather_dev(c, c->g_words.p, word_stride, 8)) return rc;
            if (any_n) {
                if (c->genome_has_n && v[R]) SCS_CUDA(c, cudaMemcpyAsync(c->g_nmask.p + (uint64_t)R * word_stride, c->genome_nmask.p, v[R] * 4, cudaMemcpyDeviceToDevice, c->st));
                else SCS_CUDA(c, cudaMemsetAsync(c->g_nmask.p + (uint64_t)R * word_stride, 0, word_stride * 4, c->st));
                if (int rc = allgather_dev(c, c->g_nmask.p, word_stride, 4)) return rc;
            }
            c->g_genome_version = version_sum; c->g_genome_stride = word_stride;
        }
        // error lists and amplicon triples change with every amplification
        if (etop) SCS_CUDA(c, cudaMemcpyAsync(c->g_errs.p + (uint64_t)R * err_stride, c->err_pool.p, etop * 4, cudaMemcpyDeviceToDevice, c->st));
        if (int rc = allgather_dev(c, c->g_errs.p, err_stride, 4)) return rc;
        uint64_t* mine = c->g_gather.p + 3 * amp_stride * (uint64_t)R;
        if (n) {
            SCS_CUDA(c, cudaMemcpyAsync(mine, c->fulls.desc.p, n * 8, cudaMemcpyDeviceToDevice, c->st));
            SCS_CUDA(c, cudaMemcpyAsync(mine + amp_stride, c->fulls.errref.p, n * 8, cudaMemcpyDeviceToDevice, c->st));
            SCS_CUDA(c, cudaMemcpyAsync(mine + 2 * amp_stride, c->full_gidx.p, n * 8, cudaMemcpyDeviceToDevice, c->st));
        }
        if (int rc = allgather_dev(c, c->g_gather.p, 3 * amp_stride, 8)) return rc;
        for (int r = 0; r < W; r++) {
            const uint64_t nr = v[3 * W + r];
            if (!nr) continue;
            scatter_amplicons_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, c->st>>>(c->g_gather.p + 3 * amp_stride * (uint64_t)r, amp_stride, nr,
                                                                                     (uint64_t)r * word_stride * 32, (uint64_t)r * err_stride, c->g_desc.p, c->g_errref.p);
            SCS_LAUNCHED(c);
        }
        // cell-wide slot prefix (identical on every rank) and this rank's range by weight
        uint64_t total_slots = 0;
        { unsigned long long last_base = 0; uint32_t last_slots = 0;
          SCS_CUDA(c, memcpy_sync(c, &last_base, sbase_g.p + (N - 1), 8, cudaMemcpyDeviceToHost));
          SCS_CUDA(c, memcpy_sync(c, &last_slots, tmp.p + (N - 1), 4, cudaMemcpyDeviceToHost));
          total_slots = last_base + last_slots; }
        SCS_CUDA(c, cudaMemcpyAsync(c->g_slot_base.p, sbase_g.p, N * 8, cudaMemcpyDeviceToDevice, c->st));
        SCS_CUDA(c, cudaMemcpyAsync(c->g_slot_base.p + N, &total_slots, 8, cudaMemcpyHostToDevice, c->st));
        std::vector<double> wts((size_t)W, 0.0); wts[R] = c->shard_weight;
        if (int rc = allreduce_f64(c, wts.data(), W)) return rc;
        double wsum = 0; for (double x : wts) wsum += x;
        double acc = 0; uint64_t lo = 0, hi = 0;
        for (int r = 0; r <= R; r++) { lo = hi; acc += wts[r]; hi = (r == W - 1) ? total_slots : (uint64_t)((long double)total_slots * (acc / wsum)); }
        c->g_slot_lo = lo; c->g_slot_hi = std::max(lo, hi); c->g_n_amp = N; c->g_bases = tot_words * 32; c->g_has_n = any_n;
        c->global_view = true;
        SCS_CUDA(c, cudaStreamSynchronize(c->st));
    }
    SCS_CUDA(c, timer.stop(&c->stats.ms_alloc));
    c->have_counts = true;
    return SCS_OK;
}

}  // namespace scs
