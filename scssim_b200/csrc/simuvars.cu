// simuvars on the device: materialise every haplotype of the simulated cell from the host-built
// edit plan (simuvars_plan.h) — the work Genome::saveSequence / Genome::generateSegment do with std::string copies,
// inserts and erases (/root/reference/lib/genome/Genome.cpp:329-691).
//
//   sv_normalize_kernel    FASTA text of one chromosome -> contiguous upper-case bases (Genome::getSubSequence's
//                          newline removal + toupper, Genome.cpp:272-278, lib/fastahack/Fasta.cpp:304-334); HBM: 2 B/base
//   sv_materialize_kernel  one thread = one aligned 16-byte vector of the output FASTA body: finds the copy run that holds
//                          its first base (binary search over run starts; a warp's lanes walk the same path, so the
//                          probes coalesce), takes 16 bases with two aligned 16-byte loads + funnel shifts, splices in
//                          the line terminator that falls inside the vector and issues one 16-byte store. Vectors that
//                          straddle a run boundary or the end of the haplotype take a bytewise path. HBM: 1 B read +
//                          1.01 B written per output base
//   sv_scatter_kernel      point substitutions (SNPs / SNVs) at their output positions: 1 byte each
//
// The output either streams to a sink through pinned slabs (the FASTA file of `scssim simuvars -o`) or stays on the
// device and is packed straight into the genreads genome (scs_simuvars_to_genome), with no text round trip.
#include <chrono>
#include <cstring>
#include <memory>

#include "ctx.h"
#include "fasta_host.h"
#include "simuvars_plan.h"

namespace scs {
namespace {

constexpr int kSvThreads = 256;

// C-locale toupper on four packed bytes
__device__ __forceinline__ uint32_t upper4(uint32_t w) {
    const uint32_t lo7 = w & 0x7f7f7f7fu;
    const uint32_t ge_a = lo7 + 0x1f1f1f1fu, gt_z = lo7 + 0x05050505u;
    const uint32_t lower = ge_a & ~gt_z & ~w & 0x80808080u;
    return w - (lower >> 2);
}

constexpr int kSvTile = kSvThreads * 16;   // bytes (bases) one CTA produces

// One thread = 16 output bases. FASTA text with a fixed line geometry: the bases sit at text[line * llen + col]; a vector
// crosses at most one line terminator (blen >= 16), which is cut out by merging two shifted 16-byte windows.
__global__ void __launch_bounds__(kSvThreads) sv_normalize_kernel(const uint8_t* __restrict__ text, uint64_t n_bases, uint32_t blen, uint32_t llen,
                                                                  uint8_t* __restrict__ out) {
    __shared__ uint64_t s_line0; __shared__ uint32_t s_col0;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kSvTile;
    if (blen && threadIdx.x == 0) { s_line0 = tile0 / blen; s_col0 = (uint32_t)(tile0 - s_line0 * blen); }   // one 64-bit division per CTA
    __syncthreads();
    const uint64_t base = tile0 + (uint64_t)threadIdx.x * 16;
    if (base >= n_bases) return;
    uint32_t X[4];
    if (blen == 0) load16(text + base, X);
    else {
        const uint32_t c = s_col0 + threadIdx.x * 16, dl = c / blen, col = c - dl * blen;
        const uint8_t* p = text + (s_line0 + dl) * llen + col;
        if (blen >= 16) {
            load16(p, X);
            const int k = (int)(blen - col);   // bases left on this line
            if (k < 16) {
                uint32_t Y[4]; load16(p + (llen - blen), Y);
#pragma unroll
                for (int m = 0; m < 4; m++) { const uint32_t lo = below_mask(k, m); X[m] = (X[m] & lo) | (Y[m] & ~lo); }
            }
        } else {   // very narrow lines: bytewise
            uint32_t cc = col;
            X[0] = X[1] = X[2] = X[3] = 0;
            for (int k = 0; k < 16; k++) {
                if (base + k < n_bases) X[k >> 2] |= (uint32_t)(*p) << (8 * (k & 3));
                if (++cc == blen) { cc = 0; p += llen - blen + 1; } else p++;
            }
        }
    }
    uint4 o; o.x = upper4(X[0]); o.y = upper4(X[1]); o.z = upper4(X[2]); o.w = upper4(X[3]);
    *reinterpret_cast<uint4*>(out + base) = o;   // the buffer is padded to a multiple of 16
}

struct SvHapArgs {
    const uint64_t* pout;   // [n_pieces + 1] first output base of each run, pout[n_pieces] = n_bases
    const uint64_t* psrc;   // [n_pieces] device address of the run's first source byte
    const uint32_t* blk;    // [n_bases / kSvBlock + 1] run that holds base k * kSvBlock (coarse index: no binary search)
    uint32_t n_pieces; uint32_t W;   // W = bases per output line, 0 = no line terminators (bases only)
    uint64_t n_bases, text_len;
    uint8_t* text;
};
constexpr uint32_t kSvBlockShift = 11;   // one coarse-index entry per 2048 output bases

__global__ void __launch_bounds__(kSvThreads) sv_materialize_kernel(const SvHapArgs A) {
    __shared__ uint64_t s_line0; __shared__ uint32_t s_col0;
    const uint32_t W = A.W;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kSvTile;
    if (W && threadIdx.x == 0) { s_line0 = tile0 / (W + 1); s_col0 = (uint32_t)(tile0 - s_line0 * (W + 1)); }
    __syncthreads();
    const uint64_t b0 = tile0 + (uint64_t)threadIdx.x * 16;
    if (b0 >= A.text_len) return;
    uint64_t i0 = b0; uint32_t col = 0; int k_nl = 16;   // k_nl: byte of the vector that holds the line terminator (>= 16: none)
    if (W) { const uint32_t c = s_col0 + threadIdx.x * 16, dl = c / (W + 1); col = c - dl * (W + 1); i0 = (s_line0 + dl) * W + col; k_nl = (int)(W - col); }
    const bool whole = W ? (b0 + 16 < A.text_len && W >= 16) : (b0 + 16 <= A.text_len);   // W: the last byte of the text is handled bytewise
    // run holding base i0, from the coarse index (clamped for the vector that is only the final line terminator)
    const uint64_t iq = i0 < A.n_bases ? i0 : A.n_bases - 1;
    uint32_t p = __ldg(A.blk + (iq >> kSvBlockShift));
    while (__ldg(A.pout + p + 1) <= iq) p++;
    const uint32_t nb = k_nl < 16 ? 15 : 16;
    if (whole && i0 + nb <= __ldg(A.pout + p + 1)) {
        uint32_t X[4];
        load16(reinterpret_cast<const uint8_t*>(__ldg(A.psrc + p) + (i0 - __ldg(A.pout + p))), X);
        // splice '\n' in at byte k_nl (branch-free; k_nl >= 16 leaves X as it is): bytes above it come from X shifted up by one
        uint32_t Y[4];
        Y[0] = X[0] << 8;
#pragma unroll
        for (int m = 1; m < 4; m++) Y[m] = __funnelshift_l(X[m - 1], X[m], 8);
        uint4 o; uint32_t* O = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const uint32_t lo = below_mask(k_nl, m), upto = below_mask(k_nl + 1, m);
            O[m] = (X[m] & lo) | (0x0A0A0A0Au & upto & ~lo) | (Y[m] & ~upto);
        }
        *reinterpret_cast<uint4*>(A.text + b0) = o;
        return;
    }
    // bytewise: run boundary inside the vector, or the tail of the haplotype
    uint64_t i = i0;
    for (int k = 0; k < 16; k++) {
        const uint64_t b = b0 + k;
        if (b >= A.text_len) break;
        uint8_t ch;
        if (W && (col == W || b == A.text_len - 1)) { ch = '\n'; col = 0; }
        else {
            while (i >= __ldg(A.pout + p + 1)) p++;
            ch = *reinterpret_cast<const uint8_t*>(__ldg(A.psrc + p) + (i - __ldg(A.pout + p)));
            i++; col++;
        }
        A.text[b] = ch;
    }
}

__global__ void __launch_bounds__(kSvThreads) sv_scatter_kernel(const uint64_t* __restrict__ subs, uint64_t n, uint32_t W, uint8_t* __restrict__ text) {
    const uint64_t i = (uint64_t)blockIdx.x * kSvThreads + threadIdx.x;
    if (i >= n) return;
    const uint64_t s = subs[i], pos = s >> 8;
    text[W ? pos + pos / W : pos] = (uint8_t)(s & 0xff);
}

struct HapJob {
    DevBuf<uint8_t> own_text; DevBuf<uint8_t>* text = &own_text;   // sink mode: one of the context's two persistent buffers
    DevBuf<uint64_t> tables;
    uint64_t text_len = 0, n_bases = 0; std::string header; size_t hap_index = 0;
    cudaEvent_t done = nullptr;
    ~HapJob() { if (done) cudaEventDestroy(done); }
};

}  // namespace

// contiguous run [lo, hi) of sequences this rank keeps: cut where the cumulative length crosses rank/world of the total
// (sequence midpoints decide), so every rank holds about the same number of bases
void shard_by_midpoint(const std::vector<uint64_t>& lens, int rank, int world, size_t* lo_out, size_t* hi_out) {
    size_t lo = 0, hi = lens.size();
    if (world > 1) {
        long double total = 0; for (uint64_t l : lens) total += l;
        long double acc = 0; lo = hi = lens.size(); bool started = false;
        for (size_t i = 0; i < lens.size(); i++) {
            long double mid = acc + lens[i] / 2.0L;
            int owner = std::min(world - 1, (int)(mid * world / (total > 0 ? total : 1)));
            if (owner == rank) { if (!started) { lo = i; started = true; } hi = i + 1; }
            acc += lens[i];
        }
        if (!started) lo = hi = 0;
    }
    *lo_out = lo; *hi_out = hi;
}

int simuvars_run(scs_ctx* c, const scs_simuvars_params& sp, const char* ref, const char* snp, const char* var, scs_sink_fn sink, void* user, bool to_genome) {
    if (!c->have_device) return c->fail(SCS_E_CUDA, "no CUDA device: simuvars has no CPU fallback");
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t_begin = clk::now();
    scs_simuvars_stats& S = c->sv_stats; S = scs_simuvars_stats{};
    const uint32_t W = to_genome ? 0u : (uint32_t)(sp.line_width > 0 ? sp.line_width : 100);

    FastaFile ff; std::string ferr;
    if (!ff.open(ref, &ferr)) return c->fail(SCS_E_IO, ferr);
    const std::vector<FaiRec>& fai = ff.fai; const char* raw = ff.data;
    fasta_write_fai(ref, fai);
    std::vector<sv::ChromIn> chroms;
    for (const FaiRec& r : fai) chroms.push_back({strip_chr_prefix(r.name), r.len});
    const auto t_read = clk::now();
    sv::Plan plan;
    if (!sv::build_plan(plan, chroms, snp, var, sp.ploidy > 0 ? sp.ploidy : 2, sp.libc_seed)) return c->fail(SCS_E_IO, plan.err);
    const auto t_plan = clk::now();
    c->sv_warnings = plan.warnings;
    S.n_chroms = chroms.size(); S.n_haps = plan.haps.size(); S.n_segments = plan.n_segments; S.n_pieces = plan.pieces.size(); S.n_subs = plan.subs.size();
    S.n_cnv = plan.n_cnv; S.n_snv = plan.n_snv; S.n_ins = plan.n_ins; S.n_del = plan.n_del; S.n_snp = plan.n_snp;
    for (auto& ch : chroms) S.ref_bases += ch.len;
    S.ms_read = secs(t_begin, t_read); S.ms_plan = secs(t_read, t_plan);

    // which haplotypes this rank materialises (everything unless the cell goes straight into a sharded genome)
    size_t own_lo = 0, own_hi = plan.haps.size();
    if (to_genome && c->P.world > 1) { std::vector<uint64_t> lens; for (auto& h : plan.haps) lens.push_back(h.len); shard_by_midpoint(lens, c->P.rank, c->P.world, &own_lo, &own_hi); }

    DevBuf<uint8_t> lit; SCS_CUDA(c, lit.reserve(plan.literals.size() + 64));
    if (!plan.literals.empty()) SCS_CUDA(c, cudaMemcpyAsync(lit.p, plan.literals.data(), plan.literals.size(), cudaMemcpyHostToDevice, c->st));

    constexpr uint64_t kSlab = 32ull << 20;
    char** pinned = c->sv_pinned; cudaEvent_t copied[2] = {nullptr, nullptr};
    struct Cleanup { cudaEvent_t* e; ~Cleanup() { for (int i = 0; i < 2; i++) if (e[i]) cudaEventDestroy(e[i]); } } cleanup{copied};
    if (!to_genome) for (int i = 0; i < 2; i++) {
        if (!pinned[i]) SCS_CUDA(c, cudaHostAlloc((void**)&pinned[i], kSlab, cudaHostAllocDefault));
        SCS_CUDA(c, cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
    }

    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;   // kernel groups on the compute stream
    struct EvFree { std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* v; ~EvFree() { for (auto& p : *v) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); } } } evfree{&timed};
    auto timed_begin = [&]() { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, c->st); timed.push_back({a, b}); };
    auto timed_end = [&]() { cudaEventRecord(timed.back().second, c->st); };

    // stream one finished haplotype to the sink: header line from the host, body through the pinned slabs
    auto drain = [&](HapJob& J) -> int {
        if (sink(user, 0, J.header.data(), J.header.size())) return c->fail(SCS_E_IO, "simuvars: sink failed");
        SCS_CUDA(c, cudaStreamWaitEvent(c->st_copy, J.done, 0));
        const uint64_t n_slabs = (J.text_len + kSlab - 1) / kSlab;
        auto issue = [&](uint64_t k) -> cudaError_t {
            const uint64_t off = k * kSlab, m = std::min(kSlab, J.text_len - off);
            cudaError_t e = cudaMemcpyAsync(pinned[k & 1], J.text->p + off, m, cudaMemcpyDeviceToHost, c->st_copy);
            return e != cudaSuccess ? e : cudaEventRecord(copied[k & 1], c->st_copy);
        };
        if (n_slabs) SCS_CUDA(c, issue(0));
        for (uint64_t k = 0; k < n_slabs; k++) {
            SCS_CUDA(c, cudaEventSynchronize(copied[k & 1]));
            if (k + 1 < n_slabs) SCS_CUDA(c, issue(k + 1));
            const uint64_t m = std::min(kSlab, J.text_len - k * kSlab);
            if (sink(user, 0, pinned[k & 1], m)) return c->fail(SCS_E_IO, "simuvars: sink failed");
            S.out_bytes += m;
        }
        S.out_bytes += J.header.size();
        return SCS_OK;
    };

    std::vector<std::unique_ptr<HapJob>> jobs;   // sink mode: at most two in flight; genome mode: all of this rank's haplotypes
    DevBuf<uint8_t>* dref = &c->sv_ref;   // one buffer: the next chromosome's normalise is stream-ordered behind the kernels that read it
    size_t n_jobs = 0;
    size_t drained = 0; long cur_chrom = -1;
    const auto t_dev0 = clk::now();
    for (size_t hi = 0; hi < plan.haps.size(); hi++) {
        const sv::Hap& H = plan.haps[hi];
        if (hi < own_lo || hi >= own_hi) continue;
        if (cur_chrom != (long)H.chrom) {
            cur_chrom = (long)H.chrom;
            // (re)load the chromosome: FASTA text up, contiguous upper-case bases out
            const FaiRec& r = fai[H.chrom];
            std::vector<char> gathered; const char* src = &raw[r.off]; uint64_t nbytes = r.len; uint32_t blen = 0, llen = 0;
            if (r.len > r.blen) {
                if (r.regular && r.llen <= (1u << 20)) { blen = r.blen; llen = r.llen; nbytes = r.len + (r.len - 1) / blen * (llen - blen); }
                else { ff.gather(H.chrom, gathered); src = gathered.data(); nbytes = gathered.size(); }
            }
            DevBuf<uint8_t>& stage = c->sv_stage; SCS_CUDA(c, stage.reserve(nbytes + 64));
            SCS_CUDA(c, dref->reserve(((r.len + 15) & ~15ull) + 64));
            if (nbytes) SCS_CUDA(c, cudaMemcpyAsync(stage.p, src, nbytes, cudaMemcpyHostToDevice, c->st));
            S.h2d_bytes += nbytes;
            if (r.len) {
                timed_begin();
                sv_normalize_kernel<<<(unsigned)((r.len + 16 * kSvThreads - 1) / (16 * kSvThreads)), kSvThreads, 0, c->st>>>((const uint8_t*)stage.p, r.len, blen, llen, dref->p);
                SCS_LAUNCHED(c); S.launches++;
                timed_end();
                S.normalize_bytes += nbytes + r.len;
            }
            SCS_CUDA(c, cudaStreamSynchronize(c->st));   // `gathered` goes out of scope
        }
        std::unique_ptr<HapJob> J(new HapJob());
        J->hap_index = hi; J->n_bases = H.len; J->header = ">" + H.name + "\n";
        J->text_len = W ? H.len + (H.len + W - 1) / W : H.len;
        const uint64_t np = H.piece_hi - H.piece_lo, ns = H.sub_hi - H.sub_lo;
        if (!to_genome) J->text = &c->sv_text[n_jobs & 1];   // the job two back has been drained (host-synchronously) by now
        n_jobs++;
        SCS_CUDA(c, J->text->reserve(((J->text_len + 15) & ~15ull) + 64));
        SCS_CUDA(c, cudaEventCreateWithFlags(&J->done, cudaEventDisableTiming));
        if (H.len) {
            const uint64_t nblk = (H.len >> kSvBlockShift) + 1;
            std::vector<uint64_t> tab(2 * np + 1 + ns + (nblk + 1) / 2);
            for (uint64_t k = 0; k < np; k++) {
                const sv::Piece& P = plan.pieces[H.piece_lo + k];
                tab[k] = P.out;
                tab[np + 1 + k] = (P.src & sv::kLiteral) ? (uint64_t)(uintptr_t)(lit.p + (P.src & ~sv::kLiteral)) : (uint64_t)(uintptr_t)(dref->p + P.src);
            }
            tab[np] = H.len;
            for (uint64_t k = 0; k < ns; k++) { const sv::Sub& s = plan.subs[H.sub_lo + k]; tab[2 * np + 1 + k] = (s.out << 8) | s.ch; }
            uint32_t* blk = reinterpret_cast<uint32_t*>(tab.data() + 2 * np + 1 + ns);   // run holding base k * kSvBlock
            for (uint64_t k = 0, pc = 0; k < nblk; k++) {
                const uint64_t at = std::min<uint64_t>(k << kSvBlockShift, H.len - 1);
                while (pc + 1 < np && plan.pieces[H.piece_lo + pc + 1].out <= at) pc++;
                blk[k] = (uint32_t)pc;
            }
            SCS_CUDA(c, J->tables.reserve(tab.size() + 2));
            SCS_CUDA(c, cudaMemcpyAsync(J->tables.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, c->st));
            SCS_CUDA(c, cudaStreamSynchronize(c->st));   // `tab` is pageable and goes out of scope
            S.h2d_bytes += tab.size() * 8;
            SvHapArgs A; A.pout = J->tables.p; A.psrc = J->tables.p + np + 1; A.blk = reinterpret_cast<const uint32_t*>(J->tables.p + 2 * np + 1 + ns); A.n_pieces = (uint32_t)np; A.W = W; A.n_bases = H.len; A.text_len = J->text_len; A.text = J->text->p;
            timed_begin();
            sv_materialize_kernel<<<(unsigned)((J->text_len + 16 * kSvThreads - 1) / (16 * kSvThreads)), kSvThreads, 0, c->st>>>(A);
            SCS_LAUNCHED(c); S.launches++;
            if (ns) { sv_scatter_kernel<<<(unsigned)((ns + kSvThreads - 1) / kSvThreads), kSvThreads, 0, c->st>>>(J->tables.p + 2 * np + 1, ns, W, J->text->p); SCS_LAUNCHED(c); S.launches++; }
            timed_end();
            S.materialize_bytes += H.len + J->text_len + 16 * np + 9 * ns + 4 * nblk;
        }
        SCS_CUDA(c, cudaEventRecord(J->done, c->st));
        S.out_bases += H.len;
        jobs.push_back(std::move(J));
        if (!to_genome && jobs.size() - drained == 2) { if (int rc = drain(*jobs[drained])) return rc; jobs[drained].reset(); drained++; }
    }
    if (!to_genome) for (; drained < jobs.size(); drained++) { if (int rc = drain(*jobs[drained])) return rc; jobs[drained].reset(); }
    SCS_CUDA(c, cudaStreamSynchronize(c->st));
    for (auto& p : timed) { float ms = 0; cudaEventElapsedTime(&ms, p.first, p.second); S.ms_kernels += ms; }
    S.ms_device = secs(t_dev0, clk::now());

    int rc = SCS_OK;
    if (to_genome) {
        std::vector<const char*> names; std::vector<SeqSrc> srcs; std::vector<std::string> keep;
        for (auto& J : jobs) keep.push_back(plan.haps[J->hap_index].name);
        for (size_t i = 0; i < jobs.size(); i++) { names.push_back(keep[i].c_str()); srcs.push_back({(const char*)jobs[i]->text->p, jobs[i]->n_bases, 0, 0, true}); }
        if (names.empty()) {
            c->seq_names.clear(); c->seq_len.clear(); c->seq_goff.clear(); c->ref_len_sum = 0; c->ref_len_half = 0; c->genome_bases = 0;
            SCS_CUDA(c, c->genome_words.reserve(2)); SCS_CUDA(c, c->genome_nmask.reserve(2));
            c->genome_has_n = 0; c->have_genome = true; c->have_frags = false; c->amplified = false; c->have_counts = false; c->genome_version++;
        } else rc = genome_from_sources(c, (int)names.size(), names.data(), srcs.data());
        SCS_CUDA(c, cudaStreamSynchronize(c->st));
    }
    S.ms_total = secs(t_begin, clk::now());
    return rc;
}

}  // namespace scs
