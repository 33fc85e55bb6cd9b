// K0 pack_genome: ASCII -> 2-bit packed genome in HBM, plus the host-side FASTA reader, the fragment
// cutter and the device exclusive scan used by the later stages.
//
// Replaces Genome::loadRefSeq + FastaReference::getSubSequence + Fragment::createSequence
// (/root/reference/lib/genome/Genome.cpp:176-198,272-278, lib/fastahack/Fasta.cpp:304-334,
// lib/fragment/Fragment.cpp:40-50): instead of one fseek+fread+toupper copy per fragment and
// strand, the whole genome is packed once (0.25 B/base) and every fragment / amplicon is an
// oriented window into it. Genome::splitToFrags (Genome.cpp:753-782) stays a host loop: it is a
// serial chain of ~2 draws per 55 kb.
#include <algorithm>
#include <cstring>

#include "ctx.h"
#include "fasta_host.h"

namespace scs {

// ---------------------------------------------------------------------------- exclusive scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* total, uint64_t* warp_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < (kScanThreads / 32) ? warp_sums[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        if (lane < (kScanThreads / 32)) warp_sums[lane] = wi - w;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    return warp_sums[warp] + inc - v;
}

// `In` is a pointer or a by-value accessor with operator[] (the GC index below counts bases of packed words on the fly)
template <class In>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(In in, uint64_t n, uint64_t* __restrict__ tile_sums) {
    __shared__ uint64_t ws[kScanThreads / 32]; __shared__ uint64_t tot;
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) if (base + k < n) s += (uint64_t)in[base + k];
    block_exclusive_scan(s, &tot, ws);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

template <class In, class TOut>
__global__ void __launch_bounds__(kScanThreads) scan_tiles(In in, uint64_t n, const uint64_t* __restrict__ tile_offs, TOut* __restrict__ out) {
    __shared__ uint64_t ws[kScanThreads / 32]; __shared__ uint64_t tot;
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint64_t v[kScanItems]; uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { v[k] = (base + k < n) ? (uint64_t)in[base + k] : 0; s += v[k]; }
    uint64_t off = block_exclusive_scan(s, &tot, ws) + (tile_offs ? tile_offs[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { if (base + k < n) out[base + k] = (TOut)off; off += v[k]; }
}

// recursive reduce-then-scan; no inter-block waiting
template <class In, class TOut> static int scan_rec(scs_ctx* c, In in, TOut* out, uint64_t n, uint64_t* total_dev) {
    if (n == 0) return SCS_OK;
    uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    DevBuf<uint64_t> sums, offs;
    SCS_CUDA(c, sums.reserve(tiles + 1));
    scan_tile_sums<In><<<(unsigned)tiles, kScanThreads, 0, c->st>>>(in, n, sums.p); SCS_LAUNCHED(c);
    if (tiles > 1) {
        SCS_CUDA(c, offs.reserve(tiles + 1));
        int rc = scan_rec<const uint64_t*, uint64_t>(c, sums.p, offs.p, tiles, total_dev);
        if (rc) return rc;
        scan_tiles<In, TOut><<<(unsigned)tiles, kScanThreads, 0, c->st>>>(in, n, offs.p, out); SCS_LAUNCHED(c);
    } else {
        scan_tiles<In, TOut><<<1, kScanThreads, 0, c->st>>>(in, n, nullptr, out); SCS_LAUNCHED(c);
        if (total_dev) SCS_CUDA(c, cudaMemcpyAsync(total_dev, sums.p, 8, cudaMemcpyDeviceToDevice, c->st));
    }
    return SCS_OK;
}

// GC index of the packed genome (amplify.cu: the GC content of a product window in four loads instead of a pass over its 32-63
// words): gc[w] = number of C/G bases in words [0, w), n[w] = number of N bases in words [0, w). Wrapping u32: only differences
// of two entries are ever used. Rebuilt when the genome changes.
struct GcCountIn { const uint64_t* w; __device__ __forceinline__ uint64_t operator[](uint64_t i) const { const uint64_t x = w[i]; return (uint64_t)__popcll((x ^ (x >> 1)) & 0x5555555555555555ull); } };
struct NCountIn { const uint32_t* m; __device__ __forceinline__ uint64_t operator[](uint64_t i) const { return (uint64_t)__popc(m[i]); } };

int build_gc_index(scs_ctx* c) {
    if (c->gcidx_version == c->genome_version) return SCS_OK;
    const uint64_t nw = c->genome_bases / 32 + 1;
    SCS_CUDA(c, c->gc_pref.reserve(nw + 1));
    if (int rc = scan_rec<GcCountIn, uint32_t>(c, GcCountIn{c->genome_words.p}, c->gc_pref.p, nw, nullptr)) return rc;
    if (c->genome_has_n) {
        SCS_CUDA(c, c->n_pref.reserve(nw + 1));
        if (int rc = scan_rec<NCountIn, uint32_t>(c, NCountIn{c->genome_nmask.p}, c->n_pref.p, nw, nullptr)) return rc;
    }
    c->gcidx_version = c->genome_version;
    return SCS_OK;
}

int exclusive_scan_u32(scs_ctx* c, const uint32_t* in, uint64_t* out, uint64_t n, uint64_t* total_host) {
    DevBuf<uint64_t> tot;
    SCS_CUDA(c, tot.reserve(1));
    SCS_CUDA(c, cudaMemsetAsync(tot.p, 0, 8, c->st));
    int rc = scan_rec<const uint32_t*, uint64_t>(c, in, out, n, tot.p);
    if (rc) return rc;
    if (total_host) { SCS_CUDA(c, cudaMemcpyAsync(total_host, tot.p, 8, cudaMemcpyDeviceToHost, c->st)); SCS_CUDA(c, cudaStreamSynchronize(c->st)); }
    return SCS_OK;
}

__global__ void __launch_bounds__(kScanThreads) scan_single_tile(const uint64_t* in, uint64_t n, uint64_t* out, uint64_t* total) {
    __shared__ uint64_t ws[kScanThreads / 32]; __shared__ uint64_t tot;
    uint64_t base = (uint64_t)threadIdx.x * kScanItems;
    uint64_t v[kScanItems]; uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
    uint64_t off = block_exclusive_scan(s, &tot, ws);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { if (base + k < n) out[base + k] = off; off += v[k]; }
    if (threadIdx.x == 0) { *total = tot; __threadfence_system(); }   // total may live in mapped host memory
}

int scan_u32_noalloc(scs_ctx* c, const uint32_t* in, uint64_t* out, uint64_t n, uint64_t* scratch, uint64_t* total_dev) {
    if (n == 0) { SCS_CUDA(c, cudaMemsetAsync(total_dev, 0, 8, c->st)); return SCS_OK; }
    const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles > (uint64_t)kScanTile) return c->fail(SCS_E_ARG, "scan_u32_noalloc: too many items");
    scan_tile_sums<const uint32_t*><<<(unsigned)tiles, kScanThreads, 0, c->st>>>(in, n, scratch); SCS_LAUNCHED(c);
    scan_single_tile<<<1, kScanThreads, 0, c->st>>>(scratch, tiles, scratch, total_dev); SCS_LAUNCHED(c);
    scan_tiles<const uint32_t*, uint64_t><<<(unsigned)tiles, kScanThreads, 0, c->st>>>(in, n, scratch, out); SCS_LAUNCHED(c);
    return SCS_OK;
}

// ------------------------------------------------------------------------------- K0 pack_genome
// One thread packs 32 bases (two 16-byte loads) into one u64 word and one u32 of N-mask. Bases other than ACGT (any
// case) become code 4 ("N", as after the reference's complement) and raise the has_n flag.
__global__ void __launch_bounds__(256) pack_genome_kernel(const uint8_t* __restrict__ ascii, uint64_t n_bases, uint64_t* __restrict__ words,
                                                          uint32_t* __restrict__ nmask, unsigned int* __restrict__ bad) {
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t n_words = (n_bases + 31) >> 5;
    if (w >= n_words) return;
    uint64_t base = w << 5;
    uint8_t b[32];
    if (base + 32 <= n_bases && ((reinterpret_cast<uintptr_t>(ascii + base) & 15) == 0)) {
        const uint4* p = reinterpret_cast<const uint4*>(ascii + base);
        uint4 a = __ldg(p), d = __ldg(p + 1);
        memcpy(b, &a, 16); memcpy(b + 16, &d, 16);
    } else {
#pragma unroll
        for (int k = 0; k < 32; k++) b[k] = (base + k < n_bases) ? ascii[base + k] : (uint8_t)'A';
    }
    uint64_t out = 0; uint32_t mask = 0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        uint32_t ch = b[k] & 0xDFu;   // upper-case (Genome::getSubSequence toupper, Genome.cpp:274)
        // A=0x41 C=0x43 G=0x47 T=0x54 -> (ch>>1)&3 = 0,1,3,2 ; fix G/T order with a xor
        uint32_t code = (ch >> 1) & 3u; code ^= (code >> 1);
        const bool bad1 = (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T');
        mask |= (uint32_t)bad1 << k;
        out |= (uint64_t)(bad1 ? 0u : code) << (2 * k);
    }
    words[w] = out; nmask[w] = mask;
    if (mask) atomicOr(bad, 1u);
}

// Same, reading straight from FASTA text with a fixed line geometry (blen bases per line, llen bytes per line incl. the
// line terminator — the .fai model of lib/fastahack/Fasta.cpp:304-334): the newlines are never stripped on the host.
__global__ void __launch_bounds__(256) pack_fasta_kernel(const uint8_t* __restrict__ text, uint64_t n_bases, uint32_t blen, uint32_t llen,
                                                         uint64_t* __restrict__ words, uint32_t* __restrict__ nmask, unsigned int* __restrict__ bad) {
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t n_words = (n_bases + 31) >> 5;
    if (w >= n_words) return;
    uint64_t base = w << 5;
    uint64_t line = base / blen; uint32_t col = (uint32_t)(base % blen);
    const uint8_t* p = text + line * llen + col;
    uint64_t out = 0; uint32_t mask = 0;
    for (int k = 0; k < 32; k++) {
        uint32_t ch = (base + k < n_bases) ? (*p & 0xDFu) : (uint32_t)'A';
        uint32_t code = (ch >> 1) & 3u; code ^= (code >> 1);
        const bool bad1 = (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T');
        mask |= (uint32_t)bad1 << k;
        out |= (uint64_t)(bad1 ? 0u : code) << (2 * k);
        if (++col == blen) { col = 0; p += llen - blen + 1; } else p++;
    }
    words[w] = out; nmask[w] = mask;
    if (mask) atomicOr(bad, 1u);
}

static std::string ref_seq_name(const std::string& header) {   // lib/fastahack/Fasta.cpp:57-68
    std::string name = header.substr(0, header.find_first_of(" \t"));
    size_t i = name.find("chrom");
    if (i == std::string::npos) { i = name.find("chr"); if (i != std::string::npos) name = name.substr(i + 3); }
    else name = name.substr(i + 5);
    return name;
}

// one sequence of the cell on the host: contiguous ASCII bases (blen == 0) or FASTA text with a fixed line geometry

int genome_from_host(scs_ctx* c, int n, const char* const* names, const char* const* seqs, const uint64_t* lens) {
    std::vector<SeqSrc> src((size_t)std::max(n, 0));
    for (int i = 0; i < n; i++) src[i] = {seqs[i], lens[i], 0, 0, false};
    return genome_from_sources(c, n, names, src.data());
}

int genome_from_sources(scs_ctx* c, int n, const char* const* names, const SeqSrc* src) {
    std::vector<uint64_t> lens_v((size_t)std::max(n, 0)); for (int i = 0; i < n; i++) lens_v[i] = src[i].len;
    const uint64_t* lens = lens_v.data();
    if (!c->have_device) return c->fail(SCS_E_CUDA, "no CUDA device: the genreads path has no CPU fallback");
    if (n <= 0) return c->fail(SCS_E_IO, "ERROR: reference sequence cannot be empty!");
    c->seq_names.clear(); c->seq_len.clear(); c->seq_goff.clear();
    uint64_t goff = 0, refLen = 0;
    for (int i = 0; i < n; i++) {
        if (lens[i] >= (1ull << 31)) return c->fail(SCS_E_ARG, "sequence longer than 2^31-1 bases (lib/fastahack/Fasta.h:36)");
        c->seq_names.push_back(ref_seq_name(names[i]));
        c->seq_len.push_back(lens[i]); c->seq_goff.push_back(goff);
        goff += (lens[i] + 31) & ~31ull;
        // Malbac::yieldReads: refLen = sum of atoi(last '_' field) / 2 (Malbac.cpp:413-419)
        const std::string& nm = c->seq_names.back();
        size_t us = nm.rfind('_');
        refLen += (uint64_t)atoi(us == std::string::npos ? nm.c_str() : nm.c_str() + us + 1);
    }
    if (goff >= (1ull << 40)) return c->fail(SCS_E_ARG, "genome too large");
    c->ref_len_sum = refLen; c->ref_len_half = refLen / 2;
    c->genome_bases = goff;
    SCS_CUDA(c, c->genome_words.reserve(goff / 32 + 2)); SCS_CUDA(c, c->genome_nmask.reserve(goff / 32 + 2));
    DevBuf<unsigned int> bad; SCS_CUDA(c, bad.reserve(1)); SCS_CUDA(c, cudaMemsetAsync(bad.p, 0, 4, c->st));
    // stage ASCII through a bounded device buffer
    const uint64_t chunk = 256ull << 20;
    DevBuf<uint8_t>& stage = c->genome_stage;
    bool need_stage = false; for (int i = 0; i < n; i++) need_stage |= !src[i].dev;
    if (need_stage) SCS_CUDA(c, stage.reserve(chunk + (chunk >> 4) + 4096));
    StageTimer timer(c);
    for (int i = 0; i < n; i++) {
        if (src[i].dev) {   // contiguous bases already in device memory (simuvars -> genome without a text round trip)
            uint64_t nw = (lens[i] + 31) >> 5;
            if (nw) { pack_genome_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, c->st>>>((const uint8_t*)src[i].p, lens[i], c->genome_words.p + (c->seq_goff[i] >> 5),
                                                                                        c->genome_nmask.p + (c->seq_goff[i] >> 5), bad.p); SCS_LAUNCHED(c); }
        } else if (src[i].blen == 0) {
            for (uint64_t off = 0; off < lens[i]; off += chunk) {
                uint64_t m = std::min(chunk, lens[i] - off);
                SCS_CUDA(c, cudaMemcpyAsync(stage.p, src[i].p + off, m, cudaMemcpyHostToDevice, c->st));
                uint64_t nw = (m + 31) >> 5;
                pack_genome_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, c->st>>>(stage.p, m, c->genome_words.p + ((c->seq_goff[i] + off) >> 5),
                                                                                    c->genome_nmask.p + ((c->seq_goff[i] + off) >> 5), bad.p);
                SCS_LAUNCHED(c);
                SCS_CUDA(c, cudaStreamSynchronize(c->st));
            }
        } else {
            // FASTA text: chunks of whole lines whose base count is a multiple of 32 (so every chunk starts on a packed word)
            const uint64_t blen = src[i].blen, llen = src[i].llen;
            uint64_t lines_per_chunk = std::max<uint64_t>(32, (chunk / llen) / 32 * 32);
            for (uint64_t l0 = 0; l0 * blen < lens[i]; l0 += lines_per_chunk) {
                const uint64_t b0 = l0 * blen, m = std::min(lines_per_chunk * blen, lens[i] - b0);
                const uint64_t nbytes = m + (m - 1) / blen * (llen - blen);   // text bytes spanned by m bases
                SCS_CUDA(c, cudaMemcpyAsync(stage.p, src[i].p + l0 * llen, nbytes, cudaMemcpyHostToDevice, c->st));
                uint64_t nw = (m + 31) >> 5;
                pack_fasta_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, c->st>>>(stage.p, m, (uint32_t)blen, (uint32_t)llen,
                                                                                   c->genome_words.p + ((c->seq_goff[i] + b0) >> 5),
                                                                                   c->genome_nmask.p + ((c->seq_goff[i] + b0) >> 5), bad.p);
                SCS_LAUNCHED(c);
                SCS_CUDA(c, cudaStreamSynchronize(c->st));
            }
        }
    }
    SCS_CUDA(c, timer.stop(&c->stats.ms_pack));
    unsigned int hbad = 0;
    SCS_CUDA(c, memcpy_sync(c, &hbad, bad.p, 4, cudaMemcpyDeviceToHost));
    c->genome_has_n = hbad ? 1 : 0;
    c->stats.n_sequences = n; c->stats.genome_bases = 0; for (auto l : c->seq_len) c->stats.genome_bases += l;
    c->have_genome = true; c->have_frags = false; c->amplified = false; c->have_counts = false; c->genome_version++;
    return SCS_OK;
}

int genome_from_fasta(scs_ctx* c, const char* path) {
    // One pass over the lines builds the .fai index (name, length, offset, bases per line, bytes per line): with a uniform
    // line geometry (what the reference's fastahack reader requires, lib/fastahack/Fasta.cpp:304-334) the text is uploaded
    // as it is and the newlines are skipped by index arithmetic on the device; otherwise the bases are gathered on the host.
    FastaFile ff; std::string ferr;
    if (!ff.open(path, &ferr)) return c->fail(SCS_E_IO, ferr);
    const std::vector<FaiRec>& fai = ff.fai; const char* raw = ff.data;
    fasta_write_fai(path, fai);
    std::vector<std::string> names; for (auto& r : fai) names.push_back(r.header);
    // irregular records: gather their bases into contiguous host buffers (slow path)
    std::vector<std::vector<char>> gathered(fai.size());
    std::vector<SeqSrc> all(fai.size());
    for (size_t i = 0; i < fai.size(); i++) {
        if (fai[i].len > 0 && fai[i].len <= fai[i].blen) { all[i] = {&raw[fai[i].off], fai[i].len, 0, 0, false}; continue; }   // one line: already contiguous
        if (fai[i].regular && fai[i].len > 0 && fai[i].llen <= (1u << 20)) { all[i] = {&raw[fai[i].off], fai[i].len, fai[i].blen, fai[i].llen, false}; continue; }
        ff.gather(i, gathered[i]);
        all[i] = {gathered[i].data(), gathered[i].size(), 0, 0, false};
    }
    // world > 1: this rank keeps a contiguous run of sequences holding about 1/world of the bases
    std::vector<uint64_t> slens(fai.size()); for (size_t i = 0; i < fai.size(); i++) slens[i] = all[i].len;
    size_t lo = 0, hi = fai.size();
    shard_by_midpoint(slens, c->P.rank, c->P.world, &lo, &hi);
    std::vector<const char*> np; std::vector<SeqSrc> sp;
    for (size_t i = lo; i < hi; i++) { np.push_back(names[i].c_str()); sp.push_back(all[i]); }
    if (np.empty()) {   // more ranks than sequences: this rank holds nothing but still takes part in the collectives
        c->seq_names.clear(); c->seq_len.clear(); c->seq_goff.clear(); c->ref_len_sum = 0; c->ref_len_half = 0; c->genome_bases = 0;
        SCS_CUDA(c, c->genome_words.reserve(2)); SCS_CUDA(c, c->genome_nmask.reserve(2));
        c->genome_has_n = 0; c->have_genome = true; c->have_frags = false; c->amplified = false; c->have_counts = false; c->genome_version++;
        return SCS_OK;
    }
    return genome_from_sources(c, (int)np.size(), np.data(), sp.data());
}

// ------------------------------------------------------------------------------ fragments (host)
DrawSrc draw_src(const scs_ctx* c, int domain) {
    DrawSrc s; s.seed = c->P.seed; s.tape[0] = s.tape[1] = nullptr; s.marks = nullptr;
    if (c->replay.on) {
        switch (domain) {
            case D_FRAG: case D_POIS: s.tape[0] = c->replay.mrand.p; s.tape[1] = c->replay.mrand.p; break;
            case D_MULTM: s.tape[0] = c->replay.mreal.p; s.tape[1] = c->replay.mreal.p; break;
            default: s.tape[0] = c->replay.wreal.p; s.tape[1] = c->replay.wint.p; break;
        }
        s.marks = c->replay.marks[domain].p;
    }
    return s;
}

uint32_t host_draw(const scs_ctx* c, int domain, int engine, uint64_t entity, uint64_t mark_index, uint64_t i) {
    if (c->replay.on) {
        const std::vector<uint32_t>& t = (domain == D_MULTM) ? c->replay.h_mreal : c->replay.h_mrand;
        const std::vector<uint64_t>& hm = c->replay.hmarks[domain];
        if (3 * mark_index + 2 >= hm.size()) return 0u;
        uint64_t off = hm[3 * mark_index + 1 + engine] + i;
        return off < t.size() ? t[off] : 0u;
    }
    uint32_t o[4];
    philox4x32_10((uint32_t)entity, (uint32_t)(entity >> 32), (uint32_t)(i >> 2), (uint32_t)(domain * 2 + engine), (uint32_t)c->P.seed,
                  (uint32_t)(c->P.seed >> 32), o);
    return o[i & 3];
}

// ---- collectives. With a communicator (scs_nccl_init) everything runs on NCCL: small host vectors are staged through a device
// ---- scratch buffer (a few words per amplification pass), device vectors are reduced / gathered in place over NVLink.
#define SCS_NCCL(ctx, call)                                                                                                        \
    do {                                                                                                                           \
        ncclResult_t r__ = (call);                                                                                                 \
        if (r__ != ncclSuccess) return (ctx)->fail(SCS_E_CUDA, std::string("NCCL error: ") + nccl_api().GetErrorString(r__) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)

static int nccl_allreduce_host(scs_ctx* c, void* v, size_t n, ncclDataType_t t) {
    SCS_CUDA(c, c->coll_dev.reserve(n + 8));
    SCS_CUDA(c, cudaMemcpyAsync(c->coll_dev.p, v, n * 8, cudaMemcpyHostToDevice, c->st));
    SCS_NCCL(c, nccl_api().AllReduce(c->coll_dev.p, c->coll_dev.p, n, t, ncclSum, c->comm, c->st));
    SCS_CUDA(c, cudaMemcpyAsync(v, c->coll_dev.p, n * 8, cudaMemcpyDeviceToHost, c->st));
    SCS_CUDA(c, cudaStreamSynchronize(c->st));
    return SCS_OK;
}

int allreduce_u64(scs_ctx* c, uint64_t* v, size_t n) {
    if (c->P.world <= 1 || n == 0) return SCS_OK;
    if (c->comm) return nccl_allreduce_host(c, v, n, ncclUint64);
    if (!c->ar_u64) return c->fail(SCS_E_STATE, "world > 1 but no collectives set (scs_nccl_init or scs_set_collectives)");
    return c->ar_u64(c->ar_user, v, n) ? c->fail(SCS_E_STATE, "allreduce callback failed") : SCS_OK;
}
int allreduce_f64(scs_ctx* c, double* v, size_t n) {
    if (c->P.world <= 1 || n == 0) return SCS_OK;
    if (c->comm) return nccl_allreduce_host(c, v, n, ncclFloat64);
    if (!c->ar_f64) return c->fail(SCS_E_STATE, "world > 1 but no collectives set (scs_nccl_init or scs_set_collectives)");
    return c->ar_f64(c->ar_user, v, n) ? c->fail(SCS_E_STATE, "allreduce callback failed") : SCS_OK;
}

int allreduce_dev_i64(scs_ctx* c, void* dev, size_t n) {
    if (c->P.world <= 1 || n == 0) return SCS_OK;
    if (c->comm) { SCS_NCCL(c, nccl_api().AllReduce(dev, dev, n, ncclUint64, ncclSum, c->comm, c->st)); return SCS_OK; }
    if (c->ar_dev_i64) {
        SCS_CUDA(c, cudaStreamSynchronize(c->st));
        return c->ar_dev_i64(c->ar_dev_user, (int64_t*)dev, n) ? c->fail(SCS_E_STATE, "device allreduce callback failed") : SCS_OK;
    }
    std::vector<uint64_t> h(n);
    SCS_CUDA(c, memcpy_sync(c, h.data(), dev, n * 8, cudaMemcpyDeviceToHost));
    if (int rc = allreduce_u64(c, h.data(), n)) return rc;
    SCS_CUDA(c, memcpy_sync(c, dev, h.data(), n * 8, cudaMemcpyHostToDevice));
    return SCS_OK;
}
int allreduce_dev_f64(scs_ctx* c, double* dev, size_t n) {
    if (c->P.world <= 1 || n == 0) return SCS_OK;
    if (c->comm) { SCS_NCCL(c, nccl_api().AllReduce(dev, dev, n, ncclFloat64, ncclSum, c->comm, c->st)); return SCS_OK; }
    if (c->ar_dev_f64) {   // caller's hook on the device buffer
        SCS_CUDA(c, cudaStreamSynchronize(c->st));
        return c->ar_dev_f64(c->ar_dev_user, dev, n) ? c->fail(SCS_E_STATE, "device allreduce callback failed") : SCS_OK;
    }
    std::vector<double> h(n);
    SCS_CUDA(c, memcpy_sync(c, h.data(), dev, n * 8, cudaMemcpyDeviceToHost));
    if (int rc = allreduce_f64(c, h.data(), n)) return rc;
    SCS_CUDA(c, memcpy_sync(c, dev, h.data(), n * 8, cudaMemcpyHostToDevice));
    return SCS_OK;
}

int allgather_dev(scs_ctx* c, void* buf, size_t count, size_t elem) {
    const int W = c->P.world, R = c->P.rank;
    if (W <= 1 || count == 0) return SCS_OK;
    const size_t block = count * elem;
    if (block % 8) return c->fail(SCS_E_ARG, "allgather_dev: block size must be a multiple of 8 bytes");
    char* base = static_cast<char*>(buf);
    if (c->comm) {   // in place: this rank's block is already at its position
        SCS_NCCL(c, nccl_api().AllGather(base + (size_t)R * block, base, block, ncclUint8, c->comm, c->st));
        return SCS_OK;
    }
    // hooks: every other block zeroed, then a sum over ranks (each element is non-zero on at most one rank)
    if (R > 0) SCS_CUDA(c, cudaMemsetAsync(base, 0, (size_t)R * block, c->st));
    if (R + 1 < W) SCS_CUDA(c, cudaMemsetAsync(base + (size_t)(R + 1) * block, 0, (size_t)(W - 1 - R) * block, c->st));
    return allreduce_dev_i64(c, buf, (size_t)W * block / 8);
}

int create_frags(scs_ctx* c) {   // Genome::splitToFrags, Genome.cpp:753-782
    if (!c->have_genome) return c->fail(SCS_E_STATE, "scs_create_frags: no genome loaded");
    const uint32_t fragMin = 10000, fragMax = 100000;   // Fragment.cpp:15-16
    const int W = std::max(1, c->P.world), R = c->P.rank;
    // global sequence numbering: rank r's sequences follow those of ranks < r
    std::vector<uint64_t> v(2 * (size_t)W + 1, 0);
    v[R] = c->seq_len.size();
    if (int rc = allreduce_u64(c, v.data(), W)) return rc;
    c->seq_global0 = 0; for (int r = 0; r < R; r++) c->seq_global0 += v[r];
    std::vector<HostFrag> all;
    for (size_t s = 0; s < c->seq_len.size(); s++) {
        int64_t chrLen = (int64_t)c->seq_len[s], start = 1; uint64_t i = 0;
        while (start <= chrLen) {
            int32_t fl = (int32_t)uni_trunc(host_draw(c, D_FRAG, E_REAL, c->seq_global0 + s, s, i++), fragMin, fragMax + 1 - fragMin);
            if (start + fl - 1 > chrLen) break;
            all.push_back({(int32_t)s, start - 1, fl, 1});
            all.push_back({(int32_t)s, start - 1, fl, -1});
            start += fl;
        }
        if (start <= chrLen) {   // tail emitted twice on strand +1 (quirk kept)
            int32_t fl = (int32_t)(chrLen - start + 1);
            all.push_back({(int32_t)s, start - 1, fl, 1});
            all.push_back({(int32_t)s, start - 1, fl, 1});
        }
    }
    // fragment numbering, total template length and the cell's reference length over all ranks
    std::fill(v.begin(), v.end(), 0);
    v[R] = all.size();
    for (auto& f : all) v[W + R] += (uint64_t)f.len;
    v[2 * W] = c->ref_len_sum;
    if (int rc = allreduce_u64(c, v.data(), v.size())) return rc;
    c->frag_global0 = 0; c->n_frags_global = 0; c->frag_len_sum_global = 0;
    for (int r = 0; r < W; r++) { if (r < R) c->frag_global0 += v[r]; c->n_frags_global += v[r]; c->frag_len_sum_global += v[W + r]; }
    c->ref_len_half = v[2 * W] / 2;
    c->stats.n_frags = all.size();
    c->frag_lo = 0; c->frag_hi = all.size();
    c->frags = all;
    uint64_t nloc = all.size();
    std::vector<uint64_t> desc(nloc);
    for (uint64_t k = 0; k < nloc; k++) {
        const HostFrag& f = all[k];
        uint64_t g = c->seq_goff[f.seq] + (uint64_t)f.start0;
        // amplification template = complement(stored): strand -1 -> genome forward, strand +1 -> reverse complement
        desc[k] = f.strand == 1 ? pack_desc(g + f.len - 1, 1, (uint32_t)f.len) : pack_desc(g, 0, (uint32_t)f.len);
    }
    SCS_CUDA(c, c->frag_desc.reserve(nloc + 1)); SCS_CUDA(c, c->frag_primers.reserve(nloc + 1));
    if (nloc) SCS_CUDA(c, cudaMemcpyAsync(c->frag_desc.p, desc.data(), nloc * 8, cudaMemcpyHostToDevice, c->st));
    SCS_CUDA(c, cudaMemsetAsync(c->frag_primers.p, 0, (nloc + 1) * 4, c->st));
    SCS_CUDA(c, cudaStreamSynchronize(c->st));
    c->have_frags = true; c->amplified = false; c->have_counts = false;
    return SCS_OK;
}

}  // namespace scs

extern "C" void scs_shard_range(uint64_t n, int rank, int world, uint64_t* lo, uint64_t* hi) {
    if (world < 1) world = 1;
    if (rank < 0) rank = 0;
    if (rank >= world) rank = world - 1;
    uint64_t q = n / (uint64_t)world, r = n % (uint64_t)world;
    *lo = q * (uint64_t)rank + std::min<uint64_t>((uint64_t)rank, r);
    *hi = *lo + q + ((uint64_t)rank < r ? 1 : 0);
}
