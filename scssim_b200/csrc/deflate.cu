// Optional block-gzip (BGZF) output of the read stage: the packed FASTQ slab is cut into 32 KiB pieces and every piece becomes one
// gzip member holding one dynamic-Huffman deflate block of literals (SURVEY.md §8f row N3 "optional block-gzip"; the reference's
// SeqWriter writes plain text only, /root/reference/lib/seqwriter/SeqWriter.cpp:41-54). ~2.1x fewer bytes over PCIe and on disk.
//
// One CTA per piece: every thread owns 128 consecutive input bytes — code lengths summed, block-wide exclusive scan of the bit
// counts, codes OR-ed into a shared-memory bit buffer (64-bit accumulator, atomics only on the two words a thread shares with
// its neighbours), CRC-32 per thread by table and combined across the CTA by multiplication with x^(8*trailing bytes) mod p (the
// identity behind zlib's crc32_combine). Members go to a fixed-stride staging area with their sizes; the same scan + realigning
// compaction kernels that pack FASTQ records pack the members. The Huffman code is fixed for the run (deflate_host.h).
#include "ctx.h"
#include "deflate_host.h"

namespace scs {

constexpr int kGzPiece = 32768;                 // uncompressed bytes per BGZF block (the format allows < 64 KiB compressed)
constexpr int kGzThreads = 256;
constexpr int kGzPerThread = kGzPiece / kGzThreads;   // 128
constexpr int kGzOutWords = 6144;               // 24 KiB bit buffer (FASTQ deflates to ~16 KiB per piece); a piece that does not fit is stored
static_assert(kGzPerThread == 128, "thread tile");

__device__ __forceinline__ uint32_t gf2_multmodp(uint32_t a, uint32_t b) {   // a(x) * b(x) mod p(x), reflected CRC-32 polynomial
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
__device__ __forceinline__ uint32_t gf2_x8n_modp(uint32_t nbytes, const uint32_t* __restrict__ x2n) {   // x^(8 * nbytes) mod p
    uint32_t p = 1u << 31; uint32_t k = 3;
    while (nbytes) { if (nbytes & 1u) p = gf2_multmodp(x2n[k & 31u], p); nbytes >>= 1; k++; }
    return p;
}

// sizes_out[k] = bytes of member k written at stage + k * stride (0 for pieces past the end of the input); piece index n_pieces gets
// the 28-byte BGZF end-of-file marker when append_eof is set.
__global__ void __launch_bounds__(kGzThreads) deflate_pieces_kernel(const char* __restrict__ in, const uint64_t* __restrict__ offs, const uint32_t* __restrict__ sizes,
                                                                    uint64_t nrec, const uint32_t* __restrict__ code_g, const uint32_t* __restrict__ prefix_g,
                                                                    uint32_t prefix_bits, const uint32_t* __restrict__ crc_g, const uint32_t* __restrict__ x2n_g,
                                                                    char* __restrict__ stage, uint64_t stride, uint32_t max_pieces, uint32_t* __restrict__ sizes_out,
                                                                    int append_eof, int* __restrict__ flags) {
    extern __shared__ __align__(16) uint32_t sm[];
    uint32_t* outw = sm;
    uint32_t* scode = outw + kGzOutWords;   // 257 entries (+3 pad)
    uint32_t* scrc = scode + 260;           // 4 x 256 (slicing-by-4)
    uint32_t* sx2n = scrc + 1024;           // 32 powers x^(2^n) + 256 tile shifts x^(1024 j)
    __shared__ uint32_t s_warp[kGzThreads / 32], s_crcw[kGzThreads / 32], s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 257; i += kGzThreads) scode[i] = code_g[i];
    for (int i = tid; i < 1024; i += kGzThreads) scrc[i] = crc_g[i];
    for (int i = tid; i < 32 + 256; i += kGzThreads) sx2n[i] = x2n_g[i];
    const uint64_t in_total = nrec ? offs[nrec - 1] + (uint64_t)sizes[nrec - 1] : 0;
    const uint64_t n_pieces = (in_total + kGzPiece - 1) / kGzPiece;
    const uint32_t prefix_words = (prefix_bits + 31) >> 5;
    __syncthreads();
    for (uint64_t k = blockIdx.x; k <= n_pieces && k < max_pieces; k += gridDim.x) {
        char* dst = stage + k * stride;
        if (k == n_pieces) {   // end-of-file marker: an empty BGZF block
            if (append_eof && tid < 28) {
                const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                dst[tid] = (char)eof[tid];
            }
            if (tid == 0) sizes_out[k] = append_eof ? 28u : 0u;
            continue;
        }
        const uint64_t start = k * (uint64_t)kGzPiece;
        const int n = (int)min((uint64_t)kGzPiece, in_total - start);
        const int beg = tid * kGzPerThread, cnt = max(0, min(kGzPerThread, n - beg));
        // ---- this thread's 128 bytes, 16 at a time (re-read from L1 in the second pass: keeps the loops small); code lengths and CRC
        const uint4* src = reinterpret_cast<const uint4*>(in + start + beg);
        uint32_t bits = 0, crc = 0xFFFFFFFFu, bad = 0;
#pragma unroll 1
        for (int v = 0; 16 * v < cnt; v++) {
            const uint4 x = __ldg(src + v);
            const uint32_t w4[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int left = cnt - (16 * v + 4 * q);   // bytes of this word that belong to the piece
                if (left >= 4) {   // CRC four bytes per step: one dependent table level instead of four
                    const uint32_t t = crc ^ w4[q];
                    crc = scrc[768 + (t & 0xFFu)] ^ scrc[512 + ((t >> 8) & 0xFFu)] ^ scrc[256 + ((t >> 16) & 0xFFu)] ^ scrc[t >> 24];
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (j < left) {
                        const uint32_t b = (w4[q] >> (j * 8)) & 0xFFu;
                        const uint32_t c = scode[b];
                        bits += c >> 24; bad |= (c >> 24) == 0;
                        if (left < 4) crc = scrc[(crc ^ b) & 0xFFu] ^ (crc >> 8);
                    }
                }
            }
        }
        bad = __syncthreads_or((int)bad);   // a byte the run's Huffman code cannot express (cannot happen for FASTQ text): the piece is stored
        // ---- exclusive scan of the bit counts over the CTA
        uint32_t inc = bits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t x = lane < kGzThreads / 32 ? s_warp[lane] : 0, xi = x;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
            if (lane < kGzThreads / 32) s_warp[lane] = xi - x;
            if (lane == kGzThreads / 32 - 1) s_total = xi;
        }
        __syncthreads();
        const uint32_t data_bits = s_total;
        uint32_t bitoff = prefix_bits + s_warp[warp] + inc - bits;
        const uint32_t eob = scode[256];
        const uint32_t end_bits = prefix_bits + data_bits + (eob >> 24);
        const uint32_t nbytes = (end_bits + 7) >> 3, size = nbytes + 8;
        const uint32_t nwords = (size + 3) >> 2;
        // ---- CRC-32 of the piece: crc(A || B) = crc(A) * x^(8|B|) + crc(B) (mod p), so every thread shifts its own and the CTA xors
        // (the bytes after thread t are `tail` bytes of the last active thread plus whole 128-byte tiles: one table entry, two products)
        uint32_t part = 0;
        {
            const int nact = (n + kGzPerThread - 1) / kGzPerThread, tail = n - (nact - 1) * kGzPerThread;
            const uint32_t xtail = tail == kGzPerThread ? sx2n[32 + 1] : gf2_x8n_modp((uint32_t)tail, sx2n);
            if (tid == nact - 1) part = crc ^ 0xFFFFFFFFu;
            else if (tid < nact - 1) part = gf2_multmodp(gf2_multmodp(sx2n[32 + nact - 2 - tid], xtail), crc ^ 0xFFFFFFFFu);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s_crcw[warp] = part;
        __syncthreads();
        uint32_t crc_all = 0;
#pragma unroll
        for (int i = 0; i < kGzThreads / 32; i++) crc_all ^= s_crcw[i];
        if (bad || nwords + 2 > (uint32_t)kGzOutWords) {
            // ---- does not fit the bit buffer (or holds a byte without a code): a STORED deflate block — valid for any input
            const uint32_t ssize = 18u + 5u + (uint32_t)n + 8u;
            uint8_t* db = reinterpret_cast<uint8_t*>(dst);
            if (tid < 18) { const uint8_t gz[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, (uint8_t)((ssize - 1) & 0xFFu), (uint8_t)((ssize - 1) >> 8)}; db[tid] = gz[tid]; }
            if (tid == 18) { db[18] = 1; db[19] = (uint8_t)(n & 0xFF); db[20] = (uint8_t)(n >> 8); db[21] = (uint8_t)(~n & 0xFF); db[22] = (uint8_t)((~n >> 8) & 0xFF); }
            for (int i = tid; i < n; i += kGzThreads) db[23 + i] = (uint8_t)in[start + i];
            if (tid < 4) { db[23 + n + tid] = (uint8_t)(crc_all >> (8 * tid)); db[27 + n + tid] = (uint8_t)((uint32_t)n >> (8 * tid)); }
            if (tid == 0) sizes_out[k] = ssize;
            __syncthreads();
            continue;
        }
        // ---- bit buffer: zero, constant prefix (member header + dynamic-block header), then every thread's codes
        for (uint32_t i = tid; i < nwords + 2; i += kGzThreads) outw[i] = i < prefix_words ? prefix_g[i] : 0u;
        __syncthreads();
        {
            unsigned long long acc = 0; uint32_t nb = bitoff & 31u; uint32_t* wp = outw + (bitoff >> 5); bool first = true;
#pragma unroll 1
            for (int v = 0; 16 * v < cnt; v++) {
                const uint4 x = __ldg(src + v);
                const uint32_t w4[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if (16 * v + j < cnt) {
                        const uint32_t c = scode[(w4[j >> 2] >> ((j & 3) * 8)) & 0xFFu];
                        acc |= (unsigned long long)(c & 0xFFFFFFu) << nb; nb += c >> 24;
                        if (nb >= 32u) {
                            if (first) { atomicOr(wp, (uint32_t)acc); first = false; } else *wp = (uint32_t)acc;   // inner words belong to this thread alone
                            wp++; acc >>= 32; nb -= 32u;
                        }
                    }
                }
            }
            if (first ? (nb > (bitoff & 31u)) : (nb > 0u)) atomicOr(wp, (uint32_t)acc);   // the word shared with the next thread (or the only one)
            if (tid == 0) {   // end-of-block code after the last literal
                const uint32_t eo = prefix_bits + data_bits;
                const unsigned long long e = (unsigned long long)(eob & 0xFFFFFFu) << (eo & 31u);
                atomicOr(outw + (eo >> 5), (uint32_t)e);
                if ((uint32_t)(e >> 32)) atomicOr(outw + (eo >> 5) + 1, (uint32_t)(e >> 32));
            }
        }
        __syncthreads();
        if (tid == 0) {
            uint8_t* ob = reinterpret_cast<uint8_t*>(outw);
            for (int i = 0; i < 4; i++) { ob[nbytes + i] = (uint8_t)(crc_all >> (8 * i)); ob[nbytes + 4 + i] = (uint8_t)((uint32_t)n >> (8 * i)); }
            ob[16] = (uint8_t)((size - 1) & 0xFFu); ob[17] = (uint8_t)((size - 1) >> 8);   // BSIZE
            sizes_out[k] = size;
        }
        __syncthreads();
        // ---- member -> its staging cell (16-byte aligned: stride is a multiple of 16)
        uint4* d4 = reinterpret_cast<uint4*>(dst); const uint4* s4 = reinterpret_cast<const uint4*>(outw);
        for (uint32_t i = tid; i < (size + 15) >> 4; i += kGzThreads) d4[i] = s4[i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- host
int gz_prepare(scs_ctx* c) {
    GzState& Z = c->gz;
    if (Z.ready) return SCS_OK;
    uint64_t hist[257]; DeflateCode D;
    fastq_model_histogram(c->prof, c->P.paired != 0, hist);
    if (!build_deflate_code(hist, D)) return c->fail(SCS_E_STATE, "gzip: could not build the Huffman code");
    uint32_t table[1024], x2n[32 + 256]; crc32_tables(table, x2n);
    SCS_CUDA(c, Z.code.reserve(260)); SCS_CUDA(c, Z.prefix.reserve(D.prefix_words.size() + 4)); SCS_CUDA(c, Z.crc.reserve(1024)); SCS_CUDA(c, Z.x2n.reserve(32 + 256));
    SCS_CUDA(c, memcpy_sync(c, Z.code.p, D.code, 257 * 4, cudaMemcpyHostToDevice));
    SCS_CUDA(c, memcpy_sync(c, Z.prefix.p, D.prefix_words.data(), D.prefix_words.size() * 4, cudaMemcpyHostToDevice));
    SCS_CUDA(c, memcpy_sync(c, Z.crc.p, table, sizeof(table), cudaMemcpyHostToDevice));
    SCS_CUDA(c, memcpy_sync(c, Z.x2n.p, x2n, sizeof(x2n), cudaMemcpyHostToDevice));
    Z.prefix_bits = D.prefix_bits;
    SCS_CUDA(c, cudaFuncSetAttribute(deflate_pieces_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gz_smem_bytes()));
    Z.ready = true;
    return SCS_OK;
}

size_t gz_smem_bytes() { return (size_t)(kGzOutWords + 260 + 1024 + 32 + 256) * 4; }
uint64_t gz_max_pieces(uint64_t slab_bytes) { return (slab_bytes + kGzPiece - 1) / kGzPiece + 1; }   // + the end-of-file marker
uint64_t gz_stage_stride() { return 65536; }

int gz_launch(scs_ctx* c, const char* plain, const uint64_t* offs, const uint32_t* sizes, uint64_t nrec, char* stage, uint32_t max_pieces, uint32_t* piece_sizes,
              int append_eof, int* flags, int sms) {
    GzState& Z = c->gz;
    const unsigned grid = (unsigned)std::min<uint64_t>(max_pieces, (uint64_t)sms * 3);
    deflate_pieces_kernel<<<grid, kGzThreads, gz_smem_bytes(), c->st>>>(plain, offs, sizes, nrec, Z.code.p, Z.prefix.p, Z.prefix_bits, Z.crc.p, Z.x2n.p, stage, gz_stage_stride(),
                                                                       max_pieces, piece_sizes, append_eof, flags);
    SCS_LAUNCHED(c);
    return SCS_OK;
}

}  // namespace scs
