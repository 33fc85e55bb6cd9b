#include "deflate_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <queue>

#include "profile_host.h"

namespace scs {

namespace {

uint32_t bit_reverse(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// LSB-first bit string writer (deflate's bit order)
struct BitWriter {
    std::vector<uint32_t> w; uint32_t nbits = 0;
    void put(uint32_t v, int n) {   // the n low bits of v, least significant first
        for (int i = 0; i < n; i++) {
            if ((nbits & 31u) == 0) w.push_back(0);
            if ((v >> i) & 1u) w.back() |= 1u << (nbits & 31u);
            nbits++;
        }
    }
    void put_huff(uint32_t code, int len) { put(bit_reverse(code, len), len); }   // Huffman codes go most significant bit first
};

}  // namespace

bool build_deflate_code(const uint64_t hist[257], DeflateCode& out) {
    constexpr int kMaxLen = 15, kSyms = 257;
    std::vector<int> used;
    for (int s = 0; s < kSyms; s++) if (hist[s]) used.push_back(s);
    if (used.size() < 2) return false;
    // plain Huffman lengths
    struct Node { uint64_t w; int left, right; };
    std::vector<Node> nodes;
    typedef std::pair<uint64_t, int> QE;
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
    for (int s : used) { nodes.push_back({hist[s], -1, -1}); pq.push({hist[s], (int)nodes.size() - 1}); }
    while (pq.size() > 1) {
        QE a = pq.top(); pq.pop(); QE b = pq.top(); pq.pop();
        nodes.push_back({a.first + b.first, a.second, b.second});
        pq.push({a.first + b.first, (int)nodes.size() - 1});
    }
    std::vector<int> depth(nodes.size(), 0);
    for (int i = (int)nodes.size() - 1; i >= 0; i--) if (nodes[i].left >= 0) { depth[nodes[i].left] = depth[i] + 1; depth[nodes[i].right] = depth[i] + 1; }
    // enforce the 15-bit limit: move the overflow to the maximum length and repair the Kraft sum (the classic deflate fix-up)
    int num[64] = {0};
    for (size_t i = 0; i < used.size(); i++) num[std::min(depth[i], 63)]++;
    for (int l = kMaxLen + 1; l < 64; l++) { num[kMaxLen] += num[l]; num[l] = 0; }
    uint64_t total = 0;
    for (int l = kMaxLen; l >= 1; l--) total += (uint64_t)num[l] << (kMaxLen - l);
    while (total != (1ull << kMaxLen)) {
        num[kMaxLen]--;
        for (int l = kMaxLen - 1; l >= 1; l--) if (num[l]) { num[l]--; num[l + 1] += 2; break; }
        total--;
    }
    // shortest codes to the most frequent symbols (ties: smaller symbol first, so the result is deterministic)
    std::vector<int> order(used.size());
    for (size_t i = 0; i < used.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hist[used[a]] > hist[used[b]]; });
    memset(out.len, 0, sizeof(out.len));
    size_t k = 0;
    for (int l = 1; l <= kMaxLen; l++) for (int c = 0; c < num[l]; c++) out.len[used[order[k++]]] = (uint8_t)l;
    // canonical codes (RFC 1951 3.2.2)
    int bl_count[kMaxLen + 1] = {0}; uint32_t next_code[kMaxLen + 2] = {0};
    for (int s = 0; s < kSyms; s++) bl_count[out.len[s]]++;
    bl_count[0] = 0;
    uint32_t code = 0;
    for (int l = 1; l <= kMaxLen; l++) { code = (code + (uint32_t)bl_count[l - 1]) << 1; next_code[l] = code; }
    double bits = 0, tot = 0;
    for (int s = 0; s < kSyms; s++) {
        out.code[s] = 0;
        if (!out.len[s]) continue;
        out.code[s] = bit_reverse(next_code[out.len[s]]++, out.len[s]) | ((uint32_t)out.len[s] << 24);
        if (s < 256) { bits += (double)hist[s] * out.len[s]; tot += (double)hist[s]; }
    }
    out.expected_bits_per_byte = tot > 0 ? bits / tot : 0;
    // ---- prefix: BGZF member header + dynamic block header
    BitWriter B;
    const uint8_t gz[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, 0, 0};   // bytes 16-17 = BSIZE, patched per block
    for (uint8_t b : gz) B.put(b, 8);
    B.put(1, 1);    // BFINAL
    B.put(2, 2);    // BTYPE = dynamic Huffman
    B.put(0, 5);    // HLIT: 257 literal/length codes
    B.put(0, 5);    // HDIST: 1 distance code (of length 0: the block holds literals only)
    B.put(15, 4);   // HCLEN: all 19 code-length-code lengths follow
    // code-length alphabet: lengths 0..15 as 5-bit codes, zero runs (17: 3-10, 18: 11-138) as 2-bit codes — a complete code
    static const int cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int cl_len[19]; for (int s = 0; s < 19; s++) cl_len[s] = (s <= 15) ? 5 : (s == 16 ? 0 : 2);
    for (int i = 0; i < 19; i++) B.put((uint32_t)cl_len[cl_order[i]], 3);
    uint32_t cl_code[19]; { int cnt[8] = {0}; uint32_t nc[8] = {0}; for (int s = 0; s < 19; s++) cnt[cl_len[s]]++; cnt[0] = 0; uint32_t c = 0;
                            for (int l = 1; l < 8; l++) { c = (c + (uint32_t)cnt[l - 1]) << 1; nc[l] = c; }
                            for (int s = 0; s < 19; s++) cl_code[s] = cl_len[s] ? nc[cl_len[s]]++ : 0; }
    std::vector<int> lens(out.len, out.len + 257); lens.push_back(0);   // + the one distance code
    for (size_t i = 0; i < lens.size();) {
        if (lens[i] == 0) {
            size_t j = i; while (j < lens.size() && lens[j] == 0) j++;
            size_t run = j - i;
            while (run >= 11) { size_t r = std::min<size_t>(run, 138); B.put_huff(cl_code[18], cl_len[18]); B.put((uint32_t)(r - 11), 7); run -= r; }
            if (run >= 3) { B.put_huff(cl_code[17], cl_len[17]); B.put((uint32_t)(run - 3), 3); run = 0; }
            while (run--) B.put_huff(cl_code[0], cl_len[0]);
            i = j;
        } else { B.put_huff(cl_code[lens[i]], cl_len[lens[i]]); i++; }
    }
    out.prefix_words = B.w; out.prefix_bits = B.nbits;
    out.prefix_words.push_back(0);
    return true;
}

void fastq_model_histogram(const HostProfile& P, bool paired, uint64_t hist[257]) {
    double h[257]; for (double& x : h) x = 0;
    const double RL = P.readLength;
    h['@'] += 1; h['#'] += 1; h['\n'] += 4; h['+'] += 1;
    if (paired) { h['/'] += 1; h['1'] += 0.5; h['2'] += 0.5; }
    for (int d = 0; d < 10; d++) h['0' + d] += 0.9;              // ~7 digits of amplicon index + ~2 of the pair number
    h['A'] += RL / 4; h['C'] += RL / 4; h['G'] += RL / 4; h['T'] += RL / 4; h['N'] += RL * 1e-3;
    // marginal quality distribution of the no-substitution pairs, averaged over bins (cdf rows are cumulative sums of the pdf)
    const int bins = P.bins;
    if (!P.qualCdf.empty() && bins > 0) {
        for (int b = 0; b < 4; b++) for (int j = 0; j < bins; j++) {
            const double* row = &P.qualCdf[((size_t)(b * 5) * bins + j) * kQualN];
            double prev = 0;
            for (int q = 0; q < kQualN; q++) { const double p = std::max(0.0, row[q] - prev); prev = row[q]; h[33 + q] += RL * p / (4.0 * bins); }
        }
    } else for (int q = 0; q < 42; q++) h[33 + q] += RL / 42;
    for (int q = 0; q < kQualN; q++) h[33 + q] += RL * 2e-5;     // substitutions / N bases draw from other rows: every Phred char stays encodable
    h[256] = 0.01;                                                // end of block: once per ~100 records
    for (int s = 0; s < 257; s++) hist[s] = h[s] > 0 ? (uint64_t)std::max(1.0, std::floor(h[s] * 1e6)) : 0;
}

void crc32_tables(uint32_t table[1024], uint32_t x2n[32 + 256]) {
    // slicing-by-4: table[k*256 + i] = CRC of byte i followed by k zero bytes
    for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : c >> 1; table[i] = c; }
    for (int k = 1; k < 4; k++) for (uint32_t i = 0; i < 256; i++) table[k * 256 + i] = (table[(k - 1) * 256 + i] >> 8) ^ table[table[(k - 1) * 256 + i] & 0xFFu];
    auto multmodp = [](uint32_t a, uint32_t b) {
        uint32_t m = 1u << 31, p = 0;
        for (;;) { if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; } m >>= 1; b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1; }
        return p;
    };
    uint32_t p = 1u << 30;   // x^1
    x2n[0] = p;
    for (int n = 1; n < 32; n++) x2n[n] = p = multmodp(p, p);
    // x2n[32 + j] = x^(8 * 128 * j) mod p: what a thread's CRC is multiplied with when j full 128-byte thread tiles follow it
    x2n[32] = 1u << 31;   // x^0
    const uint32_t x1024 = x2n[10];   // x^(2^10)
    for (int j = 1; j < 256; j++) x2n[32 + j] = multmodp(x2n[32 + j - 1], x1024);
}

}  // namespace scs
