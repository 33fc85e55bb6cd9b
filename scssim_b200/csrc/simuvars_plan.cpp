// See simuvars_plan.h. Host-only; no CUDA here.
#include "simuvars_plan.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>

#include "fasta_host.h"

namespace scs {
namespace sv {

// ------------------------------------------------------------------------------------------ libc rand()
void LibcRand::reseed(uint32_t seed) {
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        int32_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int32_t w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = w;
    }
    f = 3; b = 0;
    for (int i = 0; i < 310; i++) next();
}
uint32_t LibcRand::next() {
    uint32_t v = (uint32_t)r[f] + (uint32_t)r[b];
    r[f] = (int32_t)v;
    if (++f == 31) f = 0;
    if (++b == 31) b = 0;
    return v >> 1;
}
long LibcRand::integer(long a, long b2) { return (long)(a + (b2 - a) * (next() / 2147483648.0)); }

// ------------------------------------------------------------------------------------------ variant files
namespace {

struct Cnv { long spos, epos; float cn, mcn; };
struct PointVar { long pos; uint8_t ch; bool het; };            // SNP (het unused) or SNV
struct InsVar { long pos; uint64_t lit_off; uint32_t len; bool het; };
struct DelVar { long pos; int len; bool het; };

// variants of one chromosome in file order + an index sorted by position for range queries
template <class T> struct VarList {
    std::vector<T> v; std::vector<uint32_t> ord; bool indexed = false;
    void index() {
        ord.resize(v.size()); for (uint32_t i = 0; i < ord.size(); i++) ord[i] = i;
        std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return v[a].pos < v[b].pos; });
        indexed = true;
    }
    // file-order indices of the variants with s <= pos <= e
    void in_range(long s, long e, std::vector<uint32_t>& out) {
        if (!indexed) index();
        out.clear();
        auto lo = std::lower_bound(ord.begin(), ord.end(), s, [&](uint32_t a, long x) { return v[a].pos < x; });
        for (; lo != ord.end() && v[*lo].pos <= e; ++lo) out.push_back(*lo);
        std::sort(out.begin(), out.end());
    }
};
struct ChromVars { std::vector<Cnv> cnv; VarList<PointVar> snp, snv; VarList<InsVar> ins; VarList<DelVar> del; };

// by_chr[strip_chr_prefix(raw)] with a one-entry cache: variant files list one chromosome after the other
struct ChromLookup {
    std::map<std::string, ChromVars>& m; std::string last_raw; ChromVars* last = nullptr;
    explicit ChromLookup(std::map<std::string, ChromVars>& mm) : m(mm) {}
    ChromVars& operator()(const char* raw, size_t n) {
        if (last && last_raw.size() == n && memcmp(last_raw.data(), raw, n) == 0) return *last;
        last_raw.assign(raw, n); last = &m[strip_chr_prefix(last_raw)];
        return *last;
    }
    ChromVars& operator()(const std::string& raw) { return (*this)(raw.data(), raw.size()); }
};

// std::getline-on-stringstream field splitting (lib/split/split.cpp:3-15): no empty field after a trailing delimiter
std::vector<std::string> split_fields(const std::string& s, char d) {
    std::vector<std::string> out; size_t p = 0;
    while (p < s.size()) {
        size_t q = s.find(d, p);
        if (q == std::string::npos) { out.push_back(s.substr(p)); break; }
        out.push_back(s.substr(p, q - p)); p = q + 1;
    }
    return out;
}

bool parse_zygosity(const std::string& t, bool& het) { if (t == "het") { het = true; return true; } if (t == "homo") { het = false; return true; } return false; }

// Genome::loadAbers, Genome.cpp:35-165. The file is cut into chunks of whole lines that are parsed on threads; the records are
// merged in file order, and the first bad line in file order is the one reported (the reference stops there).
struct VarRec { char kind; std::string chrom; long a = 0, b = 0; float cn = 0, mcn = 0; uint8_t ch = 0; bool het = false; std::string seq; int dlen = 0; };
enum VarErr { VE_NONE = 0, VE_FIELDS, VE_CN, VE_SAME, VE_SNV_TYPE, VE_INS_TYPE, VE_DEL_TYPE, VE_KIND };

// one line; returns VE_NONE and fills r (r.kind = 0 for a blank / comment line)
VarErr parse_var_line(const std::string& line, VarRec& r) {
    r.kind = 0;
    if (line.empty() || line[0] == '#') return VE_NONE;
    std::vector<std::string> f = split_fields(line, '\t');
    const std::string kind = f.empty() ? std::string() : f[0];
    if (kind == "c") {
        if (f.size() != 6) return VE_FIELDS;
        float cn = (float)atof(f[4].c_str()), mcn = (float)atof(f[5].c_str());
        if (cn < mcn) return VE_CN;
        if (cn - mcn > mcn) mcn = cn - mcn;
        r.kind = 'c'; r.chrom = f[1]; r.a = atol(f[2].c_str()); r.b = atol(f[3].c_str()); r.cn = cn; r.mcn = mcn;
    } else if (kind == "s") {
        if (f.size() != 6) return VE_FIELDS;
        if (f[3].empty() || f[4].empty()) return VE_FIELDS;   // reference: std::out_of_range from at(0)
        if (f[3][0] == f[4][0]) return VE_SAME;
        if (!parse_zygosity(f[5], r.het)) return VE_SNV_TYPE;
        r.kind = 's'; r.chrom = f[1]; r.a = atol(f[2].c_str()); r.ch = (uint8_t)toupper((unsigned char)f[4][0]);
    } else if (kind == "i") {
        if (f.size() != 5) return VE_FIELDS;
        if (!parse_zygosity(f[4], r.het)) return VE_INS_TYPE;
        r.kind = 'i'; r.chrom = f[1]; r.a = atol(f[2].c_str()); r.seq = f[3];
        for (char& ch : r.seq) ch = (char)toupper((unsigned char)ch);   // the segment is upper-cased as a whole at the end, Genome.cpp:683-687
    } else if (kind == "d") {
        if (f.size() != 5) return VE_FIELDS;
        if (!parse_zygosity(f[4], r.het)) return VE_DEL_TYPE;
        r.kind = 'd'; r.chrom = f[1]; r.a = atol(f[2].c_str()); r.dlen = atoi(f[3].c_str());
    } else return VE_KIND;
    return VE_NONE;
}

bool load_variations(Plan& P, const char* path, std::map<std::string, ChromVars>& by_chr) {
    if (!path || !*path) return true;
    FILE* fp = fopen(path, "rb");
    if (!fp) { P.err = std::string("can not open file ") + path; return false; }
    fseek(fp, 0, SEEK_END); const long long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    std::string buf((size_t)std::max<long long>(sz, 0), '\0');
    const size_t got = sz > 0 ? fread(&buf[0], 1, (size_t)sz, fp) : 0;
    fclose(fp);
    buf.resize(got);
    unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("SCS_HOST_THREADS")) hw = (unsigned)std::max(1, atoi(e));
    const size_t nchunk = got >= (2u << 20) ? hw : 1;
    std::vector<size_t> cut(nchunk + 1, got); cut[0] = 0;
    for (size_t k = 1; k < nchunk; k++) {
        size_t p = std::max(cut[k - 1], got * k / nchunk);
        const size_t nl = p < got ? buf.find('\n', p) : std::string::npos;
        cut[k] = nl == std::string::npos ? got : nl + 1;
    }
    struct ChunkOut { std::vector<VarRec> recs; long n_lines = 0; VarErr err = VE_NONE; long err_line = 0; std::string err_text; };
    std::vector<ChunkOut> outs(nchunk);
    auto work = [&](size_t k) {
        ChunkOut& o = outs[k];
        size_t s = cut[k]; const size_t end = cut[k + 1];
        std::string line; VarRec r;
        while (s < end) {
            size_t e = buf.find('\n', s);
            if (e == std::string::npos || e > end) e = end;
            line.assign(buf, s, e - s);
            o.n_lines++;
            const VarErr ve = parse_var_line(line, r);
            if (ve != VE_NONE) { o.err = ve; o.err_line = o.n_lines; o.err_text = line; return; }   // later lines of this chunk do not matter
            if (r.kind) o.recs.push_back(r);
            s = e + 1;
        }
    };
    if (nchunk == 1) work(0);
    else {
        std::vector<std::thread> ts;
        for (size_t k = 0; k < nchunk; k++) ts.emplace_back(work, k);
        for (auto& t : ts) t.join();
    }
    ChromLookup chrom(by_chr);
    long line0 = 0;
    for (ChunkOut& o : outs) {
        for (VarRec& r : o.recs) {
            ChromVars& cv = chrom(r.chrom);
            switch (r.kind) {
                case 'c': cv.cnv.push_back({r.a, r.b, r.cn, r.mcn}); P.n_cnv++; break;
                case 's': cv.snv.v.push_back({r.a, r.ch, r.het}); P.n_snv++; break;
                case 'i': cv.ins.v.push_back({r.a, (uint64_t)P.literals.size(), (uint32_t)r.seq.size(), r.het}); P.literals += r.seq; P.n_ins++; break;
                default: cv.del.v.push_back({r.a, r.dlen, r.het}); P.n_del++; break;
            }
        }
        if (o.err != VE_NONE) {
            const std::string ln = std::to_string(line0 + o.err_line), in_file = std::string(" in file ") + path;
            std::string head;
            switch (o.err) {
                case VE_FIELDS: head = "ERROR: line " + ln + " has wrong number of fields" + in_file; break;
                case VE_CN: head = "ERROR: total copy number should be not lower than major copy number at line " + ln + in_file; break;
                case VE_SAME: head = "ERROR: the mutated allele should be not same as the reference allele at line " + ln + in_file; break;
                case VE_SNV_TYPE: head = "ERROR: unrecognized SNV type at line " + ln + in_file; break;
                case VE_INS_TYPE: head = "ERROR: unrecognized insert type at line " + ln + in_file; break;
                case VE_DEL_TYPE: head = "ERROR: unrecognized deletion type at line " + ln + in_file; break;
                default: head = "ERROR: unrecognized aberraton type at line " + ln + in_file; break;
            }
            P.err = head + "\n" + o.err_text;
            return false;
        }
        line0 += o.n_lines;
    }
    return true;
}

uint8_t snp_complement(uint8_t c) {   // SNP::getComplement, snp.cpp:96-110
    switch (c) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c';
        default: return 'N';
    }
}

// SNPOnChr::readSNPs + SNP::SNP, snp.cpp:13-36,147-203: id, chromosome, position, observed "X/Y", strand, reference base.
// One line [s, e) of the file (e points at the '\n' or at the end of the file); tabs are overwritten with NULs in place.
// Returns false for a line the reference warns about and skips.
struct SnpRec { const char* chrom; size_t chrom_len; PointVar v; };
bool parse_snp_line(char* s, char* e, SnpRec& out) {
    char* col[8]; int nc = 0; col[nc++] = s;
    bool too_many = false;
    for (char* p = s; p < e; p++) if (*p == '\t') { *p = 0; if (nc < 8) col[nc++] = p + 1; else too_many = true; }
    if (nc != 6 || too_many) return false;
    const char* slash = (const char*)memchr(col[3], '/', strlen(col[3]));
    if (!slash || slash == col[3] || slash[1] == 0) return false;
    const uint8_t first = (uint8_t)col[3][0], second = (uint8_t)slash[1];
    const uint8_t strand = (uint8_t)*col[4];
    uint8_t ref = (uint8_t)*col[5];   // an empty last field reads the line terminator, as the reference's fgets buffer does
    if (strand == '-') ref = snp_complement(ref);
    uint8_t nuc = (first == ref) ? second : first;
    if (strand == '-') nuc = snp_complement(nuc);
    char save = *e; *e = 0; const long pos = (long)atoll(col[2]); *e = save;
    out.chrom = col[1]; out.chrom_len = strlen(col[1]);
    out.v = {pos, (uint8_t)toupper(nuc), false};
    return true;
}

bool load_snps(Plan& P, const char* path, std::map<std::string, ChromVars>& by_chr) {
    if (!path || !*path) return true;
    FILE* fp = fopen(path, "rb");
    if (!fp) { P.err = std::string("can not open SNP file ") + path; return false; }
    fseek(fp, 0, SEEK_END); const long long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    std::vector<char> buf((size_t)std::max<long long>(sz, 0) + 2, 0);
    const size_t got = sz > 0 ? fread(buf.data(), 1, (size_t)sz, fp) : 0;
    fclose(fp);
    buf[got] = 0;
    // chunks of whole lines, one per thread; every chunk keeps its records per chromosome in file order
    unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("SCS_HOST_THREADS")) hw = (unsigned)std::max(1, atoi(e));
    const size_t nchunk = got >= (4u << 20) ? hw : 1;
    std::vector<size_t> cut(nchunk + 1, got); cut[0] = 0;
    for (size_t k = 1; k < nchunk; k++) {
        size_t p = std::max(cut[k - 1], got * k / nchunk);
        const char* nl = p < got ? (const char*)memchr(buf.data() + p, '\n', got - p) : nullptr;
        cut[k] = nl ? (size_t)(nl - buf.data()) + 1 : got;
    }
    struct ChunkOut { std::vector<std::pair<std::string, std::vector<PointVar>>> by; std::vector<long> bad_lines; long n_lines = 0; bool long_line = false; };
    std::vector<ChunkOut> outs(nchunk);
    auto work = [&](size_t k) {
        ChunkOut& o = outs[k];
        char* s = buf.data() + cut[k]; char* const end = buf.data() + cut[k + 1];
        size_t last = (size_t)-1;
        while (s < end) {
            char* e = (char*)memchr(s, '\n', (size_t)(end - s));
            if (!e) e = end;
            o.n_lines++;
            if (e - s >= 999) o.long_line = true;   // the reference's fgets(buf, 1000) would split this line
            SnpRec r;
            if (!parse_snp_line(s, e, r)) o.bad_lines.push_back(o.n_lines);
            else {
                if (last == (size_t)-1 || o.by[last].first.size() != r.chrom_len || memcmp(o.by[last].first.data(), r.chrom, r.chrom_len) != 0) {
                    last = (size_t)-1;
                    for (size_t i = 0; i < o.by.size(); i++) if (o.by[i].first.size() == r.chrom_len && memcmp(o.by[i].first.data(), r.chrom, r.chrom_len) == 0) { last = i; break; }
                    if (last == (size_t)-1) { o.by.emplace_back(std::string(r.chrom, r.chrom_len), std::vector<PointVar>()); last = o.by.size() - 1; }
                }
                o.by[last].second.push_back(r.v);
            }
            s = e + 1;
        }
    };
    if (nchunk == 1) work(0);
    else {
        std::vector<std::thread> ts;
        for (size_t k = 0; k < nchunk; k++) ts.emplace_back(work, k);
        for (auto& t : ts) t.join();
    }
    long line0 = 0;
    for (ChunkOut& o : outs) {
        if (o.long_line) { P.err = std::string("ERROR: line longer than 999 characters in SNP file ") + path; return false; }
        for (long b : o.bad_lines) if (P.warnings.size() < 4096) P.warnings += std::string("Warning: malformed snp file ") + path + ", there should be 6 fields @line " + std::to_string(line0 + b) + "\n";
        for (auto& kv : o.by) {
            std::vector<PointVar>& dst = by_chr[strip_chr_prefix(kv.first)].snp.v;
            dst.insert(dst.end(), kv.second.begin(), kv.second.end());
            P.n_snp += (long)kv.second.size();
        }
        line0 += o.n_lines;
    }
    return true;
}

// ------------------------------------------------------------------------------------------ piece table
// A std::string stand-in that stores runs instead of characters. insert/erase follow std::string: position > size
// fails (the reference dies with std::out_of_range there), erase clamps its count to the end of the string.
struct Run { uint64_t src; uint64_t q; uint32_t len; };   // q: position in the segment's pre-indel string (reference runs only)

class Rope {
    static constexpr size_t kChunk = 64;
    std::vector<std::vector<Run>> ch; std::vector<uint64_t> clen; uint64_t total = 0;

    // chunk holding position pos (pos == total -> last chunk), *base = first position of that chunk
    size_t chunk_of(uint64_t pos, uint64_t* base) const {
        uint64_t acc = 0; size_t c = 0;
        for (; c + 1 < ch.size(); c++) { if (pos < acc + clen[c]) break; acc += clen[c]; }
        *base = acc; return c;
    }
    // make sure a run starts at pos; returns (chunk, index) of it (index == size of the last chunk when pos == total)
    void boundary(uint64_t pos, size_t& ci, size_t& pi) {
        uint64_t base; ci = chunk_of(pos, &base);
        std::vector<Run>& v = ch[ci];
        uint64_t at = base; pi = 0;
        while (pi < v.size() && at + v[pi].len <= pos) { at += v[pi].len; pi++; }
        if (pi == v.size() || at == pos) {
            if (pi == v.size() && ci + 1 < ch.size()) { ci++; pi = 0; }   // boundary at the start of the next chunk
            return;
        }
        const uint32_t k = (uint32_t)(pos - at);
        Run tail = v[pi]; tail.src += k; tail.q += k; tail.len -= k;
        v[pi].len = k;
        v.insert(v.begin() + pi + 1, tail);
        pi++;
    }
    void rebalance(size_t ci) {
        if (ch[ci].size() <= kChunk) return;
        const size_t half = ch[ci].size() / 2;
        std::vector<Run> hi(ch[ci].begin() + half, ch[ci].end());
        ch[ci].resize(half);
        uint64_t l = 0; for (const Run& r : hi) l += r.len;
        clen[ci] -= l;
        ch.insert(ch.begin() + ci + 1, std::move(hi)); clen.insert(clen.begin() + ci + 1, l);
    }

  public:
    uint64_t size() const { return total; }
    void append(const Run& r) {
        if (r.len == 0) return;
        if (ch.empty() || ch.back().size() >= kChunk) { ch.emplace_back(); clen.push_back(0); }
        ch.back().push_back(r); clen.back() += r.len; total += r.len;
    }
    bool insert(uint64_t pos, const Run& r) {
        if (pos > total) return false;
        if (r.len == 0) return true;
        if (ch.empty()) { append(r); return true; }
        size_t ci, pi; boundary(pos, ci, pi);
        ch[ci].insert(ch[ci].begin() + pi, r); clen[ci] += r.len; total += r.len;
        rebalance(ci);
        return true;
    }
    bool erase(uint64_t pos, uint64_t n) {
        if (pos > total) return false;
        n = std::min(n, total - pos);
        if (n == 0) return true;
        size_t c1, p1, c2, p2;
        boundary(pos, c1, p1); boundary(pos + n, c2, p2);   // the second split lies behind the first: (c1, p1) stays valid
        if (c1 == c2) ch[c1].erase(ch[c1].begin() + p1, ch[c1].begin() + p2);
        else {
            ch[c1].resize(p1);
            ch[c2].erase(ch[c2].begin(), ch[c2].begin() + p2);
            ch.erase(ch.begin() + c1 + 1, ch.begin() + c2); clen.erase(clen.begin() + c1 + 1, clen.begin() + c2);
        }
        total -= n;
        for (size_t c = c1; c < ch.size() && c <= c1 + 1; c++) { uint64_t l = 0; for (const Run& r : ch[c]) l += r.len; clen[c] = l; }
        for (size_t c = std::min(c1 + 1, ch.size() - 1) + 1; c-- > c1;) if (ch[c].empty() && ch.size() > 1) { ch.erase(ch.begin() + c); clen.erase(clen.begin() + c); }
        return true;
    }
    template <class F> void for_each(F fn) const { for (const auto& v : ch) for (const Run& r : v) fn(r); }
};

// Fenwick tree over the sorted distinct in-segment positions of the indels: "sum of lengths recorded at positions <= x".
// std::map::insert keeps the FIRST length recorded at a position (Genome.cpp:573,600,644,675), hence `seen`.
struct OffsetIndex {
    const std::vector<int>* keys = nullptr; std::vector<long> bit; std::vector<char> seen;
    void init(const std::vector<int>* k) { keys = k; bit.assign(k->size() + 1, 0); seen.assign(k->size(), 0); }
    void record(int pos, int len) {
        size_t i = (size_t)(std::lower_bound(keys->begin(), keys->end(), pos) - keys->begin());
        if (seen[i]) return;
        seen[i] = 1;
        for (size_t x = i + 1; x < bit.size(); x += x & (~x + 1)) bit[x] += len;
    }
    long upto(int pos) const {   // sum over recorded positions <= pos
        size_t i = (size_t)(std::upper_bound(keys->begin(), keys->end(), pos) - keys->begin());
        long s = 0; for (size_t x = i; x > 0; x -= x & (~x + 1)) s += bit[x];
        return s;
    }
};

struct HapBuild { std::vector<Piece> pieces; std::vector<Sub> subs; uint64_t len = 0; };

void hap_append(HapBuild& h, uint64_t src, uint32_t len) {
    if (len == 0) return;
    if (!h.pieces.empty()) {
        Piece& p = h.pieces.back();
        if (p.src + p.len == src && ((p.src ^ src) & kLiteral) == 0 && (uint64_t)p.len + len < (1ull << 31)) { p.len += len; h.len += len; return; }
    }
    h.pieces.push_back({h.len, src, len}); h.len += len;
}

// what the random draws decide for one segment: how many copies each haplotype carries and which haplotypes are "major"
struct SegSpec { long s, e; std::vector<int> copies; std::vector<char> in_major; };

// The draws of Genome::generateSegment (Genome.cpp:404-467) for the 1-based inclusive range [s, e]. Sequential: libc rand()
// is one stream over all segments of all chromosomes. Returns false with err set; *skip = true for CN == 0 (nothing emitted).
bool assign_copies(LibcRand& rng, int ploidy, const std::string& chr, long chrLen, long s, long e, int CN, int mCN, SegSpec& out, bool* skip, std::string& err) {
    *skip = false;
    if (CN == 0) { *skip = true; return true; }
    if (s - 1 < 0 || e - s + 1 < 1) { err = "Error: cannot construct subsequence with negative offset or length < 1"; return false; }
    if (e > chrLen) { err = "ERROR: segment " + std::to_string(s) + "-" + std::to_string(e) + " lies past the end of chromosome " + chr; return false; }
    int i, j, k, n;
    std::vector<int> major, reps;
    auto has = [](const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); };
    std::vector<int> copies(ploidy, 0);   // how many copies of the segment each haplotype carries
    if (CN < ploidy) {   // :411-425: CN distinct haplotypes keep one copy, the first mCN drawn are the major ones
        for (i = 0; i < CN; i++) for (;;) { j = (int)rng.integer(0, ploidy); if (!has(reps, j)) { reps.push_back(j); break; } }
        for (i = 0; i < mCN; i++) major.push_back(reps[i]);
        for (int h : reps) copies[h] = 1;
    } else {             // :426-467
        reps.assign(ploidy, 1);
        n = CN - ploidy;
        k = (int)rng.integer(0, ploidy);
        for (i = n; i >= 0; i--) {
            if (reps[k] + i == mCN) { reps[k] += i; major.push_back(k); break; }
            else if (reps[k] + i == CN - mCN) { reps[k] += i; for (j = 0; j < ploidy; j++) if (j != k) major.push_back(j); break; }
        }
        if (i >= 0) {
            n -= i;
            if (n > 0 && ploidy < 2) { err = "ERROR: copy number cannot be distributed over one haplotype"; return false; }   // reference: endless loop
            while (n > 0) { j = (int)rng.integer(0, ploidy); if (j != k) { reps[j]++; n--; } }
        } else {
            while (n > 0) { j = (int)rng.integer(0, ploidy); reps[j]++; n--; }
            for (i = 0; i < ploidy; i++) major.push_back(i);
        }
        copies = reps;
    }
    out.s = s; out.e = e; out.copies = copies;
    out.in_major.assign(ploidy, 0); for (int h : major) if (h >= 0 && h < ploidy) out.in_major[h] = 1;
    return true;
}

// The deterministic rest of generateSegment (Genome.cpp:469-691) for one chromosome; one instance per worker thread.
struct Planner {
    int ploidy; std::string err;
    std::vector<uint32_t> hits;
    explicit Planner(int pl) : ploidy(pl) {}

    bool segment(std::vector<HapBuild>& hap, ChromVars* cv, const std::string& chr, const SegSpec& spec) {
        const long s = spec.s, e = spec.e;
        const unsigned int refSize = (unsigned int)(e - s + 1);
        const std::vector<int>& copies = spec.copies; const std::vector<char>& in_major = spec.in_major;
        int j, k, n;
        // a het variant goes to the major haplotypes when k == 0 and to the others when k == 1 (:496-499 and alike)
        auto skipped = [&](int kk, int jj) { return (kk == 0 && !in_major[jj]) || (kk == 1 && in_major[jj]); };

        std::vector<Rope> rope(ploidy);
        for (j = 0; j < ploidy; j++) for (int t = 0; t < copies[j]; t++) rope[j].append({(uint64_t)(s - 1), (uint64_t)t * refSize, refSize});
        std::vector<std::vector<std::pair<uint32_t, uint8_t>>> point(ploidy);   // (position in the segment, base), in application order

        if (cv) {
            k = 0;   // SNPs :489-508
            cv->snp.in_range(s, e, hits);
            for (uint32_t id : hits) {
                const PointVar& v = cv->snp.v[id];
                for (j = 0; j < ploidy; j++) if (!skipped(k, j) && copies[j]) point[j].push_back({(uint32_t)(v.pos - s), v.ch});
                k = (k + 1) % 2;
            }
            k = 0;   // SNVs :510-544
            cv->snv.in_range(s, e, hits);
            for (uint32_t id : hits) {
                const PointVar& v = cv->snv.v[id];
                for (j = 0; j < ploidy; j++) if (!(v.het && skipped(k, j)) && copies[j]) point[j].push_back({(uint32_t)(v.pos - s), v.ch});
                if (v.het) k = (k + 1) % 2;
            }
            // indels :546-679
            std::vector<uint32_t> ins_hits, del_hits;
            cv->ins.in_range(s, e, ins_hits); cv->del.in_range(s, e, del_hits);
            std::vector<int> ins_keys, del_keys;
            for (uint32_t id : ins_hits) ins_keys.push_back((int)(cv->ins.v[id].pos - s));
            for (uint32_t id : del_hits) del_keys.push_back((int)(cv->del.v[id].pos - s));
            std::sort(ins_keys.begin(), ins_keys.end()); ins_keys.erase(std::unique(ins_keys.begin(), ins_keys.end()), ins_keys.end());
            std::sort(del_keys.begin(), del_keys.end()); del_keys.erase(std::unique(del_keys.begin(), del_keys.end()), del_keys.end());
            std::vector<OffsetIndex> insAt(ploidy), delAt(ploidy);
            for (j = 0; j < ploidy; j++) { insAt[j].init(&ins_keys); delAt[j].init(&del_keys); }
            std::vector<int> insLen(ploidy, 0), delLen(ploidy, 0);
            k = 0;
            for (uint32_t id : ins_hits) {
                const InsVar& v = cv->ins.v[id];
                const int sindx = (int)(v.pos - s);
                for (j = 0; j < ploidy; j++) {
                    if (v.het && skipped(k, j)) continue;
                    const int offset = (int)insAt[j].upto(sindx);
                    Rope& q = rope[j];
                    if (refSize + insLen[j] == 0) { err = "ERROR: insertion at " + std::to_string(v.pos) + " on chromosome " + chr + ": empty segment"; return false; }
                    n = (int)(q.size() / (refSize + insLen[j]));
                    const int len = (int)v.len;
                    for (int t = 0; t < n; t++) {
                        const unsigned int at = sindx + offset + t * (refSize + insLen[j] + len);   // the reference's 32-bit arithmetic, :567-569
                        if (!q.insert(at, {kLiteral | v.lit_off, 0, v.len})) {
                            err = "ERROR: insertion at " + std::to_string(v.pos) + " on chromosome " + chr + " falls outside its haplotype (the reference aborts with std::out_of_range here)";
                            return false;
                        }
                    }
                    insLen[j] += len;
                    insAt[j].record(sindx, len);
                }
                if (v.het) k = (k + 1) % 2;
            }
            // k is not reset between the two loops in the reference (:608 has no `k = 0`)
            for (uint32_t id : del_hits) {
                const DelVar& v = cv->del.v[id];
                const int sindx = (int)(v.pos - s);
                const int dl = v.len;
                for (j = 0; j < ploidy; j++) {
                    if (v.het && skipped(k, j)) continue;
                    const int offset = (int)(insAt[j].upto(sindx) - delAt[j].upto(sindx));
                    if (sindx + offset < 0) continue;
                    Rope& q = rope[j];
                    if (refSize + insLen[j] - delLen[j] == 0) { err = "ERROR: deletion at " + std::to_string(v.pos) + " on chromosome " + chr + ": empty segment"; return false; }
                    n = (int)(q.size() / (refSize + insLen[j] - delLen[j]));
                    for (int t = 0; t < n; t++) {
                        const unsigned int at = sindx + offset + t * (refSize + insLen[j] - delLen[j] - dl);   // :637-639
                        // std::string::erase(pos, n) takes n as size_t: a negative length becomes "to the end"
                        if (!q.erase(at, (uint64_t)(size_t)(long)dl)) {
                            err = "ERROR: deletion at " + std::to_string(v.pos) + " on chromosome " + chr + " falls outside its haplotype (the reference aborts with std::out_of_range here)";
                            return false;
                        }
                    }
                    delLen[j] += dl;
                    delAt[j].record(sindx, dl);
                }
                if (v.het) k = (k + 1) % 2;
            }
        }
        // flatten: runs -> pieces of the haplotype; point substitutions -> output positions (a base that was deleted has none)
        for (j = 0; j < ploidy; j++) {
            HapBuild& H = hap[j];
            const uint64_t base = H.len;
            std::vector<std::pair<uint64_t, uint64_t>> qmap;   // (q, offset in this segment's output) of reference runs, q ascending
            std::vector<uint32_t> qlen;
            uint64_t at = 0;
            rope[j].for_each([&](const Run& r) {
                if (!(r.src & kLiteral)) { qmap.push_back({r.q, at}); qlen.push_back(r.len); }
                hap_append(H, r.src, r.len); at += r.len;
            });
            std::vector<std::pair<uint32_t, uint8_t>>& pv = point[j];
            if (pv.empty() || qmap.empty()) continue;
            std::stable_sort(pv.begin(), pv.end(), [](const std::pair<uint32_t, uint8_t>& a, const std::pair<uint32_t, uint8_t>& b) { return a.first < b.first; });
            for (size_t x = 0; x < pv.size(); x++) {
                if (x + 1 < pv.size() && pv[x + 1].first == pv[x].first) continue;   // a later write to the same base wins
                for (int t = 0; t < copies[j]; t++) {
                    const uint64_t q = (uint64_t)pv[x].first + (uint64_t)t * refSize;
                    size_t r = (size_t)(std::upper_bound(qmap.begin(), qmap.end(), q, [](uint64_t x, const std::pair<uint64_t, uint64_t>& e) { return x < e.first; }) - qmap.begin());
                    if (r == 0) continue;
                    r--;
                    if (q >= qmap[r].first + qlen[r]) continue;
                    H.subs.push_back({base + qmap[r].second + (q - qmap[r].first), pv[x].second});
                }
            }
        }
        return true;
    }
};

}  // namespace

bool build_plan(Plan& P, const std::vector<ChromIn>& chroms, const char* snp_file, const char* var_file, int ploidy, uint32_t libc_seed) {
    P = Plan(); P.ploidy = ploidy; P.chroms = chroms;
    if (ploidy < 1 || ploidy > 64) { P.err = "ERROR: ploidy out of range"; return false; }
    std::map<std::string, ChromVars> by_chr;
    using clk = std::chrono::steady_clock;
    auto ms_since = [](clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); };
    auto t0 = clk::now();
    if (!load_variations(P, var_file, by_chr)) return false;   // Genome::loadData order: loadAbers, loadSNPs, loadRefSeq (Genome.cpp:18-25)
    P.ms_parse_var = ms_since(t0); t0 = clk::now();
    if (!load_snps(P, snp_file, by_chr)) return false;
    P.ms_parse_snp = ms_since(t0); t0 = clk::now();
    if (chroms.empty()) { P.err = "ERROR: reference sequence cannot be empty!"; return false; }
    const int mCN = (int)ceilf((float)ploidy / 2);
    // Genome::saveSequence, Genome.cpp:329-386. Pass 1, sequential: the segment list of every chromosome with its random draws.
    LibcRand rng(libc_seed);
    const size_t nc = chroms.size();
    std::vector<std::vector<SegSpec>> specs(nc);
    std::vector<ChromVars*> cvs(nc, nullptr);
    size_t draw_err_chrom = nc; std::string draw_err;
    for (size_t c = 0; c < nc && draw_err_chrom == nc; c++) {
        const std::string& chr = chroms[c].name;
        const long chrLen = (long)chroms[c].len;
        auto it = by_chr.find(chr);
        ChromVars* cv = cvs[c] = (it == by_chr.end() ? nullptr : &it->second);
        auto add = [&](long s, long e, int CN, int mcn) {
            SegSpec sp; bool skip = false;
            if (!assign_copies(rng, ploidy, chr, chrLen, s, e, CN, mcn, sp, &skip, draw_err)) { draw_err_chrom = c; return false; }
            if (!skip) { specs[c].push_back(std::move(sp)); P.n_segments++; }
            return true;
        };
        long segStart = 1; bool ok = true;
        if (cv) for (Cnv cnv : cv->cnv) {
            if (segStart > chrLen) break;
            cnv.epos = std::min(cnv.epos, chrLen);
            if (segStart < cnv.spos && !(ok = add(segStart, cnv.spos - 1, ploidy, mCN))) break;
            if (!(ok = add(cnv.spos, cnv.epos, (int)cnv.cn, (int)cnv.mcn))) break;
            segStart = cnv.epos + 1;
        }
        if (ok && segStart <= chrLen) add(segStart, chrLen, ploidy, mCN);
    }
    const double ms_pass1 = ms_since(t0);
    // Pass 2, one task per chromosome on a few threads: variants -> piece tables -> runs and substitutions. A chromosome whose
    // draws failed in pass 1 still gets its earlier segments applied: the reference would have died at the first problem in
    // file order, and that may be an indel of an earlier segment.
    const size_t n_apply = std::min(nc, draw_err_chrom + 1);
    std::vector<std::vector<HapBuild>> built(n_apply);
    std::vector<std::string> apply_err(n_apply);
    auto apply_chrom = [&](size_t c) {
        Planner pl(ploidy);
        built[c].assign(ploidy, HapBuild());
        for (const SegSpec& sp : specs[c]) if (!pl.segment(built[c], cvs[c], chroms[c].name, sp)) { apply_err[c] = pl.err; return; }
    };
    {
        unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        if (const char* e = getenv("SCS_HOST_THREADS")) hw = (unsigned)std::max(1, atoi(e));
        size_t nt = std::min<size_t>(hw, n_apply);
        // two records with the same (stripped) name share their variant lists, whose position index is built lazily: serial then
        { std::vector<ChromVars*> u; for (ChromVars* v : cvs) if (v) u.push_back(v); std::sort(u.begin(), u.end()); if (std::adjacent_find(u.begin(), u.end()) != u.end()) nt = 1; }
        if (nt <= 1) for (size_t c = 0; c < n_apply; c++) apply_chrom(c);
        else {
            std::atomic<size_t> next{0};
            std::vector<std::thread> ts;
            for (size_t t = 0; t < nt; t++) ts.emplace_back([&] { for (size_t c; (c = next.fetch_add(1)) < n_apply;) apply_chrom(c); });
            for (auto& t : ts) t.join();
        }
    }
    const double ms_pass2 = ms_since(t0);
    for (size_t c = 0; c < n_apply; c++) if (!apply_err[c].empty()) { P.err = apply_err[c]; return false; }
    if (draw_err_chrom < nc) { P.err = draw_err; return false; }
    for (size_t c = 0; c < nc; c++) {
        const std::string& chr = chroms[c].name;
        const long chrLen = (long)chroms[c].len;
        std::vector<HapBuild>& hap = built[c];
        for (int j = 0; j < ploidy; j++) {
            if (hap[j].len >= (1ull << 32)) { P.err = "ERROR: haplotype longer than 2^32-1 bases (Genome.cpp:370 uses unsigned int)"; return false; }
            Hap h; h.chrom = (uint32_t)c; h.hap = (uint32_t)j; h.len = hap[j].len;
            h.piece_lo = P.pieces.size(); P.pieces.insert(P.pieces.end(), hap[j].pieces.begin(), hap[j].pieces.end()); h.piece_hi = P.pieces.size();
            h.sub_lo = P.subs.size(); P.subs.insert(P.subs.end(), hap[j].subs.begin(), hap[j].subs.end()); h.sub_hi = P.subs.size();
            h.name = chr + "_" + std::to_string(j + 1) + "_" + std::to_string(chrLen);
            P.haps.push_back(h);
        }
        std::vector<HapBuild>().swap(hap);
    }
    P.ms_segments = ms_since(t0);
    if (getenv("SCS_TRACE")) fprintf(stderr, "[scs trace] simuvars plan: parse var %.1f ms, parse snp %.1f ms, segments %.1f ms (draws %.1f, apply %.1f, merge %.1f)\n", P.ms_parse_var, P.ms_parse_snp, P.ms_segments, ms_pass1, ms_pass2 - ms_pass1, P.ms_segments - ms_pass2);
    return true;
}

}  // namespace sv
}  // namespace scs
