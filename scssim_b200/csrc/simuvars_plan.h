// simuvars, host side: variant files -> an EDIT PLAN per output haplotype.
//
// The reference (`scssim simuvars`: Genome::loadAbers / SNPOnChr::readSNPs / Genome::saveSequence /
// Genome::generateSegment, /root/reference/lib/genome/Genome.cpp:35-165,329-691, lib/snp/snp.cpp:13-36,147-203)
// builds every haplotype by string surgery: copy the segment CN times, overwrite SNP/SNV positions, then
// std::string::insert / erase per indel and per copy — O(length) each, O(indels^2) offset bookkeeping.
// Here the same sequence of operations is applied to a PIECE TABLE (a rope of (source, length) runs with
// std::string's insert/erase semantics), which costs O(log) per operation and touches no bases; the result is
// a list of copy runs and a list of point substitutions per haplotype that the GPU materialises in one
// gather pass (simuvars.cu). The plan is a pure function of the inputs and of libc's rand() stream, which
// is reproduced here (glibc TYPE_3 additive feedback generator; the reference never seeds it on this branch,
// src/scssim.cpp:33-38, so its output is deterministic and so is ours: byte-identical FASTA).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace scs {
namespace sv {

// glibc srandom_r/random_r, TYPE_3 (x^31 + x^3 + 1): what libc rand() returns on the systems the reference runs on.
struct LibcRand {
    int32_t r[34]; int f, b;
    explicit LibcRand(uint32_t seed = 1) { reseed(seed); }
    void reseed(uint32_t seed);
    uint32_t next();                                     // rand()
    long integer(long a, long b2);                       // randomInteger, lib/mydefine/MyDefine.cpp:290-292
};

constexpr uint64_t kLiteral = 1ull << 63;

struct Piece { uint64_t out; uint64_t src; uint32_t len; };   // out: offset in the haplotype; src: base offset in the chromosome, or kLiteral | offset in the literal pool
struct Sub { uint64_t out; uint8_t ch; };                      // haplotype[out] = ch (applied after the gather)
struct Hap {
    uint32_t chrom, hap; uint64_t len;
    uint64_t piece_lo, piece_hi, sub_lo, sub_hi;
    std::string name;                                          // <chr>_<hap+1>_<chromLen>, Genome.cpp:368
};
struct ChromIn { std::string name; uint64_t len; };            // name already stripped of chr/chrom

struct Plan {
    int ploidy = 2;
    std::vector<ChromIn> chroms;
    std::vector<Hap> haps;               // chromosome-major, ploidy per chromosome
    std::vector<Piece> pieces;
    std::vector<Sub> subs;
    std::string literals;                // inserted sequences, upper-cased
    long n_cnv = 0, n_snv = 0, n_ins = 0, n_del = 0, n_snp = 0, n_segments = 0;
    double ms_parse_var = 0, ms_parse_snp = 0, ms_segments = 0;
    std::string warnings;                // what the reference prints to stderr while loading
    std::string err;
};

// Returns false with plan.err set (messages as the reference prints them before exit()).
bool build_plan(Plan& plan, const std::vector<ChromIn>& chroms, const char* snp_file, const char* var_file, int ploidy, uint32_t libc_seed);

}  // namespace sv
}  // namespace scs
