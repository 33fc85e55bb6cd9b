// Host-side FASTA reader shared by the genreads genome loader and simuvars: one pass over the lines builds the
// .fai model of lib/fastahack/Fasta.cpp:103-191 (name, length, offset, bases per line, bytes per line).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace scs {

struct FaiRec {
    std::string header;     // full header line without '>'
    std::string name;       // first token of the header (the .fai key)
    uint64_t len = 0;       // bases
    uint64_t off = 0;       // byte offset of the first base
    uint32_t blen = 0, llen = 0;   // bases per line, bytes per line (incl. terminator)
    bool regular = true;    // uniform line geometry: newlines can be skipped by index arithmetic
    uint64_t short_lines = 0;
};

// Reads the whole file into `raw` (one extra '\n' appended) and indexes it. Returns false with *err set.
bool fasta_read_and_index(const char* path, std::vector<char>& raw, size_t& got, std::vector<FaiRec>& fai, std::string* err);
// Writes <path>.fai if it does not exist yet (side effect of FastaReference::open, Fasta.cpp:243-249).
void fasta_write_fai(const char* path, const std::vector<FaiRec>& fai);
// Bases of record i gathered into one contiguous buffer (slow path for ragged records).
void fasta_gather(const std::vector<char>& raw, size_t got, const std::vector<FaiRec>& fai, size_t i, std::vector<char>& out);
// "chr"/"chrom" prefix stripping of sequence names (lib/fastahack/Fasta.cpp:57-68, MyDefine.cpp:310-323).
std::string strip_chr_prefix(const std::string& name);

}  // namespace scs
