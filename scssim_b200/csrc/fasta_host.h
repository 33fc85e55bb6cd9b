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

// A FASTA file mapped read-only (no copy of its bytes is made on the host) and its index.
struct FastaFile {
    const char* data = nullptr; size_t size = 0;
    std::vector<FaiRec> fai;
    FastaFile() = default;
    FastaFile(const FastaFile&) = delete;
    FastaFile& operator=(const FastaFile&) = delete;
    ~FastaFile();
    // maps and indexes the file; false with *err set (messages of FastaReference::open / Genome::loadRefSeq)
    bool open(const char* path, std::string* err);
    // bases of record i gathered into one contiguous buffer (slow path for ragged records)
    void gather(size_t i, std::vector<char>& out) const;

  private:
    void* map_ = nullptr; size_t map_len_ = 0;
};
// Writes <path>.fai if it does not exist yet (side effect of FastaReference::open, Fasta.cpp:243-249).
void fasta_write_fai(const char* path, const std::vector<FaiRec>& fai);
// "chr"/"chrom" prefix stripping of sequence names (lib/fastahack/Fasta.cpp:57-68, MyDefine.cpp:310-323).
std::string strip_chr_prefix(const std::string& name);

}  // namespace scs
