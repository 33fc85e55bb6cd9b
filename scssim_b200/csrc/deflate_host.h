// Host side of the optional block-gzip output (SURVEY.md §8f row N3 "optional block-gzip"; the reference writes plain text only,
// /root/reference/lib/seqwriter/SeqWriter.cpp:41-54): the Huffman code every BGZF block of a run is written with.
//
// FASTQ text is literal-only material for deflate — bases and qualities are i.i.d.-looking, LZ77 finds next to nothing — so the
// device encoder (deflate.cu) emits one dynamic-Huffman block of literals per 32 KiB of text. The code is fixed for the run and
// derived here from the profile alone (expected byte frequencies of a record: header digits, the four bases, the profile's
// marginal quality distribution), so the compressed bytes do not depend on slab size or GPU count.
#pragma once
#include <cstdint>
#include <vector>

namespace scs {

struct HostProfile;

struct DeflateCode {
    uint8_t len[257];            // code length of literal 0..255 and of end-of-block (256); 0 = byte cannot occur
    uint32_t code[257];          // bit-reversed canonical code (LSB first, ready to OR into the stream) | len << 24
    std::vector<uint32_t> prefix_words;   // BGZF member header (18 bytes, BSIZE left 0) + dynamic-block header, LSB-first bit string
    uint32_t prefix_bits = 0;
    double expected_bits_per_byte = 0;    // under the model histogram
};

// length-limited (<= 15) canonical Huffman code for `hist` (symbols with hist == 0 get no code); false if fewer than 2 symbols
bool build_deflate_code(const uint64_t hist[257], DeflateCode& out);
// expected byte histogram of one FASTQ record of this profile (scaled to integers); every byte a record can contain is > 0
void fastq_model_histogram(const HostProfile& P, bool paired, uint64_t hist[257]);
// slicing-by-4 tables of the reflected CRC-32 (polynomial 0xEDB88320; [0..255] is the classic byte table) and x^(2^n) mod p for
// the block combine (zlib's crc32_combine)
// ([0..31]; [32 + j] = x^(1024 j) mod p, the shift over j thread tiles of 128 bytes)
void crc32_tables(uint32_t table[1024], uint32_t x2n[32 + 256]);

}  // namespace scs
