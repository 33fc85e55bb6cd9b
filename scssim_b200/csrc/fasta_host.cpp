#include "fasta_host.h"

#include <cstdio>
#include <cstring>

namespace scs {

std::string strip_chr_prefix(const std::string& in) {
    std::string name = in;
    size_t i = name.find("chrom");
    if (i == std::string::npos) { i = name.find("chr"); if (i != std::string::npos) name = name.substr(i + 3); }
    else name = name.substr(i + 5);
    return name;
}

bool fasta_read_and_index(const char* path, std::vector<char>& raw, size_t& got, std::vector<FaiRec>& fai, std::string* err) {
    FILE* f = fopen(path, "rb");
    if (!f) { if (err) *err = std::string("could not open ") + path; return false; }
    fseek(f, 0, SEEK_END); long long sz = ftell(f); fseek(f, 0, SEEK_SET);
    raw.assign((size_t)(sz > 0 ? sz : 0) + 1, 0);
    got = sz > 0 ? fread(raw.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    raw[got] = '\n';
    fai.clear();
    size_t pos = 0;
    while (pos < got) {
        char* s = &raw[pos];
        char* e = (char*)memchr(s, '\n', got + 1 - pos);
        size_t len = (size_t)(e - s);
        size_t llen = len + 1;
        if (len > 0 && s[len - 1] == '\r') len--;
        if (len > 0 && s[0] == '>') {
            FaiRec r; r.header.assign(s + 1, len - 1);
            r.name = r.header.substr(0, r.header.find_first_of(" \t"));
            r.off = (uint64_t)(pos + llen);
            fai.push_back(r);
        } else if (!fai.empty()) {
            FaiRec& a = fai.back();
            if (len == 0) { if (a.len) a.short_lines++; }          // a blank line is fine only at the very end of a record
            else {
                if (a.blen == 0) { a.blen = (uint32_t)len; a.llen = (uint32_t)llen; }
                if (a.short_lines) a.regular = false;               // bases after a short or blank line
                if (len != a.blen || llen != a.llen) { if (len > a.blen) a.regular = false; a.short_lines++; }
                a.len += len;
            }
        }
        pos += llen;
    }
    if (fai.empty()) { if (err) *err = "ERROR: reference sequence cannot be empty!"; return false; }
    return true;
}

void fasta_write_fai(const char* path, const std::vector<FaiRec>& fai) {
    std::string faiPath = std::string(path) + ".fai";
    if (FILE* t = fopen(faiPath.c_str(), "rb")) { fclose(t); return; }
    if (FILE* o = fopen(faiPath.c_str(), "wb")) {
        for (const FaiRec& r : fai) fprintf(o, "%s\t%llu\t%llu\t%u\t%u\n", r.name.c_str(), (unsigned long long)r.len, (unsigned long long)r.off, r.blen, r.llen);
        fclose(o);
    }
}

void fasta_gather(const std::vector<char>& raw, size_t got, const std::vector<FaiRec>& fai, size_t i, std::vector<char>& g) {
    g.clear(); g.reserve(fai[i].len);
    size_t q = fai[i].off, end = (i + 1 < fai.size()) ? (size_t)fai[i + 1].off : got;
    while (q < end && g.size() < fai[i].len) {
        const char* s = &raw[q]; const char* e = (const char*)memchr(s, '\n', got + 1 - q); size_t len = (size_t)(e - s); q += len + 1;
        if (len > 0 && s[len - 1] == '\r') len--;
        if (len > 0 && s[0] == '>') break;
        g.insert(g.end(), s, s + len);
    }
}

}  // namespace scs
