#include "fasta_host.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

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

FastaFile::~FastaFile() { if (map_) munmap(map_, map_len_); }

bool FastaFile::open(const char* path, std::string* err) {
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) { if (err) *err = std::string("could not open ") + path; return false; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { ::close(fd); if (err) *err = std::string("could not open ") + path; return false; }
    size = (size_t)sb.st_size;
    if (size) {
        map_ = mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (map_ == MAP_FAILED) { map_ = nullptr; ::close(fd); if (err) *err = std::string("could not map ") + path; return false; }
        map_len_ = size;
        madvise(map_, map_len_, MADV_SEQUENTIAL);
    }
    ::close(fd);
    data = (const char*)map_;
    fai.clear();
    const char* const end = data + size;
    const char* s = data;
    FaiRec* a = nullptr;
    while (s < end) {
        const char* e = (const char*)memchr(s, '\n', (size_t)(end - s));
        if (!e) e = end;                       // last line without a terminator
        size_t len = (size_t)(e - s);
        const size_t llen = len + 1;
        if (len > 0 && s[len - 1] == '\r') len--;
        if (len > 0 && s[0] == '>') {
            fai.emplace_back(); a = &fai.back();
            a->header.assign(s + 1, len - 1);
            a->name = a->header.substr(0, a->header.find_first_of(" \t"));
            a->off = (uint64_t)(s - data) + llen;
        } else if (a) {
            if (len == 0) { if (a->len) a->short_lines++; }          // a blank line is fine only at the very end of a record
            else {
                if (a->blen == 0) { a->blen = (uint32_t)len; a->llen = (uint32_t)llen; }
                if (a->short_lines) a->regular = false;               // bases after a short or blank line
                if (len != a->blen || llen != a->llen) { if (len > a->blen) a->regular = false; a->short_lines++; }
                a->len += len;
            }
        }
        s = e + 1;
    }
    if (fai.empty()) { if (err) *err = "ERROR: reference sequence cannot be empty!"; return false; }
    return true;
}

void FastaFile::gather(size_t i, std::vector<char>& g) const {
    g.clear(); g.reserve(fai[i].len);
    const char* s = data + fai[i].off;
    const char* const end = (i + 1 < fai.size()) ? data + fai[i + 1].off : data + size;
    while (s < end && g.size() < fai[i].len) {
        const char* e = (const char*)memchr(s, '\n', (size_t)(end - s));
        if (!e) e = end;
        size_t len = (size_t)(e - s);
        if (len > 0 && s[len - 1] == '\r') len--;
        if (len > 0 && s[0] == '>') break;
        g.insert(g.end(), s, s + len);
        s = e + 1;
    }
}

void fasta_write_fai(const char* path, const std::vector<FaiRec>& fai) {
    std::string faiPath = std::string(path) + ".fai";
    if (FILE* t = fopen(faiPath.c_str(), "rb")) { fclose(t); return; }
    if (FILE* o = fopen(faiPath.c_str(), "wb")) {
        for (const FaiRec& r : fai) fprintf(o, "%s\t%llu\t%llu\t%u\t%u\n", r.name.c_str(), (unsigned long long)r.len, (unsigned long long)r.off, r.blen, r.llen);
        fclose(o);
    }
}

}  // namespace scs
