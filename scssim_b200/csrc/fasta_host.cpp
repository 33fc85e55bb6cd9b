#include "fasta_host.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>

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
    // Pass 1 (threads over byte ranges): where the records start — a '>' at the beginning of a line. Pass 2 (threads over
    // records): the line-geometry state machine of FastaIndex::indexReference (Fasta.cpp:103-191) on each record's bytes.
    const char* const end = data + size;
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const size_t nchunk = size >= (64u << 20) ? hw : 1;
    std::vector<std::vector<size_t>> found(nchunk);
    auto scan_headers = [&](size_t k) {
        const size_t lo = size * k / nchunk, hi = size * (k + 1) / nchunk;
        const char* s = data + lo;
        while (s < data + hi) {
            const char* g = (const char*)memchr(s, '>', (size_t)(data + hi - s));
            if (!g) break;
            if (g == data || g[-1] == '\n') found[k].push_back((size_t)(g - data));
            s = g + 1;
        }
    };
    auto index_record = [&](FaiRec& a, const char* s, const char* rec_end) {
        // s points at '>' ; the record's lines run up to rec_end
        bool header = true;
        while (s < rec_end) {
            const char* e = (const char*)memchr(s, '\n', (size_t)(rec_end - s));
            if (!e) e = rec_end;                   // last line of the file without a terminator
            size_t len = (size_t)(e - s);
            const size_t llen = len + 1;
            if (len > 0 && s[len - 1] == '\r') len--;
            if (header) {
                a.header.assign(s + 1, len ? len - 1 : 0);
                a.name = a.header.substr(0, a.header.find_first_of(" \t"));
                a.off = (uint64_t)(s - data) + llen;
                header = false;
            } else if (len == 0) { if (a.len) a.short_lines++; }          // a blank line is fine only at the very end of a record
            else {
                if (a.blen == 0) { a.blen = (uint32_t)len; a.llen = (uint32_t)llen; }
                if (a.short_lines) a.regular = false;               // bases after a short or blank line
                if (len != a.blen || llen != a.llen) { if (len > a.blen) a.regular = false; a.short_lines++; }
                a.len += len;
            }
            s = e + 1;
        }
    };
    auto run_parallel = [&](size_t n, const std::function<void(size_t)>& fn) {
        if (n <= 1 || nchunk == 1) { for (size_t i = 0; i < n; i++) fn(i); return; }
        std::atomic<size_t> next{0};
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < std::min<size_t>(hw, n); t++) ts.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); });
        for (auto& t : ts) t.join();
    };
    run_parallel(nchunk, scan_headers);
    std::vector<size_t> starts;
    for (auto& v : found) starts.insert(starts.end(), v.begin(), v.end());
    // a '>' line with nothing after the '>' is not a header for the reader above (len == 0 there means a blank line): keep
    // the two readers identical by dropping it here as well
    starts.erase(std::remove_if(starts.begin(), starts.end(), [&](size_t o) { return o + 1 >= size || data[o + 1] == '\n' || (data[o + 1] == '\r' && (o + 2 >= size || data[o + 2] == '\n')); }), starts.end());
    fai.resize(starts.size());
    run_parallel(starts.size(), [&](size_t i) { index_record(fai[i], data + starts[i], i + 1 < starts.size() ? data + starts[i + 1] : end); });
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
