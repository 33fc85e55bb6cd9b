#include "file_sink.h"

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>

namespace scs {

static constexpr size_t kChunk = 4u << 20;

ParallelFileWriter::ParallelFileWriter(int threads) {
    threads = std::max(1, std::min(threads, 64));
    for (int i = 0; i < threads; i++) pool_.emplace_back([this] { worker(); });
}

ParallelFileWriter::~ParallelFileWriter() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_work_.notify_all();
    for (auto& t : pool_) t.join();
    for (int f = 0; f < 2; f++) if (fd_[f] >= 0) ::close(fd_[f]);
}

bool ParallelFileWriter::open(int file, const std::string& path) {
    fd_[file] = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    off_[file] = 0;
    return fd_[file] >= 0;
}

void ParallelFileWriter::worker() {
    for (;;) {
        Task t;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_work_.wait(lk, [&] { return stop_ || !queue_.empty(); });
            if (queue_.empty()) return;
            t = queue_.back(); queue_.pop_back();
        }
        bool ok = true;
        while (t.n) {
            ssize_t w = ::pwrite(t.fd, t.p, t.n, (off_t)t.off);
            if (w < 0 && errno == EINTR) continue;
            if (w <= 0) { ok = false; break; }
            t.p += w; t.n -= (size_t)w; t.off += (uint64_t)w;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (!ok) failed_ = true;
            if (--inflight_ == 0) cv_done_.notify_all();
        }
    }
}

int ParallelFileWriter::write(int file, const char* data, size_t n) {
    if (file < 0 || file > 1 || fd_[file] < 0) return 1;
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (size_t o = 0; o < n; o += kChunk) { queue_.push_back({fd_[file], data + o, std::min(kChunk, n - o), off_[file] + o}); inflight_++; }
    }
    cv_work_.notify_all();
    off_[file] += n;
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return inflight_ == 0; });
    return failed_ ? 1 : 0;
}

int ParallelFileWriter::close() {
    int rc = failed_ ? 1 : 0;
    for (int f = 0; f < 2; f++) if (fd_[f] >= 0) { if (::close(fd_[f]) != 0) rc = 1; fd_[f] = -1; }
    return rc;
}

int parallel_file_sink(void* user, int file, const char* data, size_t n) { return ((ParallelFileWriter*)user)->write(file, data, n); }

}  // namespace scs
