#include "file_sink.h"

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>

namespace scs {

static constexpr size_t kChunk = 4u << 20;
static constexpr uint64_t kBlock = 4096;

// ------------------------------------------------------------------------------------------------ AsyncFileConsumer
AsyncFileConsumer::AsyncFileConsumer(int threads, int ring, int device, bool want_direct)
    : ring_(std::max(2, std::min(ring, 64))), device_(device), want_direct_(want_direct), busy_((size_t)std::max(2, std::min(ring, 64)), 0) {
    threads = std::max(1, std::min(threads, 64));
    for (int i = 0; i < threads; i++) pool_.emplace_back([this] { worker(); });
    drain_ = std::thread([this] { drainer(); });
}

AsyncFileConsumer::~AsyncFileConsumer() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_work_.notify_all(); cv_drain_.notify_all();
    if (drain_.joinable()) drain_.join();
    for (auto& t : pool_) t.join();
    for (File& F : F_) {
        if (F.fd_direct >= 0) ::close(F.fd_direct);
        if (F.fd_buf >= 0) ::close(F.fd_buf);
        free(F.carry);
    }
}

bool AsyncFileConsumer::open(int file, const std::string& path, uint64_t base, bool create, uint64_t prealloc, bool own_end) {
    File& F = F_[file];
    F.path = path; F.base = F.off = base; F.own_end = own_end;
    F.fd_buf = ::open(path.c_str(), create ? (O_WRONLY | O_CREAT | O_TRUNC) : O_WRONLY, 0644);
    if (F.fd_buf < 0) return false;
    if (create && prealloc) (void)posix_fallocate(F.fd_buf, 0, (off_t)prealloc);   // best effort: not every file system can
    F.direct = false;
    if (want_direct_) {
        F.fd_direct = ::open(path.c_str(), O_WRONLY | O_DIRECT);
        F.direct = F.fd_direct >= 0 && posix_memalign((void**)&F.carry, kBlock, kBlock) == 0;
    }
    F.direct_lo = (base + kBlock - 1) & ~(kBlock - 1);
    F.pend_lo = F.direct_lo; F.carry_len = 0;
    return true;
}

int AsyncFileConsumer::acquire(int slot) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_slot_.wait(lk, [&] { return busy_[slot] == 0 || failed_; });
    return failed_ ? 1 : 0;
}

int AsyncFileConsumer::submit(int slot, cudaEvent_t copied, char* const p[2], const uint64_t bytes[2]) {
    Slab s; s.slot = slot; s.ev = copied;
    for (int f = 0; f < 2; f++) { s.p[f] = p[f]; s.n[f] = bytes[f]; s.off[f] = F_[f].off; F_[f].off += bytes[f]; }
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (failed_) return 1;
        busy_[slot] = 1;   // the drainer's own token; chunks in flight are counted on top
        slabs_.push_back(s);
    }
    cv_drain_.notify_one();
    return 0;
}

void AsyncFileConsumer::enqueue(int fd, const char* p, size_t n, uint64_t off, int slot) {
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (size_t o = 0; o < n; o += kChunk) { tasks_.push_back({fd, p + o, std::min(kChunk, n - o), off + o, slot}); busy_[slot]++; }
    }
    cv_work_.notify_all();
}

bool AsyncFileConsumer::write_now(int fd, const char* p, size_t n, uint64_t off) {
    while (n) {
        ssize_t w = ::pwrite(fd, p, n, (off_t)off);
        if (w < 0 && errno == EINTR) continue;
        if (w <= 0) return false;
        p += w; n -= (size_t)w; off += (uint64_t)w;
    }
    return true;
}

void AsyncFileConsumer::drainer() {
    if (device_ >= 0) cudaSetDevice(device_);
    for (;;) {
        Slab s;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_drain_.wait(lk, [&] { return stop_ || !slabs_.empty(); });
            if (slabs_.empty()) return;
            s = slabs_.front(); slabs_.pop_front();
        }
        bool ok = s.ev == nullptr || cudaEventSynchronize(s.ev) == cudaSuccess;   // the slab has landed in its pinned slot
        for (int f = 0; f < 2 && ok; f++) {
            File& F = F_[f];
            char* p = s.p[f]; uint64_t off = s.off[f], n = s.n[f];
            if (!n) continue;
            if (!F.direct) { enqueue(F.fd_buf, p, n, off, s.slot); continue; }
            if (off < F.direct_lo) {   // shares a block with the previous rank's region: exact byte range, buffered
                const uint64_t h = std::min(n, F.direct_lo - off);
                enqueue(F.fd_buf, p, h, off, s.slot);
                p += h; off += h; n -= h;
                if (!n) continue;
            }
            // the slab sits in its slot at (file offset mod 4096): the bytes carried over from the previous slab go right in front of it
            char* a = p - F.carry_len;
            if (F.carry_len) memcpy(a, F.carry, F.carry_len);
            const uint64_t avail = F.carry_len + n, al = avail & ~(kBlock - 1), rem = avail - al;
            if (rem) memcpy(F.carry, a + al, rem);
            if (al) enqueue(F.fd_direct, a, al, F.pend_lo, s.slot);
            F.pend_lo += al; F.carry_len = rem;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (!ok) failed_ = true;
            if (--busy_[s.slot] == 0 || failed_) cv_slot_.notify_all();
        }
    }
}

void AsyncFileConsumer::worker() {
    for (;;) {
        Task t;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_work_.wait(lk, [&] { return stop_ || !tasks_.empty(); });
            if (tasks_.empty()) return;
            t = tasks_.front(); tasks_.pop_front();
        }
        const bool ok = write_now(t.fd, t.p, t.n, t.off);
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (!ok) failed_ = true;
            if (--busy_[t.slot] == 0 || failed_) cv_slot_.notify_all();
        }
    }
}

int AsyncFileConsumer::finish() {
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_slot_.wait(lk, [&] {
            if (!slabs_.empty()) return false;
            for (int b : busy_) if (b) return false;
            return true;
        });
    }
    bool ok = !failed_;
    for (File& F : F_) {
        if (F.fd_buf < 0) continue;
        if (F.direct && F.carry_len) ok = write_now(F.fd_buf, F.carry, F.carry_len, F.pend_lo) && ok;   // the last partial block
        F.carry_len = 0;
        if (F.own_end && ftruncate(F.fd_buf, (off_t)F.off) != 0) ok = false;   // trim the preallocation to what was written
        if (F.fd_direct >= 0) { if (::close(F.fd_direct) != 0) ok = false; F.fd_direct = -1; }
        if (::close(F.fd_buf) != 0) ok = false;
        F.fd_buf = -1;
    }
    if (!ok) failed_ = true;
    return ok ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ ParallelFileWriter
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
