// File sink behind the pinned slabs (replaces SeqWriter::write's single mutex-guarded stream,
// /root/reference/lib/seqwriter/SeqWriter.cpp:41-54): every slab handed to the sink is cut into chunks that a small pool
// of host threads writes with pwrite() at their final file offsets — the offsets are known up front because slabs arrive
// in file order. write() returns when the slab is on its way to the page cache, so the pinned buffer can be reused;
// the _1/_2 files of a paired run stay record-aligned because both are plain byte streams in slab order.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace scs {

class ParallelFileWriter {
  public:
    explicit ParallelFileWriter(int threads);
    ~ParallelFileWriter();
    // opens (truncates) file slot `file` (0 or 1); false with errno kept on failure
    bool open(int file, const std::string& path);
    // appends n bytes to file slot `file`; returns 0 on success (all chunks written)
    int write(int file, const char* data, size_t n);
    // flushes nothing (pwrite is unbuffered) but closes the descriptors; returns 0 if every write and close succeeded
    int close();
    uint64_t bytes(int file) const { return off_[file]; }

  private:
    struct Task { int fd; const char* p; size_t n; uint64_t off; };
    void worker();
    int fd_[2] = {-1, -1}; uint64_t off_[2] = {0, 0};
    std::vector<std::thread> pool_;
    std::mutex mu_; std::condition_variable cv_work_, cv_done_;
    std::vector<Task> queue_; size_t inflight_ = 0; bool stop_ = false; bool failed_ = false;
};

// C sink adapter: user = ParallelFileWriter*
int parallel_file_sink(void* user, int file, const char* data, size_t n);

}  // namespace scs
