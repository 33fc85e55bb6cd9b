// File sinks behind the pinned slab ring (replace SeqWriter::write's single mutex-guarded ofstream,
// /root/reference/lib/seqwriter/SeqWriter.cpp:41-54, and the 50 MB per-thread buffers that feed it, Amplicon.cpp:413-424).
//
// AsyncFileConsumer — the sink of scs_yield_reads(): the launching thread never waits for the disk unless the whole ring is
//   still being written. A drainer thread waits for each slab's copy event and cuts the slab into 4 MiB chunks that a pool of
//   writer threads pwrite()s at their FINAL file offsets (slabs arrive in file order, so offsets are running sums); the `_1`/`_2`
//   files of a paired run are written concurrently and stay record-aligned because both are plain byte streams in slab order.
//   Files are opened O_DIRECT where the file system has it (no page-cache copy): the read stage places every slab in its pinned
//   slot at (file offset mod 4096), the drainer prepends the < 4 KiB carried over from the previous slab, whole blocks go out
//   directly and the last partial block — and a first one shared with the previous rank's region — through a second, buffered
//   descriptor. Space is preallocated with fallocate() and trimmed with ftruncate() at the end.
//   Several ranks can write disjoint regions of ONE file (each at its own base offset).
// ParallelFileWriter — synchronous variant used by the simuvars FASTA output and the host-side tests.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "slab_sink.h"

namespace scs {

class AsyncFileConsumer : public SlabConsumer {
  public:
    // threads: writer threads; ring: pinned slots per file; device: CUDA device the copy events belong to
    AsyncFileConsumer(int threads, int ring, int device, bool want_direct);
    ~AsyncFileConsumer() override;
    // Opens file slot `file`. create: truncate/create and preallocate `prealloc` bytes (0: nothing); otherwise the file must exist
    // (another rank created it). This rank's bytes start at `base`. `own_end`: trim the file to base + written bytes at finish.
    bool open(int file, const std::string& path, uint64_t base, bool create, uint64_t prealloc, bool own_end);
    int ring_slots() const override { return ring_; }
    uint64_t phase(int file) const override { return F_[file].direct ? (F_[file].off & 4095ull) : 0; }
    int acquire(int slot) override;
    int submit(int slot, cudaEvent_t copied, char* const p[2], const uint64_t bytes[2]) override;
    int service() override { return failed_ ? 1 : 0; }
    int finish() override;
    uint64_t bytes(int file) const { return F_[file].off - F_[file].base; }
    bool direct(int file) const { return F_[file].direct; }

  private:
    struct File {
        int fd_direct = -1, fd_buf = -1; bool direct = false, own_end = false;
        uint64_t base = 0, off = 0;        // this rank's region starts at base; off = file offset of the next slab's first byte
        uint64_t pend_lo = 0;              // first byte not handed to a writer yet (block aligned once past the shared head block)
        uint64_t direct_lo = 0;            // bytes below this offset share a block with the previous rank: buffered writes only
        char* carry = nullptr; uint64_t carry_len = 0;   // bytes [pend_lo, pend_lo + carry_len) waiting for the rest of their block
        std::string path;
    };
    struct Slab { int slot; cudaEvent_t ev; char* p[2]; uint64_t n[2]; uint64_t off[2]; };
    struct Task { int fd; const char* p; size_t n; uint64_t off; int slot; };
    void drainer();
    void worker();
    void enqueue(int fd, const char* p, size_t n, uint64_t off, int slot);
    bool write_now(int fd, const char* p, size_t n, uint64_t off);
    File F_[2];
    int ring_, device_; bool want_direct_;
    std::vector<std::thread> pool_; std::thread drain_;
    std::mutex mu_; std::condition_variable cv_work_, cv_slot_, cv_drain_;
    std::deque<Task> tasks_; std::deque<Slab> slabs_;
    std::vector<int> busy_;   // per slot: 0 free, otherwise 1 (submitted) + chunks in flight
    bool stop_ = false, failed_ = false, finishing_ = false;
};

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
