// Growable device arrays on CUDA virtual memory management: a buffer reserves a large virtual range once and maps physical
// memory behind it as it grows — no reallocation, no copy of the old contents, no second copy resident while growing, and no
// free until the buffer dies. The amplicon lists of a human-scale cell at the default primer rate grow to ~12 GB over ten
// passes; with cudaMalloc/cudaMallocAsync + copy each growth step cost 10s-100s of ms on this platform, erratically
// (profiles/r01_alloc_probe.txt). Driver entry points are fetched through cudaGetDriverEntryPoint, so the library does not
// link against libcuda and still loads on a machine without a driver (where every compute call fails loudly anyway).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <vector>

namespace scs {

struct VmmRange {
    CUdeviceptr base = 0; size_t reserved = 0, mapped = 0; int device = 0;
    struct Chunk { CUmemGenericAllocationHandle h; size_t off, size; };
    std::vector<Chunk> chunks;
    // true if the driver offers the VMM entry points (resolved once per process)
    static bool available();
    // reserve `va_bytes` of address space on the current device
    cudaError_t init(size_t va_bytes);
    // make at least `bytes` of the range usable (maps one more physical chunk if needed)
    cudaError_t grow(size_t bytes);
    // unmap and release everything; the caller has drained the streams that use the range
    void release();
};

}  // namespace scs
