#include "vmm.h"

#include <mutex>

namespace scs {
namespace {

struct Api {
    CUresult (*addressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};

Api& api() {
    static Api a; static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult st;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess && *fn != nullptr;
        };
        a.ok = get("cuMemAddressReserve", (void**)&a.addressReserve) && get("cuMemAddressFree", (void**)&a.addressFree) && get("cuMemCreate", (void**)&a.create) &&
               get("cuMemRelease", (void**)&a.release) && get("cuMemMap", (void**)&a.map) && get("cuMemUnmap", (void**)&a.unmap) &&
               get("cuMemSetAccess", (void**)&a.setAccess) && get("cuMemGetAllocationGranularity", (void**)&a.granularity);
        (void)cudaGetLastError();
    });
    return a;
}

CUmemAllocationProp prop_for(int device) {
    CUmemAllocationProp p = {};
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED; p.location.type = CU_MEM_LOCATION_TYPE_DEVICE; p.location.id = device;
    return p;
}

}  // namespace

bool VmmRange::available() { return api().ok; }

cudaError_t VmmRange::init(size_t va_bytes) {
    Api& a = api();
    if (!a.ok) return cudaErrorNotSupported;
    if (cudaGetDevice(&device) != cudaSuccess) return cudaErrorInvalidDevice;
    const size_t align = 2ull << 20;
    va_bytes = (va_bytes + align - 1) / align * align;
    if (a.addressReserve(&base, va_bytes, 0, 0, 0) != CUDA_SUCCESS) { base = 0; return cudaErrorMemoryAllocation; }
    reserved = va_bytes; mapped = 0;
    return cudaSuccess;
}

cudaError_t VmmRange::grow(size_t bytes) {
    if (bytes <= mapped) return cudaSuccess;
    if (bytes > reserved) return cudaErrorMemoryAllocation;
    Api& a = api();
    const CUmemAllocationProp prop = prop_for(device);
    size_t gran = 2ull << 20;
    a.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (gran == 0) gran = 2ull << 20;
    size_t add = (bytes - mapped + gran - 1) / gran * gran;
    if (mapped + add > reserved) add = reserved - mapped;
    Chunk c{0, mapped, add};
    if (a.create(&c.h, add, &prop, 0) != CUDA_SUCCESS) return cudaErrorMemoryAllocation;
    if (a.map(base + mapped, add, 0, c.h, 0) != CUDA_SUCCESS) { a.release(c.h); return cudaErrorMemoryAllocation; }
    CUmemAccessDesc acc = {}; acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (a.setAccess(base + mapped, add, &acc, 1) != CUDA_SUCCESS) { a.unmap(base + mapped, add); a.release(c.h); return cudaErrorMemoryAllocation; }
    chunks.push_back(c); mapped += add;
    return cudaSuccess;
}

void VmmRange::release() {
    Api& a = api();
    if (!base) return;
    for (const Chunk& c : chunks) { a.unmap(base + c.off, c.size); a.release(c.h); }
    chunks.clear();
    a.addressFree(base, reserved);
    base = 0; reserved = mapped = 0;
}

}  // namespace scs
