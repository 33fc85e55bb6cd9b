// How expensive is growing multi-GB device arrays on a cold process? (tuning probe for the amplification stage at default gamma)
//   nvcc -O2 -arch=sm_100a profiles/alloc_probe.cu -o /tmp/alloc_probe -lcuda && /tmp/alloc_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaSetDevice(0); cudaFree(0);
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 0); uint64_t keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    // 1. cudaMallocAsync growth pattern: new = old * 1.7, copy, free old (what DevBuf::reserve does), sizes 0.1 .. 12 GB
    {
        void* p = nullptr; size_t sz = 0;
        for (size_t want = 100ull << 20; want < (13ull << 30); want = want * 17 / 10) {
            double t0 = now(); void* q = nullptr; cudaError_t e = cudaMallocAsync(&q, want, st); cudaStreamSynchronize(st); double t1 = now();
            if (p) { cudaMemcpyAsync(q, p, sz, cudaMemcpyDeviceToDevice, st); cudaFreeAsync(p, st); } cudaStreamSynchronize(st); double t2 = now();
            printf("mallocAsync %6.2f GB: alloc %8.2f ms (%s), copy+free %8.2f ms\n", want / 1073741824.0, t1 - t0, cudaGetErrorString(e), t2 - t1);
            p = q; sz = want;
        }
        cudaFreeAsync(p, st); cudaStreamSynchronize(st);
        // again, warm pool
        double t0 = now(); void* q = nullptr; cudaMallocAsync(&q, 10ull << 30, st); cudaStreamSynchronize(st); printf("mallocAsync 10 GB from the warm pool: %.2f ms\n", now() - t0); cudaFreeAsync(q, st); cudaStreamSynchronize(st);
        cudaMemPoolTrimTo(pool, 0);
    }
    // 2. plain cudaMalloc / cudaFree
    for (size_t want : {1ull << 30, 4ull << 30, 16ull << 30}) {
        double t0 = now(); void* q = nullptr; cudaMalloc(&q, want); double t1 = now(); cudaMemsetAsync(q, 0, want, st); cudaStreamSynchronize(st); double t2 = now(); cudaFree(q); double t3 = now();
        printf("cudaMalloc %5.1f GB: alloc %8.2f ms, memset %8.2f ms, free %8.2f ms\n", want / 1073741824.0, t1 - t0, t2 - t1, t3 - t2);
    }
    // 3. VMM: reserve 64 GB of address space, map 512 MB chunks on demand
    {
        cuInit(0); CUdevice dev; cuDeviceGet(&dev, 0);
        CUmemAllocationProp prop = {}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
        size_t gran = 0; cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
        CUdeviceptr base = 0; double t0 = now(); CUresult r = cuMemAddressReserve(&base, 64ull << 30, 0, 0, 0); printf("VMM reserve 64 GB: %d, %.2f ms, granularity %zu\n", (int)r, now() - t0, gran);
        CUmemAccessDesc acc = {}; acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        const size_t chunk = 512ull << 20; std::vector<CUmemGenericAllocationHandle> hs;
        double tc = 0, tm = 0, ta = 0;
        for (int i = 0; i < 24; i++) {
            CUmemGenericAllocationHandle h; double a = now(); r = cuMemCreate(&h, chunk, &prop, 0); double b = now(); cuMemMap(base + i * chunk, chunk, 0, h, 0); double c2 = now(); cuMemSetAccess(base + i * chunk, chunk, &acc, 1); double d = now();
            tc += b - a; tm += c2 - b; ta += d - c2; hs.push_back(h);
            if (r != CUDA_SUCCESS) { printf("cuMemCreate failed %d\n", (int)r); break; }
        }
        printf("VMM 24 x 512 MB = 12 GB: create %.2f ms, map %.2f ms, setaccess %.2f ms\n", tc, tm, ta);
        double t1 = now(); cudaMemsetAsync((void*)base, 1, 24 * chunk, st); cudaStreamSynchronize(st); printf("VMM memset 12 GB: %.2f ms\n", now() - t1);
        for (size_t i = 0; i < hs.size(); i++) { cuMemUnmap(base + i * chunk, chunk); cuMemRelease(hs[i]); }
        cuMemAddressFree(base, 64ull << 30);
    }
    return 0;
}
