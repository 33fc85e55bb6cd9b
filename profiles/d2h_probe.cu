// D2H host-link probe (round 2): what limits FASTQ landing in host memory at N GPUs on this box?
// One thread per GPU in one process; every variant moves 256 MiB x REPS per GPU and reports GB/s per GPU and in total.
//   engine  : cudaMemcpyAsync device -> pinned host (what the slab pipeline uses)
//   wc      : the same into write-combined pinned memory (cudaHostAllocWriteCombined)
//   thp     : the same into madvise(MADV_HUGEPAGE) memory registered with cudaHostRegister
//   stores  : a kernel writing 16-byte vectors straight into mapped pinned memory (no copy engine)
// GPU sets: each GPU alone, pairs (0,j), the two halves, every second GPU, all.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o d2h_probe d2h_probe.cu -lpthread ; run: ./d2h_probe > out.json
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "CUDA %s at line %d\n", cudaGetErrorString(e__), __LINE__); exit(1); } } while (0)

static const size_t NB = 256ull << 20;
static const int REPS = 6;

__global__ void store_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

struct Gpu {
    int dev; char* d = nullptr; char* h[3] = {nullptr, nullptr, nullptr}; char* hmap_dev = nullptr; cudaStream_t st; bool thp_ok = false;
};

struct Barrier {
    std::atomic<int> count{0}; std::atomic<int> gen{0}; int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait() { int g = gen.load(); if (count.fetch_add(1) + 1 == n) { count = 0; gen++; } else while (gen.load() == g) std::this_thread::yield(); }
};

int main() {
    int ndev = 0; CK(cudaGetDeviceCount(&ndev));
    std::vector<Gpu> G(ndev);
    for (int i = 0; i < ndev; i++) {
        Gpu& g = G[i]; g.dev = i;
        CK(cudaSetDevice(i)); CK(cudaStreamCreateWithFlags(&g.st, cudaStreamNonBlocking));
        CK(cudaMalloc((void**)&g.d, NB)); CK(cudaMemset(g.d, 1, NB));
        CK(cudaHostAlloc((void**)&g.h[0], NB, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void**)&g.hmap_dev, g.h[0], 0));
        CK(cudaHostAlloc((void**)&g.h[1], NB, cudaHostAllocWriteCombined));
        void* p = mmap(nullptr, NB, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            madvise(p, NB, MADV_HUGEPAGE);
            for (size_t o = 0; o < NB; o += 4096) ((volatile char*)p)[o] = 0;
            if (cudaHostRegister(p, NB, cudaHostRegisterDefault) == cudaSuccess) { g.h[2] = (char*)p; g.thp_ok = true; } else (void)cudaGetLastError();
        }
    }
    auto run = [&](const std::vector<int>& devs, int variant) -> std::vector<double> {
        std::vector<double> rate(devs.size(), 0.0);
        Barrier bar((int)devs.size());
        std::vector<std::thread> ts;
        for (size_t k = 0; k < devs.size(); k++) ts.emplace_back([&, k] {
            Gpu& g = G[devs[k]];
            CK(cudaSetDevice(g.dev));
            auto once = [&] {
                if (variant == 3) store_kernel<<<148 * 4, 256, 0, g.st>>>((uint4*)g.hmap_dev, (const uint4*)g.d, NB / 16);
                else CK(cudaMemcpyAsync(g.h[variant], g.d, NB, cudaMemcpyDeviceToHost, g.st));
            };
            once(); CK(cudaStreamSynchronize(g.st));
            bar.wait();
            auto t0 = std::chrono::steady_clock::now();
            for (int r = 0; r < REPS; r++) once();
            CK(cudaStreamSynchronize(g.st));
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            rate[k] = (double)REPS * NB / s / 1e9;
        });
        for (auto& t : ts) t.join();
        return rate;
    };
    const char* vname[4] = {"engine", "wc", "thp", "stores"};
    std::vector<std::pair<std::string, std::vector<int>>> sets;
    for (int i = 0; i < ndev; i++) sets.push_back({"gpu" + std::to_string(i), {i}});
    for (int j = 1; j < ndev; j++) sets.push_back({"pair0_" + std::to_string(j), {0, j}});
    if (ndev >= 4) {
        std::vector<int> lo, hi, even, all;
        for (int i = 0; i < ndev; i++) { (i < ndev / 2 ? lo : hi).push_back(i); if (i % 2 == 0) even.push_back(i); all.push_back(i); }
        sets.push_back({"lower_half", lo}); sets.push_back({"upper_half", hi}); sets.push_back({"every_second", even});
        if (ndev >= 8) sets.push_back({"0_1_4_5", {0, 1, 4, 5}});
        sets.push_back({"all", all});
    }
    printf("{\"n_gpus\": %d, \"bytes_per_copy\": %zu, \"thp_registered\": %s, \"results\": [\n", ndev, NB, G[0].thp_ok ? "true" : "false");
    bool first = true;
    for (auto& s : sets) {
        for (int v = 0; v < 4; v++) {
            if (v == 2 && !G[0].thp_ok) continue;
            if (v != 0 && s.second.size() == 1 && s.second[0] != 0) continue;   // variants on single GPUs: GPU 0 only
            if (v != 0 && s.first.rfind("pair", 0) == 0) continue;
            std::vector<double> r = run(s.second, v);
            double tot = 0; for (double x : r) tot += x;
            printf("%s  {\"set\": \"%s\", \"variant\": \"%s\", \"total_GBps\": %.1f, \"per_gpu\": [", first ? "" : ",\n", s.first.c_str(), vname[v], tot);
            for (size_t k = 0; k < r.size(); k++) printf("%s%.1f", k ? ", " : "", r[k]);
            printf("]}");
            first = false;
        }
    }
    printf("\n]}\n");
    return 0;
}
