#include "nccl_shim.h"

#include <dlfcn.h>

#include <mutex>

namespace scs {

NcclApi& nccl_api() {
    static NcclApi a; static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // a copy already in the process (torch's) wins
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { a.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
        auto sym = [&](const char* n, void** fn) { *fn = dlsym(h, n); if (!*fn) a.why += std::string(" missing ") + n; return *fn != nullptr; };
        bool ok = true;
        ok &= sym("ncclGetVersion", (void**)&a.GetVersion); ok &= sym("ncclGetUniqueId", (void**)&a.GetUniqueId);
        ok &= sym("ncclCommInitRank", (void**)&a.CommInitRank); ok &= sym("ncclCommDestroy", (void**)&a.CommDestroy);
        ok &= sym("ncclCommAbort", (void**)&a.CommAbort); ok &= sym("ncclCommCount", (void**)&a.CommCount);
        ok &= sym("ncclAllReduce", (void**)&a.AllReduce); ok &= sym("ncclAllGather", (void**)&a.AllGather);
        ok &= sym("ncclGetErrorString", (void**)&a.GetErrorString);
        a.ok = ok;
    });
    return a;
}

}  // namespace scs
