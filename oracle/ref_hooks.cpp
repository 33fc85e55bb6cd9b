/* TEST INFRASTRUCTURE — see ref_hooks.h. Linked only into oracle/_ref/bin/scssim_replay. */
#include "ref_hooks.h"

#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

namespace {

enum { S_WREAL = 0, S_WINT, S_MRAND, S_MREAL, S_MINT, S_GCF, S_COUNT };
const char* kNames[S_COUNT] = {"wreal", "wint", "mrand", "mreal", "mint", "gcf"};

struct Hooks {
    FILE* f[S_COUNT];
    bool logging;
    bool seeded;
    unsigned seed;
    pthread_t main_tid;
    pthread_mutex_t mu;

    Hooks() : logging(false), seeded(false), seed(0) {
        pthread_mutex_init(&mu, NULL);
        main_tid = pthread_self();   /* static init runs on the main thread */
        for (int i = 0; i < S_COUNT; i++) f[i] = NULL;
        const char* s = getenv("SCS_SEED");
        if (s && *s) { seeded = true; seed = (unsigned)strtoul(s, NULL, 0); }
        const char* p = getenv("SCS_REPLAY_LOG");
        if (p && *p) {
            logging = true;
            for (int i = 0; i < S_COUNT; i++) {
                std::string fn = std::string(p) + "." + kNames[i] + ".bin";
                f[i] = fopen(fn.c_str(), "wb");
                if (!f[i]) { fprintf(stderr, "scs_ref_hooks: cannot open %s\n", fn.c_str()); exit(2); }
                setvbuf(f[i], NULL, _IOFBF, 1 << 22);
            }
        }
    }
    ~Hooks() {
        for (int i = 0; i < S_COUNT; i++) if (f[i]) fclose(f[i]);
    }
    void put(int s, const void* p, size_t n) {
        if (!logging) return;
        pthread_mutex_lock(&mu);
        fwrite(p, 1, n, f[s]);
        pthread_mutex_unlock(&mu);
    }
};

Hooks g;

}  // namespace

extern "C" unsigned scs_ref_seed(unsigned salt, unsigned fallback) {
    return g.seeded ? g.seed + salt : fallback;
}

extern "C" void scs_ref_log_engine(int is_int_engine, double raw) {
    uint32_t x = (uint32_t)raw;
    bool is_main = pthread_equal(pthread_self(), g.main_tid);
    int s = is_main ? (is_int_engine ? S_MINT : S_MREAL) : (is_int_engine ? S_WINT : S_WREAL);
    g.put(s, &x, 4);
}

extern "C" int scs_ref_rand(void) {
    int r = rand();
    uint32_t x = ((uint32_t)r) << 1;   /* r / 2^31 == x / 2^32 */
    g.put(S_MRAND, &x, 4);
    return r;
}

extern "C" void scs_ref_log_gc(int gc, double v) {
    (void)gc;
    g.put(S_GCF, &v, 8);
}
