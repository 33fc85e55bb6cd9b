/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Force-included (-include) into the *patched copy* of the reference that
 * oracle/build_ref.sh assembles under oracle/_ref/src. It gives the reference
 * the two things it lacks: a seed and a record of every random
 * draw, so that `scssim_replay -t 1` is deterministic and its draws can be fed
 * to the CPU restatement (oracle/) and to the CUDA path.
 *
 * Streams (one binary file per stream, `${SCS_REPLAY_LOG}.<name>.bin`):
 *   wreal  u32  worker-thread mt19937 "real" engine   ThreadPool.cpp:203-207
 *   wint   u32  worker-thread mt19937 "int" engine    ThreadPool.cpp:208-212
 *   mrand  u32  main-thread libc rand()<<1            MyDefine.cpp:285-292
 *   mreal  u32  main-thread ThreadPool::randomDouble  (randIndx_hp, MyDefine.cpp:243)
 *   mint   u32  main-thread ThreadPool::randomInteger (never hit on this path)
 *   gcf    f64  accepted GC factor                    Profile.cpp:1503-1513
 */
#ifndef SCS_REF_HOOKS_H
#define SCS_REF_HOOKS_H
#ifdef __cplusplus
extern "C" {
#endif
unsigned scs_ref_seed(unsigned salt, unsigned fallback);
void scs_ref_log_engine(int is_int_engine, double raw);
int scs_ref_rand(void);
void scs_ref_log_gc(int gc, double v);
#ifdef __cplusplus
}
#endif
#endif
