// cuda_emu.h — TEST-ONLY single-OS-thread emulation of the CUDA execution model, used to
// run the product's device code (alevin_fry_b200/csrc/*.cuh, compiled with -DAFQ_EMU by
// g++) on a machine without a GPU so kernel logic can be debugged and parity-tested in
// the CPU test tier. It is NOT a fallback: nothing in alevin_fry_b200/ builds or loads it.
//
// Model: a kernel launch runs its blocks one after another; the threads of a block are
// ucontext fibers scheduled round-robin on the calling OS thread. __syncthreads() and the
// warp collectives (__shfl_*_sync, __ballot_sync, ...) are cooperative barriers: a fiber
// that arrives yields until every fiber of the block / warp has arrived. Atomics are plain
// read-modify-writes (one OS thread). Execution is fully deterministic.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#define __shared__ static
#define __restrict__

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct uint4 { unsigned x, y, z, w; };
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
using cudaStream_t = void*;

namespace cuda_emu {

struct Fiber {
  ucontext_t ctx;
  std::vector<unsigned char> stack;
  bool done = false;
};

struct State {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  unsigned cur = 0, nthreads = 0, ndone = 0;
  // block barrier
  unsigned bar_arrived = 0, bar_gen = 0;
  // warp collectives
  unsigned warp_arrived[32] = {0}, warp_gen[32] = {0};
  uint64_t warp_xchg[32][32];
  unsigned char* dyn_smem = nullptr;
  void (*entry)(void*) = nullptr;
  void* entry_arg = nullptr;
};
extern State g;

inline void set_tid(unsigned t) { threadIdx.x = t; g.cur = t; }

// Fiber schedule: 0 = round robin (default), 1 = reverse round robin, >= 2 = pseudo-random with that
// seed (env AFQ_EMU_SCHED). Other schedules shake out code that only works because thread 0 reaches a
// barrier first (e.g. a flag read by the others after thread 0 already rewrote it).
inline unsigned sched_mode() {
  static int m = -1;
  if (m < 0) { const char* e = getenv("AFQ_EMU_SCHED"); m = e ? atoi(e) : 0; }
  return (unsigned)m;
}
inline unsigned sched_rand() {
  static unsigned long long x = 0;
  if (!x) x = 0x9E3779B97F4A7C15ull * (sched_mode() + 1);
  x ^= x << 13; x ^= x >> 7; x ^= x << 17;
  return (unsigned)(x >> 33);
}
// switch from the current fiber to another unfinished one
inline void yield() {
  const unsigned me = g.cur;
  unsigned nxt = me;
  const unsigned mode = sched_mode();
  const unsigned start = mode >= 2 ? sched_rand() % g.nthreads : 0;
  for (unsigned k = 1; k <= g.nthreads; ++k) {
    unsigned c = mode == 0 ? (me + k) % g.nthreads
               : mode == 1 ? (me + g.nthreads - k) % g.nthreads
                           : (start + k) % g.nthreads;
    if (c == me) continue;
    if (!g.fibers[c].done) { nxt = c; break; }
  }
  if (nxt == me) return;
  set_tid(nxt);
  swapcontext(&g.fibers[me].ctx, &g.fibers[nxt].ctx);
  set_tid(me);
}

inline void trampoline() {
  g.entry(g.entry_arg);
  Fiber& f = g.fibers[g.cur];
  f.done = true;
  ++g.ndone;
  if (g.ndone == g.nthreads) {
    setcontext(&g.sched);
  } else {
    const unsigned me = g.cur;
    for (unsigned k = 1; k <= g.nthreads; ++k) {
      unsigned c = (me + k) % g.nthreads;
      if (!g.fibers[c].done) { set_tid(c); setcontext(&g.fibers[c].ctx); }
    }
  }
  abort();
}

template <class F>
void run_block(unsigned nthreads, F&& body) {
  static constexpr size_t STACK = 256 * 1024;
  g.nthreads = nthreads;
  g.ndone = 0;
  g.bar_arrived = 0;
  memset(g.warp_arrived, 0, sizeof(g.warp_arrived));
  if (g.fibers.size() < nthreads) g.fibers.resize(nthreads);
  struct Thunk { F* f; };
  Thunk th{&body};
  g.entry = [](void* p) { (*static_cast<Thunk*>(p)->f)(); };
  g.entry_arg = &th;
  for (unsigned t = 0; t < nthreads; ++t) {
    Fiber& f = g.fibers[t];
    f.done = false;
    if (f.stack.size() != STACK) f.stack.resize(STACK);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.data();
    f.ctx.uc_stack.ss_size = STACK;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  const unsigned first = sched_mode() == 1 ? nthreads - 1 : (sched_mode() >= 2 ? sched_rand() % nthreads : 0);
  set_tid(first);
  swapcontext(&g.sched, &g.fibers[first].ctx);
}

// launch<<<grid, block, smem>>>: blocks run sequentially
template <class... P, class... A>
void launch(void (*kernel)(P...), unsigned grid, unsigned block, size_t smem, A... args) {
  std::vector<unsigned char> dyn(smem + 64);
  g.dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 15) & ~uintptr_t(15));
  gridDim.x = grid;
  blockDim.x = block;
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx.x = b;
    run_block(block, [&]() { kernel(args...); });
  }
  g.dyn_smem = nullptr;
}

inline unsigned warp_of(unsigned t) { return t >> 5; }
inline unsigned warp_size_of(unsigned w) {  // lanes present in warp w of the current block
  unsigned lo = w * 32, hi = lo + 32;
  if (hi > g.nthreads) hi = g.nthreads;
  return hi - lo;
}
inline void warp_barrier() {
  const unsigned w = warp_of(g.cur);
  const unsigned gen = g.warp_gen[w];
  if (++g.warp_arrived[w] == warp_size_of(w)) { g.warp_arrived[w] = 0; ++g.warp_gen[w]; return; }
  while (g.warp_gen[w] == gen) yield();
}

}  // namespace cuda_emu

inline void __syncthreads() {
  using namespace cuda_emu;
  const unsigned gen = g.bar_gen;
  if (++g.bar_arrived == g.nthreads) { g.bar_arrived = 0; ++g.bar_gen; return; }
  while (g.bar_gen == gen) yield();
}
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { cuda_emu::warp_barrier(); }

template <class T>
inline T emu_shfl_from(T v, unsigned src_lane) {
  using namespace cuda_emu;
  static_assert(sizeof(T) <= 8, "shuffle payload");
  const unsigned w = warp_of(g.cur), lane = g.cur & 31;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  g.warp_xchg[w][lane] = raw;
  warp_barrier();
  uint64_t got = (src_lane < warp_size_of(w)) ? g.warp_xchg[w][src_lane] : raw;
  warp_barrier();
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl_from(v, (unsigned)src & 31); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  unsigned lane = cuda_emu::g.cur & 31;
  return emu_shfl_from(v, lane >= d ? lane - d : lane);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  unsigned lane = cuda_emu::g.cur & 31;
  return emu_shfl_from(v, lane + d < 32 ? lane + d : lane);
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) {
  unsigned lane = cuda_emu::g.cur & 31;
  return emu_shfl_from(v, lane ^ (unsigned)m);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  using namespace cuda_emu;
  const unsigned w = warp_of(g.cur), lane = g.cur & 31;
  g.warp_xchg[w][lane] = pred ? 1 : 0;
  warp_barrier();
  unsigned m = 0;
  for (unsigned l = 0; l < warp_size_of(w); ++l) if (g.warp_xchg[w][l]) m |= 1u << l;
  warp_barrier();
  return m;
}
inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
inline int __all_sync(unsigned m, int p) {
  return __ballot_sync(m, p) == (cuda_emu::warp_size_of(cuda_emu::warp_of(cuda_emu::g.cur)) == 32 ? 0xFFFFFFFFu : ((1u << cuda_emu::warp_size_of(cuda_emu::warp_of(cuda_emu::g.cur))) - 1));
}

// ---- atomics (single OS thread => plain RMW) ----------------------------------------------
// With a pseudo-random schedule (AFQ_EMU_SCHED >= 2) a fiber may be pre-empted right BEFORE an atomic:
// other threads then get in between a plain read and the CAS that follows it, so the lost-race
// paths of the kernels' open-address tables and lists (CAS returns someone else's value) are run.
inline void emu_preempt() {
  using namespace cuda_emu;
  if (sched_mode() >= 2 && g.nthreads > 1 && (sched_rand() & 7u) == 0) yield();
}
template <class T> inline T atomicAdd(T* p, T v) { emu_preempt(); T o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T* p, T v) { emu_preempt(); T o = *p; *p = o - v; return o; }
template <class T> inline T atomicMax(T* p, T v) { emu_preempt(); T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { emu_preempt(); T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { emu_preempt(); T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { emu_preempt(); T o = *p; *p = o & v; return o; }
template <class T> inline T atomicExch(T* p, T v) { emu_preempt(); T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T cmp, T v) { emu_preempt(); T o = *p; if (o == cmp) *p = v; return o; }

// ---- intrinsics ----------------------------------------------------------------------------
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (v & (1u << i)) r |= 1u << (31 - i); return r; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline float __uint2float_rn(unsigned v) { return (float)v; }
inline void __threadfence() {}
inline void __threadfence_block() {}
