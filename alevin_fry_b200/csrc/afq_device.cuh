// afq_device.cuh — device-side building blocks shared by the per-cell kernels.
//
// Everything here is integer / byte work bounded by HBM and shared-memory traffic; no
// tensor-core path exists for it. Block-cooperative routines take
// generic pointers so the same code runs on a shared-memory arena (the binned fast
// kernels) or on a global-memory scratch arena (the giant-cell kernel).
#pragma once
#include <cstdint>
#ifdef AFQ_EMU
// test-only CPU emulation of the CUDA execution model (tests/emu/cuda_emu.h); never part
// of the product build
#include "cuda_emu.h"
#define AFQ_DYN_SMEM(name) unsigned char* name = cuda_emu::g.dyn_smem
#else
#include <cuda_runtime.h>
#define AFQ_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace afq {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;

constexpr u64 EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
constexpr u32 NONE32 = 0xFFFFFFFFu;

__device__ __forceinline__ u32 hash_key(u64 k, u32 log2cap) {
  return (u32)((k * 0x9E3779B97F4A7C15ull) >> (64 - log2cap));
}

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }

// ---- bulk copies global -> shared memory (cp.async.bulk, completion on an mbarrier; SASS: UBLKCP + SYNCS) ----------------
// One thread arms the barrier with the byte count and issues the copies; the copy engine moves the data and every thread
// waits on the barrier's phase. Source and destination must be 16-byte aligned, sizes multiples of 16 bytes.
#ifndef AFQ_EMU
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif
#ifndef AFQ_EMU
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// ---- block-wide exclusive scan of one u32 per thread (blockDim.x <= 1024) -------------
// s_warp must hold 33 u32. Returns the exclusive prefix; *total gets the block sum.
__device__ __forceinline__ u32 block_exscan(u32 v, u32* s_warp, u32* total) {
  const u32 lane = lane_id(), wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  u32 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= (u32)o) inc += t;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    u32 w = lane < nw ? s_warp[lane] : 0;
    u32 winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
      if (lane >= (u32)o) winc += t;
    }
    s_warp[lane] = winc - w;
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  u32 res = s_warp[wid] + inc - v;
  *total = s_warp[32];
  __syncthreads();
  return res;
}

// ---- in-place compaction of (keys, cnts) keeping keys != EMPTY ------------------------
// Safe in place: a chunk is fully read into registers before any thread writes, and
// writes only land at positions <= the chunk start + rank. Returns the kept count.
__device__ inline u32 block_compact_pairs(u64* keys, u32* cnts, u32 cap, u32* s_warp) {
  u32 base = 0;
  for (u32 c0 = 0; c0 < cap; c0 += blockDim.x) {
    const u32 i = c0 + threadIdx.x;
    u64 k = EMPTY_KEY;
    u32 c = 0;
    if (i < cap) { k = keys[i]; c = cnts[i]; }
    const u32 keep = (k != EMPTY_KEY) ? 1u : 0u;
    u32 tot;
    const u32 pos = block_exscan(keep, s_warp, &tot);  // contains the needed barriers
    if (keep) { keys[base + pos] = k; cnts[base + pos] = c; }
    base += tot;
    __syncthreads();
  }
  return base;
}

// in-place compaction of a u32 array keeping values != NONE32
__device__ inline u32 block_compact_u32(u32* a, u32 n, u32* s_warp) {
  u32 base = 0;
  for (u32 c0 = 0; c0 < n; c0 += blockDim.x) {
    const u32 i = c0 + threadIdx.x;
    u32 v = NONE32;
    if (i < n) v = a[i];
    const u32 keep = (v != NONE32) ? 1u : 0u;
    u32 tot;
    const u32 pos = block_exscan(keep, s_warp, &tot);
    if (keep) a[base + pos] = v;
    base += tot;
    __syncthreads();
  }
  return base;
}

// out[i] = sum_{j<i} in[j] for i in [0, n], i.e. out has n+1 entries (out[n] = total).
// in/out must not alias. All threads of the CTA must call it.
__device__ inline void block_exscan_array_small(const u32* in, u32* out, u32 n, u32* s_warp) {
  u32 base = 0;
  for (u32 c0 = 0; c0 < n; c0 += blockDim.x) {
    const u32 i = c0 + threadIdx.x;
    const u32 v = i < n ? in[i] : 0;
    u32 tot;
    const u32 ex = block_exscan(v, s_warp, &tot);
    if (i < n) out[i] = base + ex;
    base += tot;
  }
  if (threadIdx.x == 0) out[n] = base;
  __syncthreads();
}

__device__ __forceinline__ u32 next_pow2(u32 v) {
  return v <= 1 ? 1u : 1u << (32 - __clz(v - 1));
}

// ---- block bitonic sort, ascending by key, u32 payload. n must be a power of two. -----
__device__ inline void block_bitonic_pairs(u64* keys, u32* vals, u32 n) {
  for (u32 k = 2; k <= n; k <<= 1) {
    for (u32 j = k >> 1; j > 0; j >>= 1) {
      for (u32 t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const u32 p = i | j;
        const u64 a = keys[i], b = keys[p];
        const bool up = (i & k) == 0;
        if ((a > b) == up) {
          keys[i] = b; keys[p] = a;
          const u32 va = vals[i], vb = vals[p];
          vals[i] = vb; vals[p] = va;
        }
      }
      __syncthreads();
    }
  }
}

__device__ inline void block_bitonic_u32(u32* a, u32 n) {
  for (u32 k = 2; k <= n; k <<= 1) {
    for (u32 j = k >> 1; j > 0; j >>= 1) {
      for (u32 t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const u32 p = i | j;
        const u32 x = a[i], y = a[p];
        const bool up = (i & k) == 0;
        if ((x > y) == up) { a[i] = y; a[p] = x; }
      }
      __syncthreads();
    }
  }
}

// ---- USA tie rules (src/quant.rs:557-605 == src/utils.rs:688-753) ----------------------
// `best` ascending gene ids, n of them (n > 10 => drop). uo = num_rows/3, ao = 2*uo.
__device__ __forceinline__ bool d_is_spliced(u32 g) { return (g & 1u) == 0; }
__device__ __forceinline__ bool d_same_gene(u32 a, u32 b) { return (a | 1u) == (b | 1u); }

__device__ inline u32 usa_slot_for_label(const u32* best, u32 n, u32 uo, u32 ao) {
  if (n == 0) return NONE32;
  if (n == 1) return d_is_spliced(best[0]) ? (best[0] >> 1) : uo + (best[0] >> 1);
  if (n == 2) {
    const u32 g1 = best[0], g2 = best[1];
    if (d_same_gene(g1, g2)) return ao + (g1 >> 1);
    const bool s1 = d_is_spliced(g1), s2 = d_is_spliced(g2);
    if (s1 && !s2) return g1 >> 1;
    if (!s1 && s2) return g2 >> 1;
    return NONE32;
  }
  if (n <= 10) {
    int sidx = -1;
    for (u32 i = 0; i < n; ++i)
      if (d_is_spliced(best[i])) {
        if (sidx >= 0) return NONE32;
        sidx = (int)i;
      }
    if (sidx < 0) return NONE32;
    const u32 sg = best[sidx];
    if ((u32)sidx + 1 < n && d_same_gene(sg, best[sidx + 1])) return ao + (sg >> 1);
    return sg >> 1;
  }
  return NONE32;
}

}  // namespace afq
