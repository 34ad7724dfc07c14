// afq_resolve5.cuh — per-cell resolve, version 5 (experimental, AFQ_RESOLVE=5): no hash table.
//
// The cr-like result only needs the (umi, gene) pairs GROUPED: per UMI the gene(s) with the most
// reads, then per output slot the number of UMIs. v3 groups with an open-address table whose probe
// loops dominate its instruction count (lane divergence, profiles/r1d). v5 groups by
//   partition (counting sort on a hash bucket of the UMI)  +  a tiny insertion sort per bucket,
// one thread per bucket: all data-dependent loops run over a thread's own few elements, there is
// no probing, no CAS, and the arena can never overflow (pairs <= alignments <= capacity).
//
//   pass 1  records -> (umi, gene) pairs at the record's own alignment positions of the staging
//           row (no atomics for placement) + histogram of UMI-hash buckets
//   pass 2  bucket offsets (block scan)          pass 3  scatter pairs into bucket segments
//   pass 4  per bucket: insertion sort, runs = read counts, per UMI arg-max gene set -> slot
//   pass 5  compact winners                       pass 6-7  partition winners by slot range
//   pass 8  per slot bucket: insertion sort, distinct count     pass 9  emit (slot, count) ascending
#pragma once
#include "afq_kernels.cuh"

namespace afq {

__host__ __device__ constexpr u32 bin5_buckets_log2(int b) { return bin_cap_log2(b) > 8 ? bin_cap_log2(b) - 3 : 5; }
__host__ __device__ constexpr u32 bin5_threads(int b) { return b == 0 ? 32 : (b == 1 ? 128 : (b == 2 ? 256 : (b == 3 ? 512 : 1024))); }
__host__ __device__ constexpr size_t bin5_smem_bytes(int b) {
  return ((size_t)12 << bin_cap_log2(b)) + ((size_t)12 << bin5_buckets_log2(b)) + 64;
}

struct Arena5 {
  u64* K;      // [cap]  pairs grouped by bucket; later u32 S2[2*cap]
  u32* S1;     // [cap]  compacted winners
  u32* bc;     // [NB+1] bucket counters
  u32* bo;     // [NB+1] bucket offsets
  u32* bw;     // [NB+1] per-bucket winner / distinct counts, then their offsets
  u32 nb_log2;
};

__device__ __forceinline__ u32 umi_bucket(u32 umi, u32 nb_log2) { return (umi * 0x9E3779B1u) >> (32 - nb_log2); }

#define AFQ5_FOR(X, N) \
  for (u32 X##_b = 0, X = threadIdx.x; (__syncwarp(), X##_b < (N)); X##_b += blockDim.x, X += blockDim.x) \
    if (X < (N))

__device__ inline void resolve_cell5(const KArgs& a, u32 cell, const Arena5& A, CellShared* sh) {
  const u32 NB = 1u << A.nb_log2;
  const u32 T = blockDim.x, tid = threadIdx.x;
  const u64 r0 = a.cell_rec_off[cell], r1 = a.cell_rec_off[cell + 1];
  const u32 nrec = (u32)(r1 - r0);
  const u32 f0 = a.ref_off[r0];
  const u32 P = a.ref_off[r1] - f0;
  u32* st_umi = a.stage_col + f0;                                   // staging row doubles as pair scratch
  u32* st_gene = reinterpret_cast<u32*>(a.stage_val + f0);
  for (u32 i = tid; i <= NB; i += T) A.bc[i] = 0;
  if (tid == 0) { sh->red_max = 0; sh->red_cnt = 0; sh->nwin = 0; }
  __syncthreads();

  // ---- pass 1: pairs at the record's own alignment positions + bucket histogram -------------
  AFQ5_FOR(i, nrec) {
    const u64 r = r0 + i;
    const u32 umi = a.umi[r];
    const u32 o0 = a.ref_off[r], o1 = a.ref_off[r + 1];
    if (o1 == o0) continue;
    const u32 g0 = __ldg(a.t2g + a.refs[o0]);
    const u32 b = umi_bucket(umi, A.nb_log2);
    bool keep0 = true;
    u32 extra = 0;
    for (u32 k = o0 + 1; k < o1; ++k) {
      const u32 g = __ldg(a.t2g + a.refs[k]);
      bool dup = g == g0;
      if (!dup) {
        if (a.mode == MODE_TRIVIAL) { keep0 = false; }   // multi-gene record: dropped entirely
        else for (u32 j = o0 + 1; j < k; ++j) if (st_gene[j - f0] == g) { dup = true; break; }
      }
      const bool put = !dup && a.mode != MODE_TRIVIAL;
      st_umi[k - f0] = umi;
      st_gene[k - f0] = put ? g : NONE32;
      extra += put ? 1u : 0u;
    }
    st_umi[o0 - f0] = umi;
    st_gene[o0 - f0] = keep0 ? g0 : NONE32;
    const u32 cnt = (keep0 ? 1u : 0u) + extra;
    if (cnt) atomicAdd(&A.bc[b], cnt);
  }
  __syncthreads();
  // ---- pass 2-3: bucket offsets, scatter ------------------------------------------------------
  block_exscan_array_small(A.bc, A.bo, NB, sh->scan);
  AFQ5_FOR(i, P) {
    const u32 g = st_gene[i];
    if (g == NONE32) continue;
    const u32 umi = st_umi[i];
    const u32 b = umi_bucket(umi, A.nb_log2);
    A.K[A.bo[b] + atomicSub(&A.bc[b], 1u) - 1] = ((u64)umi << 32) | g;
  }
  __syncthreads();
  // ---- pass 4: per bucket: sort, run lengths, per-UMI arg-max -> winner slots (over the consumed keys)
  AFQ5_FOR(b, NB) {
    const u32 lo = A.bo[b], hi = A.bo[b + 1];
    for (u32 i = lo + 1; i < hi; ++i) {
      const u64 x = A.K[i];
      u32 j = i;
      while (j > lo && A.K[j - 1] > x) { A.K[j] = A.K[j - 1]; --j; }
      A.K[j] = x;
    }
    u32* wout = reinterpret_cast<u32*>(A.K + lo);
    u32 nw = 0, i = lo;
    while (i < hi) {
      const u32 u = (u32)(A.K[i] >> 32);
      u32 maxw = 0, nb = 0;
      u32 best[10];
      while (i < hi && (u32)(A.K[i] >> 32) == u) {
        const u64 k = A.K[i];
        u32 w = 1;
        ++i;
        while (i < hi && A.K[i] == k) { ++w; ++i; }
        if (a.mode == MODE_TRIVIAL) { wout[nw++] = (u32)k; continue; }   // every distinct (gene, umi) counts
        if (w > maxw) { maxw = w; nb = 1; best[0] = (u32)k; }
        else if (w == maxw) { if (nb < 10) best[nb] = (u32)k; ++nb; }
      }
      if (a.mode == MODE_TRIVIAL) continue;
      u32 res = NONE32;
      if (!a.usa_mode) { if (nb == 1) res = best[0]; }
      else if (nb <= 10) res = usa_slot_for_label(best, nb, a.uo, a.ao);   // genes ascend inside a UMI
      if (res != NONE32) wout[nw++] = res;
    }
    A.bw[b] = nw;
  }
  __syncthreads();
  // ---- pass 5: compact winners -------------------------------------------------------------------
  block_exscan_array_small(A.bw, A.bc, NB, sh->scan);     // bc[b] = winner offset of bucket b; bc[NB] = m
  const u32 m = A.bc[NB];
  AFQ5_FOR(b, NB) {
    const u32* win = reinterpret_cast<const u32*>(A.K + A.bo[b]);
    const u32 nw = A.bw[b], o = A.bc[b];
    for (u32 j = 0; j < nw; ++j) A.S1[o + j] = win[j];
  }
  __syncthreads();
  // ---- pass 6-7: partition winners by slot range --------------------------------------------------
  const u32 bits = 32 - __clz((int)(a.num_rows > 1 ? a.num_rows - 1 : 1));
  const u32 shift = bits > A.nb_log2 ? bits - A.nb_log2 : 0;
  for (u32 i = tid; i <= NB; i += T) A.bc[i] = 0;
  __syncthreads();
  for (u32 i = tid; i < m; i += T) atomicAdd(&A.bc[A.S1[i] >> shift], 1u);
  __syncthreads();
  block_exscan_array_small(A.bc, A.bo, NB, sh->scan);
  u32* S2 = reinterpret_cast<u32*>(A.K);
  for (u32 i = tid; i < m; i += T) { const u32 v = A.S1[i]; const u32 b = v >> shift; S2[A.bo[b] + atomicSub(&A.bc[b], 1u) - 1] = v; }
  __syncthreads();
  // ---- pass 8: per slot bucket: sort, distinct count ----------------------------------------------
  AFQ5_FOR(b, NB) {
    const u32 lo = A.bo[b], hi = A.bo[b + 1];
    for (u32 i = lo + 1; i < hi; ++i) {
      const u32 x = S2[i];
      u32 j = i;
      while (j > lo && S2[j - 1] > x) { S2[j] = S2[j - 1]; --j; }
      S2[j] = x;
    }
    u32 nd = 0;
    for (u32 i = lo; i < hi; ++i) if (i == lo || S2[i] != S2[i - 1]) ++nd;
    A.bw[b] = nd;
  }
  __syncthreads();
  block_exscan_array_small(A.bw, A.bc, NB, sh->scan);
  const u32 nnz = A.bc[NB];
  // ---- pass 9: emit ---------------------------------------------------------------------------------
  const u64 out_base = f0;
  u32 lmax = 0;
  AFQ5_FOR(b, NB) {
    const u32 lo = A.bo[b], hi = A.bo[b + 1];
    u64 o = out_base + A.bc[b];
    u32 i = lo;
    while (i < hi) {
      const u32 x = S2[i];
      u32 j = i + 1;
      while (j < hi && S2[j] == x) ++j;
      a.stage_col[o] = x;
      a.stage_val[o] = (float)(j - i);
      lmax = (j - i) > lmax ? (j - i) : lmax;
      ++o;
      i = j;
    }
  }
  if (lmax) atomicMax(&sh->red_max, lmax);
  __syncthreads();
  const float sum = (float)m;
  const float mean = sum / (float)nnz;
  u32 lover = 0;
  for (u32 i = tid; i < nnz; i += T)
    if (a.stage_val[out_base + i] > mean) ++lover;
  if (lover) atomicAdd(&sh->red_cnt, lover);
  __syncthreads();
  if (tid == 0) {
    a.sum_umi[cell] = sum;
    a.max_umi[cell] = (float)sh->red_max;
    a.num_expr[cell] = nnz;
    a.num_over_mean[cell] = sh->red_cnt;
    u8 f = 0;
    if (a.tiny_eligible && (r1 - r0) < a.small_thresh) f |= 1;
    if (nnz == 0) f |= 4;
    a.flags[cell] = f;
  }
  __syncthreads();
}

template <int BIN>
__global__ void __launch_bounds__(bin5_threads(BIN)) k_resolve5_smem(KArgs a) {
  constexpr u32 CAP = 1u << bin_cap_log2(BIN);
  constexpr u32 NB = 1u << bin5_buckets_log2(BIN);
  AFQ_DYN_SMEM(smem_raw);
  Arena5 A;
  A.K = reinterpret_cast<u64*>(smem_raw);
  A.S1 = reinterpret_cast<u32*>(A.K + CAP);
  A.bc = A.S1 + CAP;
  A.bo = A.bc + NB + 1;
  A.bw = A.bo + NB + 1;
  A.nb_log2 = bin5_buckets_log2(BIN);
  __shared__ CellShared sh;
  const u32 count = a.ctl->bin_count[BIN];
  const u32* list = a.bin_list + (u64)BIN * a.n_cells;
  for (;;) {
    if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[BIN], 1u);
    __syncthreads();
    const u32 job = sh.job;
    if (job >= count) break;
    resolve_cell5(a, list[job], A, &sh);
    __syncthreads();
  }
}

// cells with more alignments than the largest shared-memory arena: same passes on the global arena
__global__ void __launch_bounds__(1024) k_resolve5_large(KArgs a, u32 list_id) {
  __shared__ CellShared sh;
  __shared__ u32 s_buckets[3 * (2048 + 1)];
  const u64 arena = (u64)blockIdx.x << a.large_cap_log2;
  Arena5 A;
  A.K = a.large_keys + arena;
  A.S1 = a.large_cnts + arena;
  A.bc = s_buckets;
  A.bo = A.bc + 2049;
  A.bw = A.bo + 2049;
  A.nb_log2 = 11;
  const u32 count = a.ctl->bin_count[list_id];
  const u32* list = a.bin_list + (u64)list_id * a.n_cells;
  for (;;) {
    if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[list_id], 1u);
    __syncthreads();
    const u32 job = sh.job;
    if (job >= count) break;
    const u32 cell = list[job];
    const u64 r0 = a.cell_rec_off[cell], r1 = a.cell_rec_off[cell + 1];
    const u32 p = a.ref_off[r1] - a.ref_off[r0];
    if ((u64)p > (1ull << a.large_cap_log2)) {
      if (threadIdx.x == 0) {
        atomicOr(&a.ctl->error, (u32)DEV_ERR_CELL_TOO_LARGE);
        atomicMax(&a.ctl->max_cell_refs, p);
        a.sum_umi[cell] = 0; a.max_umi[cell] = 0; a.num_expr[cell] = 0; a.num_over_mean[cell] = 0; a.flags[cell] = 4;
      }
      __syncthreads();
      continue;
    }
    resolve_cell5(a, cell, A, &sh);
    __syncthreads();
  }
}

}  // namespace afq
