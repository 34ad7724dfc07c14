// afq_kernels.cuh — per-cell UMI-resolution kernels (cr-like family) and CSR assembly.
//
// Reference behaviour implemented here (paths relative to /root/reference):
//   * tiny-cell fast path           src/quant.rs:469-657, 794-846
//   * cr-like resolver              src/pugutils.rs:644-850 (+ em_optimize only_unique,
//                                   src/em.rs:499-514; USA: src/utils.rs:673-756)
//   * trivial resolver              src/pugutils.rs:852-911
//   * dense->sparse scan + stats    src/quant.rs:1150-1196
// All three resolvers reduce to one integer pipeline per cell:
//   records -> distinct (umi, gene) pairs with read counts W(u,g)   [hash combine]
//           -> sort by (umi, gene)                                   [bitonic]
//           -> per UMI the arg-max gene set B(u) -> output slot      [segment walk]
//           -> sort slots, run-length count -> (col, val) ascending  [bitonic + scan]
// The tiny path (< small_thresh records) and the <= 250-record path produce the same
// integers as the eq-class path, so cell size only selects the shared-memory arena size.
#pragma once
#include "afq_device.cuh"

namespace afq {

constexpr int NUM_SMEM_BINS = 6;                 // arena capacities below
constexpr int NUM_BINS = NUM_SMEM_BINS + 1;      // + the global-scratch (giant cell) bin
__host__ __device__ constexpr u32 bin_cap_log2(int b) {
  return b == 0 ? 8 : (b == 1 ? 10 : (b == 2 ? 11 : (b == 3 ? 12 : (b == 4 ? 13 : 14))));
}
__host__ __device__ constexpr u32 bin_threads(int b) {
  return b == 0 ? 32 : (b == 1 ? 128 : (b == 2 ? 256 : (b == 3 ? 512 : 1024)));
}

enum : u32 { MODE_CRLIKE = 0, MODE_TRIVIAL = 1 };
enum : u32 { DEV_ERR_CELL_TOO_LARGE = 1 };

constexpr int NUM_LISTS = NUM_BINS + 2;          // + two k_gene_eqc lists (big / normal cells)
struct Ctl {                       // per-batch device control block (zeroed per batch)
  u32 bin_count[NUM_LISTS + 1];
  u32 bin_cursor[NUM_LISTS + 1];
  u32 error;
  u32 max_cell_refs;
  u32 ge_max_n[2];                 // largest record / alignment count on the k_gene_eqc lists
  u32 ge_max_p[2];
  unsigned long long adj_used;     // bump pointer into the adjacency pool
};

struct KArgs {
  // input batch (device)
  u64 n_cells;
  const u64* cell_rec_off;
  const u32* umi;
  const u32* ref_off;
  const u32* refs;
  const u32* t2g;
  // config
  u32 mode, usa_mode, num_rows, uo, ao;
  u64 small_thresh;
  u32 tiny_eligible;
  // work lists
  Ctl* ctl;
  u32* bin_list;                   // [NUM_LISTS][n_cells]
  // staging + per-cell outputs
  u32* stage_col;
  float* stage_val;
  float* sum_umi;
  float* max_umi;
  u32* num_expr;
  u32* num_over_mean;
  u8* flags;
  // giant-cell scratch
  u64* large_keys;
  u32* large_cnts;
  u32 large_cap_log2;
};

struct CellShared {
  u32 scan[40];
  u32 distinct;
  u32 abort;
  u32 job;
  u32 red_max;
  u32 red_cnt;
};

// ------------------------------------------------------------------------------------
// classify cells into arena-size bins by record count
// ------------------------------------------------------------------------------------
__global__ void k_bin_cells(KArgs a, int force_bin) {
  const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n_cells) return;
  const u64 r0 = a.cell_rec_off[c], r1 = a.cell_rec_off[c + 1];
  const u64 n = r1 - r0;
  const u32 p = a.ref_off[r1] - a.ref_off[r0];
  const u64 need = n < (u64)p ? n : (u64)p;  // distinct pairs <= refs; typically << records
  int b = NUM_SMEM_BINS;
#pragma unroll
  for (int i = NUM_SMEM_BINS - 1; i >= 0; --i)
    if (need <= (1ull << bin_cap_log2(i))) b = i;
  if (force_bin >= 0 && force_bin > b) b = force_bin < NUM_SMEM_BINS ? force_bin : NUM_SMEM_BINS;
  const u32 idx = atomicAdd(&a.ctl->bin_count[b], 1u);
  a.bin_list[(u64)b * a.n_cells + idx] = (u32)c;
  if (b == NUM_SMEM_BINS) atomicMax(&a.ctl->max_cell_refs, p);
}

// open-address insert-or-increment of one (umi, gene) key
__device__ __forceinline__ void table_insert(u64* keys, u32* cnts, u64 key, u32 log2cap, u32 limit,
                                             CellShared* sh) {
  const u32 mask = (1u << log2cap) - 1;
  u32 s = hash_key(key, log2cap);
  for (;;) {
    u64 cur = keys[s];
    if (cur == EMPTY_KEY) {
      cur = atomicCAS((unsigned long long*)&keys[s], (unsigned long long)EMPTY_KEY,
                      (unsigned long long)key);
      if (cur == EMPTY_KEY) {
        if (atomicAdd(&sh->distinct, 1u) + 1 > limit) sh->abort = 1;
        cur = key;
      }
    }
    if (cur == key) { atomicAdd(&cnts[s], 1u); return; }
    if (*(volatile u32*)&sh->abort) return;  // table may be full: stop probing
    s = (s + 1) & mask;
  }
}

// ------------------------------------------------------------------------------------
// one cell, block-cooperative. keys/cnts: arena of `cap` = 2^log2cap entries.
// Returns false when the distinct-pair count exceeded `limit` (caller re-queues the cell
// on a larger arena); nothing has been written for the cell in that case.
// ------------------------------------------------------------------------------------
__device__ inline bool resolve_cell(const KArgs& a, u32 cell, u64* keys, u32* cnts, u32 log2cap,
                                    u32 limit, CellShared* sh) {
  const u32 cap = 1u << log2cap;
  const u64 r0 = a.cell_rec_off[cell], r1 = a.cell_rec_off[cell + 1];
  for (u32 i = threadIdx.x; i < cap; i += blockDim.x) { keys[i] = EMPTY_KEY; cnts[i] = 0; }
  if (threadIdx.x == 0) { sh->distinct = 0; sh->abort = 0; sh->red_max = 0; sh->red_cnt = 0; }
  __syncthreads();

  // ---- phase 1: records -> (umi, gene) pairs, combined in the open-address table -------
  for (u64 r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    if (*(volatile u32*)&sh->abort) break;
    const u32 umi = a.umi[r];
    const u32 o0 = a.ref_off[r], o1 = a.ref_off[r + 1];
    if (a.mode == MODE_TRIVIAL) {
      // src/pugutils.rs:870-881: class is multi-gene iff two consecutive refs differ in gene
      if (o1 == o0) continue;
      const u32 g0 = __ldg(a.t2g + a.refs[o0]);
      bool multi = false;
      for (u32 k = o0 + 1; k < o1; ++k)
        if (__ldg(a.t2g + a.refs[k]) != g0) { multi = true; break; }
      if (!multi) table_insert(keys, cnts, ((u64)umi << 32) | g0, log2cap, limit, sh);
      continue;
    }
    for (u32 k = o0; k < o1; ++k) {
      const u32 g = __ldg(a.t2g + a.refs[k]);
      bool dup = false;
      for (u32 j = o0; j < k; ++j)
        if (__ldg(a.t2g + a.refs[j]) == g) { dup = true; break; }
      if (dup) continue;
      table_insert(keys, cnts, ((u64)umi << 32) | g, log2cap, limit, sh);
    }
  }
  __syncthreads();
  if (sh->abort) { __syncthreads(); return false; }

  // ---- phase 2: compact + sort by (umi, gene) ------------------------------------------
  const u32 d = block_compact_pairs(keys, cnts, cap, sh->scan);
  const u32 D = next_pow2(d);
  for (u32 i = d + threadIdx.x; i < D; i += blockDim.x) keys[i] = EMPTY_KEY;
  __syncthreads();
  block_bitonic_pairs(keys, cnts, D);

  // ---- phase 3: per UMI, arg-max gene set -> output slot (written over cnts) ------------
  for (u32 i = threadIdx.x; i < d; i += blockDim.x) {
    const u64 ki = keys[i];
    if (a.mode == MODE_TRIVIAL) { cnts[i] = (u32)ki; continue; }  // every (gene, umi) counts
    const u32 u = (u32)(ki >> 32);
    if (i > 0 && (u32)(keys[i - 1] >> 32) == u) continue;  // not the first entry of this UMI
    u32 maxw = 0, nb = 0;
    u32 best[10];
    u32 j = i;
    for (; j < d; ++j) {
      const u64 kj = keys[j];
      if ((u32)(kj >> 32) != u) break;
      const u32 w = cnts[j];
      if (w > maxw) { maxw = w; nb = 1; best[0] = (u32)kj; }
      else if (w == maxw) { if (nb < 10) best[nb] = (u32)kj; ++nb; }
    }
    u32 res;
    if (!a.usa_mode) res = (nb == 1) ? best[0] : NONE32;
    else res = (nb > 10) ? NONE32 : usa_slot_for_label(best, nb, a.uo, a.ao);
    cnts[i] = res;
    for (u32 k = i + 1; k < j; ++k) cnts[k] = NONE32;
  }
  __syncthreads();

  // ---- phase 4: sort winner slots, run-length count, emit ------------------------------
  const u32 m = block_compact_u32(cnts, d, sh->scan);
  const u32 M = next_pow2(m);
  for (u32 i = m + threadIdx.x; i < M; i += blockDim.x) cnts[i] = NONE32;
  __syncthreads();
  block_bitonic_u32(cnts, M);

  const u64 out_base = a.ref_off[r0];  // this cell's staging region starts at its first ref
  u32 base = 0, lmax = 0;
  for (u32 c0 = 0; c0 < m; c0 += blockDim.x) {
    const u32 i = c0 + threadIdx.x;
    u32 start = 0, slot = 0, len = 0;
    if (i < m) {
      slot = cnts[i];
      start = (i == 0 || cnts[i - 1] != slot) ? 1u : 0u;
      if (start) {
        u32 j = i + 1;
        while (j < m && cnts[j] == slot) ++j;
        len = j - i;
      }
    }
    u32 tot;
    const u32 pos = block_exscan(start, sh->scan, &tot);
    if (start) {
      a.stage_col[out_base + base + pos] = slot;
      a.stage_val[out_base + base + pos] = (float)len;
      lmax = len > lmax ? len : lmax;
    }
    base += tot;
  }
  const u32 nnz = base;
  if (lmax) atomicMax(&sh->red_max, lmax);
  __syncthreads();
  // NumGenesOverMean (src/quant.rs:1190-1194): mean over expressed genes, f32
  const float sum = (float)m;
  const float mean = sum / (float)nnz;
  u32 lover = 0;
  for (u32 i = threadIdx.x; i < nnz; i += blockDim.x)
    if (a.stage_val[out_base + i] > mean) ++lover;
  if (lover) atomicAdd(&sh->red_cnt, lover);
  __syncthreads();
  if (threadIdx.x == 0) {
    a.sum_umi[cell] = sum;
    a.max_umi[cell] = (float)sh->red_max;
    a.num_expr[cell] = nnz;
    a.num_over_mean[cell] = sh->red_cnt;
    u8 f = 0;
    if (a.tiny_eligible && (r1 - r0) < a.small_thresh) f |= 1;  // AFQ_FLAG_TINY
    if (nnz == 0) f |= 4;                                        // AFQ_FLAG_EMPTY
    a.flags[cell] = f;
  }
  __syncthreads();
  return true;
}

// persistent kernel over one shared-memory bin: CTAs pull cells from the bin's list
template <int BIN>
__global__ void __launch_bounds__(bin_threads(BIN)) k_resolve_smem(KArgs a) {
  constexpr u32 LOG2CAP = bin_cap_log2(BIN);
  constexpr u32 CAP = 1u << LOG2CAP;
  AFQ_DYN_SMEM(smem_raw);
  u64* keys = reinterpret_cast<u64*>(smem_raw);
  u32* cnts = reinterpret_cast<u32*>(keys + CAP);
  __shared__ CellShared sh;
  const u32 count = a.ctl->bin_count[BIN];
  const u32* list = a.bin_list + (u64)BIN * a.n_cells;
  for (;;) {
    if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[BIN], 1u);
    __syncthreads();
    const u32 job = sh.job;
    if (job >= count) break;
    const u32 cell = list[job];
    const bool ok = resolve_cell(a, cell, keys, cnts, LOG2CAP, (CAP / 4) * 3, &sh);
    if (!ok && threadIdx.x == 0) {
      const u32 idx = atomicAdd(&a.ctl->bin_count[BIN + 1], 1u);
      a.bin_list[(u64)(BIN + 1) * a.n_cells + idx] = cell;
      if (BIN + 1 == NUM_SMEM_BINS) {
        const u64 r0 = a.cell_rec_off[cell], r1 = a.cell_rec_off[cell + 1];
        atomicMax(&a.ctl->max_cell_refs, a.ref_off[r1] - a.ref_off[r0]);
      }
    }
    __syncthreads();
  }
}

// giant cells: same algorithm on a per-CTA global-memory arena (L2-resident working set)
__global__ void __launch_bounds__(1024) k_resolve_large(KArgs a) {
  __shared__ CellShared sh;
  const u64 arena = (u64)blockIdx.x << a.large_cap_log2;
  u64* keys = a.large_keys + arena;
  u32* cnts = a.large_cnts + arena;
  const u32 count = a.ctl->bin_count[NUM_SMEM_BINS];
  const u32* list = a.bin_list + (u64)NUM_SMEM_BINS * a.n_cells;
  for (;;) {
    if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[NUM_SMEM_BINS], 1u);
    __syncthreads();
    const u32 job = sh.job;
    if (job >= count) break;
    const u32 cell = list[job];
    const u64 r0 = a.cell_rec_off[cell], r1 = a.cell_rec_off[cell + 1];
    const u32 p = a.ref_off[r1] - a.ref_off[r0];
    // arena sized so the table can never fill: cap >= 2 * (upper bound on pairs)
    u32 log2cap = 10;
    while (log2cap < a.large_cap_log2 && (1ull << log2cap) < 2ull * p) ++log2cap;
    if ((1ull << log2cap) < 2ull * p) {
      if (threadIdx.x == 0) {
        atomicOr(&a.ctl->error, (u32)DEV_ERR_CELL_TOO_LARGE);
        a.sum_umi[cell] = 0; a.max_umi[cell] = 0; a.num_expr[cell] = 0;
        a.num_over_mean[cell] = 0; a.flags[cell] = 4;
      }
      __syncthreads();
      continue;
    }
    resolve_cell(a, cell, keys, cnts, log2cap, 0xFFFFFFFFu, &sh);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// CSR assembly: row_ptr = exclusive scan(num_expr); gather staging rows to final CSR
// ------------------------------------------------------------------------------------
constexpr u32 SCAN_TILE = 4096;  // cells per scan tile (1024 threads x 4)

__global__ void __launch_bounds__(1024) k_scan_tile_sums(const u32* num_expr, u64 n, u64* tile_sums) {
  __shared__ u32 s_warp[40];
  const u64 t0 = (u64)blockIdx.x * SCAN_TILE;
  u32 v = 0;
  for (u32 k = 0; k < 4; ++k) {
    const u64 i = t0 + (u64)k * 1024 + threadIdx.x;
    if (i < n) v += num_expr[i];
  }
  u32 tot;
  block_exscan(v, s_warp, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_tiles(u64* tile_sums, u32 n_tiles, u64* row_ptr, u64 n) {
  // single CTA: exclusive scan of tile sums in place (u64 running carry)
  __shared__ u64 s_part[1024];
  __shared__ u64 s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (u32 c0 = 0; c0 < n_tiles; c0 += 1024) {
    const u32 i = c0 + threadIdx.x;
    const u64 v = i < n_tiles ? tile_sums[i] : 0;
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (u32 o = 1; o < 1024; o <<= 1) {
      u64 t = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
      __syncthreads();
      s_part[threadIdx.x] += t;
      __syncthreads();
    }
    const u64 inc = s_part[threadIdx.x];
    if (i < n_tiles) tile_sums[i] = s_carry + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry += inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) row_ptr[n] = s_carry;
}

__global__ void __launch_bounds__(1024) k_scan_rows(const u32* num_expr, u64 n, const u64* tile_sums,
                                                    u64* row_ptr) {
  __shared__ u32 s_warp[40];
  const u64 t0 = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * 4;
  u32 v[4];
  u32 s = 0;
#pragma unroll
  for (u32 k = 0; k < 4; ++k) { v[k] = (t0 + k < n) ? num_expr[t0 + k] : 0; s += v[k]; }
  u32 tot;
  u32 ex = block_exscan(s, s_warp, &tot);
  u64 run = tile_sums[blockIdx.x] + ex;
#pragma unroll
  for (u32 k = 0; k < 4; ++k) {
    if (t0 + k < n) row_ptr[t0 + k] = run;
    run += v[k];
  }
}

// one warp per cell copies its staging row to its CSR row
__global__ void __launch_bounds__(256) k_gather_rows(KArgs a, const u64* row_ptr, u32* col, float* val) {
  const u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= a.n_cells) return;
  const u32 lane = threadIdx.x & 31;
  const u32 nnz = a.num_expr[w];
  const u64 src = a.ref_off[a.cell_rec_off[w]];
  const u64 dst = row_ptr[w];
  for (u32 i = lane; i < nnz; i += 32) {
    col[dst + i] = a.stage_col[src + i];
    val[dst + i] = a.stage_val[src + i];
  }
}

}  // namespace afq
