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
  return b == 0 ? 32 : (b == 1 ? 128 : (b == 2 ? 256 : (b == 3 ? 512 : (b == 4 ? 512 : 1024))));
}

enum : u32 { MODE_CRLIKE = 0, MODE_TRIVIAL = 1 };
enum : u32 { DEV_ERR_CELL_TOO_LARGE = 1, DEV_ERR_POOL = 32 };   // (2..16: afq_pug.cuh)


constexpr int NUM_LISTS = NUM_BINS + 7;          // + two k_gene_eqc lists (big / normal cells) + arena-overflow list + four k_pug_smem lists
constexpr int OVF_LIST = NUM_BINS + 2;           // cells whose distinct pairs overflowed their shared-memory arena
struct Ctl {                       // per-batch device control block (zeroed per batch)
  u32 bin_count[NUM_LISTS + 1];
  u32 bin_cursor[NUM_LISTS + 1];
  u32 error;
  u32 max_cell_refs;
  u32 ge_max_n[2];                 // largest record / alignment count on the k_gene_eqc lists
  u32 ge_max_p[2];
  unsigned long long adj_used;     // bump pointer into the adjacency pool
  u32 ps3_max_n, ps3_max_p;        // largest cell on k_pug_smem<3>'s list (sizes its global arenas)
  u32 desc_count[4];               // split path: component descriptors on the lists of sizes 2 | 3-4 | 5-8 | 9-32
  u32 desc_cursor[4];              // ... and the cover kernels' work cursors
  u32 count_cursor;                // k_pug_count's / k_pug_back's work cursor over the cells of the four k_pug_build lists
  u32 back_count[4], back_cursor[4];     // k_pug_back's arena tiers: cells / work cursors
  u32 em_count[4], em_cursor[4];         // k_em_cells' arena tiers
  // per-CTA global arenas, planned ON THE DEVICE from the batch's maxima above (k_plan_arenas, afq_pipeline.cuh) so that the
  // host never reads this block back between the binning and the kernels: words / bytes per CTA and the CTAs that have one
  u32 ps3_words, ps3_blocks;             // k_pug_build<3> / k_pug_smem<3>
  u32 back_words, back_blocks;           // k_pug_back<3> / k_em_cells<3>
  u32 ge_blocks[2];                      // k_gene_eqc (big / normal list)
  unsigned long long ge_bytes[2];
};

// a CTA of a global-arena kernel beyond the CTAs its pool holds arenas for leaves; if the pool holds none at all and there is
// work, the batch is flagged (the host API grows the pool and runs the batch again)
__device__ __forceinline__ bool arena_cta_idle(Ctl* ctl, u32 blocks, u32 work) {
  if (blockIdx.x < blocks) return false;
  if (blocks == 0 && blockIdx.x == 0 && threadIdx.x == 0 && work) atomicOr(&ctl->error, (u32)DEV_ERR_POOL);
  return true;
}

struct KArgs {
  // input batch (device)
  u64 n_cells, n_records, n_refs_total;
  const u64* cell_rec_off;
  const u32* umi;
  const u32* ref_off;
  const u32* refs;
  const u32* t2g;
  // config
  u32 mode, usa_mode, num_rows, uo, ao;
  u64 small_thresh;
  u32 tiny_eligible;
  u32 prefer_ambig;                // --sa-model prefer-ambig (src/pugutils.rs:505-641)
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
  u32 nwin;
  u32 nbig;
};

// ------------------------------------------------------------------------------------
// classify cells into arena-size bins by record count
// ------------------------------------------------------------------------------------
__global__ void k_bin_cells(KArgs a, int force_bin, u32 need_shift) {
  const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n_cells) return;
  const u64 r0 = a.cell_rec_off[c], r1 = a.cell_rec_off[c + 1];
  const u64 n = r1 - r0;
  const u32 p = a.ref_off[r1] - a.ref_off[r0];
  const u64 need = (n < (u64)p ? n : (u64)p) << need_shift;  // distinct pairs <= refs; typically << records
  int b = NUM_SMEM_BINS;
#pragma unroll
  for (int i = NUM_SMEM_BINS - 1; i >= 0; --i)
    if (need <= (1ull << bin_cap_log2(i))) b = i;
  if (force_bin >= 0 && force_bin > b) b = force_bin < NUM_SMEM_BINS ? force_bin : NUM_SMEM_BINS;
  const u32 idx = atomicAdd(&a.ctl->bin_count[b], 1u);
  a.bin_list[(u64)b * a.n_cells + idx] = (u32)c;
  if (b == NUM_SMEM_BINS) atomicMax(&a.ctl->max_cell_refs, p);
}

// ------------------------------------------------------------------------------------
// Per-cell resolve.
//  * Phase 1 runs one lane per ALIGNMENT (not per record): a head bitmap over the cell's refs plus
//    word ranks give every lane its record; the warp reconverges before the table insert.
//  * One open-address table holds the distinct (umi, gene) keys with read counts; a pair starts
//    probing at home(umi) + (gene mod UMI_WINDOW), so a UMI's genes share a small window (plus the
//    occupied run behind it). The UMI's first entry from home ("leader") scans that window and
//    computes the arg-max gene set directly: no sort of the pairs.
//  * Every new entry's slot is appended to a dense list, so the later phases run over dense work
//    (bins 0-3; the two largest arenas iterate over slots instead to stay within shared memory).
//  * Counting: a presence bitmap over the output slots (in the dead key area) ranks the expressed
//    slots by prefix popcount and each winner bumps its slot's counter — flat, no sort. Cells whose
//    gene axis does not fit the key area order winners by bucket (slot >> shift) with per-bucket
//    presence masks; small cells rank-sort their winners.
//  * The record / entry / bucket loops are warp-uniform with an explicit __syncwarp() so that lanes
//    reconverge every iteration (ncu r1c: 7.5 active lanes/instruction without it). Fully
//    lock-stepped (vote-driven) probe loops were measured and are slower (r1e: +70 % warp
//    instructions from the votes and predication), so the probe loops stay plainly divergent.
//  * A lane-per-RECORD phase 1 (scan the record's refs, one insert for the ~85 % single-gene records, the refs of the
//    multi-gene ones queued for a flat second pass) was measured in round 2 and is slower than the flat form: C2 8.91 vs
//    8.18 ms (scripts/gpu_round2x.sh) — the per-record ref loads are uncoalesced and the scan loop diverges.
// ------------------------------------------------------------------------------------
__host__ __device__ constexpr u32 bin_buckets_log2(int b) { return b == 0 ? 7 : (b == 1 ? 8 : (b == 2 ? 9 : 10)); }
__host__ __device__ constexpr bool bin_has_list(int b) { return b <= 3; }
__host__ __device__ constexpr u32 bin_min_blocks(int b) { return b == 0 ? 16 : (b == 1 ? 10 : (b == 2 ? 5 : (b == 3 ? 3 : (b == 4 ? 2 : 1)))); }
__host__ __device__ constexpr size_t bin_smem_bytes(int b) {
  return ((size_t)12 << bin_cap_log2(b)) + (bin_has_list(b) ? ((size_t)3 << bin_cap_log2(b)) : 0) +
         ((size_t)12 << bin_buckets_log2(b)) + 64;
}
constexpr u32 BIG_BUCKET = 64;       // buckets longer than this are counted cooperatively
constexpr u32 RANK_SORT_MAX = 192;   // cells with at most this many winners rank-sort them

struct CellArena {
  u64* keys;   // [cap]   later reused as u32 sorted[2*cap]
  u32* cnts;   // [cap]   read counts -> winner slot (or NONE32)
  u32* list;   // [3*cap/4] slots of the distinct entries in creation order (nullptr: iterate slots)
  u32* bcnt;   // [NB]
  u32* boff;   // [NB+1]
  u32* bdo;    // [NB+1]
  u32 log2cap, nb_log2;
  bool smem;   // keys/cnts live in shared memory (the presence-bitmap form needs that)
};

__device__ __forceinline__ u32 umi_home(u32 umi, u32 log2cap) { return (umi * 0x9E3779B1u) >> (32 - log2cap); }
// A pair starts probing at its UMI's home slot plus a gene-dependent offset inside a small window:
// the genes of one UMI spread over UMI_WINDOW start slots instead of queueing behind each other
// (ncu r1l: a multi-gene record's k-th gene needed k probes, so every warp ran the probe loop
// 6-7 times with ~4 lanes). Invariant after phase 1: an entry of UMI u placed at slot p started
// at home(u)+o (o < UMI_WINDOW) and every slot of [home(u)+o, p] is occupied — so all entries of
// u lie in the fixed window [home, home+UMI_WINDOW) or in the occupied run that continues it.
#ifndef AFQ_UMI_WINDOW
#define AFQ_UMI_WINDOW 8
#endif
constexpr u32 UMI_WINDOW = AFQ_UMI_WINDOW;   // power of two; 1 = plain per-UMI clusters

// insert-or-increment (umi, gene). The probe loop only FINDS the slot; the counter bump and the
// new-entry bookkeeping sit behind the loop, where the warp has reconverged, so they issue once per
// call instead of once per probe round (ncu r1m: they ran ~6 times per warp at 4-6 lanes).
// CONV: every lane of the warp calls (lanes with nothing to insert pass on = false) and the warp is
// explicitly re-joined behind the probe loop.
template <bool LIST, bool CONV = false>
__device__ __forceinline__ void table_insert(const CellArena& A, u32 umi, u32 gene, u32 limit, CellShared* sh, bool on = true) {
  const u32 mask = (1u << A.log2cap) - 1;
  const u64 key = ((u64)umi << 32) | gene;
  u32 s = (umi_home(umi, A.log2cap) + (gene & (UMI_WINDOW - 1))) & mask;
  bool created = false, found = !on;
  if (on) for (u32 probes = 0; probes <= mask; ++probes) {
    u64 cur = A.keys[s];
    if (cur == EMPTY_KEY) {
      cur = atomicCAS((unsigned long long*)&A.keys[s], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
      if (cur == EMPTY_KEY) { created = true; found = true; break; }
    }
    if (cur == key) { found = true; break; }
    s = (s + 1) & mask;
  }
  // (a warp-aggregated claim of the list positions — one atomic per warp instead of one per new entry —
  // was measured slower: C2 8.45 vs 8.26 ms, scripts/gpu_round1zk.sh)
  if (CONV) __syncwarp();
  if (!on) return;
  if (!found) { sh->abort = 1; return; }   // table full (only reachable after the distinct limit was crossed)
  atomicAdd(&A.cnts[s], 1u);
  if (created) {
    const u32 idx = atomicAdd(&sh->distinct, 1u);
    if (idx >= limit) { sh->abort = 1; return; }
    if (LIST) A.list[idx] = s;
  }
}

// --sa-model prefer-ambig (resolve_num_molecules_crlike_from_vec_prefer_ambig, src/pugutils.rs:505-641),
// called by the leader (slot s) of UMI u in phase 2: the ids 2k / 2k+1 (spliced / unspliced of one gene)
// vote TOGETHER; the label lists the ids present of every gene attaining the largest combined count,
// ascending; returns the output slot (or NONE32) and retires the UMI's other entries. The entries are
// re-scanned instead of being copied out, and the function is kept out of line so that the hidden
// option costs the default path neither registers nor stack.
__device__ __noinline__ u32 prefer_ambig_slot(const u64* keys, u32* cnts, u32 log2cap, u32 s, u32 u, u32 usa, u32 uo, u32 ao) {
  const u32 mask = (1u << log2cap) - 1;
  const u32 home = umi_home(u, log2cap);
  auto for_entries = [&](auto f) {
    for (u32 t = s, dist = (s - home) & mask; dist <= mask; t = (t + 1) & mask, ++dist) {
      const u64 kt = keys[t];
      if (kt == EMPTY_KEY) { if (dist >= UMI_WINDOW) break; continue; }
      if ((u32)(kt >> 32) != u) continue;
      f(t, (u32)kt);
    }
  };
  auto group_weight = [&](u32 gene) {
    u32 w = 0;
    for_entries([&](u32 t, u32 g2) { if ((g2 | 1u) == (gene | 1u)) w += cnts[t]; });
    return w;
  };
  u32 maxw = 0;
  for_entries([&](u32, u32 gene) { const u32 w = group_weight(gene); maxw = w > maxw ? w : maxw; });
  u32 lab[11] = {0};
  u32 nl = 0;
  for_entries([&](u32, u32 gene) {
    if (group_weight(gene) != maxw || nl == 11) return;
    u32 q = nl;
    while (q > 0 && lab[q - 1] > gene) { lab[q] = lab[q - 1]; --q; }
    lab[q] = gene;
    ++nl;
  });
  for_entries([&](u32 t, u32) { if (t != s) cnts[t] = NONE32; });
  if (!usa) return nl == 1 ? lab[0] : NONE32;
  return nl <= 10 ? usa_slot_for_label(lab, nl, uo, ao) : NONE32;
}

// Returns false when the distinct-pair count exceeded `limit` (caller re-queues the cell on a
// larger arena); nothing has been written for the cell in that case.
template <bool LIST>
__device__ inline bool resolve_cell(const KArgs& a, u32 cell, const CellArena& A, u32 limit, CellShared* sh) {
  const u32 cap = 1u << A.log2cap, mask = cap - 1;
  const u32 NB = 1u << A.nb_log2;
  const u32 T = blockDim.x, tid = threadIdx.x;
  const u64 r0 = a.cell_rec_off[cell], r1 = a.cell_rec_off[cell + 1];
  const u32 nrec = (u32)(r1 - r0);
  // a tiny cell takes the cr-like fast path whatever -r says (src/quant.rs:780-846)
  const bool trivial = a.mode == MODE_TRIVIAL && !(a.tiny_eligible && (r1 - r0) < a.small_thresh);
  for (u32 i = tid; i < cap; i += T) { A.keys[i] = EMPTY_KEY; A.cnts[i] = 0; }
  if (tid == 0) { sh->distinct = 0; sh->abort = 0; sh->red_max = 0; sh->red_cnt = 0; sh->nwin = 0; sh->nbig = 0; }
  __syncthreads();

  // ---- phase 1: records -> distinct (umi, gene) pairs with read counts --------------------
  // Flat form (one lane per ALIGNMENT, not per record): alignment counts vary from 1 to dozens,
  // so a lane-per-record loop runs at the warp's longest record with most lanes idle (ncu r1d:
  // 9.7 active lanes per instruction). Record heads are marked in a bitmap over the cell's refs,
  // a lane finds its record by rank (prefix popcount) and the record's first ref by the highest
  // head bit at or below it; a ref is inserted iff no earlier ref of its record maps to the same
  // gene (src/pugutils.rs:774-781, 817-829: sorted-dedup gene projection). The bitmap and ranks
  // live in the (still unused) bucket area. Cells with an alignment-free record, the `trivial`
  // resolution and cells whose refs overflow the bitmap take the lane-per-record form below.
  const u64 f0 = a.ref_off[r0];
  const u32 P = (u32)(a.ref_off[r1] - f0);
  const u32 W = (P >> 5) + 1;                       // bitmap words
  u32* hb = A.bcnt;                                 // [W] record-head bits
  u32* hr = A.bcnt + W;                             // [W] heads before word w
  bool flat = !trivial && 2 * W <= 3 * NB && P > 0;
  if (flat) {
    for (u32 i = tid; i < W; i += T) hb[i] = 0;
    if (tid == 0) sh->nbig = 0;                     // "some record has no alignment"
    __syncthreads();
    for (u32 i = tid; i < nrec; i += T) {
      const u32 o0 = a.ref_off[r0 + i] - (u32)f0, o1 = a.ref_off[r0 + i + 1] - (u32)f0;
      if (o1 > o0) atomicOr(&hb[o0 >> 5], 1u << (o0 & 31));
      else sh->nbig = 1;
    }
    __syncthreads();
    flat = sh->nbig == 0;
    u32 run = 0;
    for (u32 c0 = 0; c0 < W; c0 += T) {
      const u32 i = c0 + tid;
      const u32 v = i < W ? (u32)__popc(hb[i]) : 0u;
      u32 tot;
      const u32 ex = block_exscan(v, sh->scan, &tot);
      if (i < W) hr[i] = run + ex;
      run += tot;
      __syncthreads();
    }
    if (tid == 0) sh->nbig = 0;
  }
  if (flat) {
    for (u32 base = 0; base < P; base += T) {
      // every lane runs the same control flow (inactive lanes are predicated off) and the warp
      // reconverges before the insert: lanes leave the dedup loop at different trip counts and
      // would otherwise run the whole probe sequence a few lanes at a time (ncu r1l: 3 lanes)
      const u32 i = base + tid;
      const bool act = i < P;
      bool ins = false;
      u32 g = 0, umi = 0;
      if (act) g = __ldg(a.t2g + a.refs[f0 + i]);
      const u32 lane = tid & 31;
      const u32 g1 = __shfl_up_sync(0xFFFFFFFFu, g, 1);   // gene of ref i-1 (tiles are warp-aligned)
      u32 back = 0;                                       // refs of this record before this one
      bool inv = false;
      if (act) {
        const u32 w = i >> 5;
        const u32 below = hb[w] & (0xFFFFFFFFu >> (31 - (i & 31)));   // heads at or below i in its word
        umi = a.umi[r0 + hr[w] + (u32)__popc(below) - 1];
        u32 start;
        if (below) start = (w << 5) + 31 - (u32)__clz((int)below);
        else { u32 ww = w - 1; while (hb[ww] == 0) --ww; start = (ww << 5) + 31 - (u32)__clz((int)hb[ww]); }
        back = i - start;
        inv = back >= 1 && lane >= 1 && g1 > g;
      }
      // Transcript ids ascend inside a record, and gene ids usually ascend with them: when no lane of
      // the warp sees a gene id drop, a ref repeats an earlier gene of its record iff it repeats its
      // predecessor's. Records that began in an earlier tile, and warps that do see a drop (arbitrary
      // tid -> gid maps), compare against every earlier ref of the record instead.
      const bool ascending = !__any_sync(0xFFFFFFFFu, inv);
      if (act) {
        if (ascending && back <= lane) ins = !(back >= 1 && g1 == g);
        else {
          ins = true;
          for (u32 k = 1; k <= back; ++k)
            if (__ldg(a.t2g + a.refs[f0 + i - k]) == g) { ins = false; break; }
        }
      }
      __syncwarp();
      table_insert<LIST, true>(A, umi, g, limit, sh, ins && !*(volatile u32*)&sh->abort);
    }
  } else
  for (u32 base = 0; base < nrec; base += T) {
    __syncwarp();
    const u32 i = base + tid;
    if (i >= nrec || *(volatile u32*)&sh->abort) continue;
    const u64 r = r0 + i;
    const u32 umi = a.umi[r];
    const u32 o0 = a.ref_off[r], o1 = a.ref_off[r + 1];
    if (o1 == o0) continue;
    const u32 g0 = __ldg(a.t2g + a.refs[o0]);
    if (trivial) {
      // src/pugutils.rs:870-881: a class is multi-gene iff two consecutive refs differ in gene
      bool multi = false;
      for (u32 k = o0 + 1; k < o1; ++k)
        if (__ldg(a.t2g + a.refs[k]) != g0) { multi = true; break; }
      if (!multi) table_insert<LIST>(A, umi, g0, limit, sh);
      continue;
    }
    table_insert<LIST>(A, umi, g0, limit, sh);
    for (u32 k = o0 + 1; k < o1; ++k) {
      const u32 g = __ldg(a.t2g + a.refs[k]);
      if (g == g0) continue;
      bool dup = false;
      for (u32 j = o0 + 1; j < k; ++j)
        if (__ldg(a.t2g + a.refs[j]) == g) { dup = true; break; }
      if (!dup) table_insert<LIST>(A, umi, g, limit, sh);
    }
  }
  __syncthreads();
  if (sh->abort) { __syncthreads(); return false; }
  const u32 d = sh->distinct;
  const u32 N = LIST ? d : cap;   // work items of phases 2-4: list entries or table slots
  for (u32 i = tid; i < NB; i += T) A.bcnt[i] = 0;   // (the bucket area held phase 1's bitmap)

  // ---- phase 2: per UMI (cluster leader) the arg-max gene set -> output slot, into cnts ----
  // A leader only writes cnts of entries of its own UMI and every entry's count is read by exactly
  // one thread (its UMI's leader) before being overwritten, so the phase is race-free.
  for (u32 base = 0; base < N; base += T) {
    __syncwarp();
    const u32 i = base + tid;
    bool leader = false;
    u32 s = 0, u = 0;
    if (i < N) {
      s = LIST ? A.list[i] : i;
      const u64 key = A.keys[s];
      if (!LIST && key == EMPTY_KEY) A.cnts[s] = NONE32;
      else if (trivial) A.cnts[s] = (u32)key;   // every distinct (gene, umi) counts
      else {
        u = (u32)(key >> 32);
        leader = true;
        for (u32 t = umi_home(u, A.log2cap); t != s; t = (t + 1) & mask) {   // leader = first entry of u from home
          const u64 kt = A.keys[t];
          if (kt != EMPTY_KEY && (u32)(kt >> 32) == u) { leader = false; break; }   // retired by its leader
        }
      }
    }
    __syncwarp();   // leaders start their cluster scan together
    if (!leader) continue;
    u32 maxw = 0, nb = 0, b0 = 0;
    u32 best[10];
    const bool usa = a.usa_mode != 0;
    if (a.prefer_ambig) { A.cnts[s] = prefer_ambig_slot(A.keys, A.cnts, A.log2cap, s, u, a.usa_mode, a.uo, a.ao); continue; }   // (out of line: a hidden option)
    for (u32 t = s, dist = (s - umi_home(u, A.log2cap)) & mask; dist <= mask; t = (t + 1) & mask, ++dist) {
      const u64 kt = A.keys[t];
      if (kt == EMPTY_KEY) { if (dist >= UMI_WINDOW) break; continue; }   // holes only inside the window
      if ((u32)(kt >> 32) != u) continue;
      const u32 w = A.cnts[t];
      if (w > maxw) { maxw = w; nb = 1; b0 = (u32)kt; if (usa) best[0] = (u32)kt; }
      else if (w == maxw) { if (usa && nb < 10) best[nb] = (u32)kt; ++nb; }
      if (t != s) A.cnts[t] = NONE32;
    }
    u32 res = NONE32;
    if (!usa) { if (nb == 1) res = b0; }
    else if (nb <= 10) {
      for (u32 q = 1; q < nb; ++q) { const u32 x = best[q]; u32 j = q; while (j > 0 && best[j - 1] > x) { best[j] = best[j - 1]; --j; } best[j] = x; }
      res = usa_slot_for_label(best, nb, a.uo, a.ao);
    }
    A.cnts[s] = res;
  }
  __syncthreads();

  const u64 out_base = a.ref_off[r0];             // this cell's staging rows start at its first ref
  u32 m = 0, nnz = 0, lmax = 0;
  // ---- phases 3-4, presence-bitmap form: one bit per output slot over the dead key area ---------
  // winners set their slot's bit; a prefix popcount over the bitmap words ranks every expressed slot
  // (ascending = the CSR column order), and each winner bumps the counter at its slot's rank.
  // Flat and warp-uniform: every loop runs one lane per winner or per bitmap word. Needs
  // 2 * ceil(num_rows / 32) + distinct words in the key area; other cells use the forms below.
  const u32 Wg = (a.num_rows + 31) >> 5;
  const bool bm_path = A.smem && d > RANK_SORT_MAX && (u64)2 * Wg + d <= (u64)2 * cap;
  if (bm_path) {
    u32* gbm = reinterpret_cast<u32*>(A.keys);     // [Wg] slot presence
    u32* gpre = gbm + Wg;                          // [Wg] expressed slots before word w
    u32* gcnt = gpre + Wg;                         // [nnz] molecules per expressed slot
    for (u32 i = tid; i < Wg; i += T) gbm[i] = 0;
    __syncthreads();
    u32 local_m = 0;
    for (u32 i = tid; i < N; i += T) {
      const u32 v = A.cnts[LIST ? A.list[i] : i];
      if (v != NONE32) { atomicOr(&gbm[v >> 5], 1u << (v & 31)); ++local_m; }
    }
    if (local_m) atomicAdd(&sh->nwin, local_m);
    __syncthreads();
    m = sh->nwin;
    for (u32 c0 = 0; c0 < Wg; c0 += T) {
      const u32 i = c0 + tid;
      const u32 v = i < Wg ? (u32)__popc(gbm[i]) : 0u;
      u32 tot;
      const u32 ex = block_exscan(v, sh->scan, &tot);
      if (i < Wg) gpre[i] = nnz + ex;
      nnz += tot;
      __syncthreads();
    }
    for (u32 i = tid; i < nnz; i += T) gcnt[i] = 0;
    __syncthreads();
    for (u32 i = tid; i < N; i += T) {
      const u32 v = A.cnts[LIST ? A.list[i] : i];
      if (v != NONE32) {
        const u32 rank = gpre[v >> 5] + (u32)__popc(gbm[v >> 5] & ((1u << (v & 31)) - 1u));
        if (atomicAdd(&gcnt[rank], 1u) == 0) a.stage_col[out_base + rank] = v;
      }
    }
    __syncthreads();
    for (u32 j = tid; j < nnz; j += T) {
      const u32 c = gcnt[j];
      a.stage_val[out_base + j] = (float)c;
      lmax = c > lmax ? c : lmax;
    }
  } else {
  // ---- phase 3: count winners; bucket histogram ----------------------------------------------
  const u32 bits = 32 - __clz((int)(a.num_rows > 1 ? a.num_rows - 1 : 1));
  const u32 shift = bits > A.nb_log2 ? bits - A.nb_log2 : 0;   // bucket = slot >> shift < NB
  u32 local_m = 0;
  for (u32 base = 0; base < N; base += T) {
    const u32 i = base + tid;
    if (i >= N) continue;
    const u32 v = A.cnts[LIST ? A.list[i] : i];
    if (v != NONE32) { atomicAdd(&A.bcnt[v >> shift], 1u); ++local_m; }
  }
  if (local_m) atomicAdd(&sh->nwin, local_m);
  __syncthreads();
  m = sh->nwin;
  u32* sorted = reinterpret_cast<u32*>(A.keys);   // winners, over the dead key area
  u32* tmp = sorted + cap;
  const bool rank_path = m <= RANK_SORT_MAX || shift > 8;

  if (rank_path) {
    // ---- small cells (or huge gene axes): order the winners themselves -----------------------
    for (u32 base = 0; base < N; base += T) {   // gather winners (unordered) into tmp
      const u32 i = base + tid;
      if (i >= N) continue;
      const u32 v = A.cnts[LIST ? A.list[i] : i];
      if (v != NONE32) tmp[atomicAdd(&sh->nbig, 1u)] = v;
    }
    __syncthreads();
    if (m <= 1024) {
      for (u32 i = tid; i < m; i += T) {        // rank sort: O(m^2 / T)
        const u32 x = tmp[i];
        u32 rank = 0;
        for (u32 j = 0; j < m; ++j) { const u32 y = tmp[j]; rank += (y < x || (y == x && j < i)) ? 1u : 0u; }
        sorted[rank] = x;
      }
      __syncthreads();
    } else {
      const u32 M = next_pow2(m);
      for (u32 i = tid; i < M; i += T) sorted[i] = i < m ? tmp[i] : NONE32;
      __syncthreads();
      block_bitonic_u32(sorted, M);
    }
    // run starts -> tmp (positions), then (slot, run length)
    u32 base = 0;
    for (u32 c0 = 0; c0 < m; c0 += T) {
      const u32 i = c0 + tid;
      const u32 st = (i < m && (i == 0 || sorted[i - 1] != sorted[i])) ? 1u : 0u;
      u32 tot;
      const u32 pos = block_exscan(st, sh->scan, &tot);
      if (st) tmp[base + pos] = i;
      base += tot;
    }
    nnz = base;
    __syncthreads();
    for (u32 j = tid; j < nnz; j += T) {
      const u32 i0 = tmp[j], i1 = (j + 1 < nnz) ? tmp[j + 1] : m;
      a.stage_col[out_base + j] = sorted[i0];
      a.stage_val[out_base + j] = (float)(i1 - i0);
      lmax = (i1 - i0) > lmax ? (i1 - i0) : lmax;
    }
  } else {
    // ---- bucket path ---------------------------------------------------------------------------
    block_exscan_array_small(A.bcnt, A.boff, NB, sh->scan);   // boff[b] = start of bucket b; boff[NB] = m
    for (u32 base = 0; base < N; base += T) {   // scatter winners into their bucket segments
      const u32 i = base + tid;
      if (i >= N) continue;
      const u32 v = A.cnts[LIST ? A.list[i] : i];
      if (v != NONE32) { const u32 b = v >> shift; sorted[A.boff[b] + atomicSub(&A.bcnt[b], 1u) - 1] = v; }
    }
    __syncthreads();
    const u32 lowmask = (1u << shift) - 1;
    // distinct slots per bucket from a presence mask over the low `shift` (<= 8) bits
    for (u32 base = 0; base < NB; base += T) {
      __syncwarp();
      const u32 b = base + tid;
      if (b >= NB) continue;
      const u32 lo = A.boff[b], hi = A.boff[b + 1];
      u32 nd = 0;
      if (hi - lo > BIG_BUCKET) { tmp[atomicAdd(&sh->nbig, 1u)] = b; }
      else if (hi > lo) {
        unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
        for (u32 i = lo; i < hi; ++i) {
          const u32 x = sorted[i] & lowmask;
          const unsigned long long bit = 1ull << (x & 63);
          const u32 w = x >> 6;
          if (w == 0) w0 |= bit; else if (w == 1) w1 |= bit; else if (w == 2) w2 |= bit; else w3 |= bit;
        }
        nd = (u32)(__popcll(w0) + __popcll(w1) + __popcll(w2) + __popcll(w3));
      }
      A.bcnt[b] = nd;
    }
    __syncthreads();
    const u32 nbig = sh->nbig;
    u32* bigcnt = tmp + NB;                      // [256] counters for one hot bucket at a time
    for (u32 q = 0; q < nbig; ++q) {
      const u32 b = tmp[q];
      const u32 lo = A.boff[b], hi = A.boff[b + 1];
      for (u32 i = tid; i < 256; i += T) bigcnt[i] = 0;
      if (tid == 0) sh->red_cnt = 0;
      __syncthreads();
      for (u32 i = lo + tid; i < hi; i += T) atomicAdd(&bigcnt[sorted[i] & lowmask], 1u);
      __syncthreads();
      u32 c = 0;
      for (u32 i = tid; i < 256; i += T) if (bigcnt[i]) ++c;
      if (c) atomicAdd(&sh->red_cnt, c);
      __syncthreads();
      if (tid == 0) A.bcnt[b] = sh->red_cnt;
      __syncthreads();
    }
    block_exscan_array_small(A.bcnt, A.bdo, NB, sh->scan);
    nnz = A.bdo[NB];
    // emit small buckets: one thread per bucket walks its mask in ascending order
    for (u32 base = 0; base < NB; base += T) {
      __syncwarp();
      const u32 b = base + tid;
      if (b >= NB) continue;
      const u32 lo = A.boff[b], hi = A.boff[b + 1];
      if (hi == lo || hi - lo > BIG_BUCKET) continue;
      unsigned long long wm[4] = {0, 0, 0, 0};
      for (u32 i = lo; i < hi; ++i) { const u32 x = sorted[i] & lowmask; wm[x >> 6] |= 1ull << (x & 63); }
      u64 o = out_base + A.bdo[b];
      for (u32 w = 0; w < 4; ++w) {
        unsigned long long mm = wm[w];
        while (mm) {
          const u32 low = w * 64 + (u32)__ffsll((long long)mm) - 1;
          mm &= mm - 1;
          u32 cnt = 0;
          for (u32 i = lo; i < hi; ++i) cnt += ((sorted[i] & lowmask) == low) ? 1u : 0u;
          a.stage_col[o] = (b << shift) | low;
          a.stage_val[o] = (float)cnt;
          lmax = cnt > lmax ? cnt : lmax;
          ++o;
        }
      }
    }
    for (u32 q = 0; q < nbig; ++q) {             // emit hot buckets cooperatively
      const u32 b = tmp[q];
      const u32 lo = A.boff[b], hi = A.boff[b + 1];
      __syncthreads();
      for (u32 i = tid; i < 256; i += T) bigcnt[i] = 0;
      __syncthreads();
      for (u32 i = lo + tid; i < hi; i += T) atomicAdd(&bigcnt[sorted[i] & lowmask], 1u);
      __syncthreads();
      if (tid == 0) {
        u64 o = out_base + A.bdo[b];
        for (u32 low = 0; low < 256; ++low) {
          const u32 cnt = bigcnt[low];
          if (!cnt) continue;
          a.stage_col[o] = (b << shift) | low;
          a.stage_val[o] = (float)cnt;
          lmax = cnt > lmax ? cnt : lmax;
          ++o;
        }
      }
    }
  }
  }  // !bm_path
  if (tid == 0) sh->red_cnt = 0;
  if (lmax) atomicMax(&sh->red_max, lmax);
  __syncthreads();
  // ---- statistics: NumGenesOverMean (src/quant.rs:1190-1194), mean over expressed genes, f32 ---
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
    if (a.tiny_eligible && (r1 - r0) < a.small_thresh) f |= 1;  // AFQ_FLAG_TINY
    if (nnz == 0) f |= 4;                                        // AFQ_FLAG_EMPTY
    a.flags[cell] = f;
  }
  __syncthreads();
  return true;
}

// L2 prefetch of a cell's record arrays (one 128-byte line per thread and step)
__device__ __forceinline__ void prefetch_lines_l2(const void* lo, const void* hi) {
#ifndef AFQ_EMU
  const char* p = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(lo) & ~uintptr_t(127)) + (size_t)threadIdx.x * 128;
  for (; p < reinterpret_cast<const char*>(hi); p += (size_t)blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)lo; (void)hi;
#endif
}
__device__ __forceinline__ void prefetch_cell_l2(const KArgs& a, u32 cell) {
  const u64 r0 = a.cell_rec_off[cell], r1 = a.cell_rec_off[cell + 1];
  const u32 f0 = a.ref_off[r0], f1 = a.ref_off[r1];
  prefetch_lines_l2(a.refs + f0, a.refs + f1);
  prefetch_lines_l2(a.umi + r0, a.umi + r1);
  prefetch_lines_l2(a.ref_off + r0, a.ref_off + r1 + 1);
}

// persistent kernel over one shared-memory bin: CTAs pull cells from the bin's list
template <int BIN>
__global__ void __launch_bounds__(bin_threads(BIN), bin_min_blocks(BIN)) k_resolve_smem(KArgs a) {
  constexpr u32 LOG2CAP = bin_cap_log2(BIN);
  constexpr u32 CAP = 1u << LOG2CAP;
  constexpr u32 NB = 1u << bin_buckets_log2(BIN);
  constexpr bool LIST = bin_has_list(BIN);
  AFQ_DYN_SMEM(smem_raw);
  CellArena A;
  A.keys = reinterpret_cast<u64*>(smem_raw);
  A.cnts = reinterpret_cast<u32*>(A.keys + CAP);
  A.list = LIST ? A.cnts + CAP : nullptr;
  A.bcnt = A.cnts + CAP + (LIST ? (CAP / 4) * 3 : 0);
  A.boff = A.bcnt + NB;
  A.bdo = A.boff + NB + 1;
  A.log2cap = LOG2CAP;
  A.nb_log2 = bin_buckets_log2(BIN);
  A.smem = true;
  __shared__ CellShared sh;
  const u32 count = a.ctl->bin_count[BIN];
  const u32* list = a.bin_list + (u64)BIN * a.n_cells;
#ifndef AFQ_NO_RESOLVE_PREFETCH
  // jobs are claimed one ahead: the next cell's records are prefetched into L2 while this one is resolved
  // (A/B on one box, scripts/gpu_round1zj.sh: C2 8.39 -> 8.21 ms per 100 k cells)
  if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[BIN], 1u);
  __syncthreads();
  u32 job = sh.job;
  __syncthreads();
  for (;;) {
    if (job >= count) break;
    if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[BIN], 1u);
    __syncthreads();
    const u32 next = sh.job;
    if (next < count) prefetch_cell_l2(a, list[next]);
    const u32 cell = list[job];
    job = next;
#else
  for (;;) {
    if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[BIN], 1u);
    __syncthreads();
    const u32 job = sh.job;
    if (job >= count) break;
    const u32 cell = list[job];
#endif
    const bool ok = resolve_cell<LIST>(a, cell, A, (CAP / 4) * 3, &sh);
    if (!ok && threadIdx.x == 0) {
      // distinct pairs exceeded 75 % of this arena: hand the cell to the global-arena kernel
      // (second k_resolve_large launch, after all arenas have drained)
      const u32 idx = atomicAdd(&a.ctl->bin_count[OVF_LIST], 1u);
      a.bin_list[(u64)OVF_LIST * a.n_cells + idx] = cell;
      const u64 r0 = a.cell_rec_off[cell], r1 = a.cell_rec_off[cell + 1];
      atomicMax(&a.ctl->max_cell_refs, a.ref_off[r1] - a.ref_off[r0]);
    }
    __syncthreads();
  }
}

// giant cells: same algorithm on a per-CTA global-memory arena (L2-resident working set)
__global__ void __launch_bounds__(1024) k_resolve_large(KArgs a, u32 list_id) {
  __shared__ CellShared sh;
  __shared__ u32 s_buckets[3 * 1024 + 2];
  const u64 arena = (u64)blockIdx.x << a.large_cap_log2;
  CellArena A;
  A.keys = a.large_keys + arena;
  A.cnts = a.large_cnts + arena;
  A.list = nullptr;
  A.bcnt = s_buckets;
  A.boff = A.bcnt + 1024;
  A.bdo = A.boff + 1025;
  A.nb_log2 = 10;
  A.smem = false;
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
    A.log2cap = log2cap;
    resolve_cell<false>(a, cell, A, 0xFFFFFFFFu, &sh);
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

// ---- CSR offsets from per-record alignment counts (afq_batch.rec_na8) -------------------------
__global__ void __launch_bounds__(1024) k_na_tile_sums(const u8* na, u64 n, u64* tile_sums) {
  __shared__ u32 s_warp[40];
  const u64 t0 = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * 4;
  u32 v = 0;
#pragma unroll
  for (u32 k = 0; k < 4; ++k) if (t0 + k < n) v += na[t0 + k];
  u32 tot;
  block_exscan(v, s_warp, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_na_offsets(const u8* na, u64 n, const u64* tile_sums, u32* ref_off) {
  __shared__ u32 s_warp[40];
  const u64 t0 = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * 4;
  u32 v[4];
  u32 s = 0;
#pragma unroll
  for (u32 k = 0; k < 4; ++k) { v[k] = (t0 + k < n) ? na[t0 + k] : 0; s += v[k]; }
  u32 tot;
  const u32 ex = block_exscan(s, s_warp, &tot);
  u64 run = tile_sums[blockIdx.x] + ex;
#pragma unroll
  for (u32 k = 0; k < 4; ++k) {
    if (t0 + k <= n) ref_off[t0 + k] = (u32)run;   // index n receives the closing offset
    run += v[k];
  }
}

// ---- 24-bit packed wire arrays (afq_batch.rec_umi24 / refs24) -> u32 in HBM ------------------
// One thread widens 4 values: three aligned 32-bit loads in, one 128-bit store out.
__global__ void __launch_bounds__(256) k_unpack24(const u8* src, u64 n, u32* dst) {
  const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 values
  const u64 i0 = g * 4;
  if (i0 >= n) return;
  if (i0 + 4 <= n) {
    const u32* w = reinterpret_cast<const u32*>(src) + g * 3;
    const u32 a = w[0], b = w[1], c = w[2];
    uint4 o;
    o.x = a & 0xFFFFFFu;
    o.y = (a >> 24) | ((b & 0xFFFFu) << 8);
    o.z = (b >> 16) | ((c & 0xFFu) << 16);
    o.w = c >> 8;
    *reinterpret_cast<uint4*>(dst + i0) = o;
  } else {
    for (u64 i = i0; i < n; ++i) dst[i] = (u32)src[3 * i] | ((u32)src[3 * i + 1] << 8) | ((u32)src[3 * i + 2] << 16);
  }
}

// --dump-eqclasses: one warp per cell compacts its classes (counts, label offsets, labels) from the per-cell regions
// of the dump arrays into the batch-wide CSR-of-CSR (afq_eqc_dump)
__global__ void __launch_bounds__(256) k_dump_gather(KArgs a, const u32* ncls, const u32* nlab, const u32* dcnt, const u32* doff, const u32* dlab,
                                                     const u64* cls_ptr, const u64* lab_base, u64* cls_lab_ptr, u32* labels, u32* counts) {
  const u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= a.n_cells) return;
  const u32 lane = threadIdx.x & 31;
  const u64 r0 = a.cell_rec_off[w];
  const u64 f0 = a.ref_off[r0];
  const u64 k0 = cls_ptr[w], l0 = lab_base[w];
  for (u32 j = lane; j < ncls[w]; j += 32) { counts[k0 + j] = dcnt[r0 + j]; cls_lab_ptr[k0 + j] = l0 + doff[r0 + j]; }
  for (u32 q = lane; q < nlab[w]; q += 32) labels[l0 + q] = dlab[f0 + q];
  if (w + 1 == a.n_cells && lane == 0) cls_lab_ptr[cls_ptr[w + 1]] = lab_base[w + 1];
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
