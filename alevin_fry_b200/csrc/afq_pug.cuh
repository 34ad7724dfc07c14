// afq_pug.cuh — per-cell gene-level eq-class construction for the resolutions that need it
// (cr-like-em, parsimony[-gene][-em]) and the per-cell EM, one CTA per cell.
//
// Reference behaviour implemented here (paths relative to /root/reference):
//   * transcript / gene level eq-classes   src/eq_class.rs:723-1036  (EqMap::init_from_chunk*)
//   * PUG construction                     src/pugutils.rs:65-267    (extract_graph, has_edge)
//   * weakly connected components          src/pugutils.rs:278-301
//   * greedy monochromatic cover           src/pugutils.rs:308-391, 989-1330
//   * large-component cr-like fallback     src/pugutils.rs:916-982, 1055-1072
//   * cr-like molecules (for cr-like-em)   src/pugutils.rs:644-850
//   * counting from gene eq-classes        src/em.rs:487-514 (only_unique), src/utils.rs:673-756
//   * EM                                   src/em.rs:167-582 (M1 gene mode, M2 USA subset form)
//
// Determinism contract: where the reference follows ahash iteration order we use
// the same canonical orders as the oracle — classes lexicographic by label, cover-loop start
// vertices ascending in (class label, UMI), EM accumulation in class order — so integer
// results are bit-exact and EM sums are performed in the same order with the same roundings.
//
// Data layout: every working array lives in a per-CTA arena (global memory, L2-resident;
// generic pointers so a shared-memory arena can be substituted later). The PUG is never
// stored as a matrix: UMI -> vertex chains in an open-address table give the 1-Hamming
// neighbourhood by enumerating the 3*umi_len substitutions (+ the UMI itself).
#pragma once
#include "afq_kernels.cuh"

namespace afq {

enum : u32 { GE_MODE_PUG_TXP = 0, GE_MODE_PUG_GENE = 1, GE_MODE_CRLIKE = 2 };
enum : u32 { DEV_ERR_ADJ_POOL = 2, DEV_ERR_ARENA = 4, DEV_ERR_HASH = 8, DEV_ERR_LABEL = 16 };

constexpr u32 GE_THREADS = 256;

// CTA-strided loop with a warp-uniform trip count and a __syncwarp() per iteration: lanes that took a
// long path in the body re-join their warp before the next element instead of drifting apart for the
// rest of the phase (ncu r1f: loop headers of the divergent phases ran with ~5 of 32 lanes).
#define GE_FOR(X, N) \
  for (u32 X##_b = 0, X = threadIdx.x; (__syncwarp(), X##_b < (N)); X##_b += blockDim.x, X += blockDim.x) \
    if (X < (N))
constexpr u32 SMALL_COMP = 32;      // components up to this size are covered by one thread
constexpr u32 MAX_GRAPH_THRESH = 4096;

struct GeArgs {
  u32 ge_mode;          // GE_MODE_*
  u32 only_unique;      // 1: counts from singleton classes (no EM)
  u32 em_init_uniform;
  u32 pug_exact_umi;
  u32 umi_len;
  u32 large_graph_thresh;
  u32 num_alphas;       // G (gene mode) or 3G (USA)
  // per-CTA arenas
  u8* arena;
  // adjacency pool (bump-allocated per cell, reset per batch)
  u32* adj_pool;
  u64 adj_cap;
  u64* adj_used;        // in Ctl-adjacent memory
  // which work list this launch consumes
  u32 list_id;
  u32 list_which;       // 0: big list, 1: normal list (index of the arena plan in Ctl)
  u32* ps_garena;       // k_pug_smem<3> / k_pug_build<3>: pool of per-CTA global-memory arenas (Ctl::ps3_words words each)
  u32 ps_limit_words;   // k_pug_smem: use at most this many arena words (0 = the variant's size; tests force fallbacks with it)
  // split path (k_pug_build -> k_pug_cover* -> k_pug_count, afq_pugc.cuh): per-batch global buffers
  u32* ps_win;          // [n_records]   winners (output slots) of cell c at [r0, r0 + ps_nwin[c])
  u32* ps_nwin;         // [n_cells]     molecules emitted so far (NONE32: the cell was handed back)
  u32* ps_mem;          // [4 * n_records] exported members of multi-vertex components: umi, reads << 16 | label length, label offset, gene hint
  u32* ps_desc;         // [2 * ...]     component descriptors: first member, size << 24 | cell; four size-class lists
  u32* ps_glab;         // [n_refs_total] gene-level labels (parsimony-gene); transcript-level labels are read from the input refs
  u32 ps_desc_base[4];  // first descriptor slot of the lists of sizes 2 | 3-4 | 5-8 | 9-32
  // ... EM resolutions on the split path: the cells' molecules (sorted gene labels) for k_pug_back
  u32* ps_mlab;         // [n_refs_total] labels of cell c from f0 on (a molecule's label is no longer than one of its records')
  u32* ps_nlab;         // [n_cells]      label words used
  u32* ps_moff;         // [n_records]    per molecule (cell region from r0 on): label offset inside the cell ...
  u32* ps_mlen;         // [n_records]    ... and length; ps_nwin[c] = molecules of the cell
  u32* back_list;       // [4 * n_cells]  the cells of k_pug_back's four arena tiers (k_back_bin)
  u32* back_garena;     // pool of the per-CTA global arenas of tier 3 (Ctl::back_words words each)
  u32 back_max_tier;    // largest shared-memory tier k_back_bin may choose (0..2)
  u32 classes_only;     // ge_back stops behind stage B (classes written to the dump regions): k_em_cells does stage C
  // --dump-eqclasses (src/quant.rs:1282-1307): every cell's gene eq-classes in canonical order. Cell c (records
  // [r0, r0+n), alignments [f0, f0+P)) writes class j's count / label offset at r0 + j and its labels from f0 on.
  u32* dump_ncls;       // [n_cells]  classes of the cell (zeroed per batch: tiny cells never build gene_eqc)
  u32* dump_nlab;       // [n_cells]  label words of the cell
  u32* dump_cnt;        // [n_records]
  u32* dump_off;        // [n_records] label offset of class j inside the cell
  u32* dump_lab;        // [n_refs_total]
};

__host__ __device__ inline u64 align8(u64 x) { return (x + 7) & ~7ull; }
__host__ __device__ inline u32 pow2_ge(u32 v, u32 lo) {
  u32 p = lo;
  while (p < v) p <<= 1;
  return p;
}

struct GePtrs {
  // sizes
  u32 n, P, N2, N1, P2, L3;
  // class table / UMI table (time-shared)
  u64* ctab_h; u32* ctab_r;
  // vertex table -> dense sorted vertices
  u64* vtab_k; u32* vtab_c;
  u32* rec_slot;
  u32* glab; u32* glen;
  u32* cls_rep; u32* cls_aux; u32* cls_vfirst;
  u32* cls_loff; u32* cls_llen;   // per class rank: label offset (into refs / glab) and length
  u32* vnext; u32* adj_off; u32* parent;
  u64* ckey;
  u32* cstart; u32* cbig;
  u32* vlab_off;
  u32* mol_off; u32* mol_len;
  u32* mlab;          // [2P]
  u64* mkey; u32* midx;
  u32* gcls_m; u32* gcls_cnt; u32* gcls_eoff;
  // large-component pair table
  u64* ltab_k; u32* ltab_c;
  // EM
  u32* ent_idx; u32* ent_loc; u64* tkey; u32* sup; u32* g_off;
  float* alpha_in; float* alpha_out; float* cls_inv;
  u32* sib_a; u32* sib_b;
  // mid-size component bitsets
  u32* adjm; u32* bfs; u32* cmem;
  u32 rows, W;
};

// Carve the arena for a cell with n records and P alignments. With base == nullptr only the
// total size is computed (host side sizing uses the batch maxima; the layout is monotone).
__host__ __device__ inline u64 ge_carve(u8* base, u32 n, u32 P, u32 thresh, GePtrs* o) {
  u64 off = 0;
  auto take = [&](u64 bytes) { u64 r = off; off = align8(off + bytes); return base ? base + r : (u8*)nullptr; };
  GePtrs p{};
  if (P < n) P = n;  // records without alignments still occupy table slots
  p.n = n; p.P = P;
  p.N2 = pow2_ge(2 * n + 2, 64);
  p.N1 = pow2_ge(n + 1, 64);
  p.P2 = pow2_ge(2 * P + 2, 64);
  p.L3 = pow2_ge(3 * P + 3, 64);
  p.ctab_h = (u64*)take(8ull * p.N2); p.ctab_r = (u32*)take(4ull * p.N2);
  p.vtab_k = (u64*)take(8ull * p.N2); p.vtab_c = (u32*)take(4ull * p.N2);
  p.rec_slot = (u32*)take(4ull * n);
  p.glab = (u32*)take(4ull * P); p.glen = (u32*)take(4ull * n);
  p.cls_rep = (u32*)take(4ull * p.N1); p.cls_aux = (u32*)take(4ull * p.N1); p.cls_vfirst = (u32*)take(4ull * (n + 2));
  p.cls_loff = (u32*)take(4ull * p.N1); p.cls_llen = (u32*)take(4ull * p.N1);
  p.vnext = (u32*)take(4ull * n); p.adj_off = (u32*)take(4ull * (n + 2)); p.parent = (u32*)take(4ull * n);
  p.ckey = (u64*)take(8ull * p.N1);
  p.cstart = (u32*)take(4ull * (n + 1)); p.cbig = (u32*)take(4ull * (n + 1));
  p.vlab_off = (u32*)take(4ull * (n + 2));
  p.mol_off = (u32*)take(4ull * (n + 1)); p.mol_len = (u32*)take(4ull * (n + 1));
  p.mlab = (u32*)take(8ull * P + 8);
  p.mkey = (u64*)take(8ull * p.N1); p.midx = (u32*)take(4ull * p.N1);
  p.gcls_m = (u32*)take(4ull * (n + 1)); p.gcls_cnt = (u32*)take(4ull * (n + 1)); p.gcls_eoff = (u32*)take(4ull * (n + 2));
  p.ltab_k = (u64*)take(8ull * p.P2); p.ltab_c = (u32*)take(4ull * p.P2);
  p.ent_idx = (u32*)take(4ull * P + 4); p.ent_loc = (u32*)take(4ull * P + 4);
  p.tkey = (u64*)take(8ull * p.P2);
  p.sup = (u32*)take(4ull * p.L3); p.g_off = (u32*)take(4ull * p.L3 + 8);
  p.alpha_in = (float*)take(4ull * p.L3); p.alpha_out = (float*)take(4ull * p.L3); p.cls_inv = (float*)take(4ull * (n + 1));
  p.sib_a = (u32*)take(4ull * p.L3); p.sib_b = (u32*)take(4ull * p.L3);
  p.rows = n < thresh ? n : thresh;
  if (p.rows < SMALL_COMP) p.rows = SMALL_COMP;
  p.W = (p.rows + 31) / 32;
  p.adjm = (u32*)take(4ull * p.rows * p.W);
  p.bfs = (u32*)take(4ull * 3 * p.W * GE_THREADS + 4ull * 4 * p.W);
  p.cmem = (u32*)take(4ull * p.rows + 4ull * P + 16);
  if (o) *o = p;
  return off;
}

struct GeShared {
  u32 scan[40];
  u32 job;
  u32 cnt0, cnt1, cnt2, cnt3;
  u32 flag;
  u32 best_size, best_i, best_k;
  u32 n_mol;
  u32 lab_bump;     // bump pointer into the upper half of mlab
  u32 alt;
  u32 adj_base_lo, adj_base_hi;
  u32 need_lo;
  float fsum, fmax;
};

// ---- small helpers ---------------------------------------------------------------------------
__device__ __forceinline__ u64 mix64(u64 x) {
  x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
  return x;
}
__device__ inline u64 label_hash(const u32* lab, u32 len, u32 seed) {
  u64 h = mix64(0x9E3779B97F4A7C15ull * (len + 1) + seed);
  for (u32 i = 0; i < len; ++i) h = mix64(h ^ (lab[i] + 0x9E3779B97F4A7C15ull * (i + 1)));
  if (h == EMPTY_KEY) h = 0x1234567ull;
  return h;
}
__device__ inline bool label_equal(const u32* a, u32 la, const u32* b, u32 lb) {
  if (la != lb) return false;
  for (u32 i = 0; i < la; ++i) if (a[i] != b[i]) return false;
  return true;
}
// lexicographic (std::vector operator<)
__device__ inline bool label_less(const u32* a, u32 la, const u32* b, u32 lb) {
  const u32 m = la < lb ? la : lb;
  for (u32 i = 0; i < m; ++i) { if (a[i] != b[i]) return a[i] < b[i]; }
  return la < lb;
}
__device__ inline bool sorted_contains(const u32* lab, u32 len, u32 x) {
  u32 lo = 0, hi = len;
  while (lo < hi) { u32 mid = (lo + hi) >> 1; if (lab[mid] < x) lo = mid + 1; else hi = mid; }
  return lo < len && lab[lo] == x;
}
__device__ inline bool sorted_share(const u32* a, u32 la, const u32* b, u32 lb) {
  u32 i = 0, j = 0;
  while (i < la && j < lb) { if (a[i] == b[j]) return true; if (a[i] < b[j]) ++i; else ++j; }
  return false;
}
// insertion sort + dedup of a short list in place; returns new length
__device__ inline u32 sort_dedup_small(u32* a, u32 n) {
  for (u32 i = 1; i < n; ++i) { u32 x = a[i]; u32 j = i; while (j > 0 && a[j - 1] > x) { a[j] = a[j - 1]; --j; } a[j] = x; }
  u32 m = 0;
  for (u32 i = 0; i < n; ++i) if (i == 0 || a[i] != a[m - 1]) a[m++] = a[i];
  return m;
}

// find-or-claim `key` in an open-address u64 table; returns the slot
__device__ __forceinline__ u32 tab_find_or_claim(u64* keys, u32 cap_mask, u32 log2cap, u64 key, bool* fresh) {
  u32 s = hash_key(key, log2cap);
  *fresh = false;
  for (;;) {
    u64 cur = *(volatile u64*)&keys[s];
    if (cur == EMPTY_KEY) {
      cur = atomicCAS((unsigned long long*)&keys[s], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
      if (cur == EMPTY_KEY) { *fresh = true; return s; }
    }
    if (cur == key) return s;
    s = (s + 1) & cap_mask;
  }
}
__device__ __forceinline__ u32 tab_find(const u64* keys, u32 cap_mask, u32 log2cap, u64 key) {
  u32 s = hash_key(key, log2cap);
  for (;;) {
    const u64 cur = keys[s];
    if (cur == key) return s;
    if (cur == EMPTY_KEY) return NONE32;
    s = (s + 1) & cap_mask;
  }
}
__device__ __forceinline__ u32 ilog2(u32 pow2) { return 31 - __clz((int)pow2); }

// block bitonic sort of u32 ids by a comparator, with a u32 payload. n power of two;
// ids == NONE32 sort last.
template <class Less>
__device__ inline void block_bitonic_ids(u32* ids, u32* pay, u32 n, Less less) {
  for (u32 k = 2; k <= n; k <<= 1) {
    for (u32 j = k >> 1; j > 0; j >>= 1) {
      for (u32 t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const u32 p = i | j;
        const u32 x = ids[i], y = ids[p];
        const bool up = (i & k) == 0;
        // gt = x > y under `less` with NONE32 = +inf
        bool gt;
        if (x == NONE32) gt = (y != NONE32);
        else if (y == NONE32) gt = false;
        else gt = less(y, x);
        if (gt == up) {
          ids[i] = y; ids[p] = x;
          if (pay) { const u32 a = pay[i], b = pay[p]; pay[i] = b; pay[p] = a; }
        }
      }
      __syncthreads();
    }
  }
}

__device__ inline void block_bitonic_u64(u64* a, u32 n) {
  for (u32 k = 2; k <= n; k <<= 1) {
    for (u32 j = k >> 1; j > 0; j >>= 1) {
      for (u32 t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const u32 p = i | j;
        const u64 x = a[i], y = a[p];
        const bool up = (i & k) == 0;
        if ((x > y) == up) { a[i] = y; a[p] = x; }
      }
      __syncthreads();
    }
  }
}

// ---- sorts staged through a shared-memory scratch (the arena itself is global memory, where a
// bitonic stage costs an L2 round trip; in shared memory it costs ~50 cycles) ------------------
constexpr u32 GE_SCRATCH_BYTES = 32768;
constexpr u32 GE_CLS_CACHE = 512;          // classes whose label offset/length are mirrored in shared memory
constexpr u32 GE_VCACHE = (GE_SCRATCH_BYTES / 2) / 12;   // vertices mirrored in the upper half of the scratch

__device__ inline void sort_pairs_staged(u64* keys, u32* vals, u32 n, u8* scratch, u32 budget = GE_SCRATCH_BYTES) {
  if ((u64)n * 12 <= budget && n > 1) {
    u64* sk = reinterpret_cast<u64*>(scratch);
    u32* sv = reinterpret_cast<u32*>(sk + n);
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) { sk[i] = keys[i]; sv[i] = vals[i]; }
    __syncthreads();
    block_bitonic_pairs(sk, sv, n);
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) { keys[i] = sk[i]; vals[i] = sv[i]; }
    __syncthreads();
  } else {
    block_bitonic_pairs(keys, vals, n);
  }
}
__device__ inline void sort_u64_staged(u64* a, u32 n, u8* scratch, u32 budget = GE_SCRATCH_BYTES) {
  if ((u64)n * 8 <= budget && n > 1) {
    u64* sk = reinterpret_cast<u64*>(scratch);
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) sk[i] = a[i];
    __syncthreads();
    block_bitonic_u64(sk, n);
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) a[i] = sk[i];
    __syncthreads();
  } else {
    block_bitonic_u64(a, n);
  }
}
__device__ inline void sort_u32_staged(u32* a, u32 n, u8* scratch, u32 budget = GE_SCRATCH_BYTES) {
  if ((u64)n * 4 <= budget && n > 1) {
    u32* sk = reinterpret_cast<u32*>(scratch);
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) sk[i] = a[i];
    __syncthreads();
    block_bitonic_u32(sk, n);
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) a[i] = sk[i];
    __syncthreads();
  } else {
    block_bitonic_u32(a, n);
  }
}
template <class Less>
__device__ inline void sort_ids_staged(u32* ids, u32* pay, u32 n, u8* scratch, Less less, u32 budget = GE_SCRATCH_BYTES) {
  if ((u64)n * 8 <= budget && n > 1) {
    u32* si = reinterpret_cast<u32*>(scratch);
    u32* sp = si + n;
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) { si[i] = ids[i]; sp[i] = pay[i]; }
    __syncthreads();
    block_bitonic_ids(si, sp, n, less);
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) { ids[i] = si[i]; pay[i] = sp[i]; }
    __syncthreads();
  } else {
    block_bitonic_ids(ids, pay, n, less);
  }
}

// block-wide exclusive scan over an array in place-out: out[i] = sum_{j<i} in[j]; returns total.
// in and out may alias.
__device__ inline u32 block_exscan_array(const u32* in, u32* out, u32 n, u32* s_scan) {
  u32 base = 0;
  for (u32 c0 = 0; c0 < n; c0 += blockDim.x) {
    const u32 i = c0 + threadIdx.x;
    const u32 v = i < n ? in[i] : 0;
    u32 tot;
    const u32 ex = block_exscan(v, s_scan, &tot);
    if (i < n) out[i] = base + ex;
    base += tot;
  }
  __syncthreads();
  return base;
}

// Out-of-place compaction with one contiguous index range per thread: count, one block scan,
// write. Two barriers instead of three per blockDim-sized chunk.
template <class Keep, class Write>
__device__ inline u32 block_compact_ranges(u32 N, u32* s_scan, Keep keep, Write write) {
  const u32 T = blockDim.x, K = (N + T - 1) / T;
  u32 lo = threadIdx.x * K; if (lo > N) lo = N;
  u32 hi = lo + K; if (hi > N) hi = N;
  u32 cnt = 0;
  for (u32 i = lo; i < hi; ++i) cnt += keep(i) ? 1u : 0u;
  u32 tot;
  u32 pos = block_exscan(cnt, s_scan, &tot);
  for (u32 i = lo; i < hi; ++i) if (keep(i)) write(i, pos++);
  __syncthreads();
  return tot;
}

// per-cell view used by the phases below
struct GeCell {
  __device__ GeCell(GePtrs& ptrs) : pm(&ptrs), p(ptrs) {}
  const KArgs* a;
  const GeArgs* g;
  GePtrs* pm;         // the same object, writable (thread 0 swaps table roles)
  const GePtrs& p;    // lives in shared memory: ~50 pointers that must not be re-derived or spilled per use
  u64 r0;
  u32 f0;
  bool gene_labels;   // labels are gene ids (PUG_GENE) — else transcript ids
  u8* scratch;        // GE_SCRATCH_BYTES of shared memory for staged sorts
  u32 scratch_budget; // bytes of it the sorts may use right now (the upper half may hold the vertex cache)
  const u64* vk_s;    // shared-memory copies of the dense vertex arrays while they fit (else nullptr)
  const u32* vc_s;
  const u32* cl_off_s; // shared-memory copies of the per-class label offset / length (else nullptr)
  const u32* cl_len_s;
  // label of record-local index i
  __device__ __forceinline__ const u32* rec_lab(u32 i) const {
    return gene_labels ? p.glab + (a->ref_off[r0 + i] - f0) : a->refs + a->ref_off[r0 + i];
  }
  __device__ __forceinline__ u32 rec_lab_len(u32 i) const {
    return gene_labels ? p.glen[i] : a->ref_off[r0 + i + 1] - a->ref_off[r0 + i];
  }
  // label of class rank c (through its representative record)
  __device__ __forceinline__ const u32* cls_lab(u32 c) const { return (gene_labels ? p.glab : a->refs) + (cl_off_s ? cl_off_s[c] : p.cls_loff[c]); }
  __device__ __forceinline__ u32 cls_lab_len(u32 c) const { return cl_len_s ? cl_len_s[c] : p.cls_llen[c]; }
  __device__ __forceinline__ u32 v_cls(u32 v) const { return (u32)((vk_s ? vk_s[v] : p.vtab_k[v]) >> 32); }
  __device__ __forceinline__ u32 v_umi(u32 v) const { return (u32)(vk_s ? vk_s[v] : p.vtab_k[v]); }
  __device__ __forceinline__ u32 v_cnt(u32 v) const { return vc_s ? vc_s[v] : p.vtab_c[v]; }
};

// Does the directed edge x -> y exist, given Hamming distance hd in {0,1} (has_edge,
// src/pugutils.rs:76-99)?  hd == 0: bidirected. hd == 1: x->y unless cy > 2*cx - 1.
__device__ __forceinline__ bool out_edge(u32 hd, u32 cx, u32 cy) {
  return hd == 0 || !(cy > 2 * cx - 1);
}

// Presence bitmap of the cell's UMIs (hashed) in shared memory: the 3*L substitution candidates
// of a vertex almost never exist, and the bitmap rejects them with one shared-memory load instead
// of a probe walk through the L2-resident UMI table.
constexpr u32 GE_BITMAP_LOG2 = 17;                       // 128 Kbit = 16 KB of the staged-sort scratch
__device__ __forceinline__ u32 umi_bit(u32 umi) { return (umi * 0x9E3779B1u) >> (32 - GE_BITMAP_LOG2); }

__device__ __forceinline__ u32 umi_bit2(u32 umi) { return ((umi ^ (umi >> 15)) * 0x85EBCA6Bu) >> (32 - GE_BITMAP_LOG2); }
__device__ __forceinline__ bool umi_maybe_present(const u32* bitmap, u32 umi) {   // 2-hash Bloom test
  const u32 b1 = umi_bit(umi), b2 = umi_bit2(umi);
  return ((bitmap[b1 >> 5] >> (b1 & 31)) & (bitmap[b2 >> 5] >> (b2 & 31)) & 1u) != 0;
}

// every vertex w != v carrying UMI cu whose class label shares a reference with class cv
template <class F>
__device__ __forceinline__ void visit_umi(const GeCell& c, u32 v, u32 cv, u32 cu, u32 utab_mask, u32 utab_log2, F f) {
  const u32 s = tab_find(c.p.ctab_h, utab_mask, utab_log2, (u64)cu);
  if (s == NONE32) return;
  for (u32 w = c.p.ctab_r[s]; w != NONE32; w = c.p.vnext[w]) {
    if (w == v) continue;
    const u32 cw = c.v_cls(w);
    if (cw != cv && !sorted_share(c.cls_lab(cv), c.cls_lab_len(cv), c.cls_lab(cw), c.cls_lab_len(cw))) continue;
    f(w);
  }
}

// PUG neighbours of every vertex v with want(v): f(v, w, hd) once per neighbour w — hd = 0 through
// v's own UMI, hd = 1 through one of its 3*umi_len single-base substitutions (umi_len == 0:
// --umi-edit-dist 0, own UMI only). One lane per VERTEX in both passes (ncu r1m: with a warp per
// vertex and lanes over the candidates, the own-UMI chain walk — the bulk of the edges — ran on
// one lane of the warp, a quarter of the kernel's instructions at 1 active lane):
//   pass A walks the chain of the vertex's own UMI;
//   pass B lets all lanes test the SAME substitution of their own vertex's UMI against a 2-hash
//   Bloom bitmap of the cell's UMIs in shared memory (a hit is rare) and re-joins the warp after
//   every candidate.
template <class Want, class F>
__device__ __forceinline__ void visit_neighbours(const GeCell& c, u32 V, u32 umi_len, u32 utab_mask, u32 utab_log2,
                                                 const u32* bitmap, Want want, F f) {
  const u32 T = blockDim.x, tid = threadIdx.x;
  GE_FOR(v, V) {
    if (want(v)) visit_umi(c, v, c.v_cls(v), c.v_umi(v), utab_mask, utab_log2, [&](u32 w) { f(v, w, 0u); });
  }
  if (umi_len == 0) return;
  for (u32 vb = 0; vb < V; vb += T) {
    const u32 v = vb + tid;
    const bool act = v < V && want(v);
    const u32 u = act ? c.v_umi(v) : 0u, cv = act ? c.v_cls(v) : 0u;
    for (u32 pos = 0; pos < umi_len; ++pos)
      for (u32 d = 1; d <= 3; ++d) {
        const u32 cu = u ^ (d << (2 * pos));
        const bool hit = act && umi_maybe_present(bitmap, cu);
        __syncwarp();
        if (hit) visit_umi(c, v, cv, cu, utab_mask, utab_log2, [&](u32 w) { f(v, w, 1u); });
      }
  }
}

// lock-free union-find (roots are the minimum vertex id of their set)
// (path halving: a non-root's parent is re-pointed at its grandparent with a plain store. Links are
// only ever made by CAS on ROOTS (parent[hi] == hi), and a non-root never becomes a root again, so the
// store cannot interfere with a link; it always points at an ancestor.)
__device__ inline u32 uf_find(u32* parent, u32 x) {
  for (;;) {
    const u32 p = *(volatile u32*)&parent[x];
    if (p == x) return x;
    const u32 gp = *(volatile u32*)&parent[p];
    if (gp != p) *(volatile u32*)&parent[x] = gp;
    x = gp;
  }
}
__device__ inline void uf_union(u32* parent, u32 a, u32 b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    const u32 hi = a > b ? a : b, lo = a > b ? b : a;
    if (atomicCAS(&parent[hi], hi, lo) == hi) return;
  }
}

// append a molecule whose label is the n genes at mlab + off
__device__ __forceinline__ void emit_molecule(const GeCell& c, GeShared* sh, u32 off, u32 n) {
  const u32 m = atomicAdd(&sh->n_mol, 1u);
  c.p.mol_off[m] = off;
  c.p.mol_len[m] = n;
}

// gene projection of a (sorted) reference list into dst; returns length (sorted, unique)
__device__ inline u32 project_label(const GeCell& c, const u32* lab, u32 len, u32* dst) {
  if (c.gene_labels) { for (u32 i = 0; i < len; ++i) dst[i] = lab[i]; return len; }
  for (u32 i = 0; i < len; ++i) dst[i] = __ldg(c.a->t2g + lab[i]);
  return sort_dedup_small(dst, len);
}

// ---------------------------------------------------------------------------------------------
// Greedy cover of one component with <= 32 vertices, by ONE thread, with 32-bit masks.
// mem[0..s) = vertex ids ascending (canonical visiting order, DESIGN.md determinism contract).
// Reference: get_num_molecules cover loop (src/pugutils.rs:1097-1261) + collapse_vertices
// (src/pugutils.rs:308-391).
// ---------------------------------------------------------------------------------------------
__device__ inline void cover_small_component(const GeCell& c, GeShared* sh, const u64* ckey, u32 pos, u32 s) {
  u32 mem[SMALL_COMP];
  u32 am[SMALL_COMP];
  for (u32 i = 0; i < s; ++i) mem[i] = (u32)ckey[pos + i];
  for (u32 i = 0; i < s; ++i) {
    u32 m = 0;
    const u32 v = mem[i];
    for (u32 e = c.p.adj_off[v]; e < c.p.adj_off[v + 1]; ++e) {
      const u32 w = c.g->adj_pool[((u64)sh->adj_base_hi << 32 | sh->adj_base_lo) + e];
      u32 lo = 0, hi = s;  // local index of w (binary search in mem)
      while (lo < hi) { u32 mid = (lo + hi) >> 1; if (mem[mid] < w) lo = mid + 1; else hi = mid; }
      m |= 1u << lo;
    }
    am[i] = m;
  }
  u32 unc = s == 32 ? 0xFFFFFFFFu : ((1u << s) - 1);
  while (unc) {
    u32 best_mask = 0, best_size = 0;
    const u32 remaining = (u32)__popc(unc);
    for (u32 i = 0; i < s && best_size < remaining; ++i) {
      if (!(unc >> i & 1)) continue;
      const u32 ci = c.v_cls(mem[i]);
      const u32* li = c.cls_lab(ci);
      const u32 ln = c.cls_lab_len(ci);
      for (u32 k = 0; k < ln; ++k) {
        const u32 t = li[k];
        u32 vis = 1u << i, fr = 1u << i, got = 1u << i;
        while (fr) {
          u32 nx = 0;
          u32 f = fr;
          while (f) {
            const u32 x = (u32)__ffs((int)f) - 1;
            f &= f - 1;
            u32 cand = am[x] & unc & ~vis;
            vis |= cand;
            while (cand) {
              const u32 j = (u32)__ffs((int)cand) - 1;
              cand &= cand - 1;
              const u32 cj = c.v_cls(mem[j]);
              if (cj == ci || sorted_contains(c.cls_lab(cj), c.cls_lab_len(cj), t)) { nx |= 1u << j; }
            }
          }
          got |= nx;
          fr = nx;
        }
        const u32 sz = (u32)__popc(got);
        if (sz > best_size) { best_size = sz; best_mask = got; }
        if (best_size == remaining) break;
      }
    }
    // intersect the labels of the MCC's vertices, project to genes, emit
    const u32 first = (u32)__ffs((int)best_mask) - 1;
    const u32 vfirst = mem[first];
    u32* dst = c.p.mlab + c.p.vlab_off[vfirst];
    const u32 cf = c.v_cls(vfirst);
    const u32* lf = c.cls_lab(cf);
    u32 len = c.cls_lab_len(cf);
    for (u32 i = 0; i < len; ++i) dst[i] = lf[i];
    u32 rest = best_mask & (best_mask - 1);
    while (rest) {
      const u32 j = (u32)__ffs((int)rest) - 1;
      rest &= rest - 1;
      const u32 cj = c.v_cls(mem[j]);
      if (cj == cf) continue;
      const u32* lj = c.cls_lab(cj);
      const u32 lnj = c.cls_lab_len(cj);
      u32 m2 = 0;
      for (u32 i = 0; i < len; ++i) if (sorted_contains(lj, lnj, dst[i])) dst[m2++] = dst[i];
      len = m2;
    }
    if (!c.gene_labels) {
      for (u32 i = 0; i < len; ++i) dst[i] = __ldg(c.a->t2g + dst[i]);
      len = sort_dedup_small(dst, len);
    }
    emit_molecule(c, sh, c.p.vlab_off[vfirst], len);
    unc &= ~best_mask;
  }
}

// ---------------------------------------------------------------------------------------------
// Greedy cover of one component with 32 < s <= large_graph_thresh vertices, CTA-cooperative,
// bitset rows in the arena. Threads own start vertices; the best (size, first start vertex,
// first transcript) is reduced across the CTA each round.
// ---------------------------------------------------------------------------------------------
__device__ inline void cover_mid_component(const GeCell& c, GeShared* sh, const u64* ckey, u32 pos, u32 s) {
  const u32 W = (s + 31) / 32;
  u32* mem = c.p.cmem;                 // [s] vertex ids
  u32* adjm = c.p.adjm;                // [s][W]
  u32* unc = c.p.bfs;                  // [W]
  u32* mcc = unc + W;                  // [W]
  u32* tmp = mcc + W;                  // [2W] spare
  u32* mine = tmp + 2 * W + 3 * W * threadIdx.x;  // vis / fr / nx of this thread
  const u64 adj_base = (u64)sh->adj_base_hi << 32 | sh->adj_base_lo;
  for (u32 i = threadIdx.x; i < s; i += blockDim.x) mem[i] = (u32)ckey[pos + i];
  for (u32 i = threadIdx.x; i < W; i += blockDim.x) {
    const u32 lo = i * 32;
    unc[i] = (s - lo >= 32) ? 0xFFFFFFFFu : ((1u << (s - lo)) - 1);
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < s; i += blockDim.x) {
    u32* row = adjm + (u64)i * W;
    for (u32 w = 0; w < W; ++w) row[w] = 0;
    const u32 v = mem[i];
    for (u32 e = c.p.adj_off[v]; e < c.p.adj_off[v + 1]; ++e) {
      const u32 x = c.g->adj_pool[adj_base + e];
      u32 lo = 0, hi = s;
      while (lo < hi) { u32 mid = (lo + hi) >> 1; if (mem[mid] < x) lo = mid + 1; else hi = mid; }
      row[lo >> 5] |= 1u << (lo & 31);
    }
  }
  if (threadIdx.x == 0) sh->cnt0 = s;  // remaining
  __syncthreads();

  // BFS from local vertex i along transcript t; result bitset left in `got` (= mine[0..W) reused)
  auto bfs = [&](u32 i, u32 t, u32 ci, u32* vis, u32* fr, u32* nx) -> u32 {
    for (u32 w = 0; w < W; ++w) { vis[w] = 0; fr[w] = 0; }
    vis[i >> 5] = 1u << (i & 31);
    fr[i >> 5] = 1u << (i & 31);
    u32 size = 1;
    // `vis` doubles as the visited set; matched vertices are tracked in nx-accumulated `fr` passes
    // we need the matched set at the end: keep it in tmp-free fashion by re-marking: matched = fr-union.
    // To keep memory small, matched set is rebuilt by the caller when needed (second BFS).
    for (;;) {
      bool any = false;
      for (u32 w = 0; w < W; ++w) nx[w] = 0;
      for (u32 w = 0; w < W; ++w) {
        u32 f = fr[w];
        while (f) {
          const u32 x = w * 32 + (u32)__ffs((int)f) - 1;
          f &= f - 1;
          const u32* row = adjm + (u64)x * W;
          for (u32 w2 = 0; w2 < W; ++w2) {
            u32 cand = row[w2] & unc[w2] & ~vis[w2];
            vis[w2] |= cand;
            while (cand) {
              const u32 j = w2 * 32 + (u32)__ffs((int)cand) - 1;
              cand &= cand - 1;
              const u32 cj = c.v_cls(mem[j]);
              if (cj == ci || sorted_contains(c.cls_lab(cj), c.cls_lab_len(cj), t)) {
                nx[j >> 5] |= 1u << (j & 31);
                ++size;
                any = true;
              }
            }
          }
        }
      }
      if (!any) break;
      for (u32 w = 0; w < W; ++w) fr[w] = nx[w];
    }
    return size;
  };

  while (sh->cnt0 > 0) {
    if (threadIdx.x == 0) { sh->best_size = 0; sh->best_i = NONE32; sh->best_k = NONE32; }
    __syncthreads();
    u32 my_size = 0, my_i = NONE32, my_k = NONE32;
    for (u32 i = threadIdx.x; i < s; i += blockDim.x) {
      if (!(unc[i >> 5] >> (i & 31) & 1)) continue;
      const u32 ci = c.v_cls(mem[i]);
      const u32* li = c.cls_lab(ci);
      const u32 ln = c.cls_lab_len(ci);
      for (u32 k = 0; k < ln; ++k) {
        const u32 sz = bfs(i, li[k], ci, mine, mine + W, mine + 2 * W);
        if (sz > my_size) { my_size = sz; my_i = i; my_k = k; }
      }
    }
    if (my_size) atomicMax(&sh->best_size, my_size);
    __syncthreads();
    if (my_size && my_size == sh->best_size) atomicMin(&sh->best_i, my_i);
    __syncthreads();
    if (my_size && my_size == sh->best_size && my_i == sh->best_i) atomicMin(&sh->best_k, my_k);
    __syncthreads();
    if (threadIdx.x == 0) {
      // rebuild the winning MCC as a bitset (matched vertices = start + every frontier)
      const u32 i = sh->best_i;
      const u32 ci = c.v_cls(mem[i]);
      const u32 t = c.cls_lab(ci)[sh->best_k];
      u32* vis = mine; u32* fr = mine + W; u32* nx = mine + 2 * W;
      for (u32 w = 0; w < W; ++w) { vis[w] = 0; fr[w] = 0; mcc[w] = 0; }
      vis[i >> 5] = 1u << (i & 31); fr[i >> 5] = 1u << (i & 31); mcc[i >> 5] = 1u << (i & 31);
      for (;;) {
        bool any = false;
        for (u32 w = 0; w < W; ++w) nx[w] = 0;
        for (u32 w = 0; w < W; ++w) {
          u32 f = fr[w];
          while (f) {
            const u32 x = w * 32 + (u32)__ffs((int)f) - 1;
            f &= f - 1;
            const u32* row = adjm + (u64)x * W;
            for (u32 w2 = 0; w2 < W; ++w2) {
              u32 cand = row[w2] & unc[w2] & ~vis[w2];
              vis[w2] |= cand;
              while (cand) {
                const u32 j = w2 * 32 + (u32)__ffs((int)cand) - 1;
                cand &= cand - 1;
                const u32 cj = c.v_cls(mem[j]);
                if (cj == ci || sorted_contains(c.cls_lab(cj), c.cls_lab_len(cj), t)) { nx[j >> 5] |= 1u << (j & 31); any = true; }
              }
            }
          }
        }
        if (!any) break;
        for (u32 w = 0; w < W; ++w) { fr[w] = nx[w]; mcc[w] |= nx[w]; }
      }
      // label intersection + projection + emit
      u32 first = NONE32;
      u32* dst = nullptr;
      u32 len = 0, cf = 0;
      u32 covered = 0;
      for (u32 w = 0; w < W; ++w) {
        u32 m = mcc[w];
        covered += (u32)__popc(m);
        while (m) {
          const u32 j = w * 32 + (u32)__ffs((int)m) - 1;
          m &= m - 1;
          const u32 cj = c.v_cls(mem[j]);
          if (first == NONE32) {
            first = j; cf = cj;
            dst = c.p.mlab + c.p.vlab_off[mem[j]];
            const u32* lf = c.cls_lab(cj);
            len = c.cls_lab_len(cj);
            for (u32 q = 0; q < len; ++q) dst[q] = lf[q];
          } else if (cj != cf) {
            const u32* lj = c.cls_lab(cj);
            const u32 lnj = c.cls_lab_len(cj);
            u32 m2 = 0;
            for (u32 q = 0; q < len; ++q) if (sorted_contains(lj, lnj, dst[q])) dst[m2++] = dst[q];
            len = m2;
          }
        }
      }
      if (!c.gene_labels) {
        for (u32 q = 0; q < len; ++q) dst[q] = __ldg(c.a->t2g + dst[q]);
        len = sort_dedup_small(dst, len);
      }
      emit_molecule(c, sh, c.p.vlab_off[mem[first]], len);
      for (u32 w = 0; w < W; ++w) unc[w] &= ~mcc[w];
      sh->cnt0 -= covered;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// cr-like resolution restricted to one component (> large_graph_thresh vertices), CTA-wide:
// get_num_molecules_large_component (src/pugutils.rs:916-982). One molecule per distinct UMI
// of the component, labelled with the arg-max gene set.
// Also used (whole cell) to build the molecules of `cr-like-em`.
// `count` vertices/records are supplied by the caller through `feed(insert)`.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ltab_add(const GeCell& c, u32 log2cap, u32 umi, u32 gene, u32 w) {
  bool fresh;
  const u32 s = tab_find_or_claim(c.p.ltab_k, (1u << log2cap) - 1, log2cap, ((u64)umi << 32) | gene, &fresh);
  atomicAdd(&c.p.ltab_c[s], w);
}

// after the pair table is filled: compact, sort by (umi, gene), walk each UMI, emit molecules
__device__ inline void crlike_molecules_from_ltab(const GeCell& c, GeShared* sh, u32 cap, bool prefer_ambig = false) {
  const u32 d = block_compact_pairs(c.p.ltab_k, c.p.ltab_c, cap, sh->scan);
  const u32 D = next_pow2(d);
  for (u32 i = d + threadIdx.x; i < D; i += blockDim.x) c.p.ltab_k[i] = EMPTY_KEY;
  __syncthreads();
  sort_pairs_staged(c.p.ltab_k, c.p.ltab_c, D, c.scratch, c.scratch_budget);
  for (u32 i = threadIdx.x; i < d; i += blockDim.x) {
    const u32 u = (u32)(c.p.ltab_k[i] >> 32);
    if (i > 0 && (u32)(c.p.ltab_k[i - 1] >> 32) == u) continue;
    u32 j = i;
    while (j < d && (u32)(c.p.ltab_k[j] >> 32) == u) ++j;
    // weight an entry votes with: its own count, or (--sa-model prefer-ambig, src/pugutils.rs:505-641) the
    // combined count of the ids 2k / 2k+1 of its gene, which are adjacent in the sorted segment
    auto weight = [&](u32 k) {
      u32 w = c.p.ltab_c[k];
      if (prefer_ambig) {
        const u32 gk = (u32)c.p.ltab_k[k];
        if (k > i && (((u32)c.p.ltab_k[k - 1]) | 1u) == (gk | 1u)) w += c.p.ltab_c[k - 1];
        if (k + 1 < j && (((u32)c.p.ltab_k[k + 1]) | 1u) == (gk | 1u)) w += c.p.ltab_c[k + 1];
      }
      return w;
    };
    u32 maxw = 0, nb = 0;
    for (u32 k = i; k < j; ++k) { const u32 w = weight(k); if (w > maxw) { maxw = w; nb = 1; } else if (w == maxw) ++nb; }
    const u32 off = c.p.P + atomicAdd(&sh->lab_bump, nb);  // upper half of mlab
    u32 q = 0;
    for (u32 k = i; k < j; ++k) if (weight(k) == maxw) c.p.mlab[off + q++] = (u32)c.p.ltab_k[k];
    emit_molecule(c, sh, off, nb);
  }
  __syncthreads();
}

struct GeCell;
__device__ inline void ge_back(const KArgs& a, const GeArgs& g, u32 cell, GeCell& c, GeShared* sh);

// =============================================================================================
// The kernel body for one cell. Produces this cell's sparse counts in the staging rows.
// =============================================================================================
__device__ inline void gene_eqc_cell(const KArgs& a, const GeArgs& g, u32 cell, u8* arena, GeShared* sh, u8* scratch,
                                     GePtrs* s_ptrs, u32* s_cls) {
  GeCell c(*s_ptrs);
  c.a = &a; c.g = &g; c.scratch = scratch; c.scratch_budget = GE_SCRATCH_BYTES;
  c.vk_s = nullptr; c.vc_s = nullptr; c.cl_off_s = nullptr; c.cl_len_s = nullptr;
  c.r0 = a.cell_rec_off[cell];
  const u64 r1 = a.cell_rec_off[cell + 1];
  const u32 n = (u32)(r1 - c.r0);
  c.f0 = a.ref_off[c.r0];
  const u32 P = a.ref_off[r1] - c.f0;
  c.gene_labels = g.ge_mode == GE_MODE_PUG_GENE;
  if (threadIdx.x == 0) {
    sh->need_lo = 0;
    const u64 need = ge_carve(arena, n, P, g.large_graph_thresh, s_ptrs);
    sh->need_lo = need > a.ctl->ge_bytes[g.list_which] ? 1u : 0u;
    sh->n_mol = 0; sh->lab_bump = 0; sh->alt = 0; sh->flag = 0;
    sh->cnt0 = sh->cnt1 = sh->cnt2 = sh->cnt3 = 0;
  }
  __syncthreads();
  if (sh->need_lo) {
    if (threadIdx.x == 0) {
      atomicOr(&a.ctl->error, (u32)DEV_ERR_ARENA);
      a.sum_umi[cell] = 0; a.max_umi[cell] = 0; a.num_expr[cell] = 0; a.num_over_mean[cell] = 0; a.flags[cell] = 4;
    }
    __syncthreads();
    return;
  }
  const GePtrs& p = c.p;
  const u32 T = blockDim.x, tid = threadIdx.x;

  if (g.ge_mode == GE_MODE_CRLIKE) {
    // ---------------- cr-like molecules: per UMI the arg-max gene set ------------------------
    const u32 cap = p.P2, l2 = ilog2(cap);
    GE_FOR(i, cap) { p.ltab_k[i] = EMPTY_KEY; p.ltab_c[i] = 0; }
    __syncthreads();
    GE_FOR(i, n) {
      const u32 umi = a.umi[c.r0 + i];
      const u32 o0 = a.ref_off[c.r0 + i], o1 = a.ref_off[c.r0 + i + 1];
      for (u32 k = o0; k < o1; ++k) {
        const u32 gg = __ldg(a.t2g + a.refs[k]);
        bool dup = false;
        for (u32 j = o0; j < k; ++j) if (__ldg(a.t2g + a.refs[j]) == gg) { dup = true; break; }
        if (!dup) ltab_add(c, l2, umi, gg, 1u);
      }
    }
    __syncthreads();
    crlike_molecules_from_ltab(c, sh, cap, a.prefer_ambig != 0);
  } else {
    // ---------------- phase 1: eq-classes -----------------------------------------------------
    if (c.gene_labels) {  // materialise the sorted-dedup gene projection of every record
      GE_FOR(i, n) {
        const u32 o0 = a.ref_off[c.r0 + i], o1 = a.ref_off[c.r0 + i + 1];
        u32* dst = p.glab + (o0 - c.f0);
        for (u32 k = o0; k < o1; ++k) dst[k - o0] = __ldg(a.t2g + a.refs[k]);
        p.glen[i] = sort_dedup_small(dst, o1 - o0);
      }
      __syncthreads();
    }
    const u32 N2 = p.N2, l2 = ilog2(N2), m2 = N2 - 1;
    u32 seed = 0;
    for (;; ++seed) {
      GE_FOR(i, N2) { p.ctab_h[i] = EMPTY_KEY; p.ctab_r[i] = NONE32; }
      if (tid == 0) sh->flag = 0;
      __syncthreads();
      GE_FOR(i, n) {
        bool fresh;
        const u32 s = tab_find_or_claim(p.ctab_h, m2, l2, label_hash(c.rec_lab(i), c.rec_lab_len(i), seed), &fresh);
        atomicMin(&p.ctab_r[s], i);
        p.rec_slot[i] = s;
      }
      __syncthreads();
      GE_FOR(i, n) {
        const u32 rep = p.ctab_r[p.rec_slot[i]];
        if (rep != i && !label_equal(c.rec_lab(i), c.rec_lab_len(i), c.rec_lab(rep), c.rec_lab_len(rep))) sh->flag = 1;
      }
      __syncthreads();
      const u32 bad = sh->flag;
      __syncthreads();
      if (!bad) break;
      if (seed >= 6) { if (tid == 0) atomicOr(&a.ctl->error, (u32)DEV_ERR_HASH); break; }
    }
    // compact classes -> (rep, slot), sort by label, assign ranks
    const u32 C = block_compact_ranges(N2, sh->scan, [&](u32 i) { return p.ctab_h[i] != EMPTY_KEY; },
                                       [&](u32 i, u32 pos) { p.cls_rep[pos] = p.ctab_r[i]; p.cls_aux[pos] = i; });
    const u32 Cp = next_pow2(C);
    for (u32 i = C + tid; i < Cp; i += T) { p.cls_rep[i] = NONE32; p.cls_aux[i] = NONE32; }
    __syncthreads();
    sort_ids_staged(p.cls_rep, p.cls_aux, Cp, c.scratch, [&](u32 x, u32 y) {
      return label_less(c.rec_lab(x), c.rec_lab_len(x), c.rec_lab(y), c.rec_lab_len(y));
    });
    GE_FOR(j, C) {
      p.ctab_r[p.cls_aux[j]] = j;  // slot -> rank
      const u32 rep = p.cls_rep[j];
      const u32 lo_ = c.gene_labels ? a.ref_off[c.r0 + rep] - c.f0 : a.ref_off[c.r0 + rep];
      const u32 ll_ = c.rec_lab_len(rep);
      p.cls_loff[j] = lo_;
      p.cls_llen[j] = ll_;
      if (C <= GE_CLS_CACHE) { s_cls[j] = lo_; s_cls[GE_CLS_CACHE + j] = ll_; }
    }
    __syncthreads();
    if (C <= GE_CLS_CACHE) { c.cl_off_s = s_cls; c.cl_len_s = s_cls + GE_CLS_CACHE; }

    // ---------------- phase 2: vertices = distinct (class, UMI) with read counts ---------------
    GE_FOR(i, N2) { p.vtab_k[i] = EMPTY_KEY; p.vtab_c[i] = 0; }
    __syncthreads();
    GE_FOR(i, n) {
      const u64 key = ((u64)p.ctab_r[p.rec_slot[i]] << 32) | a.umi[c.r0 + i];
      bool fresh;
      const u32 s = tab_find_or_claim(p.vtab_k, m2, l2, key, &fresh);
      atomicAdd(&p.vtab_c[s], 1u);
    }
    __syncthreads();
    // the class table is dead now (ranks are in the vertex keys): compact the vertex table into its
    // memory and swap the two tables' roles
    const u32 V = block_compact_ranges(N2, sh->scan, [&](u32 i) { return p.vtab_k[i] != EMPTY_KEY; },
                                       [&](u32 i, u32 pos) { p.ctab_h[pos] = p.vtab_k[i]; p.ctab_r[pos] = p.vtab_c[i]; });
    if (tid == 0) {
      u64* tk = c.pm->vtab_k; c.pm->vtab_k = c.pm->ctab_h; c.pm->ctab_h = tk;
      u32* tc = c.pm->vtab_c; c.pm->vtab_c = c.pm->ctab_r; c.pm->ctab_r = tc;
    }
    __syncthreads();
    const u32 Vp = next_pow2(V);
    for (u32 i = V + tid; i < Vp; i += T) p.vtab_k[i] = EMPTY_KEY;
    __syncthreads();
    sort_pairs_staged(p.vtab_k, p.vtab_c, Vp, c.scratch);   // canonical vertex order: (class rank, UMI)

    // ---------------- phase 3: UMI -> vertex chains (table time-shares the class table) ---------
    u32* bitmap = reinterpret_cast<u32*>(c.scratch);
    if (V <= GE_VCACHE) {   // mirror the dense vertex arrays in the upper half of the scratch (phases 3-6)
      u64* vk = reinterpret_cast<u64*>(c.scratch + GE_SCRATCH_BYTES / 2);
      u32* vc = reinterpret_cast<u32*>(vk + GE_VCACHE);
      GE_FOR(v, V) { vk[v] = p.vtab_k[v]; vc[v] = p.vtab_c[v]; }
      c.vk_s = vk; c.vc_s = vc;
      c.scratch_budget = GE_SCRATCH_BYTES / 2;
    }
    const u32 NU = pow2_ge(2 * V + 2, 64) < N2 ? pow2_ge(2 * V + 2, 64) : N2;   // UMI table sized by V, not by records
    const u32 lu = ilog2(NU), mu = NU - 1;
    GE_FOR(i, NU) { p.ctab_h[i] = EMPTY_KEY; p.ctab_r[i] = NONE32; }
    for (u32 i = tid; i < (1u << (GE_BITMAP_LOG2 - 5)); i += T) bitmap[i] = 0;
    __syncthreads();
    GE_FOR(v, V) {
      bool fresh;
      const u32 um = c.v_umi(v);
      const u32 s = tab_find_or_claim(p.ctab_h, mu, lu, (u64)um, &fresh);
      p.vnext[v] = atomicExch(&p.ctab_r[s], v);
      p.parent[v] = v;
      p.adj_off[v] = 0;
      p.vlab_off[v] = 0;      // fill cursor of the adjacency pass below
      const u32 bit = umi_bit(um), bit2 = umi_bit2(um);
      atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
      atomicOr(&bitmap[bit2 >> 5], 1u << (bit2 & 31));
    }
    __syncthreads();

    // ---------------- phase 4: out-degrees + union-find, then adjacency fill --------------------
    const u32 sub_len = g.pug_exact_umi ? 0u : g.umi_len;
    visit_neighbours(c, V, sub_len, mu, lu, bitmap, [](u32) { return true; }, [&](u32 v, u32 w, u32 hd) {
      if (w > v) uf_union(p.parent, v, w);
      if (out_edge(hd, c.v_cnt(v), c.v_cnt(w))) atomicAdd(&p.adj_off[v], 1u);
    });
    __syncthreads();
    const u32 E = block_exscan_array(p.adj_off, p.adj_off, V, sh->scan);
    if (tid == 0) {
      p.adj_off[V] = E;
      const u64 base = atomicAdd((unsigned long long*)g.adj_used, (unsigned long long)E);
      sh->adj_base_lo = (u32)base; sh->adj_base_hi = (u32)(base >> 32);
      sh->flag = (base + E > g.adj_cap) ? 1u : 0u;
    }
    __syncthreads();
    if (sh->flag) {
      if (tid == 0) {
        atomicOr(&a.ctl->error, (u32)DEV_ERR_ADJ_POOL);
        a.sum_umi[cell] = 0; a.max_umi[cell] = 0; a.num_expr[cell] = 0; a.num_over_mean[cell] = 0; a.flags[cell] = 4;
      }
      __syncthreads();
      return;
    }
    const u64 adj_base = (u64)sh->adj_base_hi << 32 | sh->adj_base_lo;
    if (E > 0)
      visit_neighbours(c, V, sub_len, mu, lu, bitmap, [&](u32 v) { return p.adj_off[v + 1] != p.adj_off[v]; },
                       [&](u32 v, u32 w, u32 hd) {
        if (out_edge(hd, c.v_cnt(v), c.v_cnt(w))) g.adj_pool[adj_base + p.adj_off[v] + atomicAdd(&p.vlab_off[v], 1u)] = w;
      });
    __syncthreads();
    // ---------------- phase 5: components (sorted by root, members ascending) -------------------
    GE_FOR(v, V) p.ckey[v] = ((u64)uf_find(p.parent, v) << 32) | v;
    for (u32 i = V + tid; i < Vp; i += T) p.ckey[i] = EMPTY_KEY;
    // label storage offsets per vertex (molecule labels live at their first vertex's region)
    GE_FOR(v, V) p.vlab_off[v] = c.cls_lab_len(c.v_cls(v));
    __syncthreads();
    block_exscan_array(p.vlab_off, p.vlab_off, V, sh->scan);
    sort_u64_staged(p.ckey, Vp, c.scratch, c.scratch_budget);
    // component starts
    u32 K = 0;
    {
      u32 base = 0;
      for (u32 c0 = 0; c0 < V; c0 += T) {
        const u32 i = c0 + tid;
        const u32 st = (i < V && (i == 0 || (p.ckey[i] >> 32) != (p.ckey[i - 1] >> 32))) ? 1u : 0u;
        u32 tot;
        const u32 pos = block_exscan(st, sh->scan, &tot);
        if (st) p.cstart[base + pos] = i;
        base += tot;
      }
      K = base;
      if (tid == 0) p.cstart[K] = V;
      __syncthreads();
    }
    // ---------------- phase 6: molecules -------------------------------------------------------
    GE_FOR(k, K) {
      const u32 pos = p.cstart[k], s = p.cstart[k + 1] - pos;
      if (s == 1) {
        const u32 v = (u32)p.ckey[pos];
        const u32 cv = c.v_cls(v);
        const u32 len = project_label(c, c.cls_lab(cv), c.cls_lab_len(cv), p.mlab + p.vlab_off[v]);
        emit_molecule(c, sh, p.vlab_off[v], len);
      } else if (s > g.large_graph_thresh || s > SMALL_COMP) {
        p.cbig[atomicAdd(&sh->cnt1, 1u)] = k;
      } else {
        cover_small_component(c, sh, p.ckey, pos, s);
      }
    }
    __syncthreads();
    const u32 nbig = sh->cnt1;
    if (nbig > 1) {  // deterministic processing order for the bump allocations
      const u32 Bp = next_pow2(nbig);
      for (u32 i = nbig + tid; i < Bp; i += T) p.cbig[i] = NONE32;
      __syncthreads();
      sort_u32_staged(p.cbig, Bp, c.scratch, c.scratch_budget);
    }
    for (u32 b = 0; b < nbig; ++b) {
      const u32 k = p.cbig[b];
      const u32 pos = p.cstart[k], s = p.cstart[k + 1] - pos;
      if (s > g.large_graph_thresh) {
        if (tid == 0) sh->alt = 1;
        const u32 cap = p.P2, ll2 = ilog2(cap);
        GE_FOR(i, cap) { p.ltab_k[i] = EMPTY_KEY; p.ltab_c[i] = 0; }
        __syncthreads();
        GE_FOR(i, s) {
          const u32 v = (u32)p.ckey[pos + i];
          const u32 cv = c.v_cls(v);
          u32* tmp = p.cmem + p.rows + p.vlab_off[v];  // scratch region sized like the label
          const u32 len = project_label(c, c.cls_lab(cv), c.cls_lab_len(cv), tmp);
          for (u32 q = 0; q < len; ++q) ltab_add(c, ll2, c.v_umi(v), tmp[q], c.v_cnt(v));
        }
        __syncthreads();
        crlike_molecules_from_ltab(c, sh, cap);
      } else {
        cover_mid_component(c, sh, p.ckey, pos, s);
      }
      __syncthreads();
    }
  }
  __syncthreads();

  c.vk_s = nullptr; c.vc_s = nullptr; c.scratch_budget = GE_SCRATCH_BYTES;   // vertex cache is dead from here on
  ge_back(a, g, cell, c, sh);
}

// =============================================================================================
// Stages B and C for one cell: molecules (mlab / mol_off / mol_len, sh->n_mol of them) -> gene
// eq-classes in canonical order -> counts or EM -> staging row + per-cell statistics. Shared by
// the global-arena kernel above and the shared-memory kernel (afq_pugs.cuh); the arrays it uses
// (GePtrs: mkey midx gcls_* tkey ent_* sup g_off alpha_* cls_inv sib_*) may live in either space.
// c.scratch_budget = bytes of c.scratch the staged sorts may use (0: sort in place).
// =============================================================================================
__device__ inline void ge_back(const KArgs& a, const GeArgs& g, u32 cell, GeCell& c, GeShared* sh) {
  const GePtrs& p = c.p;
  const u32 T = blockDim.x, tid = threadIdx.x;
  // =================== stage B: molecules -> gene eq-classes in canonical order ===============
  const u32 M = sh->n_mol;
  const u32 Mp = next_pow2(M);
  GE_FOR(i, Mp) {
    if (i < M) {
      const u32* lab = p.mlab + p.mol_off[i];
      const u32 len = p.mol_len[i];
      const u64 k0 = len > 0 ? (u64)lab[0] + 1 : 0;
      const u64 k1 = len > 1 ? (u64)lab[1] + 1 : 0;
      p.mkey[i] = (k0 << 32) | k1;
      p.midx[i] = i;
    } else {
      p.mkey[i] = EMPTY_KEY;
      p.midx[i] = NONE32;
    }
  }
  __syncthreads();
  sort_pairs_staged(p.mkey, p.midx, Mp, c.scratch, c.scratch_budget);
  // segments of equal 2-gene prefix; labels longer than 2 are ordered inside the segment by
  // one thread (insertion sort on the full label), then classes are counted
  auto mol_less = [&](u32 x, u32 y) {
    return label_less(p.mlab + p.mol_off[x], p.mol_len[x], p.mlab + p.mol_off[y], p.mol_len[y]);
  };
  auto mol_eq = [&](u32 x, u32 y) {
    return label_equal(p.mlab + p.mol_off[x], p.mol_len[x], p.mlab + p.mol_off[y], p.mol_len[y]);
  };
  GE_FOR(i, M) {
    if (i > 0 && p.mkey[i - 1] == p.mkey[i]) continue;
    u32 j = i + 1;
    while (j < M && p.mkey[j] == p.mkey[i]) ++j;
    bool longl = false;
    for (u32 k = i; k < j; ++k) if (p.mol_len[p.midx[k]] > 2) { longl = true; break; }
    if (longl) {
      for (u32 k = i + 1; k < j; ++k) {
        const u32 x = p.midx[k];
        u32 q = k;
        while (q > i && mol_less(x, p.midx[q - 1])) { p.midx[q] = p.midx[q - 1]; --q; }
        p.midx[q] = x;
      }
    }
  }
  __syncthreads();
  // class starts: molecule i starts a class iff its label differs from molecule i-1's
  u32 G = 0;
  {
    u32 base = 0;
    for (u32 c0 = 0; c0 < M; c0 += T) {
      const u32 i = c0 + tid;
      const u32 st = (i < M && p.mol_len[p.midx[i]] > 0 && (i == 0 || !mol_eq(p.midx[i], p.midx[i - 1]))) ? 1u : 0u;
      u32 tot;
      const u32 pos = block_exscan(st, sh->scan, &tot);
      if (st) p.gcls_m[base + pos] = i;   // position in sorted molecule order
      base += tot;
    }
    G = base;
    __syncthreads();
  }
  // zero-length labels sort first (key 0) and never start a class; count = distance to next start
  GE_FOR(j, G) {
    const u32 i = p.gcls_m[j];
    const u32 nxt = (j + 1 < G) ? p.gcls_m[j + 1] : M;
    p.gcls_cnt[j] = nxt - i;
  }
  __syncthreads();
  GE_FOR(j, G) p.gcls_m[j] = p.midx[p.gcls_m[j]];  // -> representative molecule id
  __syncthreads();
  auto cls_label = [&](u32 j) { return p.mlab + p.mol_off[p.gcls_m[j]]; };
  auto cls_len = [&](u32 j) { return p.mol_len[p.gcls_m[j]]; };
  if (g.dump_ncls) {
    // --dump-eqclasses: the cell's gene_eqc (label -> molecule count), classes in canonical (lexicographic) order.
    // A class stands for >= 1 molecule = >= 1 record whose label is at least as long, so G <= n and the labels fit P.
    u32 base = 0;
    for (u32 c0 = 0; c0 < G; c0 += T) {
      const u32 j = c0 + tid;
      const u32 ln = j < G ? cls_len(j) : 0u;
      u32 tot;
      const u32 ex = block_exscan(ln, sh->scan, &tot);
      if (j < G) {
        g.dump_off[c.r0 + j] = base + ex;
        g.dump_cnt[c.r0 + j] = p.gcls_cnt[j];
        const u32* lab = cls_label(j);
        for (u32 q = 0; q < ln; ++q) g.dump_lab[(u64)c.f0 + base + ex + q] = lab[q];
      }
      base += tot;
    }
    if (tid == 0) { g.dump_ncls[cell] = G; g.dump_nlab[cell] = base; }
    if (g.classes_only) { __syncthreads(); return; }
  }

  // =================== stage C: counts =========================================================
  const u64 out_base = c.f0;
  const bool usa = a.usa_mode != 0;
  u32 nnz = 0;
  if (g.only_unique && !usa) {
    // em_optimize(only_unique) (src/em.rs:499-514): singleton classes, ascending gene order
    u32 base = 0, lmax = 0;
    for (u32 c0 = 0; c0 < G; c0 += T) {
      const u32 j = c0 + tid;
      const u32 keep = (j < G && cls_len(j) == 1) ? 1u : 0u;
      u32 tot;
      const u32 pos = block_exscan(keep, sh->scan, &tot);
      if (keep) {
        a.stage_col[out_base + base + pos] = cls_label(j)[0];
        a.stage_val[out_base + base + pos] = (float)p.gcls_cnt[j];
        atomicAdd(&sh->cnt2, p.gcls_cnt[j]);
        lmax = p.gcls_cnt[j] > lmax ? p.gcls_cnt[j] : lmax;
      }
      base += tot;
    }
    nnz = base;
    if (lmax) atomicMax(&sh->cnt3, lmax);
    __syncthreads();
    if (tid == 0) { sh->fsum = (float)sh->cnt2; sh->fmax = (float)sh->cnt3; }
    __syncthreads();
  } else if (g.only_unique && usa) {
    // utils::extract_counts (src/utils.rs:673-756): label -> S/U/A slot, several classes may share one
    for (u32 j = tid; j < next_pow2(G); j += T) {
      u32 slot = NONE32, cnt = 0;
      if (j < G) {
        const u32 len = cls_len(j);
        slot = len > 10 ? NONE32 : usa_slot_for_label(cls_label(j), len, a.uo, a.ao);
        cnt = p.gcls_cnt[j];
      }
      p.tkey[j] = slot == NONE32 ? EMPTY_KEY : ((u64)slot << 32) | j;
      p.ent_idx[j] = cnt;
    }
    __syncthreads();
    sort_u64_staged(p.tkey, next_pow2(G), c.scratch, c.scratch_budget);
    u32 base = 0, lmax = 0;
    for (u32 c0 = 0; c0 < G; c0 += T) {
      const u32 i = c0 + tid;
      u32 st = 0, slot = 0, tot_cnt = 0;
      if (i < G && p.tkey[i] != EMPTY_KEY) {
        slot = (u32)(p.tkey[i] >> 32);
        st = (i == 0 || (u32)(p.tkey[i - 1] >> 32) != slot) ? 1u : 0u;
        if (st) for (u32 q = i; q < G && p.tkey[q] != EMPTY_KEY && (u32)(p.tkey[q] >> 32) == slot; ++q) tot_cnt += p.ent_idx[(u32)p.tkey[q]];
      }
      u32 tot;
      const u32 pos = block_exscan(st, sh->scan, &tot);
      if (st) {
        a.stage_col[out_base + base + pos] = slot;
        a.stage_val[out_base + base + pos] = (float)tot_cnt;
        atomicAdd(&sh->cnt2, tot_cnt);
        lmax = tot_cnt > lmax ? tot_cnt : lmax;
      }
      base += tot;
    }
    nnz = base;
    if (lmax) atomicMax(&sh->cnt3, lmax);
    __syncthreads();
    if (tid == 0) { sh->fsum = (float)sh->cnt2; sh->fmax = (float)sh->cnt3; }
    __syncthreads();
  } else {
    // ------------------------------- EM ---------------------------------------------------------
    // entries = (class, label position) in class order; USA rewrites gene-id labels into S/U/A
    // slot labels (extract_usa_eqmap, src/utils.rs:842-926: adjacent S,U of one gene fuse to A).
    GE_FOR(j, G) {
      u32 e = 0;
      if (!usa) e = cls_len(j);
      else {
        const u32* lab = cls_label(j);
        const u32 len = cls_len(j);
        if (len == 1) e = 1;
        else for (u32 q = 0; q < len; ++q) { ++e; if (d_is_spliced(lab[q]) && q + 1 < len && d_same_gene(lab[q], lab[q + 1])) ++q; }
      }
      p.gcls_eoff[j] = e;
    }
    __syncthreads();
    const u32 Lt = block_exscan_array(p.gcls_eoff, p.gcls_eoff, G, sh->scan);
    if (tid == 0) p.gcls_eoff[G] = Lt;
    __syncthreads();
    GE_FOR(j, G) {
      const u32* lab = cls_label(j);
      const u32 len = cls_len(j);
      u32 e = p.gcls_eoff[j];
      if (!usa) { for (u32 q = 0; q < len; ++q) p.ent_idx[e++] = lab[q]; }
      else if (len == 1) { p.ent_idx[e] = d_is_spliced(lab[0]) ? (lab[0] >> 1) : a.uo + (lab[0] >> 1); }
      else {
        for (u32 q = 0; q < len; ++q) {
          u32 idx = lab[q] >> 1;
          if (d_is_spliced(lab[q])) { if (q + 1 < len && d_same_gene(lab[q], lab[q + 1])) { idx += a.ao; ++q; } }
          else idx += a.uo;
          p.ent_idx[e++] = idx;
        }
      }
    }
    __syncthreads();
    // needs_em (src/em.rs:321-341): M2 returns the singleton tallies untouched if no class has > 1 label.
    // M1 (gene mode) always iterates.
    if (tid == 0) sh->flag = 0;
    __syncthreads();
    GE_FOR(j, G) if (p.gcls_eoff[j + 1] - p.gcls_eoff[j] > 1) sh->flag = 1;
    __syncthreads();
    const bool needs_em = !usa || sh->flag;
    // support: label indices (+ USA siblings, src/em.rs:87-113), sorted unique
    const u32 per = usa ? 3u : 1u;
    const u32 Sraw = Lt * per;
    const u32 Sp = next_pow2(Sraw);
    GE_FOR(e, Sp) p.sup[e] = NONE32;
    __syncthreads();
    GE_FOR(e, Lt) {
      const u32 idx = p.ent_idx[e];
      p.sup[e * per] = idx;
      if (usa) {
        if (idx >= a.ao) { p.sup[e * per + 1] = idx - a.uo; p.sup[e * per + 2] = idx - a.ao; }
        else if (idx >= a.uo) p.sup[e * per + 1] = idx + a.uo;
        else p.sup[e * per + 1] = idx + a.ao;
      }
    }
    __syncthreads();
    sort_u32_staged(p.sup, Sp, c.scratch, c.scratch_budget);
    u32 S = 0;
    {  // unique in place (chunked, same hazard-free pattern as the compactions)
      u32 base = 0;
      for (u32 c0 = 0; c0 < Sraw; c0 += T) {
        const u32 i = c0 + tid;
        u32 v = NONE32, keep = 0;
        if (i < Sraw) { v = p.sup[i]; keep = (v != NONE32 && (i == 0 || p.sup[i - 1] != v)) ? 1u : 0u; }
        u32 tot;
        const u32 pos = block_exscan(keep, sh->scan, &tot);
        if (keep) p.sup[base + pos] = v;
        base += tot;
        __syncthreads();
      }
      S = base;
    }
    auto loc_of = [&](u32 idx) {
      u32 lo = 0, hi = S;
      while (lo < hi) { u32 mid = (lo + hi) >> 1; if (p.sup[mid] < idx) lo = mid + 1; else hi = mid; }
      return (lo < S && p.sup[lo] == idx) ? lo : NONE32;
    };
    GE_FOR(e, Lt) p.ent_loc[e] = loc_of(p.ent_idx[e]);
    // USA sibling locations per support index: abundance (src/em.rs:167-187)
    //   A: a[U] + a[S] + a[A];  U: a[A] + a[U];  S: a[A] + a[S]
    GE_FOR(s, S) {
      u32 sa = NONE32, sb = NONE32;
      if (usa) {
        const u32 idx = p.sup[s];
        if (idx >= a.ao) { sa = loc_of(idx - a.uo); sb = loc_of(idx - a.ao); }
        else if (idx >= a.uo) sa = loc_of(idx + a.uo);
        else sa = loc_of(idx + a.ao);
      }
      p.sib_a[s] = sa; p.sib_b[s] = sb;
      p.alpha_in[s] = 0.0f; p.alpha_out[s] = 0.0f;
    }
    // transposed CSR: entries grouped by support index, class order inside a group
    const u32 Lp = next_pow2(Lt);
    GE_FOR(e, Lp) p.tkey[e] = EMPTY_KEY;
    __syncthreads();
    // key = (support index, CLASS of the entry): a class holds an index at most once, so this sorts like
    // (support index, entry) and the loops below read the class straight from the key
    GE_FOR(j, G) for (u32 e = p.gcls_eoff[j]; e < p.gcls_eoff[j + 1]; ++e) p.tkey[e] = ((u64)p.ent_loc[e] << 32) | j;
    __syncthreads();
    sort_u64_staged(p.tkey, Lp, c.scratch, c.scratch_budget);
    for (u32 s = tid; s <= S; s += T) {  // g_off[s] = first sorted entry whose support index is >= s
      u32 lo = 0, hi = Lt;
      while (lo < hi) { u32 mid = (lo + hi) >> 1; if ((u32)(p.tkey[mid] >> 32) < s) lo = mid + 1; else hi = mid; }
      p.g_off[s] = lo;
    }
    __syncthreads();
    // singleton tallies, accumulated per index in class order (integers: exact). cls_inv[j] carries what
    // class j adds to one of its indices in the M step: >= 0: abundance * cls_inv (multi-label class,
    // set by the E step); -1: nothing; <= -3: the constant -(cls_inv + 2) = its count (singleton class)
    GE_FOR(j, G) p.cls_inv[j] = (p.gcls_eoff[j + 1] - p.gcls_eoff[j] == 1) ? -((float)p.gcls_cnt[j] + 2.0f) : -1.0f;
    __syncthreads();
    GE_FOR(s, S) {
      float t = 0.0f;
      for (u32 q = p.g_off[s]; q < p.g_off[s + 1]; ++q) {
        const float w = p.cls_inv[(u32)p.tkey[q]];
        if (w < -2.0f) t = __fadd_rn(t, -w - 2.0f);
      }
      p.alpha_in[s] = t;
    }
    __syncthreads();
    if (needs_em) {
      const float uni = __fdiv_rn(1.0f, (float)g.num_alphas);
      // The EM iterations, written over (me, stride, sync) so that a SMALL problem (<= 96 support indices
      // and classes: a high-duplication cell has ~50 molecules) runs on ONE warp with __syncwarp() between
      // the steps instead of two block barriers per iteration (ncu r1zd, C5: 16 warps stalled at a barrier
      // per issuing warp). Same operations in the same order either way.
      auto em_loop = [&](const u32 me, const u32 stride, auto sync) {
        float* ain = p.alpha_in;     // ping-pong: the M step writes aout, then the two swap (no copy pass)
        float* aout = p.alpha_out;
        for (u32 b = 0; b < S; b += stride) {
          const u32 s_ = b + me;
          if (s_ < S) ain[s_] = g.em_init_uniform ? uni : __fmul_rn(__fadd_rn(ain[s_], 0.5f), 1e-3f);
        }
        if (me == 0) { sh->flag = 0; sh->cnt1 = 0; }
        sync();
        auto abund = [&](u32 s_) {
          if (!usa) return ain[s_];
          const u32 sa = p.sib_a[s_], sb = p.sib_b[s_];
          if (sb != NONE32 || p.sup[s_] >= a.ao) {  // ambiguous: U + S + A
            const float xu = sa != NONE32 ? ain[sa] : 0.0f, xs = sb != NONE32 ? ain[sb] : 0.0f;
            return __fadd_rn(__fadd_rn(xu, xs), ain[s_]);
          }
          const float xa = sa != NONE32 ? ain[sa] : 0.0f;   // U or S: A + self
          return __fadd_rn(xa, ain[s_]);
        };
        u32 it = 0;
        bool last_round = false;
        for (;;) {
          // E step, per class: inv = count / sum of abundances (label order)
          for (u32 b = 0; b < G; b += stride) {
            const u32 j = b + me;
            if (j < G) {
              const u32 e0 = p.gcls_eoff[j], e1 = p.gcls_eoff[j + 1];
              if (e1 - e0 > 1) {     // (singleton classes keep their constant)
                float inv = -1.0f;   // class contributes nothing
                float den = 0.0f;
                for (u32 e = e0; e < e1; ++e) den = __fadd_rn(den, abund(p.ent_loc[e]));
                if (den > 0.0f) inv = __fdiv_rn((float)p.gcls_cnt[j], den);
                p.cls_inv[j] = inv;
              }
            }
            __syncwarp();
          }
          sync();
          // "some index moved" flags alternate between two words: the word of iteration it+1 is cleared
          // here, behind a barrier every thread reached after it last read that word (iteration it-1)
          u32* moved = (it & 1u) ? &sh->cnt1 : &sh->flag;
          if (me == 0) *((it & 1u) ? &sh->flag : &sh->cnt1) = 0;
          // M step, per support index, contributions added in class order
          for (u32 b = 0; b < S; b += stride) {
            const u32 s_ = b + me;
            if (s_ < S) {
              float out = 0.0f;
              const float ab = abund(s_);
              for (u32 q = p.g_off[s_]; q < p.g_off[s_ + 1]; ++q) {
                const float w = p.cls_inv[(u32)p.tkey[q]];
                if (w >= 0.0f) out = __fadd_rn(out, __fmul_rn(ab, w));
                else if (w < -2.0f) out = __fadd_rn(out, -w - 2.0f);
              }
              aout[s_] = out;
              if (out > 1e-2f && fabsf(__fsub_rn(ain[s_], out)) > 1e-2f) *moved = 1;
            }
            __syncwarp();
          }
          sync();
          const bool converged = *(volatile u32*)moved == 0;
          { float* t = ain; ain = aout; aout = t; }
          ++it;
          if (!usa) {  // M1: src/em.rs:538-565
            if (!(it < 2 || (it < 100 && !converged))) break;
          } else {     // M2: src/em.rs:391-443 (clamp, then one last round)
            if (last_round) break;
            if (it >= 2 && converged) {
              for (u32 b = 0; b < S; b += stride) { const u32 s_ = b + me; if (s_ < S && ain[s_] < 0.01f) ain[s_] = 0.0f; }
              last_round = true;
              sync();
            } else if (!(it < 2 || (it < 100 && !converged))) break;
          }
        }
        for (u32 b = 0; b < S; b += stride) {
          const u32 s_ = b + me;
          if (s_ < S) { const float x = ain[s_]; p.alpha_in[s_] = x < 0.01f ? 0.0f : x; }
        }
      };
      if (S <= 96 && G <= 96) {
        if (tid < 32) em_loop(tid, 32u, [] { __syncwarp(); });
      } else {
        em_loop(tid, T, [] { __syncthreads(); });
      }
      __syncthreads();
    }
    // emit positive alphas, ascending index
    u32 base = 0;
    for (u32 c0 = 0; c0 < S; c0 += T) {
      const u32 s = c0 + tid;
      const u32 keep = (s < S && p.alpha_in[s] > 0.0f) ? 1u : 0u;
      u32 tot;
      const u32 pos = block_exscan(keep, sh->scan, &tot);
      if (keep) { a.stage_col[out_base + base + pos] = p.sup[s]; a.stage_val[out_base + base + pos] = p.alpha_in[s]; }
      base += tot;
    }
    nnz = base;
    __syncthreads();
    if (tid == 0) {  // the reference's sequential f32 scan (src/quant.rs:1156-1168)
      float fs = 0.0f, fm = 0.0f;
      for (u32 i = 0; i < nnz; ++i) { const float v = a.stage_val[out_base + i]; fs = __fadd_rn(fs, v); fm = v > fm ? v : fm; }
      sh->fsum = fs; sh->fmax = fm;
    }
    __syncthreads();
  }
  // ---- per-cell statistics (src/quant.rs:1181-1196) -------------------------------------------
  if (tid == 0) sh->cnt0 = 0;
  __syncthreads();
  const float mean = __fdiv_rn(sh->fsum, (float)nnz);
  u32 lover = 0;
  GE_FOR(i, nnz) if (a.stage_val[out_base + i] > mean) ++lover;
  if (lover) atomicAdd(&sh->cnt0, lover);
  __syncthreads();
  if (tid == 0) {
    a.sum_umi[cell] = sh->fsum;
    a.max_umi[cell] = sh->fmax;
    a.num_expr[cell] = nnz;
    a.num_over_mean[cell] = sh->cnt0;
    u8 f = 0;
    if (sh->alt) f |= 2;
    if (nnz == 0) f |= 4;
    a.flags[cell] = f;
  }
  __syncthreads();
}

// persistent kernel: CTAs pull cells from work list `list_id` (largest cells first)
#ifndef AFQ_GE_MIN_BLOCKS
#define AFQ_GE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(GE_THREADS, AFQ_GE_MIN_BLOCKS) k_gene_eqc(KArgs a, GeArgs g) {
  __shared__ GeShared sh;
  __shared__ __align__(16) u8 s_scratch[GE_SCRATCH_BYTES];
  __shared__ GePtrs s_ptrs;
  __shared__ u32 s_cls[2 * GE_CLS_CACHE];
  const u32 count = a.ctl->bin_count[g.list_id];
  if (arena_cta_idle(a.ctl, a.ctl->ge_blocks[g.list_which], count)) return;
  u8* arena = g.arena + (u64)blockIdx.x * a.ctl->ge_bytes[g.list_which];
  const u32* list = a.bin_list + (u64)g.list_id * a.n_cells;
  for (;;) {
    if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[g.list_id], 1u);
    __syncthreads();
    const u32 job = sh.job;
    __syncthreads();
    if (job >= count) break;
    gene_eqc_cell(a, g, list[job], arena, &sh, s_scratch, &s_ptrs, s_cls);
  }
}

constexpr int GE_LIST_BIG = NUM_BINS;      // bin_list row for cells > GE_BIG_RECORDS
constexpr int GE_LIST_NORMAL = NUM_BINS + 1;
constexpr u32 GE_BIG_RECORDS = 1u << 16;

}  // namespace afq
