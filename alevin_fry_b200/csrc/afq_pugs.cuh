// afq_pugs.cuh — the parsimony family (and cr-like-em) for one cell ENTIRELY IN SHARED MEMORY.
//
// Same reference behaviour as afq_pug.cuh (src/eq_class.rs:723-1036, src/pugutils.rs:65-391,
// 989-1330, src/em.rs:167-582; determinism contract of DESIGN.md), re-laid-out so that a cell's
// records are read from HBM exactly once (coalesced) and every later access is a shared-memory
// access. ncu r1s showed the global-arena kernel (k_gene_eqc) at 544 k warp instructions per C3
// cell, spread over block-wide sorts of record-sized arrays living in L2; this kernel sizes every
// structure by what the cell actually contains and never sorts records:
//
//   load      refs / record offsets (16-bit, relative) / UMIs -> arena; the NEXT cell is prefetched into L2
//   phase 1   record -> eq-class: open-address table of REPRESENTATIVE RECORD indices, labels
//             compared in place (no stored hashes, no reseeding)
//   phase 2   (class, UMI) vertices: same kind of table, read count in the entry's upper 16 bits
//   compact   table slots -> dense vertex arrays  vumi[V], vinfo[V] = class << 16 | reads,
//             vgene[V] = the class's gene when it has exactly one (emission needs no t2g gather then)
//   phase 3   UMI -> vertex chains (table of chain heads) + Bloom bitmap of the UMIs with two
//             GF(2)-linear hashes (a substitution costs one XOR with a compile-time constant)
//   phase 4   union-find (path halving) over PUG edges: own-UMI chain successors + the 3*L
//             substitutions that pass the Bloom test (one lane per vertex). The PUG is never stored.
//   phase 5   components: singletons emit directly; members of larger components hang on per-root
//             lists, the roots are counting-sorted by component size, and components are covered by
//             groups of 2 / 4 / 8 lanes (lane = start vertex; label-position bitmasks make the BFS of
//             all transcripts of a start vertex branch-free relaxations in registers) or, from 9
//             vertices / 33-transcript labels on, by one warp (16 x 16 masks in shared memory). Start
//             vertices are ranked canonically (class label lexicographic, UMI).
//   cr-like-em (gene mode): instead of phases 4-5 one lane per UMI chain takes the arg-max gene set.
//   counts    unique-only resolutions: presence bitmap over the output slots + prefix popcount
//             (sort + run-length when the gene axis does not fit the arena);
//             EM resolutions hand their molecules to ge_back (afq_pug.cuh) on arena pointers.
//
// Arena variants: 72 KB x 3 CTAs/SM, 108 KB x 2, 224 KB x 1 of shared memory, and the same code on a
// per-CTA global-memory arena (32-bit offsets, up to 65 534 records). Anything that does not fit
// after all (actual vertex count, EM back end), or holds a component with more than 32 vertices
// / beyond --large-graph-thresh, sends the WHOLE cell to the global-arena kernel's work list, which
// runs afterwards — results are identical by construction because both kernels implement the same
// canonical orders (tests force every hand-back path: AFQ_PS_LIMIT_WORDS, AFQ_NO_PS[_GLOBAL]).
#pragma once
#include "afq_pug.cuh"

namespace afq {


constexpr int PS_VARIANTS = 4;       // arenas: 72 KB x 3 CTAs/SM, 108 KB x 2, 224 KB x 1 of shared memory; variant 3 = the same
                                     // code on a per-CTA GLOBAL-memory arena (L2-resident) for cells beyond 224 KB
constexpr int PS_SMEM_VARIANTS = 3;
#ifndef AFQ_PS_V0_KWORDS
#define AFQ_PS_V0_KWORDS 18
#define AFQ_PS_V0_BLOCKS 3
#endif
#ifndef AFQ_PS_V0_THREADS
#define AFQ_PS_V0_THREADS 256
#endif
#ifndef AFQ_PS_V1_THREADS
#define AFQ_PS_V1_THREADS 512
#endif
#ifndef AFQ_PS_V2_THREADS
#define AFQ_PS_V2_THREADS 1024
#endif
#ifndef AFQ_PS_V3_THREADS
#define AFQ_PS_V3_THREADS 512
#endif
__host__ __device__ constexpr u32 ps_threads(int v) { return v == 0 ? (u32)AFQ_PS_V0_THREADS : (v == 1 ? (u32)AFQ_PS_V1_THREADS : (v == 2 ? (u32)AFQ_PS_V2_THREADS : (u32)AFQ_PS_V3_THREADS)); }
__host__ __device__ constexpr u32 ps_arena_words(int v) { return v == 0 ? AFQ_PS_V0_KWORDS * 1024u : (v == 1 ? 27u * 1024u : 56u * 1024u); }
__host__ __device__ constexpr u32 ps_min_blocks(int v) { return v == 0 ? (u32)AFQ_PS_V0_BLOCKS : (v == 1 ? 2u : (v == 2 ? 1u : 2u)); }
constexpr u32 PS_WSCR_WORDS = 64 + 256; // per-warp scratch of the warp-cooperative cover (members, masks, 16 x 16 label masks)
constexpr u32 PS_COVER_WARPS = 4;       // warps of a 256-thread CTA that run the warp form (bounds its shared scratch: 5 KB);
                                        // 8 in the larger CTAs
constexpr u32 PC_MAX_WINNERS = 16384;    // split path: molecules whose per-slot counters fit k_pug_count's shared memory (beyond: global scratch)
constexpr u32 PS_EMPTY = 0xFFFFFFFFu;
constexpr u32 PS_MULTI_GENE = 0xFFFFFFFEu;
constexpr u32 PS_MAX_RECORDS = 65535;   // record indices and read counts share a 32-bit table entry (16 bits each)
constexpr u32 PS_MAX_REFS = 65535;      // 16-bit relative record offsets in the shared-memory variants (32-bit in variant 3)

__host__ __device__ inline u32 ps_table_size(u32 n) { return pow2_ge(n + n / 4 + 8, 64); }

// Arena words a cell of n records / P alignments is EXPECTED to need (vertex count guessed at
// n/2; the kernel re-checks with the real counts and falls back if they do not fit).
__host__ __device__ inline u32 ps_need_words(u32 n, u32 P, bool gene, bool em, bool usa, bool split = false) {
  const u32 rec = n + (n + 1) / 2 + ps_table_size(n);          // UMIs, classes, table: dead after compaction
  const u32 vest = n / 2 + 16;
  u32 bw = pow2_ge(2 * vest, 64); if (bw > 4096) bw = 4096;
  // (the split form keeps no winners / cover scratch in the arena, only the staged singleton slots)
  const u32 post = pow2_ge(vest + vest / 2 + 2, 64) + bw + 2 * vest + (split ? vest : ((em ? 0 : pow2_ge(vest, 1)) + PS_COVER_WARPS * PS_WSCR_WORDS));
  u32 w = P + (n + 2) / 2 + (gene ? (n + 1) / 2 : 0) + rec + 3 * vest + (post > rec ? post - rec : 0) + 24;   // (+ alignment slack of the bulk-copied arrays)
  if (em) {
    w += 2 * vest + P / 2 + 64;
    // the EM back end (ps_back_carve) on ~n/3 molecules with ~1.4 label entries each, + molecules at the top
    const u32 M = n / 3 + 8, Lm = M + M / 2, per = usa ? 3u : 1u;
    const u32 Mp = pow2_ge(M, 1), TK = pow2_ge(Lm, 1), Sp = pow2_ge(Lm * per, 1), Sr = Lm * per + 4;
    const u32 back = 3 * Mp + 3 * (M + 2) + 3 * TK + Lm + Sp + 5 * Sr + M + 2 * vest + Lm + 64;
    if (back > w) w = back;
  }
  return w;
}
// smallest arena variant that is expected to hold the cell, or -1 (global_ok: variant 3 is available)
__host__ __device__ inline int ps_variant_for(u64 n, u64 P, bool gene, bool em, bool usa, bool global_ok, bool split = false) {
  if (n >= PS_MAX_RECORDS || P >= (1ull << 30) || n == 0) return -1;
  if (P < PS_MAX_REFS) {
    const u32 need = ps_need_words((u32)n, (u32)P, gene, em, usa, split);
    for (int v = 0; v < PS_SMEM_VARIANTS; ++v)
      if (need <= ps_arena_words(v)) return v;
  }
  return global_ok ? 3 : -1;
}
// words of global arena per CTA that hold ANY cell of up to n records / P alignments (every vertex distinct)
__host__ __device__ inline u64 ps_global_words(u32 n, u32 P, u32 num_rows) {
  if (n >= PS_MAX_RECORDS) n = PS_MAX_RECORDS - 1;
  const u64 rec = (u64)P + (n + 2) + (n + 1) / 2 + n + (n + 1) / 2 + ps_table_size(n);
  const u64 post = (u64)pow2_ge(n + n / 2 + 2, 64) + 4096 + 3ull * n + pow2_ge(n, 1) + 2ull * PS_COVER_WARPS * PS_WSCR_WORDS + n / 2 + 2 +
                   2ull * ((num_rows + 31) / 32);   // chain table, bitmap, vnext / parent / nxt, winners, cover scratch, re-route list, slot bitmap
  return (rec + 3ull * n + post + 2ull * n + P + 4096 + 3) & ~3ull;   // + EM: molecule offsets / lengths and labels; 16-byte multiple
}

struct PsExtra {
  u32 fail;       // the cell must be redone by the global-arena kernel
  u32 n_mlist;    // vertices in components of size > 1
  u32 n_win;      // winners (unique-only resolutions)
  u32 szc[SMALL_COMP + 4];   // multi-vertex components per size (counting sort of their roots)
  u32 n_over;     // components of <= 8 vertices re-routed to the warp-cooperative cover (a label > 32 transcripts)
  u32 next_w, next_g;   // cover work queues: warp-form components / group passes handed out to the warps
  u32 dbase[4];         // SPLIT: this cell's first descriptor on each size-class list
  u32 n_bulk;           // bulk-copy loads issued so far (the mbarrier's phase parity)
  unsigned long long bar;   // mbarrier of the bulk-copy loads
};

// all lanes of a warp call: the warp claims the next item of a shared-memory work counter
__device__ __forceinline__ u32 warp_claim(u32* counter) {
  u32 it = 0;
  if (lane_id() == 0) it = atomicAdd(counter, 1u);
  return __shfl_sync(0xFFFFFFFFu, it, 0);
}

// all 32 lanes of the warp call; lanes with pred get consecutive indices from *counter
__device__ __forceinline__ u32 warp_bump(u32* counter, bool pred) {
  const u32 m = __ballot_sync(0xFFFFFFFFu, pred);
  if (!m) return 0;
  const u32 lane = lane_id();
  const u32 leader = (u32)__ffs((int)m) - 1;
  u32 base = 0;
  if (lane == leader) base = atomicAdd(counter, (u32)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, (int)leader);
  return base + (u32)__popc(m & ((1u << lane) - 1u));
}

struct PsCell {
  const KArgs* a;
  u32* refs;            // [P] transcript ids (gene ids after the in-place projection, PUG_GENE)
  const u16* roff;      // [n+1] record offsets into refs (shared-memory variants) ...
  const u32* roff32;    // ... or 32-bit offsets (variant 3; then roff == nullptr)
  const u16* rlen;      // [n] label lengths (PUG_GENE) or nullptr (length = offset difference)
  const u32* vumi;      // [V]
  const u32* vinfo;     // [V] class (representative record) << 16 | read count
  const u32* vgene;     // [V] the class's gene if it has exactly one, PS_MULTI_GENE, or NONE32 (empty label)
  bool gene;            // labels already are gene ids
  __device__ __forceinline__ u32 off(u32 r) const { return roff ? (u32)roff[r] : roff32[r]; }
  __device__ __forceinline__ const u32* lab(u32 r) const { return refs + off(r); }
  __device__ __forceinline__ u32 len(u32 r) const { return rlen ? (u32)rlen[r] : off(r + 1) - off(r); }
  __device__ __forceinline__ u32 vcls(u32 v) const { return vinfo[v] >> 16; }
  __device__ __forceinline__ u32 vcnt(u32 v) const { return vinfo[v] & 0xFFFFu; }
  __device__ __forceinline__ u32 gene_of(u32 x) const { return gene ? x : __ldg(a->t2g + x); }
  __device__ __forceinline__ bool related(u32 cv, u32 cw) const {   // classes share a reference
    return cv == cw || sorted_share(lab(cv), len(cv), lab(cw), len(cw));
  }
};

// where molecules go
struct PsSink {
  u32 mode;             // 0 unique-only gene mode, 1 unique-only USA, 2 labels for ge_back (EM)
  u32 uo, ao;
  u32* A;               // arena base (mode 2: labels are stored at word offsets from it)
  u32 lab_lo, lab_hi;   // mode 2: label bump region [lab_lo, lab_hi)
  u32* mol_off; u32* mol_len;
  GeShared* sh;         // n_mol / lab_bump
  PsExtra* ex;
  // mode 3 (split path, EM resolutions): molecules go to the cell's regions of the GLOBAL molecule pool
  u32* g_lab;           // labels of the cell: pool + f0 (at most P words, see afq_pugc.cuh)
  u32* g_nlab;          // label words used so far
  u32* g_nmol;          // molecules emitted so far
  u32* g_moff; u32* g_mlen;   // per molecule: label offset / length (cell region: + r0)
};

// store one molecule label of `want` reserved words; returns the destination (mode 2: arena, labels grow down; mode 3:
// the cell's global pool region, labels grow up) or nullptr when the arena is exhausted
__device__ __forceinline__ u32* ps_label_alloc(const PsSink& sk, u32 want, u32* off_out) {
  if (sk.mode == 3) {
    const u32 off = atomicAdd(sk.g_nlab, want);
    *off_out = off;
    return sk.g_lab + off;
  }
  const u32 used = atomicAdd(&sk.sh->lab_bump, want) + want;     // labels grow DOWN from lab_hi
  if (used > sk.lab_hi - sk.lab_lo) { sk.ex->fail = 1; return nullptr; }
  *off_out = sk.lab_hi - used;
  return sk.A + *off_out;
}
__device__ __forceinline__ void ps_label_commit(const PsSink& sk, u32 off, u32 len) {
  if (sk.mode == 3) {
    const u32 id = atomicAdd(sk.g_nmol, 1u);
    sk.g_moff[id] = off; sk.g_mlen[id] = len;
    return;
  }
  const u32 id = atomicAdd(&sk.sh->n_mol, 1u);
  sk.mol_off[id] = off;
  sk.mol_len[id] = len;
}

// One molecule whose transcript label is { t = l0[k], k in [0, n0) : keep(k, t) } (ascending). Returns the
// output slot for the unique-only modes (NONE32: contributes nothing); mode 2 stores the sorted
// gene label and returns NONE32.
template <class Keep>
__device__ __forceinline__ u32 ps_emit(const PsCell& c, const PsSink& sk, const u32* l0, u32 n0, Keep keep) {
  if (sk.mode == 0) {
    // em_optimize(only_unique), src/em.rs:499-514: only single-gene labels count
    u32 g0 = NONE32;
    for (u32 k = 0; k < n0; ++k) {
      const u32 t = l0[k];
      if (!keep(k, t)) continue;
      const u32 gg = c.gene_of(t);
      if (g0 == NONE32) g0 = gg;
      else if (gg != g0) return NONE32;
    }
    return g0;
  }
  if (sk.mode == 1) {
    // utils::extract_counts, src/utils.rs:673-756: sorted-unique gene label of <= 10 ids -> S/U/A slot
    u32 best[11];
    u32 nb = 0;
    for (u32 k = 0; k < n0; ++k) {
      const u32 t = l0[k];
      if (!keep(k, t)) continue;
      const u32 gg = c.gene_of(t);
      u32 q = 0;
      while (q < nb && best[q] < gg) ++q;
      if (q < nb && best[q] == gg) continue;
      if (nb == 11) continue;                       // already too long to matter
      for (u32 j = nb; j > q; --j) best[j] = best[j - 1];
      best[q] = gg;
      ++nb;
    }
    if (nb > 10) return NONE32;
    return usa_slot_for_label(best, nb, sk.uo, sk.ao);
  }
  // mode 2: store the sorted-unique gene label. It is built in a small thread-local buffer first so
  // that exactly its length is taken from the arena (a 40-transcript label is usually 1-2 genes)
  u32 buf[32];
  u32 nb = 0;
  bool big = false;
  for (u32 k = 0; k < n0 && !big; ++k) {
    const u32 t = l0[k];
    if (!keep(k, t)) continue;
    const u32 gg = c.gene_of(t);
    u32 q = nb;
    if (!c.gene) { q = 0; while (q < nb && buf[q] < gg) ++q; if (q < nb && buf[q] == gg) continue; }
    if (nb == 32) { big = true; break; }
    for (u32 j = nb; j > q; --j) buf[j] = buf[j - 1];
    buf[q] = gg;
    ++nb;
  }
  const u32 want = big ? n0 : nb;
  u32 off;
  u32* dst = ps_label_alloc(sk, want, &off);
  if (!dst) return NONE32;
  u32 m = nb;
  if (!big) { for (u32 q = 0; q < nb; ++q) dst[q] = buf[q]; }
  else {      // more than 32 genes: project in place
    m = 0;
    for (u32 k = 0; k < n0; ++k) {
      const u32 t = l0[k];
      if (keep(k, t)) dst[m++] = c.gene_of(t);
    }
    if (!c.gene) m = sort_dedup_small(dst, m);
  }
  ps_label_commit(sk, off, m);
  return NONE32;
}


constexpr u32 PS_CRL_GENES = 24;     // cr-like-em: candidate genes of one UMI held by a thread

// one molecule whose (ascending, distinct) gene label is given
__device__ __forceinline__ u32 ps_emit_genes(const PsSink& sk, const u32* genes, u32 nb) {
  if (sk.mode == 0) return nb == 1 ? genes[0] : NONE32;
  if (sk.mode == 1) return nb <= 10 ? usa_slot_for_label(genes, nb, sk.uo, sk.ao) : NONE32;
  u32 off;
  u32* dst = ps_label_alloc(sk, nb, &off);
  if (!dst) return NONE32;
  for (u32 q = 0; q < nb; ++q) dst[q] = genes[q];
  ps_label_commit(sk, off, nb);
  return NONE32;
}

// ---------------------------------------------------------------------------------------------
// Greedy monochromatic cover of one component with 2..32 vertices (get_num_molecules cover loop,
// src/pugutils.rs:1097-1261, + collapse_vertices, src/pugutils.rs:308-391). The members hang on
// the root's list (head / nxt). Start vertices are visited in ascending (class label, UMI) order;
// the first strictly larger MCC wins, i.e. the largest MCC of the earliest (start vertex,
// transcript). Two forms: one THREAD per component (sizes 2..PS_WARP_COMP-1, the bulk) and one
// WARP per component (lane i = start vertex i), because a component's cover costs O(s^2 |label|)
// dependent shared-memory loads per round and a single thread on a 12-vertex component kept the
// whole CTA waiting at the barrier (ncu r1u: 39 % of all stall samples).
// ---------------------------------------------------------------------------------------------
constexpr u32 PS_WARP_COMP = 9;       // components of this many vertices and more take the warp form (a 16-lane group form
                                      // was measured: register arrays of 16 spill 1.2 KB per thread and C5 got 25 % slower, r1x)

__device__ __forceinline__ bool ps_canon_less(const PsCell& c, u32 x, u32 y) {   // (class label lexicographic, UMI)
  const u32 cx = c.vcls(x), cy = c.vcls(y);
  return cx == cy ? c.vumi[x] < c.vumi[y] : label_less(c.lab(cx), c.len(cx), c.lab(cy), c.len(cy));
}
// out-neighbour mask of member i (has_edge, src/pugutils.rs:76-99) over members j in [j0, s)
__device__ __forceinline__ u32 ps_out_mask(const PsCell& c, const u32* mem, u32 s, u32 i, bool exact) {
  const u32 ui = c.vumi[mem[i]], ci = c.vcls(mem[i]), ni = c.vcnt(mem[i]);
  u32 m = 0;
  for (u32 j = 0; j < s; ++j) {
    if (j == i) continue;
    const u32 x = ui ^ c.vumi[mem[j]];
    const u32 hd = (u32)__popc((x | (x >> 1)) & 0x55555555u);
    if (exact ? hd != 0 : hd > 1) continue;
    if (!c.related(ci, c.vcls(mem[j]))) continue;
    if (out_edge(hd, ni, c.vcnt(mem[j]))) m |= 1u << j;
  }
  return m;
}
// vertices reachable from member i over out-edges through uncovered vertices whose label holds t
__device__ __forceinline__ u32 ps_bfs(const PsCell& c, const u32* mem, const u32* am, u32 unc, u32 i, u32 ci, u32 t) {
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
        const u32 cj = c.vcls(mem[j]);
        if (cj == ci || sorted_contains(c.lab(cj), c.len(cj), t)) nx |= 1u << j;
      }
    }
    got |= nx;
    fr = nx;
  }
  return got;
}
// molecule of one MCC: label = intersection of its class labels (src/pugutils.rs:1161-1188) -> genes
__device__ __forceinline__ void ps_emit_mcc(const PsCell& c, const PsSink& sk, u32* winners, u32* gbm, const u32* mem, u32 mask) {
  const u32 first = (u32)__ffs((int)mask) - 1;
  const u32 cf = c.vcls(mem[first]);
  const u32 rest = mask & (mask - 1);
  const u32 slot = ps_emit(c, sk, c.lab(cf), c.len(cf), [&](u32, u32 t) {
    u32 r = rest;
    while (r) {
      const u32 j = (u32)__ffs((int)r) - 1;
      r &= r - 1;
      const u32 cj = c.vcls(mem[j]);
      if (cj != cf && !sorted_contains(c.lab(cj), c.len(cj), t)) return false;
    }
    return true;
  });
  if (sk.mode < 2 && slot != NONE32) {
    winners[atomicAdd(&sk.ex->n_win, 1u)] = slot;
    if (gbm) atomicOr(&gbm[slot >> 5], 1u << (slot & 31));
  }
}

// G lanes per component (G = 2, 4 or 8, sizes <= G), 32/G components per warp pass; every lane of the
// warp calls. Lane `sub` of a group owns start vertex `sub`. Everything a BFS needs is precomputed as
// bitmasks RELATIVE TO THE LANE'S OWN LABEL: M[j] = positions k of my label whose transcript is in
// vertex j's label. Then the BFS of ALL my transcripts at once is a fixed number of branch-free
// relaxations reach[j] |= reach[x] & M[j] over the out-edges x -> j, in registers.
// Components with a label longer than 32 transcripts are appended to `olist` for the warp form.
template <int G>
__device__ inline void ps_cover_group(const PsCell& c, const PsSink& sk, u32* winners, u32* gbm, const u32* head, const u32* nxt,
                                      const u32* clist, u32 kb, u32 k1, bool exact, u32* olist, u32* n_over) {
  const u32 lane = lane_id(), sub = lane % G, gbase = lane - sub;
  {     // one warp pass: components clist[kb .. kb + 32/G) (warp-uniform kb; the caller hands out the passes)
    const u32 k = kb + lane / G;
    const bool valid = k < k1;
    const u32 r = valid ? clist[k] : 0u;
    u32 v = valid ? head[r] : PS_EMPTY;
    for (u32 q = 0; q < sub && v != PS_EMPTY; ++q) v = nxt[v];
    bool has = v != PS_EMPTY;
    u32 ci = 0, ui = 0, ni = 0, ln = 0;
    const u32* li = nullptr;
    if (has) { ci = c.vcls(v); ui = c.vumi[v]; ni = c.vcnt(v); li = c.lab(ci); ln = c.len(ci); }
    __syncwarp();
    // a label that does not fit a 32-bit position mask: the whole component takes the warp form
    const u32 overm = __ballot_sync(0xFFFFFFFFu, has && ln > 32);
    if ((overm >> gbase) & ((1u << G) - 1u)) {
      if (sub == 0 && valid) olist[atomicAdd(n_over, 1u)] = r;
      has = false;
    }
    const u32 full = ln >= 32 ? 0xFFFFFFFFu : ((1u << ln) - 1u);
    u32 M[G];
    u32 am = 0, rank = 0;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const u32 vj = __shfl_sync(0xFFFFFFFFu, v, (int)gbase + j);
      const bool hj = __shfl_sync(0xFFFFFFFFu, (u32)has, (int)gbase + j) != 0;
      const u32 cj = __shfl_sync(0xFFFFFFFFu, ci, (int)gbase + j);
      const u32 uj = __shfl_sync(0xFFFFFFFFu, ui, (int)gbase + j);
      const u32 nj = __shfl_sync(0xFFFFFFFFu, ni, (int)gbase + j);
      u32 m = 0;
      if (has && hj) {
        if ((u32)j == sub || cj == ci) m = full;
        else {
          const u32* lj = c.lab(cj);      // positions of my label present in j's: one merge of the two sorted lists
          const u32 lnj = c.len(cj);
          for (u32 qa = 0, qb = 0; qa < ln && qb < lnj;) {
            const u32 x = li[qa], y = lj[qb];
            if (x == y) { m |= 1u << qa; ++qa; ++qb; }
            else if (x < y) ++qa;
            else ++qb;
          }
        }
        if ((u32)j != sub) {
          const u32 x = ui ^ uj;
          const u32 hd = (u32)__popc((x | (x >> 1)) & 0x55555555u);
          // has_edge (src/pugutils.rs:76-99): classes must share a reference (m != 0), Hamming distance <= 1
          if ((exact ? hd == 0 : hd <= 1) && m != 0 && out_edge(hd, ni, nj)) am |= 1u << j;
          if (cj == ci ? uj < ui : label_less(c.lab(cj), c.len(cj), li, ln)) ++rank;   // canonical order (class label, UMI)
          (void)vj;
        }
      }
      M[j] = m;
      __syncwarp();
    }
    u32 amx[G];
#pragma unroll
    for (int x = 0; x < G; ++x) amx[x] = __shfl_sync(0xFFFFFFFFu, am, (int)gbase + x);
    u32 unc = (__ballot_sync(0xFFFFFFFFu, has) >> gbase) & ((1u << G) - 1u);
    while (__any_sync(0xFFFFFFFFu, unc != 0)) {
      u32 my_size = 0, my_mask = 1u << sub, my_k = 0;
      const bool start = has && ((unc >> sub) & 1u);
      if (start) {
        u32 reach[G];
#pragma unroll
        for (int j = 0; j < G; ++j) reach[j] = (u32)j == sub ? full : 0u;
        for (int round = 0; round < G - 1; ++round) {     // a path has at most G - 1 edges; usually 1-2 rounds
          u32 changed = 0;
#pragma unroll
          for (int x = 0; x < G; ++x) {
            const u32 ax = amx[x] & unc;
#pragma unroll
            for (int j = 0; j < G; ++j) {
              const u32 add = ((ax >> j) & 1u) ? (reach[x] & M[j] & ~reach[j]) : 0u;
              reach[j] |= add;
              changed |= add;
            }
          }
          if (!changed) break;
        }
        for (u32 q = 0; q < ln; ++q) {            // first transcript with the largest reachable set
          u32 sz = 0, mk = 0;
#pragma unroll
          for (int j = 0; j < G; ++j) { const u32 b = (reach[j] >> q) & 1u; sz += b; mk |= b << j; }
          if (sz > my_size) { my_size = sz; my_mask = mk; my_k = q; }
        }
      }
      __syncwarp();
      // group arg-max: largest MCC, earliest start vertex in canonical order
      const u32 key = start ? ((my_size << 8) | (255u - rank)) + 1u : 0u;
      u32 best = key;
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) { const u32 t = __shfl_xor_sync(0xFFFFFFFFu, best, o); best = t > best ? t : best; }
      const u32 wm = (__ballot_sync(0xFFFFFFFFu, start && key == best) >> gbase) & ((1u << G) - 1u);
      const u32 wl = wm ? (u32)__ffs((int)wm) - 1 : 0u;
      const u32 mask = __shfl_sync(0xFFFFFFFFu, my_mask, (int)(gbase + wl));
      if (start && wm && sub == wl) {
        // label = intersection of the MCC's class labels (src/pugutils.rs:1161-1188), as positions of mine
        u32 inter = full;
#pragma unroll
        for (int j = 0; j < G; ++j) if ((mask >> j) & 1u) inter &= M[j];
        (void)my_k;
        const u32 gv = c.vgene[v];
        const u32 slot = (gv < PS_MULTI_GENE && inter != 0) ? ps_emit_genes(sk, &gv, 1u)
                                                            : ps_emit(c, sk, li, ln, [&](u32 q, u32) { return ((inter >> q) & 1u) != 0; });
        if (sk.mode < 2 && slot != NONE32) {
          winners[atomicAdd(&sk.ex->n_win, 1u)] = slot;
          if (gbm) atomicOr(&gbm[slot >> 5], 1u << (slot & 31));
        }
      }
      if (wm) unc &= ~mask;
      __syncwarp();
    }
  }
}

// one warp per component (lane i = start vertex i). Per-warp shared scratch: wmem[32] members in
// canonical order, wam[32] out-neighbour masks, Mw[16 x 16] label-position masks. Every lane calls.
// Components of <= 16 vertices whose labels hold <= 32 transcripts (nearly all) use the MASK form:
// Mw[i][j] = positions of i's label present in j's label, built once by the whole warp (one merge
// per pair); edges, every BFS and the label intersection are then bit operations on shared-memory
// words. The BFS form with per-candidate label searches handles the rest (ncu r1z, C5: the BFS form
// on 9-16 vertex components with 18-transcript labels kept 28 % of the stall samples at the barrier).
__device__ inline void ps_cover_warp(const PsCell& c, const PsSink& sk, u32* winners, const u32* head, const u32* nxt, u32 r, bool exact,
                                     u32* gbm, u32* wmem, u32* wam) {
  const u32 lane = lane_id();
  u32* Mw = wam + 32;
  u32 s = 0;
  if (lane == 0) for (u32 x = head[r]; x != PS_EMPTY; x = nxt[x]) wam[s++] = x;     // unordered members
  s = __shfl_sync(0xFFFFFFFFu, s, 0);
  __syncwarp();
  u32 x = 0, rank = 0;
  if (lane < s) {   // rank sort into canonical order (a strict total order: (class, UMI) is unique)
    x = wam[lane];
    for (u32 j = 0; j < s; ++j) if (j != lane && ps_canon_less(c, wam[j], x)) ++rank;
  }
  __syncwarp();
  if (lane < s) wmem[rank] = x;
  __syncwarp();
  u32 ci = 0, ln = 0;
  const u32* li = nullptr;
  if (lane < s) { ci = c.vcls(wmem[lane]); li = c.lab(ci); ln = c.len(ci); }
  const bool masked = s <= 16 && !__any_sync(0xFFFFFFFFu, lane < s && ln > 32);
  if (masked) {
    for (u32 p = lane; p < s * s; p += 32) {      // all (i, j) pairs across the warp
      const u32 i = p / s, j = p - i * s;
      const u32 cI = c.vcls(wmem[i]), cJ = c.vcls(wmem[j]);
      const u32 lnI = c.len(cI);
      u32 m = 0;
      if (i == j || cI == cJ) m = lnI >= 32 ? 0xFFFFFFFFu : ((1u << lnI) - 1u);
      else {
        const u32* lI = c.lab(cI);
        const u32* lJ = c.lab(cJ);
        const u32 lnJ = c.len(cJ);
        for (u32 qa = 0, qb = 0; qa < lnI && qb < lnJ;) {
          const u32 a_ = lI[qa], b_ = lJ[qb];
          if (a_ == b_) { m |= 1u << qa; ++qa; ++qb; }
          else if (a_ < b_) ++qa;
          else ++qb;
        }
      }
      Mw[i * 16 + j] = m;
    }
    __syncwarp();
  }
  u32 my_am = 0;
  if (lane < s) {
    if (!masked) my_am = ps_out_mask(c, wmem, s, lane, exact);
    else {
      const u32 ui = c.vumi[wmem[lane]], ni = c.vcnt(wmem[lane]);
      for (u32 j = 0; j < s; ++j) {
        if (j == lane) continue;
        const u32 xx = ui ^ c.vumi[wmem[j]];
        const u32 hd = (u32)__popc((xx | (xx >> 1)) & 0x55555555u);
        if ((exact ? hd == 0 : hd <= 1) && Mw[lane * 16 + j] != 0 && out_edge(hd, ni, c.vcnt(wmem[j]))) my_am |= 1u << j;
      }
    }
  }
  wam[lane] = my_am;
  __syncwarp();
  u32 unc = s == 32 ? 0xFFFFFFFFu : ((1u << s) - 1);
  while (unc) {
    const u32 remaining = (u32)__popc(unc);
    u32 my_size = 0, my_mask = 0;
    if (lane < s && (unc >> lane & 1)) {
      for (u32 k = 0; k < ln; ++k) {
        u32 got;
        if (!masked) got = ps_bfs(c, wmem, wam, unc, lane, ci, li[k]);
        else {
          u32 ck = 0;                               // vertices whose label holds my transcript k
          for (u32 j = 0; j < s; ++j) ck |= ((Mw[lane * 16 + j] >> k) & 1u) << j;
          const u32 ok = ck & unc;
          u32 fr = 1u << lane;
          got = fr;
          while (fr) {
            u32 nx = 0;
            for (u32 f = fr; f; f &= f - 1) nx |= wam[(u32)__ffs((int)f) - 1];
            nx &= ok & ~got;
            got |= nx;
            fr = nx;
          }
        }
        const u32 sz = (u32)__popc(got);
        if (sz > my_size) { my_size = sz; my_mask = got; }
        if (my_size == remaining) break;
      }
    }
    u32 best = my_size;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const u32 t = __shfl_xor_sync(0xFFFFFFFFu, best, o); best = t > best ? t : best; }
    u32 mask;
    u32 winner;
    if (best == 0) { winner = (u32)__ffs((int)unc) - 1; mask = 1u << winner; }      // empty labels only: cover the first alone
    else {
      const u32 cand = __ballot_sync(0xFFFFFFFFu, my_size == best);
      winner = (u32)__ffs((int)cand) - 1;                                           // earliest start vertex with the largest MCC
      mask = __shfl_sync(0xFFFFFFFFu, my_mask, (int)winner);
    }
    if (lane == winner) {
      if (!masked) ps_emit_mcc(c, sk, winners, gbm, wmem, mask);
      else {
        u32 inter = 0xFFFFFFFFu;
        for (u32 mm = mask; mm; mm &= mm - 1) inter &= Mw[lane * 16 + (u32)__ffs((int)mm) - 1];
        const u32 gv = c.vgene[wmem[lane]];
        const u32 slot = (gv < PS_MULTI_GENE && inter != 0) ? ps_emit_genes(sk, &gv, 1u)
                                                            : ps_emit(c, sk, li, ln, [&](u32 q, u32) { return ((inter >> q) & 1u) != 0; });
        if (sk.mode < 2 && slot != NONE32) {
          winners[atomicAdd(&sk.ex->n_win, 1u)] = slot;
          if (gbm) atomicOr(&gbm[slot >> 5], 1u << (slot & 31));
        }
      }
    }
    unc &= ~mask;
    __syncwarp();
  }
}

// Bloom bitmap of the cell's UMIs with two GF(2)-LINEAR hashes: H(u ^ delta) = H(u) ^ H(delta), so
// the 3*L single-base substitutions of a UMI cost one XOR with a compile-time constant each instead
// of a multiplicative hash (ncu r1t: the Bloom tests were 15 % of the kernel's instructions).
__host__ __device__ constexpr u32 ps_h1(u32 u) { return u ^ (u >> 6) ^ (u >> 15); }
__host__ __device__ constexpr u32 ps_h2(u32 u) { return (u >> 3) ^ (u >> 10) ^ (u >> 20) ^ (u << 7); }
__device__ __forceinline__ u32 ps_umi_slot(u32 umi, u32 log2cap) { return (umi * 0x9E3779B1u) >> (32 - log2cap); }

// carve the ge_back arrays for M molecules / Lm label words from words [lo, hi) of the arena;
// returns false if they do not fit
__device__ inline bool ps_back_carve(u32* A, u32 lo, u32 hi, u32 M, u32 Lm, u32 per, GePtrs* o) {
  u64 off = (u64)lo * 4;
  auto take = [&](u64 bytes) { off = align8(off); u8* r = reinterpret_cast<u8*>(A) + off; off += bytes; return r; };
  const u32 Mp = next_pow2(M ? M : 1), Lp = next_pow2(Lm ? Lm : 1), TK = Mp > Lp ? Mp : Lp;
  const u32 Sp = next_pow2(Lm * per ? Lm * per : 1);
  o->mkey = (u64*)take(8ull * Mp); o->midx = (u32*)take(4ull * Mp);
  o->gcls_m = (u32*)take(4ull * (M + 1)); o->gcls_cnt = (u32*)take(4ull * (M + 1)); o->gcls_eoff = (u32*)take(4ull * (M + 2));
  o->tkey = (u64*)take(8ull * TK);
  o->ent_idx = (u32*)take(4ull * (TK + 1)); o->ent_loc = (u32*)take(4ull * (Lm + 1));
  const u32 Sr = Lm * per + 2;     // support indices before de-duplication (only `sup` is sorted: power of two)
  o->sup = (u32*)take(4ull * Sp); o->g_off = (u32*)take(4ull * (Sr + 2));
  o->alpha_in = (float*)take(4ull * Sr); o->alpha_out = (float*)take(4ull * Sr); o->cls_inv = (float*)take(4ull * (M + 1));
  o->sib_a = (u32*)take(4ull * Sr); o->sib_b = (u32*)take(4ull * Sr);
  return align8(off) <= (u64)hi * 4;
}

// stage B only (k_pug_back with classes_only): molecule sort keys + class arrays
__device__ inline bool ps_back_carve_b(u32* A, u32 lo, u32 hi, u32 M, GePtrs* o) {
  u64 off = (u64)lo * 4;
  auto take = [&](u64 bytes) { off = align8(off); u8* r = reinterpret_cast<u8*>(A) + off; off += bytes; return r; };
  const u32 Mp = next_pow2(M ? M : 1);
  o->mkey = (u64*)take(8ull * Mp); o->midx = (u32*)take(4ull * Mp);
  o->gcls_m = (u32*)take(4ull * (M + 1)); o->gcls_cnt = (u32*)take(4ull * (M + 1)); o->gcls_eoff = (u32*)take(4ull * (M + 2));
  return align8(off) <= (u64)hi * 4;
}
__host__ __device__ inline u64 ps_back_words_b(u64 M) {
  u64 Mp = 1; while (Mp < (M ? M : 1)) Mp <<= 1;
  return (12 * Mp + 12 * (M + 2) + 8 * 6 + 3) / 4;
}

// words of arena ps_back_carve needs for M molecules / Lm label words (same arithmetic)
__host__ __device__ inline u64 ps_back_words(u64 M, u64 Lm, u32 per) {
  auto p2 = [](u64 v) { u64 p = 1; while (p < v) p <<= 1; return p; };
  const u64 Mp = p2(M ? M : 1), Lp = p2(Lm ? Lm : 1), TK = Mp > Lp ? Mp : Lp, Sp = p2(Lm * per ? Lm * per : 1), Sr = Lm * per + 2;
  const u64 bytes = 8 * Mp + 4 * Mp + 4 * (M + 1) * 2 + 4 * (M + 2) + 8 * TK + 4 * (TK + 1) + 4 * (Lm + 1) + 4 * Sp + 4 * (Sr + 2) +
                    4 * Sr * 2 + 4 * (M + 1) + 4 * Sr * 2 + 8 * 16;   // (+ the align8 padding of the 15 arrays)
  return (bytes + 3) / 4;
}

// L2 prefetch of a cell's record arrays (one 128-byte line per thread and step)
__device__ __forceinline__ void ps_prefetch_lines(const void* lo, const void* hi) {
#ifndef AFQ_EMU
  const char* p = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(lo) & ~uintptr_t(127)) + (size_t)threadIdx.x * 128;
  for (; p < reinterpret_cast<const char*>(hi); p += (size_t)blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)lo; (void)hi;
#endif
}
__device__ __forceinline__ void ps_prefetch_cell(const KArgs& a, u32 cell) {
  const u64 r0 = a.cell_rec_off[cell], r1 = a.cell_rec_off[cell + 1];
  const u32 f0 = a.ref_off[r0], f1 = a.ref_off[r1];
  ps_prefetch_lines(a.refs + f0, a.refs + f1);
  ps_prefetch_lines(a.umi + r0, a.umi + r1);
  ps_prefetch_lines(a.ref_off + r0, a.ref_off + r1 + 1);
}

// =============================================================================================
// One cell. Returns false when the cell has to be redone by the global-arena kernel (nothing has
// been written for it in that case).
// =============================================================================================
// SPLIT (unique-only parsimony resolutions): the cell is only BUILT here — classes, vertices, components; singleton
// components are emitted straight into the cell's winner list in global memory and every multi-vertex component
// is exported (members + a descriptor on its size class's list) for the flat cover kernels of afq_pugc.cuh.
template <bool WIDE, bool SPLIT = false>
__device__ inline bool ps_cell(const KArgs& a, const GeArgs& g, u32 cell, u32* A, u32 AW, GeShared* sh, PsExtra* ex,
                               GePtrs* s_ptrs) {
  const u32 T = blockDim.x, tid = threadIdx.x;
  const u64 r0 = a.cell_rec_off[cell], r1 = a.cell_rec_off[cell + 1];
  const u32 n = (u32)(r1 - r0);
  const u32 f0 = a.ref_off[r0];
  const u32 P = a.ref_off[r1] - f0;
  const bool gene = g.ge_mode == GE_MODE_PUG_GENE;
  const bool usa = a.usa_mode != 0;
  // molecules (gene labels) instead of output slots: the EM resolutions, and --dump-eqclasses (the classes come out of ge_back).
  // In-kernel back end (em) or, on the split path, the global molecule pool consumed by k_pug_back (sem).
  const bool want_mol = g.only_unique == 0 || g.dump_ncls != nullptr;
  const bool em = !SPLIT && want_mol, sem = SPLIT && want_mol;

  // ---- arena layout, phase A --------------------------------------------------------------------
  // (shared-memory variants: refs / UMIs / record offsets arrive by bulk copy from 16-byte aligned addresses, so the
  // arrays start `lead` words into their 16-byte aligned regions)
#ifndef AFQ_EMU
  const bool bulk = !WIDE && (u64)f0 + P + 4 <= a.n_refs_total && r1 + 5 <= a.n_records && P > 0;
#else
  const bool bulk = false;
#endif
  const u32 lead_f = bulk ? (f0 & 3u) : 0u, lead_r = bulk ? (u32)(r0 & 3u) : 0u;
  const u32 refs_words = bulk ? ((lead_f + P + 3) & ~3u) : P;
  const u32 umi_words = bulk ? ((lead_r + n + 3) & ~3u) : n;
  const u32 roff_raw_words = (lead_r + n + 1 + 3) & ~3u;       // raw 32-bit offsets, staged in the (still unused) table area
  u32 off = 0;
  u32* refs_raw = A + off; off += refs_words;
  u32* refs = refs_raw + lead_f;
  u16* roff = WIDE ? nullptr : reinterpret_cast<u16*>(A + off);
  u32* roff32 = WIDE ? A + off : nullptr;
  off += WIDE ? n + 1 : (n + 2) / 2;
  u16* rlen = nullptr;
  if (gene) { rlen = reinterpret_cast<u16*>(A + off); off += (n + 1) / 2; }
  off = (off + 3) & ~3u;
  const u32 off_rec = off;                       // everything from here is re-carved after compaction
  u32* umi_raw = A + off; off += umi_words;
  u32* umi = umi_raw + lead_r;
  u16* rcls = reinterpret_cast<u16*>(A + off); off += (n + 1) / 2;
  off = (off + 3) & ~3u;
  const u32 TN = ps_table_size(n), tl2 = ilog2(TN), tmask = TN - 1;
  u32* tab = A + off; off += TN;
  const u32 off_dense = off;
  if (tid == 0) { ex->fail = (off_dense > AW || n >= PS_MAX_RECORDS || (!WIDE && P >= PS_MAX_REFS)) ? 1u : 0u; ex->n_mlist = 0; ex->n_win = 0;
                  sh->n_mol = 0; sh->lab_bump = 0; sh->alt = 0; sh->flag = 0; sh->cnt0 = sh->cnt1 = sh->cnt2 = sh->cnt3 = 0; }
  __syncthreads();
  if (ex->fail) { __syncthreads(); return false; }
  // ---- load: the cell's records are read from HBM once ----------------------------------------------
#ifndef AFQ_EMU
  if (bulk) {
    // three bulk copies (cp.async.bulk global -> shared, one thread issues them, the copy engine moves the data); every
    // thread waits on the mbarrier. ncu r2c: the three strided load loops held 10 % of k_pug_build's stall samples.
    if (tid == 0) {
      fence_proxy_async();                         // the previous cell's generic-proxy accesses of the arena come first
      mbar_expect_tx((u64*)&ex->bar, 4u * (refs_words + umi_words + roff_raw_words));
      bulk_g2s(refs_raw, a.refs + ((u64)f0 - lead_f), 4u * refs_words, (u64*)&ex->bar);
      bulk_g2s(umi_raw, a.umi + (r0 - lead_r), 4u * umi_words, (u64*)&ex->bar);
      bulk_g2s(tab, a.ref_off + (r0 - lead_r), 4u * roff_raw_words, (u64*)&ex->bar);
    }
    mbar_wait((u64*)&ex->bar, ex->n_bulk & 1u);
    for (u32 i = tid; i <= n; i += T) roff[i] = (u16)(tab[lead_r + i] - f0);
    __syncthreads();
    if (tid == 0) ex->n_bulk++;
    for (u32 i = tid; i < TN; i += T) tab[i] = PS_EMPTY;
    __syncthreads();
  } else
#endif
  {
    for (u32 i = tid; i < P; i += T) refs[i] = a.refs[(u64)f0 + i];
    for (u32 i = tid; i <= n; i += T) { if (WIDE) roff32[i] = a.ref_off[r0 + i] - f0; else roff[i] = (u16)(a.ref_off[r0 + i] - f0); }
    for (u32 i = tid; i < n; i += T) umi[i] = a.umi[r0 + i];
    for (u32 i = tid; i < TN; i += T) tab[i] = PS_EMPTY;
    __syncthreads();
  }
  PsCell c;
  c.a = &a; c.refs = refs; c.roff = roff; c.roff32 = roff32; c.rlen = rlen; c.gene = gene; c.vumi = nullptr; c.vinfo = nullptr; c.vgene = nullptr;
  if (gene) {   // sorted-dedup gene projection of every record, in place (src/eq_class.rs:742-744)
    GE_FOR(i, n) {
      u32* dst = refs + c.off(i);
      const u32 ln = c.off(i + 1) - c.off(i);
      for (u32 k = 0; k < ln; ++k) dst[k] = __ldg(a.t2g + dst[k]);
      const u32 gl = sort_dedup_small(dst, ln);
      if (gl > 0xFFFFu) ex->fail = 1;               // (16-bit label lengths; only reachable in variant 3)
      rlen[i] = (u16)gl;
    }
    __syncthreads();
    if (ex->fail) { __syncthreads(); return false; }
  }
  // ---- phase 1: record -> eq-class (representative record) ----------------------------------------
  GE_FOR(i, n) {
    const u32* li = c.lab(i);
    const u32 ln = c.len(i);
    u32 h = (ln + 1) * 0x9E3779B1u;
    for (u32 k = 0; k < ln; ++k) { h = (h ^ li[k]) * 0x85EBCA6Bu; h ^= h >> 13; }
    u32 s = (h * 0x9E3779B1u) >> (32 - tl2);
    for (;;) {
      u32 cur = *(volatile u32*)&tab[s];
      if (cur == PS_EMPTY) {
        cur = atomicCAS(&tab[s], PS_EMPTY, i);
        if (cur == PS_EMPTY) { rcls[i] = (u16)i; break; }
      }
      if (c.len(cur) == ln && label_equal(c.lab(cur), ln, li, ln)) { rcls[i] = (u16)cur; break; }
      s = (s + 1) & tmask;
    }
  }
  __syncthreads();
  for (u32 i = tid; i < TN; i += T) tab[i] = PS_EMPTY;
  __syncthreads();
  // ---- phase 2: (class, UMI) vertices; entry = representative record | reads << 16 ---------------
  GE_FOR(i, n) {
    const u32 cc = rcls[i], u = umi[i];
    u32 h = (u ^ (cc * 0x9E3779B1u)) * 0x85EBCA6Bu;
    h ^= h >> 15;
    u32 s = (h * 0xC2B2AE35u) >> (32 - tl2);
    for (;;) {
      u32 cur = *(volatile u32*)&tab[s];
      if (cur == PS_EMPTY) {
        cur = atomicCAS(&tab[s], PS_EMPTY, i | (1u << 16));
        if (cur == PS_EMPTY) break;
      }
      const u32 r = cur & 0xFFFFu;
      if (rcls[r] == cc && umi[r] == u) { atomicAdd(&tab[s], 1u << 16); break; }
      s = (s + 1) & tmask;
    }
  }
  __syncthreads();
  // ---- compaction: table slots -> dense vertices (appended behind the table) ----------------------
  u32 V;
  {
    const u32 K = ((TN + T - 1) / T) | 1u;        // odd chunk length: the threads' strided reads hit distinct banks
    u32 lo = tid * K; if (lo > TN) lo = TN;
    u32 hi = lo + K; if (hi > TN) hi = TN;
    u32 cnt = 0;
    for (u32 i = lo; i < hi; ++i) cnt += tab[i] != PS_EMPTY ? 1u : 0u;
    u32 pos = block_exscan(cnt, sh->scan, &V);
    // dense arrays + (EM) molecule offsets / lengths at the top of the arena must fit
    const u32 top = em ? AW - 2 * V : AW;
    if (off_dense + 3 * V > top || 3 * V > AW) return false;            // uniform: V is a block-wide value
    u32* vumi_w = A + off_dense;
    u32* vinfo_w = vumi_w + V;
    for (u32 i = lo; i < hi; ++i) {
      const u32 e = tab[i];
      if (e == PS_EMPTY) continue;
      const u32 r = e & 0xFFFFu;
      vumi_w[pos] = umi[r];
      vinfo_w[pos] = ((u32)rcls[r] << 16) | (e >> 16);
      ++pos;
    }
    c.vumi = vumi_w; c.vinfo = vinfo_w;
  }
  __syncthreads();
  const u32* vumi = c.vumi;
  // the gene of every vertex whose class maps to ONE gene (the usual case), looked up here with all
  // lanes busy: the emission inside the serialised cover rounds then needs no tid_to_gid gather
  // (ncu r1w: the winner lane's dependent __ldg chain sat on the cover's critical path)
  u32* vgene = A + off_dense + 2 * V;
  GE_FOR(v, V) {
    const u32 cv = c.vcls(v);
    const u32* lv = c.lab(cv);
    const u32 ln = c.len(cv);
    u32 g0 = NONE32;
    for (u32 k = 0; k < ln; ++k) {
      const u32 gg = c.gene_of(lv[k]);
      if (g0 == NONE32) g0 = gg;
      else if (gg != g0) { g0 = PS_MULTI_GENE; break; }
    }
    vgene[v] = g0;
    if (SPLIT && ln > 0xFFFFu) ex->fail = 1;          // (16-bit label lengths in the exported member records)
  }
  c.vgene = vgene;
  __syncthreads();

  // ---- re-carve: chains, union-find, Bloom bitmap, winners go into the dead record arrays
  // [off_rec, off_dense) first and into the free space behind the dense vertices after that ------
  u32 segA = off_rec, segB = off_dense + 3 * V;
  const u32 topB = em ? AW - 2 * V : AW;
  bool fits = true;
  auto alloc = [&](u32 words) -> u32* {
    if (segA + words <= off_dense) { u32* r = A + segA; segA += words; return r; }
    if (segB + words <= topB) { u32* r = A + segB; segB += words; return r; }
    fits = false;
    return A;
  };
  const u32 NU = pow2_ge(V + V / 2 + 2, 64), ul2 = ilog2(NU), umask = NU - 1;
  u32 BW = pow2_ge(2 * V, 64); if (BW > 4096) BW = 4096;
  u32* utab = alloc(NU);                              // later: the multi-vertex list, run starts
  u32* bloom = alloc(BW);
  u32* vnext = alloc(V);                              // later: root of every vertex
  u32* parent = alloc(V);                             // later: component sizes
  u32* winners = (em || SPLIT) ? A : alloc(next_pow2(V ? V : 1));
  u32* nxt = BW >= V ? bloom : alloc(V);              // component member lists (the bitmap is dead by then)
  // per-warp scratch of the warp-cooperative cover: 8 warps in the larger CTAs when that still fits
  u32 ncw = PS_COVER_WARPS;
  if (T >= 512 && (segA + 2 * PS_COVER_WARPS * PS_WSCR_WORDS <= off_dense || segB + 2 * PS_COVER_WARPS * PS_WSCR_WORDS + V + V / 2 + 2 * ((a.num_rows + 31) >> 5) + 64 <= topB))
    ncw = 2 * PS_COVER_WARPS;
  u32* swin = SPLIT ? alloc(V) : nullptr;             // SPLIT: the singleton components' slots, staged until the cell is known to succeed
  u32* wscr = SPLIT ? A : alloc(ncw * PS_WSCR_WORDS);
  u32* olist = SPLIT ? A : alloc(V / 2 + 2);          // components re-routed from the group cover to the warp cover
  if (!fits) return false;                            // uniform (V is block-wide)
  const u32 Wg = (a.num_rows + 31) >> 5;
  u32* gbm = nullptr;                                 // unique-only: presence bitmap + prefix over the output slots
  if (!em && !SPLIT) { gbm = alloc(2 * Wg); if (!fits) gbm = nullptr; }
  // EM: molecule labels grow down from the molecule offsets / lengths at the top of the arena
  PsSink sk;
  sk.mode = sem ? 3u : (em ? 2u : (usa ? 1u : 0u));
  if (sem) {
    sk.g_lab = g.ps_mlab + f0; sk.g_nlab = g.ps_nlab + cell; sk.g_nmol = g.ps_nwin + cell;
    sk.g_moff = g.ps_moff + r0; sk.g_mlen = g.ps_mlen + r0;
    if (tid == 0) { g.ps_nwin[cell] = 0; g.ps_nlab[cell] = 0; }      // (a barrier follows before the first molecule)
  }
  sk.uo = a.uo; sk.ao = a.ao; sk.A = A; sk.sh = sh; sk.ex = ex;
  sk.lab_lo = segB; sk.lab_hi = em ? topB : segB;
  sk.mol_off = A + (AW - V); sk.mol_len = A + (AW - 2 * V);
  // ---- phase 3: UMI -> vertex chains, Bloom bitmap -------------------------------------------------
  const u32 bmask = (32u * BW) - 1;
  for (u32 i = tid; i < NU; i += T) utab[i] = PS_EMPTY;
  for (u32 i = tid; i < BW; i += T) bloom[i] = 0;
  if (gbm) for (u32 i = tid; i < 2 * Wg; i += T) gbm[i] = 0;
  if (tid < 36) ex->szc[tid] = 0;
  __syncthreads();
  GE_FOR(v, V) {
    const u32 um = vumi[v];
    u32 s = ps_umi_slot(um, ul2);
    for (;;) {
      u32 cur = *(volatile u32*)&utab[s];
      if (cur == PS_EMPTY) {
        cur = atomicCAS(&utab[s], PS_EMPTY, v);
        if (cur == PS_EMPTY) { vnext[v] = PS_EMPTY; break; }
      }
      if (vumi[cur] == um) { vnext[v] = atomicExch(&utab[s], v); break; }
      s = (s + 1) & umask;
    }
    parent[v] = v;
    const u32 b1 = ps_h1(um) & bmask, b2 = ps_h2(um) & bmask;
    atomicOr(&bloom[b1 >> 5], 1u << (b1 & 31));
    atomicOr(&bloom[b2 >> 5], 1u << (b2 & 31));
  }
  __syncthreads();
  if (g.ge_mode == GE_MODE_CRLIKE) {
    // ---- cr-like molecules (cr-like-em): per UMI the genes with the largest read count -----------
    // get_num_molecules_cell_ranger_like (src/pugutils.rs:799-850) + resolver (src/pugutils.rs:644-749):
    // W(u, g) = reads of u in classes whose gene projection holds g; label = arg-max gene set. One
    // lane per UMI chain; a UMI with more than PS_CRL_GENES candidate genes sends the cell back.
    GE_FOR(sl, NU) {
      const u32 hd = utab[sl];
      if (hd == PS_EMPTY) continue;
      u32 gs[PS_CRL_GENES], ws[PS_CRL_GENES];
      u32 ng = 0;
      bool over = false;
      for (u32 w = hd; w != PS_EMPTY && !over; w = vnext[w]) {
        const u32 cw = c.vcls(w), nw_ = c.vcnt(w);
        const u32* lw = c.lab(cw);
        const u32 ln = c.len(cw);
        u32 touched = 0;                                   // genes of THIS class already credited (sorted-dedup projection)
        for (u32 k = 0; k < ln; ++k) {
          const u32 gg = c.gene_of(lw[k]);
          u32 q = 0;
          while (q < ng && gs[q] != gg) ++q;
          if (q == ng) {
            if (ng == PS_CRL_GENES) { over = true; break; }
            gs[ng] = gg; ws[ng] = 0; ++ng;
          }
          if (!(touched >> q & 1u)) { ws[q] += nw_; touched |= 1u << q; }
        }
      }
      if (over) { ex->fail = 1; continue; }
      u32 maxw = 0;
      for (u32 q = 0; q < ng; ++q) maxw = ws[q] > maxw ? ws[q] : maxw;
      u32 nb = 0;
      for (u32 q = 0; q < ng; ++q)
        if (ws[q] == maxw) {                               // keep the winners, ascending
          const u32 x = gs[q];
          u32 j = nb;
          while (j > 0 && gs[j - 1] > x) { gs[j] = gs[j - 1]; --j; }
          gs[j] = x;
          ++nb;
        }
      const u32 slot = ps_emit_genes(sk, gs, nb);
      if (slot != NONE32) {
        winners[atomicAdd(&ex->n_win, 1u)] = slot;
        if (gbm) atomicOr(&gbm[slot >> 5], 1u << (slot & 31));
      }
    }
  } else {
  // ---- phase 4: union-find over the PUG's edges (any edge type connects; src/pugutils.rs:278-301) --
  // (a) same UMI, another class sharing a reference: each chain pair once, from its earlier member
  GE_FOR(v, V) {
    const u32 cv = c.vcls(v);
    for (u32 w = vnext[v]; w != PS_EMPTY; w = vnext[w])
      if (c.related(cv, c.vcls(w))) uf_union(parent, v, w);
  }
  // (b) Hamming distance 1: every lane tests the 3*L substitutions of its own vertex's UMI against
  // the Bloom bitmap in straight-line code (hits collected in a mask), then the rare hits are
  // looked up in the chain table, the warp re-joining at a vote per round
  const u32 sub_len = g.pug_exact_umi ? 0u : g.umi_len;
  if (sub_len)
    for (u32 vb = 0; vb < V; vb += T) {
      const u32 v = vb + tid;
      const bool act = v < V;
      const u32 u = act ? vumi[v] : 0u, cv = act ? c.vcls(v) : 0u;
      const u32 h1 = ps_h1(u), h2 = ps_h2(u);
      unsigned long long hits = 0;
#pragma unroll
      for (u32 pos = 0; pos < 16; ++pos) {
        if (pos < sub_len) {
#pragma unroll
          for (u32 d = 1; d <= 3; ++d) {
            const u32 b1 = (h1 ^ ps_h1(d << (2 * pos))) & bmask, b2 = (h2 ^ ps_h2(d << (2 * pos))) & bmask;
            const u32 bit = (bloom[b1 >> 5] >> (b1 & 31)) & (bloom[b2 >> 5] >> (b2 & 31)) & 1u;
            hits |= (unsigned long long)bit << (pos * 3 + d - 1);
          }
        }
      }
      if (!act) hits = 0;
      while (__any_sync(0xFFFFFFFFu, hits != 0)) {
        if (hits) {
          const u32 k = (u32)__ffsll((long long)hits) - 1;
          hits &= hits - 1;
          const u32 cu = u ^ ((k % 3 + 1) << (2 * (k / 3)));
          u32 s = ps_umi_slot(cu, ul2);
          for (;;) {
            const u32 cur = utab[s];
            if (cur == PS_EMPTY) break;
            if (vumi[cur] == cu) {
              for (u32 w = cur; w != PS_EMPTY; w = vnext[w])
                if (w > v && c.related(cv, c.vcls(w))) uf_union(parent, v, w);   // each pair once
              break;
            }
            s = (s + 1) & umask;
          }
        }
      }
    }
  __syncthreads();
  // ---- phase 5: components ------------------------------------------------------------------------
  // members of a component hang on a list at its root; the roots of components with > 1 vertex are
  // counting-sorted by component size so that the lanes of a warp cover components of (nearly)
  // equal size (ncu r1t: one lane per sorted-list run start left 2-3 lanes active in the cover)
  u32* root = vnext;
  GE_FOR(v, V) root[v] = uf_find(parent, v);
  __syncthreads();
  u32* csz = parent;
  u32* head = utab;                 // [V]   (NU >= V + V/2 + 2)
  u32* clist = utab + V;            // [<= V/2] roots of the multi-vertex components, by size
  GE_FOR(v, V) { csz[v] = 0; head[v] = PS_EMPTY; }
  __syncthreads();
  GE_FOR(v, V) atomicAdd(&csz[root[v]], 1u);
  __syncthreads();
  for (u32 vb = 0; vb < V; vb += T) {
    const u32 v = vb + tid;
    u32 slot = NONE32;
    if (v < V) {
      const u32 r = root[v];
      const u32 sz = csz[r];
      if (sz == 1) {   // singleton component: the class label itself (src/pugutils.rs:1262-1322)
        const u32 gv = vgene[v];
        if (gv < PS_MULTI_GENE) slot = ps_emit_genes(sk, &gv, 1u);
        else {
          const u32 cv = c.vcls(v);
          slot = ps_emit(c, sk, c.lab(cv), c.len(cv), [](u32, u32) { return true; });
        }
      } else {
        if (!SPLIT) nxt[v] = atomicExch(&head[r], v);
        if (v == r) atomicAdd(&ex->szc[sz > SMALL_COMP ? SMALL_COMP + 1 : sz], 1u);
      }
    }
    const u32 wi = warp_bump(&ex->n_win, slot != NONE32);
    if (slot != NONE32) {
      if (SPLIT) swin[wi] = slot;     // (staged in shared memory: nothing may reach global memory before the cell is known to succeed)
      else { winners[wi] = slot; if (gbm) atomicOr(&gbm[slot >> 5], 1u << (slot & 31)); }
    }
  }
  __syncthreads();
  if (tid == 0) {
    u32 run = 0;
    for (u32 z = 0; z <= SMALL_COMP + 1; ++z) { const u32 t = ex->szc[z]; ex->szc[z] = run; run += t; }
    ex->n_mlist = run;
    // a component of more than 32 vertices (or beyond --large-graph-thresh): global-arena kernel
    if (ex->szc[SMALL_COMP + 1] != run) ex->fail = 1;
    for (u32 z = 2; z <= SMALL_COMP; ++z) if (z > g.large_graph_thresh && ex->szc[z] != (z == SMALL_COMP ? ex->szc[SMALL_COMP + 1] : ex->szc[z + 1])) ex->fail = 1;
  }
  __syncthreads();
  if (ex->fail) { __syncthreads(); return false; }
  const u32 K = ex->n_mlist;
  if (SPLIT) {
    // ---- export (nothing can fail from here on) ------------------------------------------------------
    const u32 r0w = (u32)r0;
    const u32 nsingle = sem ? 0u : ex->n_win;
    for (u32 i = tid; i < nsingle; i += T) g.ps_win[r0w + i] = swin[i];
    if (K) {
      GE_FOR(v, V) if (root[v] == v && csz[v] > 1) clist[atomicAdd(&ex->szc[csz[v]], 1u)] = v;
      __syncthreads();
      // after the scatter szc[z] = END of size z's range: size classes 2 | 3-4 | 5-8 | 9-32 are contiguous in clist
      const u32 Kc[5] = {0u, ex->szc[2], ex->szc[4], ex->szc[8], K};
      if (tid < 4) {   // reserve this cell's descriptor ranges on the four global lists
        const u32 cnt = Kc[tid + 1] - Kc[tid];
        ex->dbase[tid] = cnt ? atomicAdd(&a.ctl->desc_count[tid], cnt) : 0u;
      }
      // member offsets of the components inside the cell's member region [r0, r0 + n), in clist order
      u32 base = 0;
      for (u32 c0 = 0; c0 < K; c0 += T) {
        const u32 k = c0 + tid;
        const u32 sz = k < K ? csz[clist[k]] : 0u;
        u32 tot;
        const u32 exs = block_exscan(sz, sh->scan, &tot);
        if (k < K) { head[clist[k]] = base + exs; nxt[clist[k]] = 0; }
        base += tot;
      }
      __syncthreads();
      GE_FOR(k, K) {
        const u32 r = clist[k];
        const u32 cls = k < Kc[1] ? 0u : (k < Kc[2] ? 1u : (k < Kc[3] ? 2u : 3u));
        const u64 pos = (u64)g.ps_desc_base[cls] + ex->dbase[cls] + (k - Kc[cls]);
        g.ps_desc[2 * pos] = r0w + head[r];
        g.ps_desc[2 * pos + 1] = (csz[r] << 24) | cell;
      }
      const u32* labsrc = refs;
      GE_FOR(v, V) {
        const u32 r = root[v];
        if (csz[r] <= 1) continue;
        const u32 idx = r0w + head[r] + atomicAdd(&nxt[r], 1u);
        const u32 cv = c.vcls(v);
        const u32 ln = c.len(cv), lo = f0 + c.off(cv);
        u32* m = g.ps_mem + 4ull * idx;
        m[0] = vumi[v]; m[1] = (c.vcnt(v) << 16) | ln; m[2] = lo; m[3] = vgene[v];
        if (gene) for (u32 q = 0; q < ln; ++q) g.ps_glab[(u64)lo + q] = labsrc[c.off(cv) + q];   // projected labels (same values from every writer)
      }
    }
    __syncthreads();
    if (tid == 0 && !sem) g.ps_nwin[cell] = nsingle;      // (sem: the molecule counter has been running since the singletons)
    __syncthreads();
    return true;
  }
  if (K) {
    GE_FOR(v, V) if (root[v] == v && csz[v] > 1) clist[atomicAdd(&ex->szc[csz[v]], 1u)] = v;
    __syncthreads();
    // after the scatter szc[z] = END of size z's range
    const u32 K2 = ex->szc[2], K4 = ex->szc[4], K8 = ex->szc[8];
    if (tid == 0) { ex->n_over = 0; ex->next_w = 0; ex->next_g = 0; }
    __syncthreads();
    // The cover work is handed out to the warps from two shared-memory queues, most expensive items first
    // (warp-form components by decreasing size, then passes of the 8-, 4- and 2-lane group form): a static
    // split left warps 0-3 with three passes and warps 6-7 with one, and the barrier behind the cover held
    // 20 % of all stall samples (ncu r2a).
    const bool exact = g.pug_exact_umi != 0;
    const u32 wid = tid >> 5, nw = ncw;
    u32* wmem = wscr + (wid < nw ? wid : 0u) * PS_WSCR_WORDS;
    if (wid < nw)
      for (;;) {
        const u32 it = warp_claim(&ex->next_w);
        if (it >= K - K8) break;
        ps_cover_warp(c, sk, winners, head, nxt, clist[K - 1 - it], exact, gbm, wmem, wmem + 32);   // largest components first
      }
    const u32 n8 = (K8 - K4 + 3) / 4, n4 = (K4 - K2 + 7) / 8, n2 = (K2 + 15) / 16;
    for (;;) {
      const u32 it = warp_claim(&ex->next_g);
      if (it >= n8 + n4 + n2) break;
      if (it < n8) ps_cover_group<8>(c, sk, winners, gbm, head, nxt, clist, K4 + it * 4, K8, exact, olist, &ex->n_over);
      else if (it < n8 + n4) ps_cover_group<4>(c, sk, winners, gbm, head, nxt, clist, K2 + (it - n8) * 8, K4, exact, olist, &ex->n_over);
      else ps_cover_group<2>(c, sk, winners, gbm, head, nxt, clist, (it - n8 - n4) * 16, K2, exact, olist, &ex->n_over);
    }
    __syncthreads();
    const u32 KO = ex->n_over;                                   // small components with a long label
    if (tid == 0) ex->next_w = 0;
    __syncthreads();
    if (KO && wid < nw)
      for (;;) {
        const u32 it = warp_claim(&ex->next_w);
        if (it >= KO) break;
        ps_cover_warp(c, sk, winners, head, nxt, olist[it], exact, gbm, wmem, wmem + 32);
      }
  }
  }
  __syncthreads();
  if (ex->fail) { __syncthreads(); return false; }
  if (SPLIT) return true;        // (cr-like-em on the split path: the per-UMI molecules are in the global pool)

  const u64 out_base = f0;
  if (em) {
      // ---- EM resolutions: hand the molecules to the shared back end on arena pointers --------------
    const u32 M = sh->n_mol, Lm = sh->lab_bump;
    if (tid == 0) {
      GePtrs pp{};
      const bool ok = ps_back_carve(A, 0, sk.lab_hi - Lm, M, Lm, usa ? 3u : 1u, &pp);
      pp.mlab = A; pp.mol_off = sk.mol_off; pp.mol_len = sk.mol_len;
      *s_ptrs = pp;
      // (a separate flag word: ex->fail was just read by the other threads without a barrier behind)
      sh->flag = ok ? 0u : 1u;
    }
    __syncthreads();
    if (sh->flag) { __syncthreads(); return false; }
    GeCell gc(*s_ptrs);
    gc.a = &a; gc.g = &g; gc.scratch = nullptr; gc.scratch_budget = 0;
    gc.vk_s = nullptr; gc.vc_s = nullptr; gc.cl_off_s = nullptr; gc.cl_len_s = nullptr;
    gc.r0 = r0; gc.f0 = f0; gc.gene_labels = gene;
    ge_back(a, g, cell, gc, sh);
    return true;
  }
  // ---- unique-only resolutions: count the winners per output slot -----------------------------------
  const u32 m = ex->n_win;
  u32 nnz = 0, lmax = 0;
  if (m && gbm) {
    // presence bitmap over the output slots (bits were set as the winners were produced): a prefix
    // popcount ranks the expressed slots (= CSR column order), every winner bumps its slot's counter
    u32* gpre = gbm + Wg;
    u32* gcnt = utab;                  // [nnz <= m <= V] (lists are dead)
    {
      const u32 Kw = ((Wg + T - 1) / T) | 1u;      // (odd: conflict-free)
      u32 lo = tid * Kw; if (lo > Wg) lo = Wg;
      u32 hi = lo + Kw; if (hi > Wg) hi = Wg;
      u32 cnt = 0;
      for (u32 i = lo; i < hi; ++i) cnt += (u32)__popc(gbm[i]);
      u32 pos = block_exscan(cnt, sh->scan, &nnz);
      for (u32 i = lo; i < hi; ++i) { gpre[i] = pos; pos += (u32)__popc(gbm[i]); }
    }
    for (u32 i = tid; i < nnz; i += T) gcnt[i] = 0;
    __syncthreads();
    for (u32 i = tid; i < m; i += T) {
      const u32 v = winners[i];
      const u32 rank = gpre[v >> 5] + (u32)__popc(gbm[v >> 5] & ((1u << (v & 31)) - 1u));
      if (atomicAdd(&gcnt[rank], 1u) == 0) a.stage_col[out_base + rank] = v;
    }
    __syncthreads();
    const float mean = __fdiv_rn((float)m, (float)nnz);
    u32 lover = 0;
    for (u32 j = tid; j < nnz; j += T) {
      const u32 cn = gcnt[j];
      a.stage_val[out_base + j] = (float)cn;
      lmax = cn > lmax ? cn : lmax;
      if ((float)cn > mean) ++lover;
    }
    if (lmax) atomicMax(&sh->cnt3, lmax);
    if (lover) atomicAdd(&sh->cnt0, lover);
    __syncthreads();
  } else if (m) {
    // gene axis too wide for the arena: sort the winners, run-length count
    const u32 Mp = next_pow2(m);
    for (u32 i = m + tid; i < Mp; i += T) winners[i] = NONE32;
    __syncthreads();
    block_bitonic_u32(winners, Mp);
    u32* starts = utab;                // run starts (the lists are dead)
    u32 base = 0;
    for (u32 c0 = 0; c0 < m; c0 += T) {
      const u32 i = c0 + tid;
      const u32 st = (i < m && (i == 0 || winners[i - 1] != winners[i])) ? 1u : 0u;
      u32 tot;
      const u32 pos = block_exscan(st, sh->scan, &tot);
      if (st) starts[base + pos] = i;
      base += tot;
    }
    nnz = base;
    __syncthreads();
    for (u32 j = tid; j < nnz; j += T) {
      const u32 i0 = starts[j], i1 = (j + 1 < nnz) ? starts[j + 1] : m;
      a.stage_col[out_base + j] = winners[i0];
      a.stage_val[out_base + j] = (float)(i1 - i0);
      lmax = (i1 - i0) > lmax ? (i1 - i0) : lmax;
    }
    if (lmax) atomicMax(&sh->cnt3, lmax);
    __syncthreads();
    // NumGenesOverMean (src/quant.rs:1190-1194)
    const float mean = __fdiv_rn((float)m, (float)nnz);
    u32 lover = 0;
    for (u32 j = tid; j < nnz; j += T) {
      const u32 i0 = starts[j], i1 = (j + 1 < nnz) ? starts[j + 1] : m;
      if ((float)(i1 - i0) > mean) ++lover;
    }
    if (lover) atomicAdd(&sh->cnt0, lover);
    __syncthreads();
  }
  if (tid == 0) {
    a.sum_umi[cell] = (float)m;
    a.max_umi[cell] = (float)sh->cnt3;
    a.num_expr[cell] = nnz;
    a.num_over_mean[cell] = sh->cnt0;
    a.flags[cell] = nnz == 0 ? 4 : 0;
  }
  __syncthreads();
  return true;
}

constexpr int PS_LIST0 = NUM_BINS + 3;      // bin_list rows of the three arena variants

template <int VAR>
__global__ void __launch_bounds__(ps_threads(VAR), ps_min_blocks(VAR)) k_pug_smem(KArgs a, GeArgs g) {
  AFQ_DYN_SMEM(smem_raw);
  if (VAR >= PS_SMEM_VARIANTS && arena_cta_idle(a.ctl, a.ctl->ps3_blocks, a.ctl->bin_count[PS_LIST0 + VAR])) return;
  const u32 gwords = VAR < PS_SMEM_VARIANTS ? 0u : a.ctl->ps3_words;
  u32* A = VAR < PS_SMEM_VARIANTS ? reinterpret_cast<u32*>(smem_raw) : g.ps_garena + (u64)blockIdx.x * gwords;
  const u32 AWmax = VAR < PS_SMEM_VARIANTS ? ps_arena_words(VAR) : gwords;
  __shared__ GeShared sh;
  __shared__ PsExtra ex;
  __shared__ GePtrs s_ptrs;
  if (threadIdx.x == 0) {
    ex.n_bulk = 0;
#ifndef AFQ_EMU
    mbar_init((u64*)&ex.bar, 1);
#endif
  }
  const u32 count = a.ctl->bin_count[PS_LIST0 + VAR];
  const u32* list = a.bin_list + (u64)(PS_LIST0 + VAR) * a.n_cells;
  // jobs are claimed one ahead so that the NEXT cell's records can be prefetched into L2 while this
  // one is resolved (ncu r1y: 7 % of the stall samples sat on the three load loops, DRAM latency)
  if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[PS_LIST0 + VAR], 1u);
  __syncthreads();
  u32 job = sh.job;
  __syncthreads();
  while (job < count) {
    if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[PS_LIST0 + VAR], 1u);
    __syncthreads();
    const u32 next = sh.job;
    __syncthreads();
    if (next < count) ps_prefetch_cell(a, list[next]);
    const u32 cell = list[job];
    const u32 AW = (g.ps_limit_words && g.ps_limit_words < AWmax) ? g.ps_limit_words : AWmax;
    const bool ok = ps_cell<(VAR >= PS_SMEM_VARIANTS)>(a, g, cell, A, AW, &sh, &ex, &s_ptrs);
    if (!ok && threadIdx.x == 0) {
      const u32 idx = atomicAdd(&a.ctl->bin_count[GE_LIST_NORMAL], 1u);
      a.bin_list[(u64)GE_LIST_NORMAL * a.n_cells + idx] = cell;
    }
    __syncthreads();
    job = next;
  }
}

// The build-only form (SPLIT) of the same persistent loop: unique-only parsimony resolutions. Without the cover, the
// counting and the EM back end the kernel needs fewer registers and a fraction of the code (instruction cache).
__host__ __device__ constexpr u32 ps_build_min_blocks(int v) { return v == 0 ? (u32)AFQ_PS_V0_BLOCKS : (v == 1 ? 2u : (v == 2 ? 1u : 2u)); }
template <int VAR>
__global__ void __launch_bounds__(ps_threads(VAR), ps_build_min_blocks(VAR)) k_pug_build(KArgs a, GeArgs g) {
  AFQ_DYN_SMEM(smem_raw);
  if (VAR >= PS_SMEM_VARIANTS && arena_cta_idle(a.ctl, a.ctl->ps3_blocks, a.ctl->bin_count[PS_LIST0 + VAR])) return;
  const u32 gwords = VAR < PS_SMEM_VARIANTS ? 0u : a.ctl->ps3_words;
  u32* A = VAR < PS_SMEM_VARIANTS ? reinterpret_cast<u32*>(smem_raw) : g.ps_garena + (u64)blockIdx.x * gwords;
  const u32 AWmax = VAR < PS_SMEM_VARIANTS ? ps_arena_words(VAR) : gwords;
  __shared__ GeShared sh;
  __shared__ PsExtra ex;
  if (threadIdx.x == 0) {
    ex.n_bulk = 0;
#ifndef AFQ_EMU
    mbar_init((u64*)&ex.bar, 1);
#endif
  }
  const u32 count = a.ctl->bin_count[PS_LIST0 + VAR];
  const u32* list = a.bin_list + (u64)(PS_LIST0 + VAR) * a.n_cells;
  if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[PS_LIST0 + VAR], 1u);
  __syncthreads();
  u32 job = sh.job;
  __syncthreads();
  while (job < count) {
    if (threadIdx.x == 0) sh.job = atomicAdd(&a.ctl->bin_cursor[PS_LIST0 + VAR], 1u);
    __syncthreads();
    const u32 next = sh.job;
    __syncthreads();
    if (next < count) ps_prefetch_cell(a, list[next]);
    const u32 cell = list[job];
    const u32 AW = (g.ps_limit_words && g.ps_limit_words < AWmax) ? g.ps_limit_words : AWmax;
    const bool ok = ps_cell<(VAR >= PS_SMEM_VARIANTS), true>(a, g, cell, A, AW, &sh, &ex, nullptr);
    if (!ok && threadIdx.x == 0) {
      g.ps_nwin[cell] = NONE32;        // k_pug_count skips the cell: k_gene_eqc redoes it
      const u32 idx = atomicAdd(&a.ctl->bin_count[GE_LIST_NORMAL], 1u);
      a.bin_list[(u64)GE_LIST_NORMAL * a.n_cells + idx] = cell;
    }
    __syncthreads();
    job = next;
  }
}

// classify cells for the gene-eq-class resolutions: tiny cells go to the cr-like arenas
// (src/quant.rs:794-846); cells expected to fit a shared-memory arena go to k_pug_smem's lists
// (ps_mode bit 0: enabled, bit 1: gene-level labels, bit 2: EM, bit 3: global-arena variant available, bit 4: split form); the rest to the k_gene_eqc lists.
// ge_max_n / ge_max_p cover every non-tiny cell of their size class because k_pug_smem may hand
// any of its cells back to the k_gene_eqc list.
__global__ void k_bin_cells_ge(KArgs a, int force_bin, u32 big_records, u32 need_shift, u32 ps_mode) {
  const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n_cells) return;
  const u64 r0 = a.cell_rec_off[c], r1 = a.cell_rec_off[c + 1];
  const u64 n = r1 - r0;
  const u32 p = a.ref_off[r1] - a.ref_off[r0];
  int b;
  if (a.tiny_eligible && n < a.small_thresh) {
    const u64 need = (n < (u64)p ? n : (u64)p) << need_shift;
    b = NUM_SMEM_BINS;
#pragma unroll
    for (int i = NUM_SMEM_BINS - 1; i >= 0; --i)
      if (need <= (1ull << bin_cap_log2(i))) b = i;
    if (force_bin >= 0 && force_bin > b) b = force_bin < NUM_SMEM_BINS ? force_bin : NUM_SMEM_BINS;
    if (b == NUM_SMEM_BINS) atomicMax(&a.ctl->max_cell_refs, p);
  } else {
    const int w = n > big_records ? 0 : 1;
    b = w == 0 ? GE_LIST_BIG : GE_LIST_NORMAL;
    atomicMax(&a.ctl->ge_max_n[w], (u32)n);
    atomicMax(&a.ctl->ge_max_p[w], p);
    if (w == 1 && (ps_mode & 1u)) {
      const int v = ps_variant_for(n, p, (ps_mode & 2u) != 0, (ps_mode & 4u) != 0, a.usa_mode != 0, (ps_mode & 8u) != 0, (ps_mode & 16u) != 0);
      if (v >= 0) b = PS_LIST0 + v;
      if (v == 3) { atomicMax(&a.ctl->ps3_max_n, (u32)n); atomicMax(&a.ctl->ps3_max_p, p); }
    }
  }
  const u32 idx = atomicAdd(&a.ctl->bin_count[b], 1u);
  a.bin_list[(u64)b * a.n_cells + idx] = (u32)c;
}

}  // namespace afq
