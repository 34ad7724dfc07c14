// afq_pugc.cuh — the split form of the parsimony path for the unique-only resolutions (parsimony, parsimony-gene):
//
//   k_pug_build<v>   (afq_pugs.cuh, ps_cell<.., SPLIT>)  one CTA per cell in shared memory: classes, vertices, union-find.
//                    Singleton components go straight to the cell's winner list; every multi-vertex component is
//                    EXPORTED: its members (UMI, reads, label offset / length, single-gene hint: 16 B each) into the
//                    cell's region of a global member pool and a descriptor onto one of four global lists by size.
//   k_pug_cover2     sizes 2:      one THREAD per component — a 2-vertex component is always ONE molecule whose label is the
//                                  intersection of the two class labels (whatever the edge direction, has_edge
//                                  src/pugutils.rs:76-99), so no cover has to run
//   k_pug_cover_g<G> sizes 3-4 / 5-8: G lanes per component (lane = start vertex, label-position bitmasks in registers)
//   k_pug_cover_w    sizes 9-32:   one warp per component
//                    — FLAT over all cells of the batch: no per-cell barrier, no idle warps behind a cell's largest
//                    component (ncu r2a: the barrier behind the in-CTA cover held 20 % of k_pug_smem's stall samples and
//                    the cover's register arrays set the whole kernel's 80 registers / 3 CTAs per SM)
//   k_pug_count      one CTA per cell: winners -> presence bitmap over the output slots -> prefix popcount ranks ->
//                    counts, staging row, featureDump statistics
//
// Reference behaviour: get_num_molecules' cover loop (src/pugutils.rs:1097-1261) + collapse_vertices
// (src/pugutils.rs:308-391), canonical start-vertex order (class label lexicographic, UMI) as in afq_pugs.cuh; the result
// is identical to k_pug_smem / k_gene_eqc by construction (tests compare all three against the oracle).
#pragma once
#include "afq_pugs.cuh"

namespace afq {

constexpr u32 PC_WSCR_WORDS = 160 + 32 + 256;     // per-warp scratch of the warp form: member arrays, out masks, 16 x 16 label masks
constexpr u32 PC_THREADS = 256;

struct PcMember { u32 umi, cnt, len, cls, gene; };   // cls = label offset = class identity
__device__ __forceinline__ PcMember pc_load(const u32* mem, u32 idx) {
  const uint4 m = *reinterpret_cast<const uint4*>(mem + 4ull * idx);
  PcMember r;
  r.umi = m.x; r.cnt = m.y >> 16; r.len = m.y & 0xFFFFu; r.cls = m.z; r.gene = m.w;
  return r;
}

struct PcCtx {
  const KArgs* a; const GeArgs* g;
  const u32* lab;      // label pool: the input refs (transcript level) or the projected gene labels
  bool exact;
  PsCell c;            // only gene_of / gene are used (ps_emit)
  PsSink sk;           // mode 0 (gene) / 1 (USA), uo, ao
};
__device__ __forceinline__ PcCtx pc_ctx(const KArgs& a, const GeArgs& g) {
  PcCtx x;
  x.a = &a; x.g = &g;
  const bool gene = g.ge_mode == GE_MODE_PUG_GENE;
  x.lab = gene ? g.ps_glab : a.refs;
  x.exact = g.pug_exact_umi != 0;
  x.c.a = &a; x.c.refs = nullptr; x.c.roff = nullptr; x.c.roff32 = nullptr; x.c.rlen = nullptr;
  x.c.vumi = nullptr; x.c.vinfo = nullptr; x.c.vgene = nullptr; x.c.gene = gene;
  const bool em = g.only_unique == 0 || g.dump_ncls != nullptr;     // molecules for k_pug_back instead of output slots
  x.sk.mode = em ? 3u : (a.usa_mode ? 1u : 0u); x.sk.uo = a.uo; x.sk.ao = a.ao;
  x.sk.g_lab = nullptr; x.sk.g_nlab = nullptr; x.sk.g_nmol = nullptr; x.sk.g_moff = nullptr; x.sk.g_mlen = nullptr;
  x.sk.A = nullptr; x.sk.lab_lo = x.sk.lab_hi = 0; x.sk.mol_off = x.sk.mol_len = nullptr; x.sk.sh = nullptr; x.sk.ex = nullptr;
  return x;
}
__device__ __forceinline__ void pc_put(const PcCtx& x, u32 cell, u32 slot) {
  if (slot == NONE32) return;
  const u32 r0 = (u32)x.a->cell_rec_off[cell];
  x.g->ps_win[r0 + atomicAdd(&x.g->ps_nwin[cell], 1u)] = slot;
}
// the sink of one cell: output slots (unique-only) or the cell's regions of the global molecule pool (EM)
__device__ __forceinline__ PsSink pc_sink(const PcCtx& x, u32 cell) {
  PsSink sk = x.sk;
  if (sk.mode == 3) {
    const u64 r0 = x.a->cell_rec_off[cell];
    const u32 f0 = x.a->ref_off[r0];
    sk.g_lab = x.g->ps_mlab + f0; sk.g_nlab = x.g->ps_nlab + cell; sk.g_nmol = x.g->ps_nwin + cell;
    sk.g_moff = x.g->ps_moff + r0; sk.g_mlen = x.g->ps_mlen + r0;
  }
  return sk;
}
// one molecule of cell `cell` whose transcript label is { li[k] : keep(k, li[k]) }; gv = its single gene if known
template <class Keep>
__device__ __forceinline__ void pc_emit(const PcCtx& x, u32 cell, const u32* li, u32 ln, u32 gv, bool gv_ok, Keep keep) {
  const PsSink sk = pc_sink(x, cell);
  const u32 slot = (gv < PS_MULTI_GENE && gv_ok) ? ps_emit_genes(sk, &gv, 1u) : ps_emit(x.c, sk, li, ln, keep);
  if (sk.mode < 2) pc_put(x, cell, slot);
}
// ... given as the positions `inter` of label (li, ln)
__device__ __forceinline__ void pc_emit_masked(const PcCtx& x, u32 cell, const u32* li, u32 ln, u32 inter, u32 gv) {
  pc_emit(x, cell, li, ln, gv, inter != 0, [&](u32 q, u32) { return ((inter >> q) & 1u) != 0; });
}

// ---- sizes 2: one thread per component -----------------------------------------------------------------------------
__global__ void __launch_bounds__(PC_THREADS) k_pug_cover2(KArgs a, GeArgs g) {
  const PcCtx x = pc_ctx(a, g);
  const u32 count = a.ctl->desc_count[0];
  const u32* dl = g.ps_desc + 2ull * g.ps_desc_base[0];
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const u32 ms = dl[2ull * i], cell = dl[2ull * i + 1] & 0xFFFFFFu;
    const PcMember A = pc_load(g.ps_mem, ms), B = pc_load(g.ps_mem, ms + 1);
    if (A.gene == B.gene && A.gene < PS_MULTI_GENE) pc_emit(x, cell, x.lab + A.cls, A.len, A.gene, true, [](u32, u32) { return true; });
    else if (A.cls == B.cls) pc_emit(x, cell, x.lab + A.cls, A.len, NONE32, false, [](u32, u32) { return true; });
    else {
      const u32* lb = x.lab + B.cls;
      const u32 nb = B.len;
      pc_emit(x, cell, x.lab + A.cls, A.len, NONE32, false, [&](u32, u32 t) { return sorted_contains(lb, nb, t); });
    }
  }
}

// ---- sizes 9-32 (and any component with a label beyond 32 transcripts): one warp per component ------------------------
__device__ __forceinline__ bool pc_canon_less(const PcCtx& x, u32 cx, u32 lx, u32 ux, u32 cy, u32 ly, u32 uy) {
  return cx == cy ? ux < uy : label_less(x.lab + cx, lx, x.lab + cy, ly);
}
// every lane of the warp calls; scr = PC_WSCR_WORDS words of per-warp shared scratch
__device__ inline void pc_cover_warp(const PcCtx& x, u32 ms, u32 s, u32 cell, u32* scr) {
  const u32 lane = lane_id();
  u32* s_cls = scr; u32* s_len = scr + 32; u32* s_umi = scr + 64; u32* s_cnt = scr + 96; u32* s_gen = scr + 128;
  u32* wam = scr + 160;
  u32* Mw = scr + 192;
  u32* t_cls = Mw; u32* t_len = Mw + 32; u32* t_umi = Mw + 64;       // unsorted copies for the rank sort (Mw is built afterwards)
  PcMember me{0, 0, 0, 0, 0};
  if (lane < s) { me = pc_load(x.g->ps_mem, ms + lane); t_cls[lane] = me.cls; t_len[lane] = me.len; t_umi[lane] = me.umi; }
  __syncwarp();
  u32 rank = 0;
  if (lane < s)   // rank sort into canonical order (a strict total order: (class, UMI) is unique)
    for (u32 j = 0; j < s; ++j) if (j != lane && pc_canon_less(x, t_cls[j], t_len[j], t_umi[j], me.cls, me.len, me.umi)) ++rank;
  __syncwarp();
  if (lane < s) { s_cls[rank] = me.cls; s_len[rank] = me.len; s_umi[rank] = me.umi; s_cnt[rank] = me.cnt; s_gen[rank] = me.gene; }
  __syncwarp();
  u32 ci = 0, ln = 0;
  const u32* li = nullptr;
  if (lane < s) { ci = s_cls[lane]; ln = s_len[lane]; li = x.lab + ci; }
  const bool masked = s <= 16 && !__any_sync(0xFFFFFFFFu, lane < s && ln > 32);
  if (masked) {
    for (u32 p = lane; p < s * s; p += 32) {      // all (i, j) pairs across the warp
      const u32 i = p / s, j = p - i * s;
      const u32 cI = s_cls[i], cJ = s_cls[j];
      const u32 lnI = s_len[i];
      u32 m = 0;
      if (i == j || cI == cJ) m = lnI >= 32 ? 0xFFFFFFFFu : ((1u << lnI) - 1u);
      else {
        const u32* lI = x.lab + cI;
        const u32* lJ = x.lab + cJ;
        const u32 lnJ = s_len[j];
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
    const u32 ui = s_umi[lane], ni = s_cnt[lane];
    for (u32 j = 0; j < s; ++j) {
      if (j == lane) continue;
      const u32 xx = ui ^ s_umi[j];
      const u32 hd = (u32)__popc((xx | (xx >> 1)) & 0x55555555u);
      if (x.exact ? hd != 0 : hd > 1) continue;
      const bool rel = masked ? Mw[lane * 16 + j] != 0
                              : (ci == s_cls[j] || sorted_share(li, ln, x.lab + s_cls[j], s_len[j]));
      if (rel && out_edge(hd, ni, s_cnt[j])) my_am |= 1u << j;
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
        if (!masked) {   // BFS over out-edges through uncovered vertices whose label holds my transcript k
          const u32 t = li[k];
          u32 vis = 1u << lane, fr = 1u << lane;
          got = fr;
          while (fr) {
            u32 nx = 0;
            for (u32 f = fr; f; f &= f - 1) {
              u32 cand = wam[(u32)__ffs((int)f) - 1] & unc & ~vis;
              vis |= cand;
              for (; cand; cand &= cand - 1) {
                const u32 j = (u32)__ffs((int)cand) - 1;
                if (s_cls[j] == ci || sorted_contains(x.lab + s_cls[j], s_len[j], t)) nx |= 1u << j;
              }
            }
            got |= nx;
            fr = nx;
          }
        } else {
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
      if (masked) {
        u32 inter = 0xFFFFFFFFu;
        for (u32 mm = mask; mm; mm &= mm - 1) inter &= Mw[lane * 16 + (u32)__ffs((int)mm) - 1];
        pc_emit_masked(x, cell, li, ln, inter, s_gen[lane]);
      } else {      // label = intersection of the MCC's class labels (src/pugutils.rs:1161-1188)
        const u32 first = (u32)__ffs((int)mask) - 1;
        const u32 cf = s_cls[first];
        const u32 rest = mask & (mask - 1);
        pc_emit(x, cell, x.lab + cf, s_len[first], NONE32, false, [&](u32, u32 t) {
          for (u32 r = rest; r; r &= r - 1) {
            const u32 j = (u32)__ffs((int)r) - 1;
            if (s_cls[j] != cf && !sorted_contains(x.lab + s_cls[j], s_len[j], t)) return false;
          }
          return true;
        });
      }
    }
    unc &= ~mask;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(PC_THREADS) k_pug_cover_w(KArgs a, GeArgs g) {
  __shared__ u32 s_scr[(PC_THREADS / 32) * PC_WSCR_WORDS];
  const PcCtx x = pc_ctx(a, g);
  const u32 count = a.ctl->desc_count[3];
  const u32* dl = g.ps_desc + 2ull * g.ps_desc_base[3];
  u32* scr = s_scr + (threadIdx.x >> 5) * PC_WSCR_WORDS;
  const u32 wpb = blockDim.x >> 5;
  for (u32 i = blockIdx.x * wpb + (threadIdx.x >> 5); i < count; i += gridDim.x * wpb) {
    // largest components first (the list is in arrival order; sizes 9-32 are rare)
    const u32 ms = dl[2ull * i], w1 = dl[2ull * i + 1];
    pc_cover_warp(x, ms, w1 >> 24, w1 & 0xFFFFFFu, scr);
    __syncwarp();
  }
}

// ---- sizes 3-4 (G = 4) and 5-8 (G = 8): G lanes per component, 32 / G components per warp pass ---------------------
// Lane `sub` of a group owns start vertex `sub`. Everything a BFS needs is precomputed as bitmasks RELATIVE TO THE
// LANE'S OWN LABEL: M[j] = positions k of my label whose transcript is in vertex j's label. The BFS of ALL my
// transcripts at once is then a fixed number of branch-free relaxations reach[j] |= reach[x] & M[j] over the
// out-edges x -> j, in registers.
template <int G, int CLS>
__global__ void __launch_bounds__(PC_THREADS) k_pug_cover_g(KArgs a, GeArgs g) {
  __shared__ u32 s_scr[(PC_THREADS / 32) * PC_WSCR_WORDS];
  const PcCtx x = pc_ctx(a, g);
  const u32 count = a.ctl->desc_count[CLS];
  const u32* dl = g.ps_desc + 2ull * g.ps_desc_base[CLS];
  u32* scr = s_scr + (threadIdx.x >> 5) * PC_WSCR_WORDS;
  const u32 lane = lane_id(), sub = lane % G, gbase = lane - sub;
  constexpr u32 PER = 32 / G;
  const u32 wpb = blockDim.x >> 5;
  const u32 passes = (count + PER - 1) / PER;
  for (u32 pass = blockIdx.x * wpb + (threadIdx.x >> 5); pass < passes; pass += gridDim.x * wpb) {   // warp-uniform
    const u32 k = pass * PER + lane / G;
    const bool valid = k < count;
    const u32 ms = valid ? dl[2ull * k] : 0u, w1 = valid ? dl[2ull * k + 1] : 0u;
    const u32 size = w1 >> 24, cell = w1 & 0xFFFFFFu;
    bool has = valid && sub < size;
    PcMember me{0, 0, 0, 0, 0};
    if (has) me = pc_load(g.ps_mem, ms + sub);
    const u32 ci = me.cls, ui = me.umi, ni = me.cnt, ln = me.len;
    const u32* li = x.lab + ci;
    __syncwarp();
    // a label that does not fit a 32-bit position mask: the whole component takes the warp form (below)
    const u32 overm = __ballot_sync(0xFFFFFFFFu, has && ln > 32);
    u32 over_groups = 0;                                   // bit q: group q of this pass goes to the warp form
#pragma unroll
    for (u32 q = 0; q < PER; ++q) if ((overm >> (q * G)) & ((1u << G) - 1u)) over_groups |= 1u << q;
    if ((over_groups >> (lane / G)) & 1u) has = false;
    const u32 full = ln >= 32 ? 0xFFFFFFFFu : ((1u << ln) - 1u);
    u32 M[G];
    u32 am = 0, rank = 0;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const bool hj = __shfl_sync(0xFFFFFFFFu, (u32)has, (int)gbase + j) != 0;
      const u32 cj = __shfl_sync(0xFFFFFFFFu, ci, (int)gbase + j);
      const u32 uj = __shfl_sync(0xFFFFFFFFu, ui, (int)gbase + j);
      const u32 nj = __shfl_sync(0xFFFFFFFFu, ni, (int)gbase + j);
      const u32 lnj = __shfl_sync(0xFFFFFFFFu, ln, (int)gbase + j);
      u32 m = 0;
      if (has && hj) {
        if ((u32)j == sub || cj == ci) m = full;
        else {
          const u32* lj = x.lab + cj;      // positions of my label present in j's: one merge of the two sorted lists
          for (u32 qa = 0, qb = 0; qa < ln && qb < lnj;) {
            const u32 xa = li[qa], yb = lj[qb];
            if (xa == yb) { m |= 1u << qa; ++qa; ++qb; }
            else if (xa < yb) ++qa;
            else ++qb;
          }
        }
        if ((u32)j != sub) {
          const u32 xx = ui ^ uj;
          const u32 hd = (u32)__popc((xx | (xx >> 1)) & 0x55555555u);
          // has_edge (src/pugutils.rs:76-99): classes must share a reference (m != 0), Hamming distance <= 1
          if ((x.exact ? hd == 0 : hd <= 1) && m != 0 && out_edge(hd, ni, nj)) am |= 1u << j;
          if (cj == ci ? uj < ui : label_less(x.lab + cj, lnj, li, ln)) ++rank;   // canonical order (class label, UMI)
        }
      }
      M[j] = m;
      __syncwarp();
    }
    u32 amx[G];
#pragma unroll
    for (int xj = 0; xj < G; ++xj) amx[xj] = __shfl_sync(0xFFFFFFFFu, am, (int)gbase + xj);
    u32 unc = (__ballot_sync(0xFFFFFFFFu, has) >> gbase) & ((1u << G) - 1u);
    while (__any_sync(0xFFFFFFFFu, unc != 0)) {
      u32 my_size = 0, my_mask = 1u << sub;
      const bool start = has && ((unc >> sub) & 1u);
      if (start) {
        u32 reach[G];
#pragma unroll
        for (int j = 0; j < G; ++j) reach[j] = (u32)j == sub ? full : 0u;
        for (int round = 0; round < G - 1; ++round) {     // a path has at most G - 1 edges; usually 1-2 rounds
          u32 changed = 0;
#pragma unroll
          for (int xj = 0; xj < G; ++xj) {
            const u32 ax = amx[xj] & unc;
#pragma unroll
            for (int j = 0; j < G; ++j) {
              const u32 add = ((ax >> j) & 1u) ? (reach[xj] & M[j] & ~reach[j]) : 0u;
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
          if (sz > my_size) { my_size = sz; my_mask = mk; }
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
        pc_emit_masked(x, cell, li, ln, inter, me.gene);
      }
      if (wm) unc &= ~mask;
      __syncwarp();
    }
    // components with a long label: the whole warp covers them one after the other (rare)
    for (u32 og = over_groups; og; og &= og - 1) {
      const u32 q = (u32)__ffs((int)og) - 1;
      const u32 oms = __shfl_sync(0xFFFFFFFFu, ms, (int)(q * G)), ow1 = __shfl_sync(0xFFFFFFFFu, w1, (int)(q * G));
      pc_cover_warp(x, oms, ow1 >> 24, ow1 & 0xFFFFFFu, scr);
      __syncwarp();
    }
  }
}

// ---- counting: one CTA per cell ----------------------------------------------------------------------------------------
// winners -> presence bitmap over the output slots -> prefix popcount ranks (= CSR column order) -> per-slot counters.
// Dynamic shared memory: 2 * ceil(num_rows / 32) + PC_MAX_WINNERS words.
__host__ __device__ inline size_t pc_count_smem_bytes(u32 num_rows) { return 4ull * (2ull * ((num_rows + 31) / 32) + PC_MAX_WINNERS + 8); }
__global__ void __launch_bounds__(PC_THREADS) k_pug_count(KArgs a, GeArgs g) {
  AFQ_DYN_SMEM(smem_raw);
  u32* gbm = reinterpret_cast<u32*>(smem_raw);
  const u32 Wg = (a.num_rows + 31) >> 5;
  u32* gpre = gbm + Wg;
  u32* gcnt = gpre + Wg;
  __shared__ u32 s_scan[40];
  __shared__ u32 s_job, s_max, s_over;
  const u32 T = blockDim.x, tid = threadIdx.x;
  // the cells of the four k_pug_build lists, in list order
  u32 cum[PS_VARIANTS + 1];
  cum[0] = 0;
  for (int v = 0; v < PS_VARIANTS; ++v) cum[v + 1] = cum[v] + a.ctl->bin_count[PS_LIST0 + v];
  for (;;) {
    if (tid == 0) s_job = atomicAdd(&a.ctl->count_cursor, 1u);
    __syncthreads();
    const u32 job = s_job;
    __syncthreads();
    if (job >= cum[PS_VARIANTS]) break;
    int v = 0;
    while (job >= cum[v + 1]) ++v;
    const u32 cell = a.bin_list[(u64)(PS_LIST0 + v) * a.n_cells + (job - cum[v])];
    const u32 m = g.ps_nwin[cell];
    if (m == NONE32) continue;                     // handed back to k_gene_eqc (block-uniform)
    const u64 r0 = a.cell_rec_off[cell];
    const u32* win = g.ps_win + r0;
    const u64 out_base = a.ref_off[r0];
    for (u32 i = tid; i < Wg; i += T) gbm[i] = 0;
    if (tid == 0) { s_max = 0; s_over = 0; }
    __syncthreads();
    for (u32 i = tid; i < m; i += T) { const u32 s = win[i]; atomicOr(&gbm[s >> 5], 1u << (s & 31)); }
    __syncthreads();
    u32 nnz = 0;
    {   // prefix popcount: one chunk of words per thread (odd length: conflict-free), ONE block scan
      const u32 Kw = ((Wg + T - 1) / T) | 1u;
      u32 lo = tid * Kw; if (lo > Wg) lo = Wg;
      u32 hi = lo + Kw; if (hi > Wg) hi = Wg;
      u32 pc = 0;
      for (u32 i = lo; i < hi; ++i) pc += (u32)__popc(gbm[i]);
      u32 pos = block_exscan(pc, s_scan, &nnz);
      for (u32 i = lo; i < hi; ++i) { gpre[i] = pos; pos += (u32)__popc(gbm[i]); }
    }
    // the per-slot counters: shared memory, or for a giant cell (more molecules than PC_MAX_WINNERS) the cell's region of
    // the member pool, which is dead once the cover kernels are done (4 words per record >= nnz)
    u32* cnt = m <= PC_MAX_WINNERS ? gcnt : g.ps_mem + 4ull * r0;
    for (u32 i = tid; i < nnz; i += T) cnt[i] = 0;
    __syncthreads();
    for (u32 i = tid; i < m; i += T) {
      const u32 s = win[i];
      const u32 rank = gpre[s >> 5] + (u32)__popc(gbm[s >> 5] & ((1u << (s & 31)) - 1u));
      if (atomicAdd(&cnt[rank], 1u) == 0) a.stage_col[out_base + rank] = s;
    }
    __syncthreads();
    const float mean = __fdiv_rn((float)m, (float)nnz);     // NumGenesOverMean (src/quant.rs:1190-1194)
    u32 lmax = 0, lover = 0;
    for (u32 j = tid; j < nnz; j += T) {
      const u32 cn = cnt[j];
      a.stage_val[out_base + j] = (float)cn;
      lmax = cn > lmax ? cn : lmax;
      if ((float)cn > mean) ++lover;
    }
    if (lmax) atomicMax(&s_max, lmax);
    if (lover) atomicAdd(&s_over, lover);
    __syncthreads();
    if (tid == 0) {
      a.sum_umi[cell] = (float)m;
      a.max_umi[cell] = (float)s_max;
      a.num_expr[cell] = nnz;
      a.num_over_mean[cell] = s_over;
      a.flags[cell] = nnz == 0 ? 4 : 0;
    }
    __syncthreads();
  }
}

// ---- EM resolutions: one CTA per cell runs the shared back end (ge_back, afq_pug.cuh: molecules -> gene eq-classes in
// canonical order -> counts / EM -> staging row + statistics) on the cell's molecules in the global pool. The back end's
// arrays are carved from an arena sized by what the cell holds: k_back_bin sorts the cells into four tiers by
// ps_back_words(molecules, label words) — 48 KB x 4 CTAs per SM, 100 KB x 2, 224 KB x 1 of shared memory, and per-CTA
// global arenas for the rest — and the four k_pug_back<tier> launches run side by side on lanes.
constexpr u32 PB_THREADS = 256;
constexpr int PB_TIERS = 4;
__host__ __device__ constexpr u32 pb_arena_words(int tier) { return tier == 0 ? 12u * 1024u : (tier == 1 ? 25u * 1024u : 56u * 1024u); }

__global__ void __launch_bounds__(256) k_back_bin(KArgs a, GeArgs g) {
  u32 cum[PS_VARIANTS + 1];
  cum[0] = 0;
  for (int v = 0; v < PS_VARIANTS; ++v) cum[v + 1] = cum[v] + a.ctl->bin_count[PS_LIST0 + v];
  const u32 per = a.usa_mode ? 3u : 1u;
  for (u32 job = blockIdx.x * blockDim.x + threadIdx.x; job < cum[PS_VARIANTS]; job += gridDim.x * blockDim.x) {
    int v = 0;
    while (job >= cum[v + 1]) ++v;
    const u32 cell = a.bin_list[(u64)(PS_LIST0 + v) * a.n_cells + (job - cum[v])];
    const u32 M = g.ps_nwin[cell];
    if (M == NONE32) continue;                     // handed back by k_pug_build
    const u64 need = g.classes_only ? ps_back_words_b(M) : ps_back_words(M, g.ps_nlab[cell], per);
    // shared-memory tiers up to g.back_max_tier; beyond: per-CTA global arenas at full occupancy (a 100 / 224 KB arena
    // means 2 / 1 CTAs per SM, which only pays when the cell's EM state would otherwise thrash L2)
    const int mt = (int)g.back_max_tier;
    const int tier = need <= pb_arena_words(0) ? 0 : ((mt >= 1 && need <= pb_arena_words(1)) ? 1 : ((mt >= 2 && need <= pb_arena_words(2)) ? 2 : 3));
    g.back_list[(u64)tier * a.n_cells + atomicAdd(&a.ctl->back_count[tier], 1u)] = cell;
  }
}

template <int TIER>
__global__ void __launch_bounds__(PB_THREADS) k_pug_back(KArgs a, GeArgs g) {
  AFQ_DYN_SMEM(smem_raw);
  if (TIER == 3 && arena_cta_idle(a.ctl, a.ctl->back_blocks, a.ctl->back_count[3])) return;
  const u32 gwords = TIER < 3 ? 0u : a.ctl->back_words;
  u32* A = TIER < 3 ? reinterpret_cast<u32*>(smem_raw) : g.back_garena + (u64)blockIdx.x * gwords;
  const u32 AW = TIER < 3 ? pb_arena_words(TIER) : gwords;
  __shared__ GeShared sh;
  __shared__ GePtrs s_ptrs;
  __shared__ u32 s_job;
  const u32 tid = threadIdx.x;
  const u32 total = a.ctl->back_count[TIER];
  const u32* list = g.back_list + (u64)TIER * a.n_cells;
  for (;;) {
    if (tid == 0) s_job = atomicAdd(&a.ctl->back_cursor[TIER], 1u);
    __syncthreads();
    const u32 job = s_job;
    __syncthreads();
    if (job >= total) break;
    const u32 cell = list[job];
    const u32 M = g.ps_nwin[cell], Lm = g.ps_nlab[cell];
    const u64 r0 = a.cell_rec_off[cell];
    const u32 f0 = a.ref_off[r0];
    if (tid == 0) {
      GePtrs pp{};
      const bool ok = g.classes_only ? ps_back_carve_b(A, 0, AW, M, &pp) : ps_back_carve(A, 0, AW, M, Lm, a.usa_mode ? 3u : 1u, &pp);
      pp.mlab = g.ps_mlab + f0; pp.mol_off = g.ps_moff + r0; pp.mol_len = g.ps_mlen + r0;
      s_ptrs = pp;
      sh.flag = ok ? 0u : 1u;
      sh.n_mol = M; sh.lab_bump = Lm; sh.alt = 0;
      sh.cnt0 = sh.cnt1 = sh.cnt2 = sh.cnt3 = 0;
    }
    __syncthreads();
    if (sh.flag) {      // (cannot happen: the tier was chosen with the same arithmetic; kept as a safety net) -> k_gene_eqc
      if (tid == 0) {
        g.ps_nwin[cell] = NONE32;
        const u32 idx = atomicAdd(&a.ctl->bin_count[GE_LIST_NORMAL], 1u);
        a.bin_list[(u64)GE_LIST_NORMAL * a.n_cells + idx] = cell;
      }
      __syncthreads();
      continue;
    }
    GeCell gc(s_ptrs);
    gc.a = &a; gc.g = &g; gc.scratch = nullptr; gc.scratch_budget = 0;
    gc.vk_s = nullptr; gc.vc_s = nullptr; gc.cl_off_s = nullptr; gc.cl_len_s = nullptr;
    gc.r0 = r0; gc.f0 = f0; gc.gene_labels = g.ge_mode == GE_MODE_PUG_GENE;
    ge_back(a, g, cell, gc, &sh);
    __syncthreads();
  }
}

}  // namespace afq
