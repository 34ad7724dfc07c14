// afq_emc.cuh — stage C of the EM resolutions on the split path, as a kernel of its own: k_em_cells.
//
// k_pug_back (afq_pugc.cuh) turns a cell's molecules into gene eq-classes in canonical order (stage B of ge_back) and
// leaves them in the cell's class regions (count / label offset at r0 + j, labels from f0 on — the --dump-eqclasses
// layout). This kernel takes it from there, one CTA per cell, with the algorithm of k_em_subset (afq_infer.cuh): support
// genes marked in a gene-axis bitmap whose prefix popcount is the local index (ascending gene order = CSR column order,
// no sort), E step one thread per class, M step one thread per gene gathering through a transposed list sorted by class
// position, so that every f32 sum follows the reference's order. Compared with ge_back's stage C (three block-wide
// bitonic sorts over power-of-two padded arrays, ~45 % of k_pug_back's samples on C4, ncu r2k) it sorts nothing and keeps
// ~2.5x less state, which is what lets USA cells stay in shared memory at 3 CTAs per SM.
//
// Reference semantics per mode:
//   gene mode, EM     em_optimize (M1, src/em.rs:487-582): always iterates, no last round
//   USA mode,  EM     extract_usa_eqmap (src/utils.rs:842-926) + em_optimize_subset (M2, src/em.rs:251-456)
//   gene mode, unique em_optimize(only_unique) (src/em.rs:499-514): tallies of the single-gene classes
//   USA mode,  unique extract_counts (src/utils.rs:673-756): the S | U | A tie rules on labels of <= 10 ids
#pragma once
#include "afq_pugc.cuh"
#include "afq_infer.cuh"     // EM thresholds (src/em.rs:28-34)

namespace afq {

constexpr u32 EC_THREADS = 256;
constexpr int EC_TIERS = 4;
__host__ __device__ constexpr u32 ec_arena_words(int tier) { return tier == 0 ? 6u * 1024u : (tier == 1 ? 13u * 1024u : 24u * 1024u); }
__host__ __device__ inline size_t ec_smem_bytes(int tier, u32 num_rows) { return 4ull * (2ull * ((num_rows + 31) / 32) + ec_arena_words(tier) + 8); }
// arena words of a cell with C classes, E label words, at most S support slots
__host__ __device__ inline u64 ec_need_words(u64 C, u64 E, u64 S, bool usa) { return (C + 2) + 2 * C + 3 * E + (usa ? 5 : 3) * S + (S + 2) + 64; }
__host__ __device__ inline u64 ec_support_bound(u64 E, u32 num_rows, bool usa) { const u64 s = E * (usa ? 3 : 1); return s < num_rows ? s : num_rows; }

// sort the cells of the k_pug_build lists into the arena tiers of k_em_cells
__global__ void __launch_bounds__(256) k_em_bin(KArgs a, GeArgs g) {
  u32 cum[PS_VARIANTS + 1];
  cum[0] = 0;
  for (int v = 0; v < PS_VARIANTS; ++v) cum[v + 1] = cum[v] + a.ctl->bin_count[PS_LIST0 + v];
  const bool usa = a.usa_mode != 0;
  for (u32 job = blockIdx.x * blockDim.x + threadIdx.x; job < cum[PS_VARIANTS]; job += gridDim.x * blockDim.x) {
    int v = 0;
    while (job >= cum[v + 1]) ++v;
    const u32 cell = a.bin_list[(u64)(PS_LIST0 + v) * a.n_cells + (job - cum[v])];
    if (g.ps_nwin[cell] == NONE32) continue;       // handed back to k_gene_eqc
    const u64 C = g.dump_ncls[cell], E = g.dump_nlab[cell];
    // the support is at most 3 slots per label entry in USA mode and typically ~2.4: the tier is chosen with an estimate and a
    // cell that turns out larger (exact size known once its bitmap is built) moves to the global-arena tier, which runs last
    u64 Sest = usa ? (5 * E) / 2 + 8 : E;
    const u64 Sb = ec_support_bound(E, a.num_rows, usa);
    if (Sest > Sb) Sest = Sb;
    const u64 need = ec_need_words(C, E, Sest, usa);
    const int tier = need <= ec_arena_words(0) ? 0 : (need <= ec_arena_words(1) ? 1 : (need <= ec_arena_words(2) ? 2 : 3));
    g.back_list[(u64)tier * a.n_cells + atomicAdd(&a.ctl->em_count[tier], 1u)] = cell;
  }
}

template <int TIER>
__global__ void __launch_bounds__(EC_THREADS) k_em_cells(KArgs a, GeArgs g) {
  AFQ_DYN_SMEM(smem_raw);
  u32* gbm = reinterpret_cast<u32*>(smem_raw);                 // [Wg] presence bitmap over the output slots
  const u32 Wg = (a.num_rows + 31) >> 5;
  u32* gpre = gbm + Wg;                                        // [Wg] support slots before word w
  if (TIER == 3 && arena_cta_idle(a.ctl, a.ctl->back_blocks, a.ctl->em_count[3])) return;
  const u32 gwords = TIER < 3 ? 0u : a.ctl->back_words;
  u32* A = TIER < 3 ? gpre + Wg : g.back_garena + (u64)blockIdx.x * gwords;
  __shared__ u32 s_scan[40];
  __shared__ u32 s_job, s_flag, s_needs_em, s_over, s_conv[2];
  __shared__ float s_sum, s_max;
  const u32 T = blockDim.x, tid = threadIdx.x;
  const bool usa = a.usa_mode != 0, only_unique = g.only_unique != 0;
  const u32 uo = a.uo, ao = a.ao;
  const u32 total = a.ctl->em_count[TIER];
  const u32* list = g.back_list + (u64)TIER * a.n_cells;
  for (;;) {
    if (tid == 0) s_job = atomicAdd(&a.ctl->em_cursor[TIER], 1u);
    __syncthreads();
    const u32 job = s_job;
    __syncthreads();
    if (job >= total) break;
    const u32 cell = list[job];
    const u64 r0 = a.cell_rec_off[cell];
    const u32 f0 = a.ref_off[r0];
    const u64 out_base = f0;
    const u32 C = g.dump_ncls[cell], Eraw = g.dump_nlab[cell];
    const u32* ccnt = g.dump_cnt + r0;
    const u32* roff = g.dump_off + r0;
    const u32* rlab = g.dump_lab + f0;
    const u32 AW = TIER < 3 ? ec_arena_words(TIER) : gwords;
    u32 off = 0;
    u32* coff = A + off; off += C + 2;                         // class -> first (re-mapped) label entry
    float* cinv = reinterpret_cast<float*>(A + off); off += C;
    u32* clen_raw = A + off; off += C;                         // raw label lengths
    u32* elab = A + off; off += Eraw;                          // label entries as OUTPUT SLOTS
    u32* eloc = A + off; off += Eraw;
    u32* tr = A + off; off += Eraw;
    if (TIER < 3 && off > AW) {      // (cannot happen: k_em_bin counted these words) -> global-arena tier
      if (tid == 0) g.back_list[3ull * a.n_cells + atomicAdd(&a.ctl->em_count[3], 1u)] = cell;
      __syncthreads();
      continue;
    }
    // ---- labels as output slots --------------------------------------------------------------------------------------------
    // gene mode: the gene ids themselves. USA, EM: extract_usa_eqmap — S id 2k -> k, U id 2k+1 -> G + k, an adjacent
    // (S, U) pair of one gene -> 2G + k. USA, unique-only: the label's single slot by the tie rules, or nothing.
    for (u32 j = tid; j < C; j += T) {
      const u32 lo = roff[j], hi = j + 1 < C ? roff[j + 1] : Eraw;
      clen_raw[j] = hi - lo;
    }
    if (tid == 0) { s_needs_em = 0; s_flag = 0; }
    __syncthreads();
    {
      u32 base = 0;
      for (u32 c0 = 0; c0 < C; c0 += T) {
        const u32 j = c0 + tid;
        u32 ln = 0;
        if (j < C) {
          const u32 n = clen_raw[j];
          const u32* lab = rlab + roff[j];
          if (!usa) ln = n;
          else if (only_unique) ln = (n <= 10 && usa_slot_for_label(lab, n, uo, ao) != NONE32) ? 1u : 0u;
          else if (n == 1) ln = 1;
          else for (u32 i = 0; i < n; ++i) { ++ln; if (d_is_spliced(lab[i]) && i + 1 < n && d_same_gene(lab[i], lab[i + 1])) ++i; }
        }
        u32 tot;
        const u32 ex = block_exscan(ln, s_scan, &tot);
        if (j < C) coff[j] = base + ex;
        base += tot;
      }
      if (tid == 0) coff[C] = base;
    }
    for (u32 i = tid; i < Wg; i += T) gbm[i] = 0;
    __syncthreads();
    auto mark = [&](u32 s) { atomicOr(&gbm[s >> 5], 1u << (s & 31)); };
    for (u32 j = tid; j < C; j += T) {
      const u32 n = clen_raw[j], e0 = coff[j], ln = coff[j + 1] - e0;
      const u32* lab = rlab + roff[j];
      if (ln > 1) s_needs_em = 1;
      if (!usa) { for (u32 k = 0; k < n; ++k) elab[e0 + k] = lab[k]; }
      else if (only_unique) { if (ln) elab[e0] = usa_slot_for_label(lab, n, uo, ao); }
      else if (n == 1) elab[e0] = d_is_spliced(lab[0]) ? (lab[0] >> 1) : uo + (lab[0] >> 1);
      else {
        u32 q = e0;
        for (u32 i = 0; i < n; ++i) {
          const u32 gn = lab[i];
          u32 idx = gn >> 1;
          if (d_is_spliced(gn)) { if (i + 1 < n && d_same_gene(gn, lab[i + 1])) { idx += ao; ++i; } }
          else idx += uo;
          elab[q++] = idx;
        }
      }
      for (u32 k = 0; k < ln; ++k) {
        const u32 s = elab[e0 + k];
        mark(s);
        if (usa && !only_unique) {      // src/em.rs:87-113: the sibling slots take part in get_abundance_for
          if (s >= ao) { mark(s - uo); mark(s - ao); }
          else if (s >= uo) mark(s + uo);
          else mark(s + ao);
        }
      }
    }
    __syncthreads();
    u32 S = 0;
    {   // prefix popcount over the bitmap words: one chunk per thread (odd length: conflict-free), ONE block scan
      const u32 Kw = ((Wg + T - 1) / T) | 1u;
      u32 lo = tid * Kw; if (lo > Wg) lo = Wg;
      u32 hi = lo + Kw; if (hi > Wg) hi = Wg;
      u32 cnt = 0;
      for (u32 i = lo; i < hi; ++i) cnt += (u32)__popc(gbm[i]);
      u32 pos = block_exscan(cnt, s_scan, &S);
      for (u32 i = lo; i < hi; ++i) { gpre[i] = pos; pos += (u32)__popc(gbm[i]); }
    }
    // the support-sized arrays, carved with the exact size
    if (TIER < 3 && ec_need_words(C, Eraw, S, usa) > AW) {      // larger than k_em_bin's estimate: the global-arena tier (runs last)
      if (tid == 0) g.back_list[3ull * a.n_cells + atomicAdd(&a.ctl->em_count[3], 1u)] = cell;
      __syncthreads();
      continue;
    }
    u32* gidx = A + off; off += S;
    float* a0 = reinterpret_cast<float*>(A + off); off += S;
    float* a1 = reinterpret_cast<float*>(A + off); off += S;
    u32* tr_off = A + off; off += S + 2;
    u32* tr_cur = reinterpret_cast<u32*>(a1);                  // (fill cursors: only before the first iteration)
    u32* sibA = nullptr; u32* sibB = nullptr;
    if (usa) { sibA = A + off; off += S; sibB = A + off; off += S; }
    float* a_in = a0;
    __syncthreads();
    auto rank_of = [&](u32 s) { return gpre[s >> 5] + (u32)__popc(gbm[s >> 5] & ((1u << (s & 31)) - 1u)); };
    for (u32 i = tid; i < Wg; i += T) {
      u32 w = gbm[i], r = gpre[i];
      while (w) { const u32 b = (u32)__ffs((int)w) - 1; w &= w - 1; gidx[r++] = (i << 5) + b; }
    }
    for (u32 s = tid; s < S; s += T) { a0[s] = 0.0f; tr_off[s] = 0; tr_cur[s] = 0; }
    __syncthreads();
    // ---- local indices, unique tallies, transposed lists ---------------------------------------------------------------
    for (u32 j = tid; j < C; j += T) {
      const u32 e0 = coff[j], ln = coff[j + 1] - e0;
      for (u32 k = 0; k < ln; ++k) {
        const u32 s = rank_of(elab[e0 + k]);
        eloc[e0 + k] = s;
        atomicAdd(&tr_off[s], 1u);
      }
      if (ln == 1) atomicAdd(&a_in[eloc[e0]], (float)ccnt[j]);     // whole numbers: exact, order-free
    }
    if (usa && !only_unique)
      for (u32 s = tid; s < S; s += T) {
        const u32 x = gidx[s];
        if (x >= ao) { sibA[s] = rank_of(x - uo); sibB[s] = rank_of(x - ao); }
        else if (x >= uo) { sibA[s] = rank_of(x + uo); sibB[s] = NONE32; }
        else { sibA[s] = rank_of(x + ao); sibB[s] = NONE32; }
      }
    __syncthreads();
    // M1 (gene mode) always iterates; M2 (USA) returns the tallies untouched when no class is ambiguous (src/em.rs:339-341)
    const bool run_em = !only_unique && (usa ? s_needs_em != 0 : true) && S > 0;
    if (run_em) {
      {
        u32 base = 0;
        for (u32 c0 = 0; c0 < S; c0 += T) {
          const u32 s = c0 + tid;
          const u32 v = s < S ? tr_off[s] : 0u;
          u32 tot;
          const u32 ex = block_exscan(v, s_scan, &tot);
          if (s < S) tr_off[s] = base + ex;
          base += tot;
        }
        if (tid == 0) tr_off[S] = base;
        __syncthreads();
      }
      for (u32 j = tid; j < C; j += T) {
        const u32 e0 = coff[j], ln = coff[j + 1] - e0;
        for (u32 k = 0; k < ln; ++k) { const u32 s = eloc[e0 + k]; tr[tr_off[s] + atomicAdd(&tr_cur[s], 1u)] = j; }
      }
      __syncthreads();
      const float uni = __fdiv_rn(1.0f, (float)g.num_alphas);
      for (u32 s = tid; s < S; s += T) {
        // every slot's classes in canonical class order (insertion sort of a short list): the f32 sums follow the reference's order
        const u32 b = tr_off[s], e = tr_off[s + 1];
        for (u32 i = b + 1; i < e; ++i) { const u32 x = tr[i]; u32 q = i; while (q > b && tr[q - 1] > x) { tr[q] = tr[q - 1]; --q; } tr[q] = x; }
        a_in[s] = g.em_init_uniform ? uni : __fmul_rn(__fadd_rn(a_in[s], 0.5f), 1e-3f);
      }
      __syncthreads();
      u32 it = 0;
      bool converged = true, last_round = false;
      float* a_out = a1;
      auto abund = [&](const float* av, u32 s) {       // get_abundance_for (src/em.rs:167-187)
        if (!usa) return av[s];
        if (gidx[s] >= ao) return __fadd_rn(__fadd_rn(av[sibA[s]], av[sibB[s]]), av[s]);
        return __fadd_rn(av[sibA[s]], av[s]);
      };
      if (tid == 0) { s_conv[0] = 1; s_conv[1] = 1; }
      __syncthreads();
      while (it < INF_MIN_ITER || (it < INF_MAX_ITER && !converged) || last_round) {
        // TWO barriers per iteration: the alphas ping-pong between two arrays and the "converged" flag alternates
        // between two words (the other one is re-armed while this one is in use).
        // E step: one thread per ambiguous class
        for (u32 j = tid; j < C; j += T) {
          const u32 e0 = coff[j], ln = coff[j + 1] - e0;
          if (ln > 1) {
            float denom = 0.0f;
            for (u32 k = 0; k < ln; ++k) denom = __fadd_rn(denom, abund(a_in, eloc[e0 + k]));
            cinv[j] = denom > 0.0f ? __fdiv_rn((float)ccnt[j], denom) : -1.0f;
          }
        }
        __syncthreads();
        if (tid == 0) s_conv[(it + 1) & 1u] = 1;
        for (u32 s = tid; s < S; s += T) {       // M step: every slot gathers its classes' contributions in class order
          float sum = 0.0f;
          const float ef = abund(a_in, s);
          for (u32 t = tr_off[s]; t < tr_off[s + 1]; ++t) {
            const u32 j = tr[t];
            if (coff[j + 1] - coff[j] == 1) sum = __fadd_rn(sum, (float)ccnt[j]);
            else if (cinv[j] >= 0.0f) sum = __fadd_rn(sum, __fmul_rn(ef, cinv[j]));
          }
          if (sum > INF_ALPHA_CHECK_CUTOFF && fabsf(__fadd_rn(a_in[s], -sum)) > INF_REL_DIFF_TOLERANCE) s_conv[it & 1u] = 0;
          a_out[s] = sum;
        }
        __syncthreads();
        converged = s_conv[it & 1u] != 0;
        { float* t = a_in; a_in = a_out; a_out = t; }
        ++it;
        if (usa) {     // M2: clamp, then one last round (src/em.rs:391-443)
          if (last_round) break;
          if (it >= INF_MIN_ITER && converged) {
            for (u32 s = tid; s < S; s += T) if (a_in[s] < INF_MIN_OUTPUT_ALPHA) a_in[s] = 0.0f;
            last_round = true;
            __syncthreads();
          }
        }
      }
      for (u32 s = tid; s < S; s += T) if (a_in[s] < INF_MIN_OUTPUT_ALPHA) a_in[s] = 0.0f;
      __syncthreads();
    }
    // ---- output: positive values, ascending slot ----------------------------------------------------------------------------
    u32 nnz = 0;
    for (u32 c0 = 0; c0 < S; c0 += T) {
      const u32 s = c0 + tid;
      const float v = s < S ? a_in[s] : 0.0f;
      const u32 keep = v > 0.0f ? 1u : 0u;
      u32 tot;
      const u32 pos = block_exscan(keep, s_scan, &tot);
      if (keep) { a.stage_col[out_base + nnz + pos] = gidx[s]; a.stage_val[out_base + nnz + pos] = v; }
      nnz += tot;
    }
    __syncthreads();
    if (tid == 0) {  // the reference's sequential f32 scan (src/quant.rs:1156-1168)
      float fs = 0.0f, fm = 0.0f;
      for (u32 i = 0; i < nnz; ++i) { const float v = a.stage_val[out_base + i]; fs = __fadd_rn(fs, v); fm = v > fm ? v : fm; }
      s_sum = fs; s_max = fm; s_over = 0;
    }
    __syncthreads();
    const float mean = __fdiv_rn(s_sum, (float)nnz);
    u32 lover = 0;
    for (u32 i = tid; i < nnz; i += T) if (a.stage_val[out_base + i] > mean) ++lover;
    if (lover) atomicAdd(&s_over, lover);
    __syncthreads();
    if (tid == 0) {
      a.sum_umi[cell] = s_sum; a.max_umi[cell] = s_max;
      a.num_expr[cell] = nnz; a.num_over_mean[cell] = s_over; a.flags[cell] = nnz == 0 ? 4 : 0;
    }
    __syncthreads();
  }
}

}  // namespace afq
