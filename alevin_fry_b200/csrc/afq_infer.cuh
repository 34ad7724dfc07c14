// afq_infer.cuh — EM over a GLOBAL gene-eq-class table: the kernel behind `alevin-fry infer` (src/infer.rs:31-241) and the
// C-ABI call afq_infer. One CTA per cell runs em_optimize_subset (src/em.rs:251-456, thresholds src/em.rs:28-34):
//
//   rows     the cell's (eq-class id, count) row of the count matrix is staged into shared memory with ONE bulk-copy
//            (cp.async.bulk global -> shared, completion on an mbarrier: the copy engine moves the row, no thread issues loads)
//   support  the genes the cell's classes touch (+ their spliced / unspliced / ambiguous siblings in USA mode,
//            src/em.rs:87-113) are marked in a presence bitmap over the gene axis; a prefix popcount gives every gene its
//            LOCAL index, ascending in gene id — so the result row comes out in CSR column order without a sort
//   E step   one thread per multi-gene class: denominator = sum of the effective abundances of its genes, in label order
//   M step   one thread per support gene GATHERS its contributions through a transposed (gene -> classes) list sorted
//            by class position: the f32 sums are taken in exactly the order of the reference's scatter loop
//            (classes in row order, genes in label order), with one rounding per multiply / add / divide
//   output   genes with alpha > 0 (after the MIN_OUTPUT_ALPHA clamp), ascending -> staging row + statistics
//
// Cells whose arrays do not fit the shared-memory arena run the same code on a per-CTA global arena.
#pragma once
#include "afq_kernels.cuh"

namespace afq {

constexpr u32 INF_THREADS = 256;
constexpr u32 INF_ARENA_WORDS = 24 * 1024;        // 96 KB of dynamic shared memory behind the gene-axis bitmap
constexpr float INF_MIN_OUTPUT_ALPHA = 0.01f, INF_ALPHA_CHECK_CUTOFF = 1e-2f, INF_REL_DIFF_TOLERANCE = 1e-2f;
constexpr u32 INF_MIN_ITER = 2, INF_MAX_ITER = 100;

struct InferArgs {
  u64 n_cells;
  const u64* cell_off;     // [n_cells + 1] rows of the count matrix (CSR over cells)
  const u32* cell_eq;      // eq-class ids, in row order
  const u32* cell_cnt;     // counts
  const u32* lab_off;      // [n_classes + 1]
  const u32* labels;       // gene indices of every class (USA: already in the S | U | A column space)
  u32 num_alphas, usa, uo, ao, init_uniform, only_unique;
  const u64* stage_off;    // [n_cells + 1] staging offsets (an upper bound of every row's length)
  u32* stage_col; float* stage_val;
  float* sum_umi; float* max_umi; u32* num_expr; u32* num_over_mean; u8* flags;
  u32* cursor;             // work cursor (zeroed by the host)
  u32* garena; u64 garena_words;   // per-CTA global arenas for the cells that exceed INF_ARENA_WORDS
};

// words of arena a cell with C classes, E label entries and at most S support genes needs
__host__ __device__ inline u64 inf_need_words(u64 C, u64 E, u64 S, bool usa) {
  return 2 * (C + 12) + (C + 2) + C + 2 * E + (usa ? 6 : 4) * S + (S + 2) + 64;
}
__host__ __device__ inline size_t inf_smem_bytes(u32 num_alphas) { return 4ull * (2ull * ((num_alphas + 31) / 32) + INF_ARENA_WORDS + 8); }


__global__ void __launch_bounds__(INF_THREADS) k_em_subset(InferArgs p) {
  AFQ_DYN_SMEM(smem_raw);
  u32* gbm = reinterpret_cast<u32*>(smem_raw);                 // [Wg] presence bitmap over the gene axis
  const u32 Wg = (p.num_alphas + 31) >> 5;
  u32* gpre = gbm + Wg;                                        // [Wg] support genes before word w
  u32* sarena = gpre + Wg + ((4 - ((2 * Wg) & 3)) & 3);        // 16-byte aligned
  __shared__ u32 s_scan[40];
  __shared__ u32 s_job, s_flag, s_needs_em, s_nnz, s_max_bits, s_over;
  __shared__ float s_sum;
  __shared__ __align__(8) u64 s_bar;
  const u32 T = blockDim.x, tid = threadIdx.x;
  u32 phase = 0;
#ifndef AFQ_EMU
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
#endif
  for (;;) {
    if (tid == 0) s_job = atomicAdd(p.cursor, 1u);
    __syncthreads();
    const u32 cell = s_job;
    __syncthreads();
    if (cell >= p.n_cells) break;
    const u64 o0 = p.cell_off[cell], o1 = p.cell_off[cell + 1];
    const u32 C = (u32)(o1 - o0);
    const u64 out_base = p.stage_off[cell];
    if (C == 0) {
      if (tid == 0) { p.sum_umi[cell] = 0; p.max_umi[cell] = 0; p.num_expr[cell] = 0; p.num_over_mean[cell] = 0; p.flags[cell] = 4; }
      continue;
    }
    // ---- arena: shared memory when the cell fits (decided from the staging bound the host computed the same way) ----
    const u64 Sb = p.stage_off[cell + 1] - out_base;             // upper bound of the support size
    // E is not known yet: the host guarantees need <= INF_ARENA_WORDS for cells it marked small by giving them
    // stage bounds computed from E; here the exact E is derived first with the class lengths read from global memory
    u32 Eloc = 0;
    for (u32 j = tid; j < C; j += T) { const u32 q = p.cell_eq[o0 + j]; Eloc += p.lab_off[q + 1] - p.lab_off[q]; }
    u32 Etot;
    block_exscan(Eloc, s_scan, &Etot);
    const u32 E = Etot;
    const bool in_smem = inf_need_words(C, E, Sb, p.usa != 0) <= INF_ARENA_WORDS;
    u32* A = in_smem ? sarena : p.garena + (u64)blockIdx.x * p.garena_words;
    u32 off = 0;
    const u32 lead = (u32)(o0 & 3);                              // rows are staged from a 16-byte aligned address
    const u32 rowlen = (lead + C + 3) & ~3u;
    u32* ceq_raw = A + off; off += rowlen + 4;
    u32* ccnt_raw = A + off; off += rowlen + 4;
    u32* coff = A + off; off += C + 2;
    float* cinv = reinterpret_cast<float*>(A + off); off += C;
    u32* eloc = A + off; off += E;
    u32* tr = A + off; off += E;
    u32* gidx = A + off; off += (u32)Sb;
    float* a_in = reinterpret_cast<float*>(A + off); off += (u32)Sb;
    float* eff = reinterpret_cast<float*>(A + off); off += (u32)Sb;
    u32* tr_off = A + off; off += (u32)Sb + 2;
    u32* tr_cur = A + off; off += (u32)Sb;
    u32* sibA = nullptr; u32* sibB = nullptr;
    if (p.usa) { sibA = A + off; off += (u32)Sb; sibB = A + off; off += (u32)Sb; }
    // ---- stage the row ---------------------------------------------------------------------------------------------------
#ifndef AFQ_EMU
    if (in_smem) {
      if (tid == 0) {
        fence_proxy_async();                             // earlier generic-proxy accesses of the arena are ordered first
        mbar_expect_tx(&s_bar, 2 * rowlen * 4);
        bulk_g2s(ceq_raw, p.cell_eq + (o0 - lead), rowlen * 4, &s_bar);
        bulk_g2s(ccnt_raw, p.cell_cnt + (o0 - lead), rowlen * 4, &s_bar);
      }
      mbar_wait(&s_bar, phase);
      phase ^= 1;
    } else
#endif
    {
      for (u32 j = tid; j < C; j += T) { ceq_raw[lead + j] = p.cell_eq[o0 + j]; ccnt_raw[lead + j] = p.cell_cnt[o0 + j]; }
      __syncthreads();
    }
    const u32* ceq = ceq_raw + lead;
    const u32* ccnt = ccnt_raw + lead;
    // ---- class offsets, support bitmap ---------------------------------------------------------------------------------
    {
      u32 base = 0;
      for (u32 c0 = 0; c0 < C; c0 += T) {
        const u32 j = c0 + tid;
        const u32 ln = j < C ? p.lab_off[ceq[j] + 1] - p.lab_off[ceq[j]] : 0u;
        u32 tot;
        const u32 ex = block_exscan(ln, s_scan, &tot);
        if (j < C) coff[j] = base + ex;
        base += tot;
      }
      if (tid == 0) { coff[C] = base; s_needs_em = 0; s_flag = 0; }
    }
    for (u32 i = tid; i < Wg; i += T) gbm[i] = 0;
    __syncthreads();
    auto mark = [&](u32 g) { atomicOr(&gbm[g >> 5], 1u << (g & 31)); };
    for (u32 j = tid; j < C; j += T) {
      const u32 lo = p.lab_off[ceq[j]], ln = coff[j + 1] - coff[j];
      if (ln > 1) s_needs_em = 1;
      for (u32 k = 0; k < ln; ++k) {
        const u32 g = p.labels[lo + k];
        mark(g);
        if (p.usa) {      // src/em.rs:87-113: the sibling slots of a gene take part in get_abundance_for
          if (g >= p.ao) { mark(g - p.uo); mark(g - p.ao); }
          else if (g >= p.uo) mark(g + p.uo);
          else mark(g + p.ao);
        }
      }
    }
    __syncthreads();
    u32 S = 0;
    {   // prefix popcount: one chunk of words per thread (odd length: conflict-free), ONE block scan
      const u32 Kw = ((Wg + T - 1) / T) | 1u;
      u32 lo = tid * Kw; if (lo > Wg) lo = Wg;
      u32 hi = lo + Kw; if (hi > Wg) hi = Wg;
      u32 pc = 0;
      for (u32 i = lo; i < hi; ++i) pc += (u32)__popc(gbm[i]);
      u32 pos = block_exscan(pc, s_scan, &S);
      for (u32 i = lo; i < hi; ++i) { gpre[i] = pos; pos += (u32)__popc(gbm[i]); }
    }
    __syncthreads();       // (the prefix words are read by other threads from here on)
    auto rank_of = [&](u32 g) { return gpre[g >> 5] + (u32)__popc(gbm[g >> 5] & ((1u << (g & 31)) - 1u)); };
    for (u32 i = tid; i < Wg; i += T) {
      u32 w = gbm[i], r = gpre[i];
      while (w) { const u32 b = (u32)__ffs((int)w) - 1; w &= w - 1; gidx[r++] = (i << 5) + b; }
    }
    for (u32 s = tid; s < S; s += T) { a_in[s] = 0.0f; tr_off[s] = 0; tr_cur[s] = 0; }
    __syncthreads();
    // ---- local indices, unique tallies, transposed lists ---------------------------------------------------------------
    for (u32 j = tid; j < C; j += T) {
      const u32 lo = p.lab_off[ceq[j]], e0 = coff[j], ln = coff[j + 1] - e0;
      for (u32 k = 0; k < ln; ++k) {
        const u32 s = rank_of(p.labels[lo + k]);
        eloc[e0 + k] = s;
        atomicAdd(&tr_off[s], 1u);
      }
      if (ln == 1) atomicAdd(&a_in[eloc[e0]], (float)ccnt[j]);     // whole numbers: exact, order-free
    }
    if (p.usa)
      for (u32 s = tid; s < S; s += T) {
        const u32 g = gidx[s];
        if (g >= p.ao) { sibA[s] = rank_of(g - p.uo); sibB[s] = rank_of(g - p.ao); }
        else if (g >= p.uo) { sibA[s] = rank_of(g + p.uo); sibB[s] = NONE32; }
        else { sibA[s] = rank_of(g + p.ao); sibB[s] = NONE32; }
      }
    __syncthreads();
    const bool run_em = s_needs_em && !p.only_unique;
    if (run_em) {
      {   // exclusive scan of the per-gene entry counts (in place: tr_off[s] = first entry of gene s)
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
      const float uni = __fdiv_rn(1.0f, (float)p.num_alphas);
      for (u32 s = tid; s < S; s += T) {
        // every gene's classes in row order (insertion sort of a short list): the f32 sums below follow the reference's order
        const u32 b = tr_off[s], e = tr_off[s + 1];
        for (u32 i = b + 1; i < e; ++i) { const u32 x = tr[i]; u32 q = i; while (q > b && tr[q - 1] > x) { tr[q] = tr[q - 1]; --q; } tr[q] = x; }
        a_in[s] = p.init_uniform ? uni : __fmul_rn(__fadd_rn(a_in[s], 0.5f), 1e-3f);
      }
      __syncthreads();
      u32 it = 0;
      bool converged = true, last_round = false;
      while (it < INF_MIN_ITER || (it < INF_MAX_ITER && !converged) || last_round) {
        for (u32 s = tid; s < S; s += T) {       // get_abundance_for (src/em.rs:167-187)
          float v;
          if (!p.usa) v = a_in[s];
          else {
            const u32 g = gidx[s];
            if (g >= p.ao) v = __fadd_rn(__fadd_rn(a_in[sibA[s]], a_in[sibB[s]]), a_in[s]);
            else v = __fadd_rn(a_in[sibA[s]], a_in[s]);
          }
          eff[s] = v;
        }
        if (tid == 0) s_flag = 1;                // "converged"
        __syncthreads();
        for (u32 j = tid; j < C; j += T) {
          const u32 e0 = coff[j], ln = coff[j + 1] - e0;
          if (ln > 1) {
            float denom = 0.0f;
            for (u32 k = 0; k < ln; ++k) denom = __fadd_rn(denom, eff[eloc[e0 + k]]);
            cinv[j] = denom > 0.0f ? __fdiv_rn((float)ccnt[j], denom) : -1.0f;
          }
        }
        __syncthreads();
        for (u32 s = tid; s < S; s += T) {
          float sum = 0.0f;
          const float ef = eff[s];
          for (u32 t = tr_off[s]; t < tr_off[s + 1]; ++t) {
            const u32 j = tr[t];
            if (coff[j + 1] - coff[j] == 1) sum = __fadd_rn(sum, (float)ccnt[j]);
            else if (cinv[j] >= 0.0f) sum = __fadd_rn(sum, __fmul_rn(ef, cinv[j]));
          }
          if (sum > INF_ALPHA_CHECK_CUTOFF && fabsf(__fadd_rn(a_in[s], -sum)) > INF_REL_DIFF_TOLERANCE) s_flag = 0;
          a_in[s] = sum;
        }
        __syncthreads();
        converged = s_flag != 0;
        __syncthreads();
        ++it;
        if (last_round) break;
        if (it >= INF_MIN_ITER && converged) {
          for (u32 s = tid; s < S; s += T) if (a_in[s] < INF_MIN_OUTPUT_ALPHA) a_in[s] = 0.0f;
          last_round = true;
          __syncthreads();
        }
      }
      for (u32 s = tid; s < S; s += T) if (a_in[s] < INF_MIN_OUTPUT_ALPHA) a_in[s] = 0.0f;
      __syncthreads();
    }
    // ---- output: positive alphas, ascending gene id ---------------------------------------------------------------------
    if (tid == 0) { s_sum = 0.0f; s_max_bits = 0; s_over = 0; }
    u32 nnz = 0;
    float lsum = 0.0f, lmax = 0.0f;
    for (u32 c0 = 0; c0 < S; c0 += T) {
      const u32 s = c0 + tid;
      const float v = s < S ? a_in[s] : 0.0f;
      const u32 keep = v > 0.0f ? 1u : 0u;
      u32 tot;
      const u32 pos = block_exscan(keep, s_scan, &tot);
      if (keep) { p.stage_col[out_base + nnz + pos] = gidx[s]; p.stage_val[out_base + nnz + pos] = v; lsum += v; lmax = fmaxf(lmax, v); }
      nnz += tot;
    }
    if (lsum != 0.0f) atomicAdd(&s_sum, lsum);
    if (lmax > 0.0f) atomicMax(&s_max_bits, __float_as_uint(lmax));   // positive floats order like their bit patterns
    __syncthreads();
    const float mean = __fdiv_rn(s_sum, (float)nnz);
    u32 lover = 0;
    for (u32 i = tid; i < nnz; i += T) if (p.stage_val[out_base + i] > mean) ++lover;
    if (lover) atomicAdd(&s_over, lover);
    __syncthreads();
    if (tid == 0) {
      p.sum_umi[cell] = s_sum; p.max_umi[cell] = __uint_as_float(s_max_bits);
      p.num_expr[cell] = nnz; p.num_over_mean[cell] = s_over; p.flags[cell] = nnz == 0 ? 4 : 0;
    }
    __syncthreads();
  }
}

// one warp per cell copies its staging row to its CSR row (explicit staging offsets)
__global__ void __launch_bounds__(256) k_gather_rows_at(u64 n_cells, const u64* stage_off, const u32* num_expr, const u32* stage_col,
                                                        const float* stage_val, const u64* row_ptr, u32* col, float* val) {
  const u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_cells) return;
  const u32 lane = threadIdx.x & 31;
  const u32 nnz = num_expr[w];
  const u64 src = stage_off[w], dst = row_ptr[w];
  for (u32 i = lane; i < nnz; i += 32) { col[dst + i] = stage_col[src + i]; val[dst + i] = stage_val[src + i]; }
}

}  // namespace afq
