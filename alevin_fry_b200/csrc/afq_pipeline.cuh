// afq_pipeline.cuh — the per-batch launch sequence, written once against a small Launcher
// concept so that the CUDA library (afq_cuda.cu, real <<<>>> launches on a stream) and the
// test-only CPU emulator harness (tests/emu) enqueue exactly the same kernels in the same
// order with the same arguments.
//
// Launcher concept:
//   template <class... P, class... A>
//   void launch(int kid, void (*kernel)(P...), unsigned grid, unsigned block, size_t smem, A... args);
//   int  memset_zero(void* dev, size_t bytes);                // 0 on success
//   bool sync_sizing();                                       // read the control block back after the binning and size the global arenas exactly
//   int  read_ctl(const Ctl* dev, Ctl* host);                 // synchronising read-back (sync_sizing() only)
//   int  grid_for_bin(int bin);                               // persistent grid size
//   int  ge_blocks(int which);                                // 0 big-cell grid, 1 normal grid
//   u8*  ge_arena(int which, u64 min_bytes, u64* cap_bytes);  // grow-only pool of k_gene_eqc's arenas (min 0: the learnt size), nullptr on failure
//   u32* adj_pool(u64 entries);                               // grow-only, nullptr on failure
//   void region_begin(); void region_end(int kid);            // wall time of a forked region (profiling)
//   void fork(int lanes); void lane(int i); void join();      // independent launches may overlap:
//       between fork and join, launches go to lane i's stream (CUDA) / run in order (emulator)
//   u32  need_shift();                                        // arena-size bias learnt from overflows
//   int  ps_grid(int variant);                                // persistent grid of k_pug_smem<variant>; 0 = disabled
//   u32* ps_garena(u64 min_words, u64* cap_words);            // grow-only pool of the global arenas of k_pug_smem<3> / k_pug_build<3>, nullptr on failure
//   u32  ps_limit_words();                                    // 0, or a smaller arena for k_pug_smem (tests: forces fallbacks)
//   bool ps_split(u64 n_records, u64 n_refs, u64 n_cells, bool gene_labels, bool molecules, PsSplitBufs* out);
//                                                              // grow-only global buffers of the split parsimony path; false = not available
//   int  pc_grid(int which, size_t smem);                     // grid of the flat cover kernels (0) / k_pug_count (1) / k_pug_back tiers 0..3 (2..5) / k_em_cells tiers 0..3 (6..9)
//   u32* back_garena(u64 min_words, u64* cap_words);          // grow-only pool of the global arenas of k_pug_back<3> / k_em_cells<3>, nullptr on failure
//   u32  back_max_tier();                                     // largest shared-memory tier of k_pug_back (AFQ_BACK_MAX_TIER)
//   bool em_split();                                          // stage C in k_em_cells (default) or inside k_pug_back (AFQ_NO_EM_SPLIT=1)
//   bool cls_bufs(u64 n_records, u64 n_refs, u64 n_cells, ClsBufs* out);   // internal class regions (no --dump-eqclasses)
#pragma once
#include <string>

#include "../../include/afq.h"
#include "afq_kernels.cuh"
#include "afq_pug.cuh"
#include "afq_pugs.cuh"
#include "afq_pugc.cuh"
#include "afq_emc.cuh"

namespace afq {

enum KernelId : int {
  KID_BIN = 0, KID_SMEM0 = 1, KID_LARGE = 7, KID_SCAN_SUMS = 8, KID_SCAN_TILES = 9, KID_SCAN_ROWS = 10,
  KID_GATHER = 11, KID_GENE_EQC = 12, KID_GENE_EQC_BIG = 13, KID_BIN_GE = 14, KID_REGION = 15, KID_NA_OFFSETS = 16, KID_UNPACK24 = 17,
  KID_PUG_SMEM0 = 18, KID_PUG_REGION = 22, KID_PUG_BUILD0 = 23, KID_PUG_COVER2 = 27, KID_PUG_COVER4 = 28,
  KID_PUG_COVER8 = 29, KID_PUG_COVERW = 30, KID_PUG_COUNT = 31, KID_COVER_REGION = 32, KID_PUG_BACK = 33, KID_BACK_REGION = 38, KID_EM_CELLS = 39, KID_EM_REGION = 44, KID_PLAN = 45, NUM_KID = 46
};
static const char* const KID_NAMES[NUM_KID] = {
    "k_bin_cells", "k_resolve_smem<0>", "k_resolve_smem<1>", "k_resolve_smem<2>", "k_resolve_smem<3>",
    "k_resolve_smem<4>", "k_resolve_smem<5>", "k_resolve_large", "k_scan_tile_sums", "k_scan_tiles",
    "k_scan_rows", "k_gather_rows", "k_gene_eqc", "k_gene_eqc(big cells)", "k_bin_cells_ge", "resolve_region(wall)",
    "k_na_offsets(+tile sums)", "k_unpack24", "k_pug_smem<0>", "k_pug_smem<1>", "k_pug_smem<2>", "k_pug_smem<3>(global arena)", "pug_region(wall)",
    "k_pug_build<0>", "k_pug_build<1>", "k_pug_build<2>", "k_pug_build<3>(global arena)", "k_pug_cover2", "k_pug_cover_g<4>", "k_pug_cover_g<8>",
    "k_pug_cover_w", "k_pug_count", "cover_region(wall)", "k_pug_back<0>", "k_pug_back<1>", "k_pug_back<2>", "k_pug_back<3>(global arena)", "k_back_bin", "back_region(wall)",
    "k_em_cells<0>", "k_em_cells<1>", "k_em_cells<2>", "k_em_cells<3>(global arena)", "k_em_bin", "em_region(wall)", "k_plan_arenas"};

struct PipeBufs {  // device scratch owned by the caller (one set per stream-ordered pipeline)
  Ctl* ctl;
  u32* bin_list;       // [NUM_LISTS][n_cells]
  u32* stage_col;      // [n_refs_total + 1]
  float* stage_val;
  u64* tile_sums;      // [n_cells / SCAN_TILE + 2]
  u64* large_keys;     // giant-cell arenas for k_resolve_large
  u32* large_cnts;
  u32 large_cap_log2, large_blocks;
  // --dump-eqclasses (nullptr: off): per-cell regions, see GeArgs
  u32* dump_ncls = nullptr; u32* dump_nlab = nullptr; u32* dump_cnt = nullptr; u32* dump_off = nullptr; u32* dump_lab = nullptr;
};

struct ClsBufs { u32* ncls; u32* nlab; u32* cnt; u32* off; u32* lab; };
struct PsSplitBufs { u32* win; u32* nwin; u32* mem; u32* desc; u32* glab; u32* mlab; u32* nlab; u32* moff; u32* mlen; u32* back_list; };

inline bool res_is_pug(int r) {
  return r == AFQ_RES_PARSIMONY || r == AFQ_RES_PARSIMONY_EM || r == AFQ_RES_PARSIMONY_GENE || r == AFQ_RES_PARSIMONY_GENE_EM;
}
inline bool res_is_em(int r) {
  return r == AFQ_RES_CR_LIKE_EM || r == AFQ_RES_PARSIMONY_EM || r == AFQ_RES_PARSIMONY_GENE_EM;
}

template <int BIN, class L>
inline void launch_smem_bin(L& l, const KArgs& a) {
  const size_t smem = bin_smem_bytes(BIN);
  l.launch(KID_SMEM0 + BIN, k_resolve_smem<BIN>, (unsigned)l.grid_for_bin(BIN), bin_threads(BIN), smem, a);
}

// The arena kernels are independent of each other (overflowing cells go to a separate list that is
// drained afterwards), so they are forked onto per-arena lanes and overlap on the device: small
// CTAs fill the SMs while the few big-arena CTAs are still running.
template <class L>
inline void launch_crlike_bins(L& l, const KArgs& a, const PipeBufs& pb) {
  l.region_begin();
  l.fork(NUM_BINS);
  l.lane(5); launch_smem_bin<5>(l, a);
  l.lane(6); l.launch(KID_LARGE, k_resolve_large, pb.large_blocks, 1024u, (size_t)0, a, (u32)NUM_SMEM_BINS);
  l.lane(4); launch_smem_bin<4>(l, a);
  l.lane(3); launch_smem_bin<3>(l, a);
  l.lane(2); launch_smem_bin<2>(l, a);
  l.lane(1); launch_smem_bin<1>(l, a);
  l.lane(0); launch_smem_bin<0>(l, a);
  l.join();
  l.launch(KID_LARGE, k_resolve_large, pb.large_blocks, 1024u, (size_t)0, a, (u32)OVF_LIST);
  l.region_end(KID_REGION);
}

// 24-bit packed wire array -> u32 (dst must be 16-byte aligned with room for n values)
template <class L>
inline void enqueue_unpack24(L& l, const u8* src, u64 n, u32* dst) {
  if (!n) return;
  const u64 groups = (n + 3) / 4;
  l.launch(KID_UNPACK24, k_unpack24, (unsigned)((groups + 255) / 256), 256u, (size_t)0, src, n, dst);
}

// rec_na8 -> rec_ref_offsets on the device (three small launches). `na_tiles` needs
// n_records / SCAN_TILE + 2 entries; a dummy u64 receives the (unused) grand total.
template <class L>
inline void enqueue_na8_offsets(L& l, const u8* na8, u64 n_records, u32* ref_off, u64* na_tiles, u64* total_slot) {
  const u32 n_tiles = (u32)(n_records / SCAN_TILE + 1);  // covers index n_records (the closing offset)
  l.launch(KID_NA_OFFSETS, k_na_tile_sums, n_tiles, 1024u, (size_t)0, na8, n_records, na_tiles);
  l.launch(KID_NA_OFFSETS, k_scan_tiles, 1u, 1024u, (size_t)0, na_tiles, n_tiles, total_slot, (u64)0);
  l.launch(KID_NA_OFFSETS, k_na_offsets, n_tiles, 1024u, (size_t)0, na8, n_records, (const u64*)na_tiles, ref_off);
}

// ---- per-CTA global arenas, planned on the device ----------------------------------------------------------------------------
// The global-arena kernels (k_pug_build<3> / k_pug_smem<3>, k_pug_back<3> / k_em_cells<3>, k_gene_eqc) give every CTA an arena that
// holds the largest cell of the batch on their list. Those maxima exist only on the device (k_bin_cells_ge): k_plan_arenas turns
// them into a stride and the number of CTAs the pool has arenas for, in the control block, and the kernels read both there. With
// l.sync_sizing() the host reads the block back first and sizes the pools exactly (device API, emulator); without, nothing is read
// back — afq_submit returns while the previous batch is still computing — the pools keep the size learnt from earlier batches, a
// pool that holds fewer arenas than the grid runs fewer CTAs, and one that holds none flags DEV_ERR_POOL (the host grows it and
// runs the batch again).
struct ArenaPlan { u64 ps3_words, back_words, ge_bytes[2]; };
struct PlanArgs {
  u32 em_split, ps_on;
  u32 grid_ps3, grid_back, grid_ge[2];
  u64 pool_ps3, pool_back;      // words
  u64 pool_ge[2];               // bytes
};
__host__ __device__ inline ArenaPlan plan_arenas(const Ctl& h, u32 num_rows, bool usa, u32 lgt, bool em_split) {
  ArenaPlan p;
  p.ps3_words = ps_global_words(h.ps3_max_n, h.ps3_max_p, num_rows);
  // tier 3 of the back end / of k_em_cells holds any cell of the batch (molecules <= records, label words <= alignments)
  u64 gw = em_split ? ps_back_words_b(h.ge_max_n[1]) : ps_back_words(h.ge_max_n[1], h.ge_max_p[1], usa ? 3u : 1u);
  if (em_split) {
    const u64 ew = ec_need_words(h.ge_max_n[1], h.ge_max_p[1], ec_support_bound(h.ge_max_p[1], num_rows, usa), usa);
    if (ew > gw) gw = ew;
  }
  p.back_words = (gw + 15) & ~3ull;
  for (int w = 0; w < 2; ++w) p.ge_bytes[w] = align8(ge_carve(nullptr, h.ge_max_n[w], h.ge_max_p[w], lgt, nullptr)) + 64;
  return p;
}
__host__ __device__ inline u32 plan_ge_grid(const Ctl& h, int which, u32 grid, bool ps_on) {
  if (which == 1 && ps_on && grid > h.bin_count[GE_LIST_NORMAL] + 148u) grid = h.bin_count[GE_LIST_NORMAL] + 148u;   // hand-backs are rare
  return grid;
}
__global__ void k_plan_arenas(KArgs a, GeArgs g, PlanArgs pa) {
  if (blockIdx.x || threadIdx.x) return;
  Ctl& h = *a.ctl;
  const ArenaPlan p = plan_arenas(h, a.num_rows, a.usa_mode != 0, g.large_graph_thresh, pa.em_split != 0);
  auto fit = [](u64 pool, u64 stride, u32 grid, u64 limit) -> u32 {
    if (stride >= limit) return 0u;
    const u64 n = pool / (stride ? stride : 1);
    return (u32)(n < grid ? n : grid);
  };
  h.ps3_words = (u32)(p.ps3_words < 0xFFFFFFF0ull ? p.ps3_words : 0xFFFFFFF0ull);
  h.ps3_blocks = fit(pa.pool_ps3, p.ps3_words, pa.grid_ps3, 0xFFFFFFF0ull);
  h.back_words = (u32)(p.back_words < 0xFFFFFFF0ull ? p.back_words : 0xFFFFFFF0ull);
  h.back_blocks = fit(pa.pool_back, p.back_words, pa.grid_back, 0xFFFFFFF0ull);
  for (int w = 0; w < 2; ++w) {
    h.ge_bytes[w] = p.ge_bytes[w];
    h.ge_blocks[w] = fit(pa.pool_ge[w], p.ge_bytes[w], plan_ge_grid(h, w, pa.grid_ge[w], pa.ps_on != 0), ~0ull);
  }
  // A pool without a single arena for a kernel that may get work: flag the batch now and empty the work lists, so that no later
  // kernel runs on half-built state (the tiny cells' cr-like rows are done; every other row stays empty: num_expr was zeroed).
  u32 ps_total = 0;
  for (int v = 0; v < PS_VARIANTS; ++v) ps_total += h.bin_count[PS_LIST0 + v];
  const bool starved = (pa.grid_ps3 && h.bin_count[PS_LIST0 + 3] && !h.ps3_blocks) || (pa.grid_back && ps_total && !h.back_blocks) ||
                       (h.bin_count[GE_LIST_BIG] && !h.ge_blocks[0]) || ((h.bin_count[GE_LIST_NORMAL] || (pa.ps_on && ps_total)) && !h.ge_blocks[1]);
  if (starved) {
    h.error |= (u32)DEV_ERR_POOL;
    for (int v = 0; v < PS_VARIANTS; ++v) h.bin_count[PS_LIST0 + v] = 0;
    h.bin_count[GE_LIST_BIG] = 0; h.bin_count[GE_LIST_NORMAL] = 0;
  }
}

// Enqueue the whole pipeline for one batch. All batch/out pointers are device pointers.
template <class L>
int enqueue_batch(L& l, const afq_config& cfg, int force_bin, const PipeBufs& pb, const afq_batch& b,
                  const afq_device_out& o, std::string& err) {
  if (b.n_cells == 0) return l.memset_zero(o.row_ptr, sizeof(u64)) ? AFQ_ERR_CUDA : AFQ_OK;
  KArgs a{};
  a.n_cells = b.n_cells; a.n_records = b.n_records; a.n_refs_total = b.n_refs_total;
  a.cell_rec_off = b.cell_rec_offsets;
  a.umi = b.rec_umi32;
  a.ref_off = b.rec_ref_offsets;
  a.refs = b.refs;
  a.t2g = nullptr;  // filled by the launcher owner (see set_t2g)
  a.mode = (cfg.resolution == AFQ_RES_TRIVIAL) ? MODE_TRIVIAL : MODE_CRLIKE;
  a.usa_mode = cfg.usa_mode ? 1u : 0u;
  a.num_rows = cfg.num_rows;
  a.uo = cfg.usa_mode ? cfg.num_rows / 3 : 0;
  a.ao = 2 * a.uo;
  a.small_thresh = cfg.small_thresh;
  a.tiny_eligible = cfg.sa_model == AFQ_SA_WINNER_TAKE_ALL ? 1u : 0u;
  a.prefer_ambig = cfg.sa_model == AFQ_SA_PREFER_AMBIG ? 1u : 0u;
  a.ctl = pb.ctl;
  a.bin_list = pb.bin_list;
  a.stage_col = pb.stage_col;
  a.stage_val = pb.stage_val;
  a.sum_umi = o.sum_umi;
  a.max_umi = o.max_umi;
  a.num_expr = o.num_expr;
  a.num_over_mean = o.num_over_mean;
  a.flags = o.flags;
  a.large_keys = pb.large_keys;
  a.large_cnts = pb.large_cnts;
  a.large_cap_log2 = pb.large_cap_log2;
  a.t2g = l.t2g();

  if (l.memset_zero(pb.ctl, sizeof(Ctl))) { err = "memset(ctl) failed"; return AFQ_ERR_CUDA; }
  const int res = cfg.resolution;
  const unsigned bin_grid = (unsigned)((b.n_cells + 255) / 256);
  const bool dump = pb.dump_ncls != nullptr;      // the cells' gene eq-classes are wanted: every non-tiny cell builds them (ge_back)
  if (dump && (l.memset_zero(pb.dump_ncls, 4 * b.n_cells) || l.memset_zero(pb.dump_nlab, 4 * b.n_cells))) { err = "memset(dump) failed"; return AFQ_ERR_CUDA; }
  if ((res == AFQ_RES_CR_LIKE && !dump) || res == AFQ_RES_TRIVIAL) {
    l.launch(KID_BIN, k_bin_cells, bin_grid, 256u, (size_t)0, a, force_bin, l.need_shift());
    launch_crlike_bins(l, a, pb);
  } else {
    // cr-like-em and the parsimony family: tiny cells keep the cr-like fast path
    // (src/quant.rs:794-846, regardless of -r); every other cell builds gene eq-classes.
    if (cfg.large_graph_thresh > MAX_GRAPH_THRESH && res_is_pug(res)) {
      err = "--large-graph-thresh above 4096 is not supported on the CUDA path";
      return AFQ_ERR_UNSUPPORTED;
    }
    if (l.memset_zero(o.num_expr, 4 * b.n_cells)) { err = "memset(num_expr) failed"; return AFQ_ERR_CUDA; }   // (see k_plan_arenas)
    GeArgs g{};
    g.ge_mode = res_is_pug(res) ? ((res == AFQ_RES_PARSIMONY_GENE || res == AFQ_RES_PARSIMONY_GENE_EM) ? GE_MODE_PUG_GENE : GE_MODE_PUG_TXP)
                                : GE_MODE_CRLIKE;
    g.only_unique = res_is_em(res) ? 0u : 1u;
    g.ps_limit_words = l.ps_limit_words();
    g.dump_ncls = pb.dump_ncls; g.dump_nlab = pb.dump_nlab; g.dump_cnt = pb.dump_cnt; g.dump_off = pb.dump_off; g.dump_lab = pb.dump_lab;
    // cells expected to fit a shared-memory arena take k_pug_smem (parsimony family, cr-like-em); it hands cells
    // it cannot finish (arena too small after all, a component of more than 32 vertices) back to
    // the k_gene_eqc list, which is drained afterwards
    // (cr-like-em in USA mode stays on k_gene_eqc: its EM back end over 3 slots per gene needs the
    // 224 KB / global arenas for ordinary cells, where one CTA per SM iterates slower than k_gene_eqc's
    // four — measured r1x: C4 76.4 ms with k_pug_smem vs 64.2 ms without)
    // The SPLIT form (afq_pugc.cuh): k_pug_build per cell -> flat k_pug_cover* over the batch -> k_pug_count (unique-only)
    // or k_pug_back (EM resolutions / --dump-eqclasses: the shared back end on the cells' molecules in a global pool).
    const bool want_mol = !g.only_unique || dump;
    PsSplitBufs sb{};
    const bool split = l.ps_grid(0) > 0 && !a.prefer_ambig && b.n_cells < (1ull << 24) &&
                       (g.ge_mode == GE_MODE_CRLIKE ? want_mol : cfg.large_graph_thresh >= 2) &&
                       (want_mol || pc_count_smem_bytes(cfg.num_rows) <= 200 * 1024) &&
                       l.ps_split(b.n_records, b.n_refs_total, b.n_cells, g.ge_mode == GE_MODE_PUG_GENE, want_mol, &sb);
    const bool ps_on = split || (l.ps_grid(0) > 0 && (g.ge_mode == GE_MODE_CRLIKE ? (!cfg.usa_mode && !a.prefer_ambig) : cfg.large_graph_thresh >= 2));
    // (bit 2: the in-kernel arena also holds the molecules and the EM back end — not on the split path)
    const u32 ps_mode = ps_on ? (1u | (g.ge_mode == GE_MODE_PUG_GENE ? 2u : 0u) | ((want_mol && !split) ? 4u : 0u) | (l.ps_grid(3) > 0 ? 8u : 0u) | (split ? 16u : 0u)) : 0u;
    l.launch(KID_BIN_GE, k_bin_cells_ge, bin_grid, 256u, (size_t)0, a, force_bin, GE_BIG_RECORDS, l.need_shift(), ps_mode);
    launch_crlike_bins(l, a, pb);
    g.em_init_uniform = cfg.em_init_uniform ? 1u : 0u;
    g.pug_exact_umi = cfg.pug_exact_umi ? 1u : 0u;
    g.umi_len = cfg.umi_len;
    g.large_graph_thresh = (u32)(cfg.large_graph_thresh > MAX_GRAPH_THRESH ? MAX_GRAPH_THRESH : cfg.large_graph_thresh);
    g.num_alphas = cfg.usa_mode ? cfg.num_rows : cfg.num_gene_ids;
    g.adj_cap = 4ull * b.n_records + (1ull << 20);
    g.adj_pool = l.adj_pool(g.adj_cap);
    g.adj_used = (u64*)&pb.ctl->adj_used;
    if (!g.adj_pool) { err = "adjacency pool allocation failed"; return AFQ_ERR_CUDA; }
    // ---- the global arenas (see k_plan_arenas) ----
    const bool sync = l.sync_sizing();
    const bool usa = cfg.usa_mode != 0;
    const bool em_split = split && want_mol && l.em_split();
    const u32 all_cells = (u32)b.n_cells;
    Ctl h{};
    PlanArgs pa{};
    pa.em_split = em_split ? 1u : 0u; pa.ps_on = ps_on ? 1u : 0u;
    pa.grid_ps3 = ps_on ? (u32)l.ps_grid(3) : 0u;
    pa.grid_back = (split && want_mol) ? (u32)(l.pc_grid(5, 0) > l.pc_grid(9, 0) ? l.pc_grid(5, 0) : l.pc_grid(9, 0)) : 0u;
    pa.grid_ge[0] = (u32)l.ge_blocks(0); pa.grid_ge[1] = (u32)l.ge_blocks(1);
    u64 want_ps3 = 0, want_back = 0, want_ge[2] = {0, 0};     // (0: whatever the launcher has learnt)
    if (sync) {
      if (l.read_ctl(pb.ctl, &h)) { err = "reading the control block failed"; return AFQ_ERR_CUDA; }
      const ArenaPlan p = plan_arenas(h, cfg.num_rows, usa, g.large_graph_thresh, em_split);
      if (p.ps3_words >= 0xFFFFFFF0ull || p.back_words >= 0xFFFFFFF0ull) { err = "a cell needs a global arena of more than 2^32 words"; return AFQ_ERR_UNSUPPORTED; }
      const u32 c3 = h.bin_count[PS_LIST0 + 3];
      want_ps3 = p.ps3_words * (c3 < pa.grid_ps3 ? c3 : pa.grid_ps3) + 16;
      want_back = p.back_words * pa.grid_back + 16;
      for (int w = 0; w < 2; ++w) {
        const u32 cells = h.bin_count[w == 0 ? GE_LIST_BIG : GE_LIST_NORMAL] + (w == 1 && ps_on ? all_cells : 0u);   // every k_pug_smem cell may come back
        u32 blocks = plan_ge_grid(h, w, pa.grid_ge[w], ps_on);
        if (blocks > cells) blocks = cells;
        want_ge[w] = p.ge_bytes[w] * blocks + 64;
      }
    }
    if (pa.grid_ps3) { g.ps_garena = l.ps_garena(want_ps3, &pa.pool_ps3); if (!g.ps_garena) { err = "k_pug_smem global arena allocation failed"; return AFQ_ERR_CUDA; } }
    if (pa.grid_back) { g.back_garena = l.back_garena(want_back, &pa.pool_back); if (!g.back_garena) { err = "k_pug_back global arena allocation failed"; return AFQ_ERR_CUDA; } }
    u8* ge_pool[2];
    for (int w = 0; w < 2; ++w) {
      ge_pool[w] = l.ge_arena(w, want_ge[w], &pa.pool_ge[w]);
      if (!ge_pool[w]) { err = "gene-eq-class arena allocation failed (" + std::to_string(want_ge[w]) + " B)"; return AFQ_ERR_CUDA; }
    }
    l.launch(KID_PLAN, k_plan_arenas, 1u, 32u, (size_t)0, a, g, pa);
    // without the read-back the list sizes are unknown here: every kernel gets its full grid and finds its list on the device
    auto list_cells = [&](int list) { return sync ? h.bin_count[list] : all_cells; };
    // the arena variants of k_pug_smem are independent of each other (cells they cannot finish go to k_gene_eqc's
    // list, drained afterwards): forked onto lanes like the cr-like arenas so that every persistent kernel's tail
    // overlaps the others instead of idling the chip (14.5 + 7.3 + 10.3 + 5.5 ms back to back)
    u32 ps_cells = 0;
    // unique-only parsimony resolutions take the SPLIT form (afq_pugc.cuh): build per cell, cover flat over the batch, count per cell
    if (split) {
      g.ps_win = sb.win; g.ps_nwin = sb.nwin; g.ps_mem = sb.mem; g.ps_desc = sb.desc; g.ps_glab = sb.glab;
      g.ps_mlab = sb.mlab; g.ps_nlab = sb.nlab; g.ps_moff = sb.moff; g.ps_mlen = sb.mlen; g.back_list = sb.back_list;
      const u64 nr = b.n_records;
      g.ps_desc_base[0] = 0;
      g.ps_desc_base[1] = (u32)(nr / 2 + 1);
      g.ps_desc_base[2] = g.ps_desc_base[1] + (u32)(nr / 3 + 1);
      g.ps_desc_base[3] = g.ps_desc_base[2] + (u32)(nr / 5 + 1);
    }
    l.region_begin();
    l.fork(PS_VARIANTS);
    for (int v = PS_VARIANTS - 1; ps_on && v >= 0; --v) {   // biggest cells first
      const u32 cnt = list_cells(PS_LIST0 + v);
      u32 blocks = (u32)l.ps_grid(v);
      if (!cnt || !blocks) continue;
      ps_cells = sync ? ps_cells + cnt : all_cells;
      if (blocks > cnt) blocks = cnt;
      l.lane(v);
      if (v == 3) {
        if (split) l.launch(KID_PUG_BUILD0 + 3, k_pug_build<3>, blocks, ps_threads(3), (size_t)0, a, g);
        else l.launch(KID_PUG_SMEM0 + 3, k_pug_smem<3>, blocks, ps_threads(3), (size_t)0, a, g);
        continue;
      }
      const size_t smem = (size_t)ps_arena_words(v) * 4;
      if (split) {
        if (v == 0) l.launch(KID_PUG_BUILD0 + 0, k_pug_build<0>, blocks, ps_threads(0), smem, a, g);
        else if (v == 1) l.launch(KID_PUG_BUILD0 + 1, k_pug_build<1>, blocks, ps_threads(1), smem, a, g);
        else l.launch(KID_PUG_BUILD0 + 2, k_pug_build<2>, blocks, ps_threads(2), smem, a, g);
      } else {
        if (v == 0) l.launch(KID_PUG_SMEM0 + 0, k_pug_smem<0>, blocks, ps_threads(0), smem, a, g);
        else if (v == 1) l.launch(KID_PUG_SMEM0 + 1, k_pug_smem<1>, blocks, ps_threads(1), smem, a, g);
        else l.launch(KID_PUG_SMEM0 + 2, k_pug_smem<2>, blocks, ps_threads(2), smem, a, g);
      }
    }
    l.join();
    l.region_end(KID_PUG_REGION);
    if (split && ps_cells) {
      if (g.ge_mode != GE_MODE_CRLIKE) {
        // the four size classes are independent: forked onto lanes, rarest / most expensive first
        const unsigned cg = (unsigned)l.pc_grid(0, 0);
        l.region_begin();
        l.fork(4);
        l.lane(3); l.launch(KID_PUG_COVERW, k_pug_cover_w, cg, PC_THREADS, (size_t)0, a, g);
        l.lane(2); l.launch(KID_PUG_COVER8, k_pug_cover_g<8, 2>, cg, PC_THREADS, (size_t)0, a, g);
        l.lane(1); l.launch(KID_PUG_COVER4, k_pug_cover_g<4, 1>, cg, PC_THREADS, (size_t)0, a, g);
        l.lane(0); l.launch(KID_PUG_COVER2, k_pug_cover2, cg, PC_THREADS, (size_t)0, a, g);
        l.join();
        l.region_end(KID_COVER_REGION);
      }
      if (!want_mol) {
        const size_t csm = pc_count_smem_bytes(cfg.num_rows);
        unsigned cb = (unsigned)l.pc_grid(1, csm);
        if (cb > ps_cells) cb = ps_cells;
        l.launch(KID_PUG_COUNT, k_pug_count, cb, PC_THREADS, csm, a, g);
      } else {
        // Stage B (molecules -> gene eq-classes, canonical order) in k_pug_back, stage C (counts / EM) in k_em_cells; the
        // classes travel through the dump regions (the caller's when --dump-eqclasses is on, internal ones otherwise).
        // AFQ_NO_EM_SPLIT=1: k_pug_back runs ge_back's own stage C instead (A/B).
        if (em_split && !dump) {
          ClsBufs cb{};
          if (!l.cls_bufs(b.n_records, b.n_refs_total, b.n_cells, &cb)) { err = "class region allocation failed"; return AFQ_ERR_CUDA; }
          g.dump_ncls = cb.ncls; g.dump_nlab = cb.nlab; g.dump_cnt = cb.cnt; g.dump_off = cb.off; g.dump_lab = cb.lab;
        }
        g.classes_only = em_split ? 1u : 0u;
        const unsigned g3 = (unsigned)l.pc_grid(5, 0);
        g.back_max_tier = l.back_max_tier();
        l.launch(KID_PUG_BACK + 4, k_back_bin, (unsigned)((ps_cells + 255) / 256), 256u, (size_t)0, a, g);
        l.region_begin();
        l.fork(PB_TIERS);
        l.lane(3); l.launch(KID_PUG_BACK + 3, k_pug_back<3>, g3, PB_THREADS, (size_t)0, a, g);
        l.lane(2); l.launch(KID_PUG_BACK + 2, k_pug_back<2>, (unsigned)l.pc_grid(4, 0), PB_THREADS, (size_t)pb_arena_words(2) * 4, a, g);
        l.lane(1); l.launch(KID_PUG_BACK + 1, k_pug_back<1>, (unsigned)l.pc_grid(3, 0), PB_THREADS, (size_t)pb_arena_words(1) * 4, a, g);
        l.lane(0); l.launch(KID_PUG_BACK + 0, k_pug_back<0>, (unsigned)l.pc_grid(2, 0), PB_THREADS, (size_t)pb_arena_words(0) * 4, a, g);
        l.join();
        l.region_end(KID_BACK_REGION);
        if (em_split) {
          l.launch(KID_EM_CELLS + 4, k_em_bin, (unsigned)((ps_cells + 255) / 256), 256u, (size_t)0, a, g);
          l.region_begin();
          l.fork(3);
          l.lane(2); l.launch(KID_EM_CELLS + 2, k_em_cells<2>, (unsigned)l.pc_grid(8, 0), EC_THREADS, ec_smem_bytes(2, cfg.num_rows), a, g);
          l.lane(1); l.launch(KID_EM_CELLS + 1, k_em_cells<1>, (unsigned)l.pc_grid(7, 0), EC_THREADS, ec_smem_bytes(1, cfg.num_rows), a, g);
          l.lane(0); l.launch(KID_EM_CELLS + 0, k_em_cells<0>, (unsigned)l.pc_grid(6, 0), EC_THREADS, ec_smem_bytes(0, cfg.num_rows), a, g);
          l.join();
          // the global-arena tier runs last: it also takes the cells whose exact support exceeded their tier's arena
          l.launch(KID_EM_CELLS + 3, k_em_cells<3>, (unsigned)l.pc_grid(9, 0), EC_THREADS, ec_smem_bytes(3, cfg.num_rows) - 4ull * ec_arena_words(2), a, g);
          l.region_end(KID_EM_REGION);
          g.classes_only = 0;      // (k_gene_eqc behind runs the whole back end for the handed-back cells)
        }
      }
    }
    for (int which = 0; which < 2; ++which) {
      const int list = which == 0 ? GE_LIST_BIG : GE_LIST_NORMAL;
      const u32 cells = sync ? h.bin_count[list] + (which == 1 ? ps_cells : 0u) : all_cells;   // upper bound: every k_pug_smem cell may come back
      if (cells == 0) continue;
      u32 blocks = plan_ge_grid(h, which, pa.grid_ge[which], sync && ps_on);
      if (blocks > cells) blocks = cells;
      g.arena = ge_pool[which];
      g.list_id = (u32)list;
      g.list_which = (u32)which;
      l.launch(which == 0 ? KID_GENE_EQC_BIG : KID_GENE_EQC, k_gene_eqc, blocks, GE_THREADS, (size_t)0, a, g);
    }
  }
  const u32 n_tiles = (u32)((b.n_cells + SCAN_TILE - 1) / SCAN_TILE);
  l.launch(KID_SCAN_SUMS, k_scan_tile_sums, n_tiles, 1024u, (size_t)0, (const u32*)o.num_expr, (u64)b.n_cells, pb.tile_sums);
  l.launch(KID_SCAN_TILES, k_scan_tiles, 1u, 1024u, (size_t)0, pb.tile_sums, n_tiles, (u64*)o.row_ptr, (u64)b.n_cells);
  l.launch(KID_SCAN_ROWS, k_scan_rows, n_tiles, 1024u, (size_t)0, (const u32*)o.num_expr, (u64)b.n_cells, (const u64*)pb.tile_sums, (u64*)o.row_ptr);
  l.launch(KID_GATHER, k_gather_rows, (unsigned)((b.n_cells * 32 + 255) / 256), 256u, (size_t)0, a, (const u64*)o.row_ptr, (u32*)o.col, (float*)o.val);
  return AFQ_OK;
}

inline const char* device_error_string(const Ctl& h, std::string& buf) {
  if (h.error & DEV_ERR_CELL_TOO_LARGE) buf = "a cell has " + std::to_string(h.max_cell_refs) + " alignments, more than the giant-cell arena holds (the host API grows the arena and retries; with afq_quant_device raise AFQ_LARGE_CAP_LOG2)";
  else if (h.error & DEV_ERR_POOL) buf = "the pool of a global-arena kernel holds no arena for this batch's largest cell (the host API grows the pool and retries)";
  else if (h.error & DEV_ERR_ADJ_POOL) buf = "PUG adjacency pool exhausted";
  else if (h.error & DEV_ERR_ARENA) buf = "a cell does not fit the gene-eq-class arena";
  else if (h.error & DEV_ERR_HASH) buf = "eq-class label hash collision not resolved after reseeding";
  else if (h.error & DEV_ERR_LABEL) buf = "empty covering label in the PUG cover";
  else buf = "device-side error flags " + std::to_string(h.error);
  return buf.c_str();
}

}  // namespace afq
