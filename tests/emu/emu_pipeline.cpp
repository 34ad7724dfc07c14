// emu_pipeline.cpp — TEST-ONLY: runs the product's kernels (alevin_fry_b200/csrc/*.cuh,
// compiled for the host with -DAFQ_EMU) through the shared launch sequence of
// afq_pipeline.cuh on the CPU emulator (cuda_emu.h). Used by tests/test_emu_parity.py to
// check kernel logic against the oracle without a GPU, and for debugging. Never loaded by
// the product: the product path is libafq.so on a real sm_100 device or an error.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "cuda_emu.h"
emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
namespace cuda_emu { State g; }

#include "../../alevin_fry_b200/csrc/afq_pipeline.cuh"
#include "../../alevin_fry_b200/csrc/afq_infer.cuh"

using namespace afq;

namespace {
struct EmuLauncher {
  const u32* t2g_ = nullptr;
  std::vector<u8> arena[2];
  std::vector<u32> adj, garena, sp_win, sp_nwin, sp_mem, sp_desc, sp_glab, sp_mlab, sp_nlab, sp_moff, sp_mlen, sp_big, sp_bga, c_ncls, c_nlab, c_cnt, c_off, c_lab;
  u64 launches = 0;
  int ge_threads_override = 0;
  const u32* t2g() const { return t2g_; }
  template <class... P, class... A>
  void launch(int, void (*k)(P...), unsigned grid, unsigned block, size_t smem, A... args) {
    ++launches;
    // persistent kernels pull work from lists: one emulated CTA is enough (and much faster)
    cuda_emu::launch(k, grid, block, smem, args...);
  }
  int memset_zero(void* p, size_t n) { memset(p, 0, n); return 0; }
  // AFQ_EMU_ASYNC=1: the afq_submit form of the pipeline — no read-back, fixed pools of AFQ_EMU_POOL_WORDS words
  bool sync_sizing() { const char* s = getenv("AFQ_EMU_ASYNC"); return !(s && atoi(s)); }
  u64 pool_words(u64 min_words) {
    if (sync_sizing()) return min_words;
    const char* s = getenv("AFQ_EMU_POOL_WORDS");
    const u64 fixed = s ? (u64)atoll(s) : (4ull << 20);
    return fixed > min_words ? fixed : min_words;
  }
  int read_ctl(const Ctl* d, Ctl* h) { *h = *d; return 0; }
  int grid_for_bin(int) { return 1; }
  int ge_blocks(int) { return 2; }
  u8* ge_arena(int which, u64 min_bytes, u64* cap) {
    *cap = 4 * pool_words((min_bytes + 3) / 4);
    arena[which].assign((size_t)*cap + 64, 0xCD);
    return arena[which].data();
  }
  u32* adj_pool(u64 n) { adj.assign((size_t)n, 0xCDCDCDCDu); return adj.data(); }
  void fork(int) {}
  void lane(int) {}
  void join() {}
  void region_begin() {}
  void region_end(int) {}
  u32 need_shift() { const char* s = getenv("AFQ_NEED_SHIFT"); return s ? (u32)atoi(s) : 0; }
  u32* ps_garena(u64 min_words, u64* cap) { *cap = pool_words(min_words); garena.assign((size_t)*cap + 16, 0xCDCDCDCDu); return garena.data(); }
  u32 ps_limit_words() { const char* s = getenv("AFQ_PS_LIMIT_WORDS"); return s ? (u32)atoi(s) : 0; }
  bool ps_split(u64 n_records, u64 n_refs, u64 n_cells, bool gene_labels, bool molecules, PsSplitBufs* o) {
    const char* s = getenv("AFQ_NO_PS_SPLIT");
    if (s && atoi(s)) return false;
    sp_win.assign(n_records + 4, 0xCDCDCDCDu); sp_nwin.assign(n_cells + 4, 0xCDCDCDCDu);
    sp_mem.assign(4 * (size_t)n_records + 16, 0xCDCDCDCDu);
    sp_desc.assign(2 * (size_t)(n_records / 2 + n_records / 3 + n_records / 5 + n_records / 9 + 8), 0xCDCDCDCDu);
    sp_glab.assign(gene_labels ? n_refs + 4 : 4, 0xCDCDCDCDu);
    // (16-byte alignment of the member pool: std::vector<u32> storage from operator new is 16-byte aligned)
    sp_mlab.assign(molecules ? n_refs + 4 : 4, 0xCDCDCDCDu); sp_nlab.assign(n_cells + 4, 0xCDCDCDCDu);
    sp_moff.assign(molecules ? n_records + 4 : 4, 0xCDCDCDCDu); sp_mlen.assign(molecules ? n_records + 4 : 4, 0xCDCDCDCDu); sp_big.assign(4 * n_cells + 4, 0xCDCDCDCDu);
    o->win = sp_win.data(); o->nwin = sp_nwin.data(); o->mem = sp_mem.data(); o->desc = sp_desc.data(); o->glab = gene_labels ? sp_glab.data() : nullptr;
    o->mlab = sp_mlab.data(); o->nlab = sp_nlab.data(); o->moff = sp_moff.data(); o->mlen = sp_mlen.data(); o->back_list = sp_big.data();
    return true;
  }
  int pc_grid(int, size_t) { return 1; }
  bool em_split() { const char* s = getenv("AFQ_NO_EM_SPLIT"); return !(s && atoi(s)); }
  bool cls_bufs(u64 n_records, u64 n_refs, u64 n_cells, ClsBufs* o) {
    c_ncls.assign(n_cells + 4, 0xCDCDCDCDu); c_nlab.assign(n_cells + 4, 0xCDCDCDCDu); c_cnt.assign(n_records + 4, 0xCDCDCDCDu);
    c_off.assign(n_records + 4, 0xCDCDCDCDu); c_lab.assign(n_refs + 4, 0xCDCDCDCDu);
    o->ncls = c_ncls.data(); o->nlab = c_nlab.data(); o->cnt = c_cnt.data(); o->off = c_off.data(); o->lab = c_lab.data();
    return true;
  }
  u32 back_max_tier() { const char* s = getenv("AFQ_BACK_MAX_TIER"); return s ? (u32)atoi(s) : 0u; }
  u32* back_garena(u64 min_words, u64* cap) { *cap = pool_words(min_words); sp_bga.assign((size_t)*cap + 16, 0xCDCDCDCDu); return sp_bga.data(); }
  int ps_grid(int v) {
    const char* s = getenv("AFQ_NO_PS");
    if (s && atoi(s)) return 0;
    const char* gq = getenv("AFQ_NO_PS_GLOBAL");
    return (v == 3 && gq && atoi(gq)) ? 0 : 1;
  }
};

struct EmuResult {
  std::vector<u64> cls_ptr, cls_lab_ptr;     // --dump-eqclasses
  std::vector<u32> cls_labels, cls_counts;
  std::vector<u64> row_ptr;
  std::vector<u32> col, num_expr, num_over_mean;
  std::vector<float> val, sum_umi, max_umi;
  std::vector<u8> flags;
};
}  // namespace

static Ctl g_last_ctl;

extern "C" {

// bin_count[] of the last afq_emu_quant call (which kernel took how many cells; tests)
void afq_emu_last_counts(uint32_t* out, int n) { for (int i = 0; i < n && i <= NUM_LISTS; ++i) out[i] = g_last_ctl.bin_count[i]; }

// k_scan_* and k_bin_* use grids of many CTAs: honoured as is (sequential CTAs).
int afq_emu_quant_dump(const afq_config* cfg, const uint32_t* tid_to_gid, uint64_t n_refs, const afq_batch* b,
                       afq_result* out, void** handle, char* errbuf, size_t errlen, uint32_t* dev_error, afq_eqc_dump* dump);
int afq_emu_quant(const afq_config* cfg, const uint32_t* tid_to_gid, uint64_t n_refs, const afq_batch* b,
                  afq_result* out, void** handle, char* errbuf, size_t errlen, uint32_t* dev_error) {
  return afq_emu_quant_dump(cfg, tid_to_gid, n_refs, b, out, handle, errbuf, errlen, dev_error, nullptr);
}
int afq_emu_quant_dump(const afq_config* cfg, const uint32_t* tid_to_gid, uint64_t n_refs, const afq_batch* b,
                       afq_result* out, void** handle, char* errbuf, size_t errlen, uint32_t* dev_error, afq_eqc_dump* dump) {
  (void)n_refs;
  EmuLauncher l;
  l.t2g_ = tid_to_gid;
  const u64 nc = b->n_cells, nf = b->n_refs_total;
  std::vector<Ctl> ctl(1);
  std::vector<u32> bin_list((size_t)NUM_LISTS * (nc + 1)), stage_col(nf + 1);
  std::vector<float> stage_val(nf + 1);
  std::vector<u64> tile_sums(nc / SCAN_TILE + 2);
  const u32 large_cap_log2 = 16, large_blocks = 1;
  std::vector<u64> large_keys((size_t)large_blocks << large_cap_log2);
  std::vector<u32> large_cnts((size_t)large_blocks << large_cap_log2);
  PipeBufs pb{ctl.data(), bin_list.data(), stage_col.data(), stage_val.data(), tile_sums.data(),
              large_keys.data(), large_cnts.data(), large_cap_log2, large_blocks};
  auto* r = new EmuResult();
  r->row_ptr.assign(nc + 1, 0); r->col.assign(nf + 1, 0); r->val.assign(nf + 1, 0);
  r->sum_umi.assign(nc + 1, 0); r->max_umi.assign(nc + 1, 0); r->num_expr.assign(nc + 1, 0);
  r->num_over_mean.assign(nc + 1, 0); r->flags.assign(nc + 1, 0);
  afq_device_out o{};
  o.row_ptr = r->row_ptr.data(); o.cap_cells = nc + 1;
  o.col = r->col.data(); o.val = r->val.data(); o.cap_nnz = nf + 1;
  o.sum_umi = r->sum_umi.data(); o.max_umi = r->max_umi.data();
  o.num_expr = r->num_expr.data(); o.num_over_mean = r->num_over_mean.data(); o.flags = r->flags.data();
  std::string err;
  int force_bin = -1;
  if (const char* s = getenv("AFQ_FORCE_BIN")) force_bin = atoi(s);
  afq_batch bb = *b;
  std::vector<u32> na_off;
  std::vector<u64> na_tiles;
  std::vector<u32> umi_wide, refs_wide;
  if (!bb.rec_umi32 && bb.rec_umi24) {
    umi_wide.assign(bb.n_records + 4, 0xCDCDCDCDu);
    enqueue_unpack24(l, bb.rec_umi24, bb.n_records, umi_wide.data());
    bb.rec_umi32 = umi_wide.data();
  }
  if (!bb.refs && bb.refs24) {
    refs_wide.assign(bb.n_refs_total + 4, 0xCDCDCDCDu);
    enqueue_unpack24(l, bb.refs24, bb.n_refs_total, refs_wide.data());
    bb.refs = refs_wide.data();
  }
  if (!bb.rec_ref_offsets && bb.rec_na8) {   // compact alignment counts -> offsets (device scan kernels)
    na_off.assign(bb.n_records + 2, 0xCDCDCDCDu);
    na_tiles.assign(bb.n_records / SCAN_TILE + 4, 0);
    enqueue_na8_offsets(l, bb.rec_na8, bb.n_records, na_off.data(), na_tiles.data() + 1, na_tiles.data());
    bb.rec_ref_offsets = na_off.data();
  }
  std::vector<u32> d_ncls, d_nlab, d_cnt, d_off, d_lab;
  if (dump && cfg->resolution != AFQ_RES_TRIVIAL && nc) {
    d_ncls.assign(nc + 1, 0xCDCDCDCDu); d_nlab.assign(nc + 1, 0xCDCDCDCDu);
    d_cnt.assign(bb.n_records + 1, 0xCDCDCDCDu); d_off.assign(bb.n_records + 1, 0xCDCDCDCDu); d_lab.assign(nf + 1, 0xCDCDCDCDu);
    pb.dump_ncls = d_ncls.data(); pb.dump_nlab = d_nlab.data(); pb.dump_cnt = d_cnt.data(); pb.dump_off = d_off.data(); pb.dump_lab = d_lab.data();
  }
  int rc = enqueue_batch(l, *cfg, force_bin, pb, bb, o, err);
  if (rc == AFQ_OK && dump) {
    r->cls_ptr.assign(nc + 1, 0);
    r->cls_lab_ptr.assign(1, 0);
    for (u64 c = 0; c < nc && pb.dump_ncls; ++c) {
      const u64 r0 = bb.cell_rec_offsets[c], f0 = bb.rec_ref_offsets[r0];
      r->cls_ptr[c + 1] = r->cls_ptr[c] + d_ncls[c];
      for (u32 j = 0; j < d_ncls[c]; ++j) {
        const u32 lo = d_off[r0 + j], hi = j + 1 < d_ncls[c] ? d_off[r0 + j + 1] : d_nlab[c];
        r->cls_labels.insert(r->cls_labels.end(), d_lab.begin() + f0 + lo, d_lab.begin() + f0 + hi);
        r->cls_lab_ptr.push_back(r->cls_labels.size());
        r->cls_counts.push_back(d_cnt[r0 + j]);
      }
    }
    dump->n_cells = nc; dump->n_classes = r->cls_counts.size(); dump->n_labels = r->cls_labels.size();
    dump->cell_cls_ptr = r->cls_ptr.data(); dump->cls_lab_ptr = r->cls_lab_ptr.data();
    dump->labels = r->cls_labels.data(); dump->counts = r->cls_counts.data();
  }
  g_last_ctl = ctl[0];
  if (dev_error) *dev_error = ctl[0].error;
  if (rc == AFQ_OK && ctl[0].error) {
    std::string buf;
    err = device_error_string(ctl[0], buf);
    rc = AFQ_ERR_INTERNAL;
  }
  if (errbuf && errlen) { strncpy(errbuf, err.c_str(), errlen - 1); errbuf[errlen - 1] = 0; }
  if (rc != AFQ_OK) { delete r; return rc; }
  out->n_cells = nc;
  out->nnz = r->row_ptr[nc];
  out->row_ptr = r->row_ptr.data(); out->col = r->col.data(); out->val = r->val.data();
  out->sum_umi = r->sum_umi.data(); out->max_umi = r->max_umi.data();
  out->num_expr = r->num_expr.data(); out->num_over_mean = r->num_over_mean.data(); out->flags = r->flags.data();
  *handle = r;
  return AFQ_OK;
}

void afq_emu_release(void* h) { delete static_cast<EmuResult*>(h); }

// k_em_subset (afq_infer.cuh) under emulation: same argument preparation as afq_infer in afq_cuda.cu
int afq_emu_infer(uint32_t num_alphas, int usa, int init_uniform, uint64_t n_classes, const uint32_t* lab_off, const uint32_t* labels,
                  uint64_t n_cells, const uint64_t* cell_off, const uint32_t* cell_eq, const uint32_t* cell_cnt, afq_result* out, void** handle,
                  uint32_t force_global) {
  (void)n_classes;
  auto* r = new EmuResult();
  std::vector<u64> stage_off(n_cells + 1, 0);
  u64 garena_words = 0;
  for (u64 cc = 0; cc < n_cells; ++cc) {
    u64 E = 0;
    for (u64 k = cell_off[cc]; k < cell_off[cc + 1]; ++k) E += lab_off[cell_eq[k] + 1] - lab_off[cell_eq[k]];
    u64 Sb = E * (usa ? 3 : 1);
    if (Sb > num_alphas) Sb = num_alphas;
    stage_off[cc + 1] = stage_off[cc] + Sb;
    const u64 need = inf_need_words(cell_off[cc + 1] - cell_off[cc], E, Sb, usa != 0);
    if ((need > INF_ARENA_WORDS || force_global) && need > garena_words) garena_words = need;
  }
  const u64 nnz_in = n_cells ? cell_off[n_cells] : 0;
  std::vector<u32> eq(cell_eq, cell_eq + nnz_in), cnt(cell_cnt, cell_cnt + nnz_in);
  eq.resize(nnz_in + 8, 0); cnt.resize(nnz_in + 8, 0);
  std::vector<u32> stage_col(stage_off[n_cells] + 8), cursor(4, 0), garena(garena_words + 16, 0xCDCDCDCDu);
  std::vector<float> stage_val(stage_off[n_cells] + 8);
  r->row_ptr.assign(n_cells + 1, 0);
  r->sum_umi.assign(n_cells + 1, 0); r->max_umi.assign(n_cells + 1, 0); r->num_expr.assign(n_cells + 1, 0);
  r->num_over_mean.assign(n_cells + 1, 0); r->flags.assign(n_cells + 1, 0);
  InferArgs p{};
  p.n_cells = n_cells; p.cell_off = cell_off; p.cell_eq = eq.data(); p.cell_cnt = cnt.data(); p.lab_off = lab_off; p.labels = labels;
  p.num_alphas = num_alphas; p.usa = usa ? 1u : 0u; p.uo = usa ? num_alphas / 3 : 0; p.ao = 2 * p.uo; p.init_uniform = init_uniform ? 1u : 0u;
  p.stage_off = stage_off.data(); p.stage_col = stage_col.data(); p.stage_val = stage_val.data();
  p.sum_umi = r->sum_umi.data(); p.max_umi = r->max_umi.data(); p.num_expr = r->num_expr.data(); p.num_over_mean = r->num_over_mean.data();
  p.flags = r->flags.data(); p.cursor = cursor.data(); p.garena = garena.data(); p.garena_words = garena_words;
  if (force_global) p.garena_words = garena_words;
  if (n_cells) cuda_emu::launch(k_em_subset, 1u, INF_THREADS, inf_smem_bytes(num_alphas), p);
  for (u64 cc = 0; cc < n_cells; ++cc) r->row_ptr[cc + 1] = r->row_ptr[cc] + r->num_expr[cc];
  r->col.assign(r->row_ptr[n_cells] + 1, 0); r->val.assign(r->row_ptr[n_cells] + 1, 0);
  for (u64 cc = 0; cc < n_cells; ++cc)
    for (u32 i = 0; i < r->num_expr[cc]; ++i) { r->col[r->row_ptr[cc] + i] = stage_col[stage_off[cc] + i]; r->val[r->row_ptr[cc] + i] = stage_val[stage_off[cc] + i]; }
  out->n_cells = n_cells; out->nnz = r->row_ptr[n_cells];
  out->row_ptr = r->row_ptr.data(); out->col = r->col.data(); out->val = r->val.data();
  out->sum_umi = r->sum_umi.data(); out->max_umi = r->max_umi.data(); out->num_expr = r->num_expr.data();
  out->num_over_mean = r->num_over_mean.data(); out->flags = r->flags.data();
  *handle = r;
  return AFQ_OK;
}

}  // extern "C"
