// afq_cuda.cu — C-ABI implementation (include/afq.h) over the sm_100a kernels.
//
// One afq_ctx per GPU. The device API (afq_quant_device) enqueues the whole per-batch
// pipeline on the caller's stream without host synchronisation:
//   memset(ctl) -> k_bin_cells -> k_resolve_smem<0..5> (persistent, one launch per arena
//   size) -> k_resolve_large -> row scan (3 launches) -> k_gather_rows
// The host API (afq_submit / afq_wait) wraps it with pinned-buffer H2D / D2H copies on a
// separate copy stream so that batch k+1 uploads while batch k computes.
// There is NO CPU fallback: without a usable CUDA device afq_create fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/afq.h"
#include "afq_pipeline.cuh"
#include "afq_infer.cuh"

using namespace afq;

namespace {

thread_local std::string g_create_err;

#define CUDA_TRY(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                   \
      return AFQ_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

template <class T>
struct DBuf {  // grow-only device buffer
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    cap = (e == cudaSuccess) ? want : 0;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T>
struct HBuf {  // grow-only pinned host buffer
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaHostAlloc((void**)&p, want * sizeof(T), cudaHostAllocPortable);
    cap = (e == cudaSuccess) ? want : 0;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Work {  // per-pipeline scratch (ordered on one stream)
  DBuf<Ctl> ctl;
  DBuf<u32> bin_list;
  DBuf<u32> stage_col;
  DBuf<float> stage_val;
  DBuf<u64> tile_sums;
  DBuf<u8> ge_arena[2];
  DBuf<u32> adj_pool;
  DBuf<u32> ps_garena;     // k_pug_smem<3> arenas
  DBuf<u32> ps_win, ps_nwin, ps_mem, ps_desc, ps_glab, ps_mlab, ps_nlab, ps_moff, ps_mlen, ps_back_list, ps_back_garena, cls_ncls, cls_nlab, cls_cnt, cls_off, cls_lab;   // split parsimony path (afq_pugc.cuh, afq_emc.cuh)
  DBuf<u64> na_tiles;      // rec_na8 -> offsets scan
  DBuf<u32> na_ref_off;    // device API: offsets derived from rec_na8
  DBuf<u32> umi_wide, refs_wide;   // rec_umi24 / refs24 widened to u32
  bool sync_sizing = true;         // read the control block back after the binning (device API) or plan the global arenas blind (afq_submit)
  // every pipeline has its own stream, lanes and giant-cell arenas, so that two of them can run side by side on the device
  cudaStream_t st = nullptr;       // host pipelines only (the device API runs on the caller's stream)
  cudaStream_t lanes[NUM_BINS] = {nullptr};
  cudaEvent_t ev_fork = nullptr, ev_lane[NUM_BINS] = {nullptr};
  u64* large_keys = nullptr;       // giant-cell scratch of k_resolve_large
  u32* large_cnts = nullptr;
  u32 large_cap_log2 = 0, large_blocks = 0;     // what the two arrays above were allocated for
  cudaError_t create_streams(bool own_stream) {
    cudaError_t e;
    if (own_stream && (e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming)) != cudaSuccess) return e;
    for (int i = 0; i < NUM_BINS; ++i) {     // (stream priorities for the lanes of the larger arenas were tried, r2v: no effect)
      if ((e = cudaStreamCreateWithFlags(&lanes[i], cudaStreamNonBlocking)) != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&ev_lane[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  // (re)allocate the giant-cell arenas for `blocks` CTAs of 2^cap_log2 entries; cudaFree waits for the device, so no batch is using them
  cudaError_t ensure_large(u32 cap_log2, u32 blocks) {
    if (large_keys && cap_log2 == large_cap_log2 && blocks == large_blocks) return cudaSuccess;
    if (large_keys) cudaFree(large_keys);
    if (large_cnts) cudaFree(large_cnts);
    large_keys = nullptr; large_cnts = nullptr;
    const size_t entries = (size_t)blocks << cap_log2;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&large_keys, entries * sizeof(u64))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&large_cnts, entries * sizeof(u32))) != cudaSuccess) return e;
    large_cap_log2 = cap_log2; large_blocks = blocks;
    return cudaSuccess;
  }
  cudaError_t ensure(u64 n_cells, u64 n_refs) {
    cudaError_t e;
    if ((e = ctl.ensure(1)) != cudaSuccess) return e;
    if ((e = bin_list.ensure((size_t)NUM_LISTS * n_cells)) != cudaSuccess) return e;
    if ((e = stage_col.ensure(n_refs + 1)) != cudaSuccess) return e;
    if ((e = stage_val.ensure(n_refs + 1)) != cudaSuccess) return e;
    if ((e = tile_sums.ensure(n_cells / SCAN_TILE + 2)) != cudaSuccess) return e;
    return cudaSuccess;
  }
  void release() {
    ctl.release(); bin_list.release(); stage_col.release(); stage_val.release();
    tile_sums.release(); ge_arena[0].release(); ge_arena[1].release(); adj_pool.release(); ps_garena.release();
    ps_win.release(); ps_nwin.release(); ps_mem.release(); ps_desc.release(); ps_glab.release();
    ps_mlab.release(); ps_nlab.release(); ps_moff.release(); ps_mlen.release(); ps_back_list.release(); ps_back_garena.release();
    cls_ncls.release(); cls_nlab.release(); cls_cnt.release(); cls_off.release(); cls_lab.release();
    na_tiles.release(); na_ref_off.release(); umi_wide.release(); refs_wide.release();
    if (large_keys) cudaFree(large_keys);
    if (large_cnts) cudaFree(large_cnts);
    large_keys = nullptr; large_cnts = nullptr;
    for (int i = 0; i < NUM_BINS; ++i) { if (lanes[i]) cudaStreamDestroy(lanes[i]); if (ev_lane[i]) cudaEventDestroy(ev_lane[i]); lanes[i] = nullptr; ev_lane[i] = nullptr; }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (st) cudaStreamDestroy(st);
    ev_fork = nullptr; st = nullptr;
  }
};

struct Slot {  // one in-flight host batch
  DBuf<u64> cell_rec_off;
  DBuf<u32> umi, ref_off, refs;
  DBuf<u8> na8, umi24, refs24;
  DBuf<u64> row_ptr;
  DBuf<u32> col;
  DBuf<float> val, sum_umi, max_umi;
  DBuf<u32> num_expr, num_over_mean;
  DBuf<u8> flags;
  HBuf<u64> h_row_ptr;
  HBuf<u32> h_col, h_num_expr, h_num_over_mean;
  HBuf<float> h_val, h_sum, h_max;
  HBuf<u8> h_flags;
  HBuf<Ctl> h_ctl;
  // --dump-eqclasses
  DBuf<u32> dump_ncls, dump_nlab, dump_cnt, dump_off, dump_lab, dq_counts, dq_labels;
  DBuf<u64> dq_cls_ptr, dq_lab_base, dq_cls_lab_ptr;
  HBuf<u64> h_cls_ptr, h_lab_base, h_cls_lab_ptr;
  HBuf<u32> h_dq_counts, h_dq_labels;
  u64 n_records = 0;
  bool has_dump = false;
  u32 retried = 0;          // re-runs done for this batch (bit 0: giant-cell arenas grown, bit 1: arena pools grown)
  afq_batch db{};             // the batch with DEVICE pointers (giant-cell retry)
  afq_device_out dout{};
  cudaEvent_t ev_h2d = nullptr, ev_done = nullptr, ev_d2h = nullptr;
  u64 n_cells = 0, n_refs = 0, ticket = 0;
  int pipe = 0;               // which host pipeline (Work + stream) runs this batch
  bool busy = false;
  void release() {
    cell_rec_off.release(); umi.release(); ref_off.release(); refs.release(); na8.release(); umi24.release(); refs24.release();
    row_ptr.release(); col.release(); val.release(); sum_umi.release(); max_umi.release();
    num_expr.release(); num_over_mean.release(); flags.release();
    h_row_ptr.release(); h_col.release(); h_num_expr.release(); h_num_over_mean.release();
    h_val.release(); h_sum.release(); h_max.release(); h_flags.release(); h_ctl.release();
    dump_ncls.release(); dump_nlab.release(); dump_cnt.release(); dump_off.release(); dump_lab.release(); dq_counts.release(); dq_labels.release();
    dq_cls_ptr.release(); dq_lab_base.release(); dq_cls_lab_ptr.release();
    h_cls_ptr.release(); h_lab_base.release(); h_cls_lab_ptr.release(); h_dq_counts.release(); h_dq_labels.release();
    if (ev_h2d) cudaEventDestroy(ev_h2d);
    if (ev_done) cudaEventDestroy(ev_done);
    if (ev_d2h) cudaEventDestroy(ev_d2h);
  }
};

struct ProfRec { int kid; cudaEvent_t a, b; };

}  // namespace

struct afq_ctx {
  afq_config cfg{};
  int device = 0;
  int num_sms = 0;
  u32* d_t2g = nullptr;
  u64 n_refs = 0;
  std::string err;
  u64 launches = 0;
  u64 reruns = 0;
  // giant-cell scratch (every pipeline holds arenas of this shape, Work::ensure_large)
  u32 large_cap_log2 = 21;
  u32 large_blocks = 0;   // 0 = one per SM
  int force_bin = -1;
  int grid_smem[NUM_SMEM_BINS] = {0};
  int ge_grid = 0;
  int grid_ps[PS_VARIANTS] = {0};
  u32 ps_limit_words = 0;      // AFQ_PS_LIMIT_WORDS: smaller k_pug_smem arena (tests: forces fallbacks to k_gene_eqc)
  bool no_ps_global = false;   // AFQ_NO_PS_GLOBAL=1: cells beyond the shared-memory arenas take k_gene_eqc
  bool no_ps = false;          // AFQ_NO_PS=1: parsimony cells all take the global-arena kernel (A/B experiments)
  u32 need_shift = 0;          // arena-size bias, raised when a batch overflowed many arenas
  bool no_lanes = false;       // AFQ_NO_LANES=1: launch the arena kernels back to back on the caller's stream
  bool no_ps_split = false;    // AFQ_NO_PS_SPLIT=1: unique-only parsimony stays on the single-kernel k_pug_smem (A/B, tests)
  int grid_count = 0;          // k_pug_count's persistent grid (0: its shared memory does not fit => no split path)
  int grid_back[4] = {0, 0, 0, 0};   // k_pug_back<tier>
  int grid_em[4] = {0, 0, 0, 0};     // k_em_cells<tier>
  bool no_em_split = false;    // AFQ_NO_EM_SPLIT=1: k_pug_back runs ge_back's own stage C (A/B)
  // pools of the per-CTA global arenas on the afq_submit path (no read-back, see k_plan_arenas): sizes learnt from the batches seen
  // so far — k_pug_build<3> words, k_pug_back<3> / k_em_cells<3> words, k_gene_eqc bytes (big / normal list)
  u64 pool_hint[4] = {0, 0, 0, 0};
  u64 pool_default_bytes = 256ull << 20;   // AFQ_POOL_MB: every pool's initial size
  u64 pool_budget_bytes = 8ull << 30;      // no pool is grown beyond this for full occupancy (it always holds one arena)
  u32 back_max_tier = 0;       // AFQ_BACK_MAX_TIER: largest shared-memory arena tier of k_pug_back (r2j: 0 is fastest on C3-em / C4 / C5)
  // pipelines: the device API's, and two for afq_submit — consecutive host batches alternate between them, so that the
  // next batch's kernels fill the SMs that the tail of this batch's persistent kernels leaves idle (a batch of 15 k cells
  // runs 25 % below the rate of one of 125 k on C4, r2r). AFQ_HOST_PIPES=1: one pipeline, batches back to back.
  static constexpr int NPIPE = 2;
  Work work_dev, work_host[NPIPE];
  int host_pipes = NPIPE;
  static constexpr int NSLOT = 3;
  Slot slots[NSLOT];
  cudaStream_t s_copy = nullptr, s_d2h = nullptr;
  u64 next_ticket = 1;
  std::mutex mu;
  // profiling
  bool profiling = false;
  std::vector<ProfRec> prof;
  double kid_ms[NUM_KID] = {0};
  u64 kid_launches[NUM_KID] = {0};
  Ctl* h_ctl_dev = nullptr;  // pinned, for afq_device_finish
  // afq_infer: result arrays (valid until the next call)
  std::vector<u64> inf_row_ptr;
  std::vector<u32> inf_col, inf_num_expr, inf_num_over_mean;
  std::vector<float> inf_val, inf_sum, inf_max;
  std::vector<u8> inf_flags;
  int grid_infer = 0;
};

namespace {

struct ProfScope {
  afq_ctx* c; int kid; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(afq_ctx* c_, int kid_, cudaStream_t st_) : c(c_), kid(kid_), st(st_) {
    c->launches++;
    c->kid_launches[kid]++;
    if (c->profiling) {
      cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a, st);
    }
  }
  ~ProfScope() {
    if (c->profiling) { cudaEventRecord(b, st); c->prof.push_back({kid, a, b}); }
  }
};

template <int VAR>
int setup_ps(afq_ctx* c) {
  const size_t smem = VAR < PS_SMEM_VARIANTS ? (size_t)ps_arena_words(VAR) * 4 : 0;
  if (smem) CUDA_TRY(c, cudaFuncSetAttribute(k_pug_smem<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (smem) CUDA_TRY(c, cudaFuncSetAttribute(k_pug_build<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pug_smem<VAR>, (int)ps_threads(VAR), smem));
  if (occ < 1) { c->err = "k_pug_smem variant does not fit an SM"; return AFQ_ERR_CUDA; }
  c->grid_ps[VAR] = occ * c->num_sms;
  return AFQ_OK;
}

template <int BIN>
int setup_bin(afq_ctx* c) {
  const size_t smem = bin_smem_bytes(BIN);
  CUDA_TRY(c, cudaFuncSetAttribute(k_resolve_smem<BIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  int occ = 0;
  CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_resolve_smem<BIN>,
                                                            (int)bin_threads(BIN), smem));
  if (occ < 1) occ = 1;
  c->grid_smem[BIN] = occ * c->num_sms;
  return AFQ_OK;
}

// Launcher for afq_pipeline.cuh: real launches on a CUDA stream, timed with event pairs when
// profiling is on.
struct CudaLauncher {
  afq_ctx* c;
  Work* w;
  cudaStream_t st;          // the caller's stream
  cudaStream_t cur;         // stream launches currently go to (st, or a lane between fork/join)
  int nlanes = 0;
  cudaEvent_t reg_a = nullptr, reg_b = nullptr;
  const u32* t2g() const { return c->d_t2g; }
  template <class... P, class... A>
  void launch(int kid, void (*k)(P...), unsigned grid, unsigned block, size_t smem, A... args) {
    ProfScope ps(c, kid, cur);
    k<<<grid, block, smem, cur>>>(args...);
  }
  int memset_zero(void* p, size_t n) { return cudaMemsetAsync(p, 0, n, st) != cudaSuccess; }
  bool sync_sizing() { return w->sync_sizing; }
  template <class T>
  T* pool(DBuf<T>& buf, int which, u64 min_elems, u64* cap) {
    u64 want = min_elems;
    if (!w->sync_sizing) {
      const u64 learnt = c->pool_hint[which] > c->pool_default_bytes / sizeof(T) ? c->pool_hint[which] : c->pool_default_bytes / sizeof(T);
      if (learnt > want) want = learnt;
    }
    if (buf.ensure((size_t)want + 64) != cudaSuccess) return nullptr;
    *cap = buf.cap - 64;
    return buf.p;
  }
  int read_ctl(const Ctl* d, Ctl* h) {
    if (cudaMemcpyAsync(c->h_ctl_dev, d, sizeof(Ctl), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    *h = *c->h_ctl_dev;
    return 0;
  }
  int grid_for_bin(int b) { return c->grid_smem[b]; }
  int ge_blocks(int which) { return which == 0 ? 16 : c->ge_grid; }
  u8* ge_arena(int which, u64 min_bytes, u64* cap) { return pool(w->ge_arena[which], 2 + which, min_bytes, cap); }
  u32* adj_pool(u64 n) { return w->adj_pool.ensure((size_t)n) == cudaSuccess ? w->adj_pool.p : nullptr; }
  u32 need_shift() { return c->need_shift; }
  int ps_grid(int v) { return (c->no_ps || (v == 3 && c->no_ps_global)) ? 0 : c->grid_ps[v]; }
  u32* ps_garena(u64 min_words, u64* cap) { return pool(w->ps_garena, 0, min_words, cap); }
  u32 ps_limit_words() { return c->ps_limit_words; }
  bool ps_split(u64 n_records, u64 n_refs, u64 n_cells, bool gene_labels, bool molecules, PsSplitBufs* o) {
    if (c->no_ps_split || (molecules ? c->grid_back[0] <= 0 : c->grid_count <= 0)) return false;
    const size_t desc_slots = (size_t)(n_records / 2 + n_records / 3 + n_records / 5 + n_records / 9 + 8);
    bool ok = w->ps_nwin.ensure(n_cells + 4) == cudaSuccess && w->ps_mem.ensure(4 * (size_t)n_records + 16) == cudaSuccess &&
              w->ps_desc.ensure(2 * desc_slots) == cudaSuccess && (!gene_labels || w->ps_glab.ensure(n_refs + 4) == cudaSuccess);
    if (ok && !molecules) ok = w->ps_win.ensure(n_records + 4) == cudaSuccess;
    if (ok && molecules)
      ok = w->ps_mlab.ensure(n_refs + 4) == cudaSuccess && w->ps_nlab.ensure(n_cells + 4) == cudaSuccess &&
           w->ps_moff.ensure(n_records + 4) == cudaSuccess && w->ps_mlen.ensure(n_records + 4) == cudaSuccess &&
           w->ps_back_list.ensure(4 * (size_t)n_cells + 4) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();      // out of memory for the split buffers: the single-kernel path still works
      return false;
    }
    o->win = w->ps_win.p; o->nwin = w->ps_nwin.p; o->mem = w->ps_mem.p; o->desc = w->ps_desc.p; o->glab = gene_labels ? w->ps_glab.p : nullptr;
    o->mlab = w->ps_mlab.p; o->nlab = w->ps_nlab.p; o->moff = w->ps_moff.p; o->mlen = w->ps_mlen.p; o->back_list = w->ps_back_list.p;
    return true;
  }
  int pc_grid(int which, size_t) { return which == 0 ? c->num_sms * 8 : (which == 1 ? c->grid_count : (which < 6 ? c->grid_back[which - 2] : c->grid_em[which - 6])); }
  u32 back_max_tier() { return c->back_max_tier; }
  bool em_split() { return !c->no_em_split && c->grid_em[0] > 0; }
  bool cls_bufs(u64 n_records, u64 n_refs, u64 n_cells, ClsBufs* o) {
    if (w->cls_ncls.ensure(n_cells + 4) != cudaSuccess || w->cls_nlab.ensure(n_cells + 4) != cudaSuccess || w->cls_cnt.ensure(n_records + 4) != cudaSuccess ||
        w->cls_off.ensure(n_records + 4) != cudaSuccess || w->cls_lab.ensure(n_refs + 4) != cudaSuccess) return false;
    o->ncls = w->cls_ncls.p; o->nlab = w->cls_nlab.p; o->cnt = w->cls_cnt.p; o->off = w->cls_off.p; o->lab = w->cls_lab.p;
    return true;
  }
  u32* back_garena(u64 min_words, u64* cap) { return pool(w->ps_back_garena, 1, min_words, cap); }
  // fork / join: lanes are the pipeline's own non-blocking streams ordered after / before the caller stream
  void fork(int n) {
    if (c->no_lanes) return;
    nlanes = n;
    cudaEventRecord(w->ev_fork, st);
    for (int i = 0; i < n; ++i) cudaStreamWaitEvent(w->lanes[i], w->ev_fork, 0);
  }
  void lane(int i) { cur = c->no_lanes ? st : w->lanes[i]; }
  void join() {
    for (int i = 0; i < nlanes; ++i) { cudaEventRecord(w->ev_lane[i], w->lanes[i]); cudaStreamWaitEvent(st, w->ev_lane[i], 0); }
    cur = st;
    nlanes = 0;
  }
  void region_begin() {
    if (c->profiling) { cudaEventCreate(&reg_a); cudaEventCreate(&reg_b); cudaEventRecord(reg_a, st); }
  }
  void region_end(int kid) {
    if (c->profiling) { cudaEventRecord(reg_b, st); c->prof.push_back({kid, reg_a, reg_b}); c->kid_launches[kid]++; }
  }
};

// Enqueue the full device pipeline for one batch. All pointers are device pointers.
int run_pipeline(afq_ctx* c, Work& w, const afq_batch& b, const afq_device_out& o, cudaStream_t st, Slot* dump_slot = nullptr) {
  if (b.n_cells >= 0xFFFFFFF0ull || b.n_refs_total >= 0xFFFFFFF0ull || b.n_records >= 0xFFFFFFF0ull) {
    c->err = "batch too large: n_cells, n_records and n_refs_total must be < 2^32 (split the batch)";
    return AFQ_ERR_INVALID;
  }
  if (o.cap_cells < b.n_cells + 1 || o.cap_nnz < b.n_refs_total) {
    c->err = "afq_device_out capacities too small (need n_cells+1 rows and n_refs_total nnz)";
    return AFQ_ERR_INVALID;
  }
  if (b.n_cells) CUDA_TRY(c, w.ensure(b.n_cells, b.n_refs_total));
  CUDA_TRY(c, w.ensure_large(c->large_cap_log2, c->large_blocks));
  afq_batch bb = b;
  CudaLauncher l{c, &w, st, st};
  if (!bb.rec_umi32 && bb.n_records) {
    if (!bb.rec_umi24) { c->err = "afq_batch needs rec_umi32 or rec_umi24"; return AFQ_ERR_INVALID; }
    if (c->cfg.umi_len > 12) { c->err = "rec_umi24 needs umi_len <= 12"; return AFQ_ERR_INVALID; }
    if ((uintptr_t)bb.rec_umi24 & 3) { c->err = "rec_umi24 must be 4-byte aligned"; return AFQ_ERR_INVALID; }
    CUDA_TRY(c, w.umi_wide.ensure(bb.n_records + 4));
    enqueue_unpack24(l, bb.rec_umi24, bb.n_records, w.umi_wide.p);
    bb.rec_umi32 = w.umi_wide.p;
  }
  if (!bb.refs && bb.n_refs_total) {
    if (!bb.refs24) { c->err = "afq_batch needs refs or refs24"; return AFQ_ERR_INVALID; }
    if ((uintptr_t)bb.refs24 & 3) { c->err = "refs24 must be 4-byte aligned"; return AFQ_ERR_INVALID; }
    CUDA_TRY(c, w.refs_wide.ensure(bb.n_refs_total + 4));
    enqueue_unpack24(l, bb.refs24, bb.n_refs_total, w.refs_wide.p);
    bb.refs = w.refs_wide.p;
  }
  if (!bb.rec_ref_offsets) {
    if (!bb.rec_na8) { c->err = "afq_batch needs rec_ref_offsets or rec_na8"; return AFQ_ERR_INVALID; }
    CUDA_TRY(c, w.na_tiles.ensure(bb.n_records / SCAN_TILE + 4));
    CUDA_TRY(c, w.na_ref_off.ensure(bb.n_records + 2));
    enqueue_na8_offsets(l, bb.rec_na8, bb.n_records, w.na_ref_off.p, w.na_tiles.p + 1, w.na_tiles.p);
    bb.rec_ref_offsets = w.na_ref_off.p;
  }
  PipeBufs pb{w.ctl.p, w.bin_list.p, w.stage_col.p, w.stage_val.p, w.tile_sums.p,
              w.large_keys, w.large_cnts, w.large_cap_log2, w.large_blocks};
  Slot* ds = (dump_slot && b.n_cells) ? dump_slot : nullptr;
  if (ds) {
    CUDA_TRY(c, ds->dump_ncls.ensure(b.n_cells + 1)); CUDA_TRY(c, ds->dump_nlab.ensure(b.n_cells + 1));
    CUDA_TRY(c, ds->dump_cnt.ensure(b.n_records + 1)); CUDA_TRY(c, ds->dump_off.ensure(b.n_records + 1)); CUDA_TRY(c, ds->dump_lab.ensure(b.n_refs_total + 1));
    CUDA_TRY(c, ds->dq_cls_ptr.ensure(b.n_cells + 2)); CUDA_TRY(c, ds->dq_lab_base.ensure(b.n_cells + 2));
    CUDA_TRY(c, ds->dq_cls_lab_ptr.ensure(b.n_records + 2)); CUDA_TRY(c, ds->dq_counts.ensure(b.n_records + 1)); CUDA_TRY(c, ds->dq_labels.ensure(b.n_refs_total + 1));
    pb.dump_ncls = ds->dump_ncls.p; pb.dump_nlab = ds->dump_nlab.p; pb.dump_cnt = ds->dump_cnt.p; pb.dump_off = ds->dump_off.p; pb.dump_lab = ds->dump_lab.p;
  }
  int rc = enqueue_batch(l, c->cfg, c->force_bin, pb, bb, o, c->err);
  if (rc != AFQ_OK) return rc;
  if (ds) {   // compact the cells' classes: two exclusive scans (classes, label words) + one gather
    const u32 n_tiles = (u32)((b.n_cells + SCAN_TILE - 1) / SCAN_TILE);
    for (int k = 0; k < 2; ++k) {
      const u32* src = k == 0 ? ds->dump_ncls.p : ds->dump_nlab.p;
      u64* dst = k == 0 ? ds->dq_cls_ptr.p : ds->dq_lab_base.p;
      l.launch(KID_SCAN_SUMS, k_scan_tile_sums, n_tiles, 1024u, (size_t)0, src, (u64)b.n_cells, w.tile_sums.p);
      l.launch(KID_SCAN_TILES, k_scan_tiles, 1u, 1024u, (size_t)0, w.tile_sums.p, n_tiles, dst, (u64)b.n_cells);
      l.launch(KID_SCAN_ROWS, k_scan_rows, n_tiles, 1024u, (size_t)0, src, (u64)b.n_cells, (const u64*)w.tile_sums.p, dst);
    }
    KArgs ka{};
    ka.n_cells = b.n_cells; ka.cell_rec_off = bb.cell_rec_offsets; ka.ref_off = bb.rec_ref_offsets;
    l.launch(KID_GATHER, k_dump_gather, (unsigned)((b.n_cells * 32 + 255) / 256), 256u, (size_t)0, ka, (const u32*)ds->dump_ncls.p,
             (const u32*)ds->dump_nlab.p, (const u32*)ds->dump_cnt.p, (const u32*)ds->dump_off.p, (const u32*)ds->dump_lab.p,
             (const u64*)ds->dq_cls_ptr.p, (const u64*)ds->dq_lab_base.p, ds->dq_cls_lab_ptr.p, ds->dq_labels.p, ds->dq_counts.p);
  }
  CUDA_TRY(c, cudaGetLastError());
  return AFQ_OK;
}

// learn the arena-size bias: if > 2 % of a batch's cells overflowed their arena, later batches
// start every cell one arena size up
void learn_from(afq_ctx* c, const Ctl& h) {
  u64 total = 0;
  for (int i = 0; i < NUM_LISTS; ++i) if (i != OVF_LIST) total += h.bin_count[i];
  if (total >= 64 && (u64)h.bin_count[OVF_LIST] * 50 > total && c->need_shift < 2) c->need_shift++;
  // the arena pools of the afq_submit path: room for every CTA of the grid at this batch's strides (within the budget, and
  // never less than one arena), so that later batches with cells like these run at full occupancy
  if (c->cfg.resolution == AFQ_RES_CR_LIKE || c->cfg.resolution == AFQ_RES_TRIVIAL) return;
  auto learn = [&](int which, u64 stride, u64 blocks, u64 elem_bytes) {
    u64 wantv = stride * (blocks ? blocks : 1);
    const u64 budget = c->pool_budget_bytes / elem_bytes;
    if (wantv > budget) wantv = budget;
    if (wantv < stride) wantv = stride;
    wantv += 64;
    if (wantv > c->pool_hint[which]) c->pool_hint[which] = wantv;
  };
  const bool starved = (h.error & DEV_ERR_POOL) != 0;      // (k_plan_arenas emptied the lists: learn every stride)
  if (h.bin_count[PS_LIST0 + 3] || starved) learn(0, h.ps3_words, (u64)c->grid_ps[3], 4);
  learn(1, h.back_words, (u64)std::max(c->grid_back[3], c->grid_em[3]), 4);
  if (h.bin_count[GE_LIST_BIG] || starved) learn(2, h.ge_bytes[0], 16, 1);
  learn(3, h.ge_bytes[1], std::min<u64>((u64)c->ge_grid, (u64)h.bin_count[GE_LIST_NORMAL] + 148), 1);
}

int check_device_error(afq_ctx* c, const Ctl& h) {
  learn_from(c, h);
  if (getenv("AFQ_DEBUG_CTL")) {     // work-list sizes of the batch (diagnostics)
    fprintf(stderr, "[afq ctl] lists:");
    for (int i = 0; i < NUM_LISTS; ++i) fprintf(stderr, " %u", h.bin_count[i]);
    fprintf(stderr, " | descs: %u %u %u %u | arenas: ps3 %u w x %u, back %u w x %u, ge %llu B x %u / %llu B x %u | error %u\n", h.desc_count[0], h.desc_count[1],
            h.desc_count[2], h.desc_count[3], h.ps3_words, h.ps3_blocks, h.back_words, h.back_blocks, h.ge_bytes[0], h.ge_blocks[0], h.ge_bytes[1], h.ge_blocks[1], h.error);
  }
  if (!h.error) return AFQ_OK;
  std::string buf;
  c->err = device_error_string(h, buf);
  return (h.error & (DEV_ERR_CELL_TOO_LARGE | DEV_ERR_ARENA | DEV_ERR_ADJ_POOL)) ? AFQ_ERR_UNSUPPORTED : AFQ_ERR_INTERNAL;
}

// the device pipeline of one host batch + the D2H of its small per-cell results, on the compute stream
int enqueue_slot(afq_ctx* c, Slot& s) {
  const u64 nc = s.n_cells;
  const bool dump = c->cfg.dump_eq != 0 && c->cfg.resolution != AFQ_RES_TRIVIAL;
  Work& w = c->work_host[s.pipe];
  cudaStream_t st = w.st;
  int rc = run_pipeline(c, w, s.db, s.dout, st, dump ? &s : nullptr);
  if (rc != AFQ_OK) return rc;
  s.has_dump = dump && nc > 0;
  if (s.has_dump) {
    CUDA_TRY(c, s.h_cls_ptr.ensure(nc + 2)); CUDA_TRY(c, s.h_lab_base.ensure(nc + 2));
    CUDA_TRY(c, cudaMemcpyAsync(s.h_cls_ptr.p, s.dq_cls_ptr.p, (nc + 1) * sizeof(u64), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(s.h_lab_base.p, s.dq_lab_base.p, (nc + 1) * sizeof(u64), cudaMemcpyDeviceToHost, st));
  }
  // small per-cell results come back on the compute stream right behind the kernels
  CUDA_TRY(c, cudaMemcpyAsync(s.h_row_ptr.p, s.row_ptr.p, (nc + 1) * sizeof(u64), cudaMemcpyDeviceToHost, st));
  if (nc) {
    CUDA_TRY(c, cudaMemcpyAsync(s.h_sum.p, s.sum_umi.p, nc * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(s.h_max.p, s.max_umi.p, nc * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(s.h_num_expr.p, s.num_expr.p, nc * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(s.h_num_over_mean.p, s.num_over_mean.p, nc * sizeof(u32), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(s.h_flags.p, s.flags.p, nc * sizeof(u8), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(s.h_ctl.p, w.ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
  } else {
    memset(s.h_ctl.p, 0, sizeof(Ctl));
  }
  return AFQ_OK;
}

// A cell had more alignments than the giant-cell arenas hold (DEV_ERR_CELL_TOO_LARGE): grow the arenas to fit it — fewer,
// larger ones within the same memory budget — and run the batch again. Its inputs are still in the slot's device buffers.
int grow_large_arena_and_rerun(afq_ctx* c, Slot& s, u32 max_cell_refs) {
  u32 log2cap = c->large_cap_log2;
  while (log2cap < 31 && (1ull << log2cap) < 2ull * max_cell_refs) ++log2cap;
  if ((1ull << log2cap) < 2ull * max_cell_refs) { c->err = "a cell has more than 2^30 alignments"; return AFQ_ERR_UNSUPPORTED; }
  const size_t budget = (size_t)c->large_blocks << c->large_cap_log2;     // entries held today
  const u32 blocks = (u32)std::max<size_t>(1, std::min<size_t>(c->large_blocks, budget >> log2cap));
  c->large_cap_log2 = log2cap; c->large_blocks = blocks;                  // (every pipeline re-allocates at its next batch)
  int rc = enqueue_slot(c, s);
  if (rc != AFQ_OK) return rc;
  CUDA_TRY(c, cudaStreamSynchronize(c->work_host[s.pipe].st));
  return AFQ_OK;
}

// A global-arena kernel found no arena in its pool for the batch's largest cell (DEV_ERR_POOL, afq_submit path): the pools were
// sized from earlier batches. learn_from() has the strides of this one; run it again on pools that hold them.
int grow_pools_and_rerun(afq_ctx* c, Slot& s) {
  const Ctl& h = *s.h_ctl.p;
  if (h.ps3_words >= 0xFFFFFFF0u || h.back_words >= 0xFFFFFFF0u) { c->err = "a cell needs a global arena of more than 2^32 words"; return AFQ_ERR_UNSUPPORTED; }
  learn_from(c, h);
  int rc = enqueue_slot(c, s);          // (a pool that grows is re-allocated behind a cudaFree, which waits for the device)
  if (rc != AFQ_OK) return rc;
  CUDA_TRY(c, cudaStreamSynchronize(c->work_host[s.pipe].st));
  return AFQ_OK;
}

}  // namespace

// The pipeline keeps ~10 streams busy (copy, compute, read-back, up to 7 lanes). With the default of 8 hardware work queues
// several of them share one, and a host that runs batches ahead gets its uploads queued behind the lane kernels of the batch
// before (measured r2u: C4 e2e 83.6 ms at 8 queues, 70.9 ms at 32). The setting is read when the CUDA context is created, so it
// is made here, when the library is loaded — unless the application has chosen a value itself.
__attribute__((constructor)) static void afq_more_hardware_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

extern "C" {

int afq_abi_version(void) { return AFQ_ABI_VERSION; }

const char* afq_last_error(const afq_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

int afq_create(const afq_config* cfg, const uint32_t* tid_to_gid, uint64_t n_refs, afq_ctx** out) {
  if (!cfg || !tid_to_gid || !out || n_refs == 0) { g_create_err = "afq_create: null argument"; return AFQ_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no CUDA device available (") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                   "); the afq product path has no CPU fallback";
    return AFQ_ERR_NO_DEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "afq_create: bad device ordinal"; return AFQ_ERR_INVALID; }
  if (cfg->umi_len > 16) { g_create_err = "umi_len > 16 is not supported on the CUDA path (UMI packed in 32 bits)"; return AFQ_ERR_UNSUPPORTED; }
  if (cfg->umi_len == 0 && !cfg->pug_exact_umi && res_is_pug(cfg->resolution)) {
    g_create_err = "umi_len = 0 with a parsimony resolution: the 1-edit UMI neighbourhood needs the UMI length (set umi_len, or pug_exact_umi)";
    return AFQ_ERR_INVALID;
  }
  if (cfg->sa_model != AFQ_SA_WINNER_TAKE_ALL && cfg->sa_model != AFQ_SA_PREFER_AMBIG) { g_create_err = "bad sa_model"; return AFQ_ERR_INVALID; }
  if (cfg->resolution < AFQ_RES_TRIVIAL || cfg->resolution > AFQ_RES_PARSIMONY_GENE) { g_create_err = "bad resolution"; return AFQ_ERR_INVALID; }
  if (cfg->usa_mode && (cfg->num_rows % 3 != 0)) { g_create_err = "USA mode needs num_rows = 3G"; return AFQ_ERR_INVALID; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return AFQ_ERR_CUDA; }
  if (prop.major < 10) {
    g_create_err = std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                   "; this library is built for sm_100a only";
    return AFQ_ERR_NO_DEVICE;
  }
  auto* c = new afq_ctx();
  c->cfg = *cfg;
  c->device = cfg->device;
  c->num_sms = prop.multiProcessorCount;
  c->n_refs = n_refs;
  if (const char* s = getenv("AFQ_LARGE_CAP_LOG2")) c->large_cap_log2 = (u32)atoi(s);
  if (const char* s = getenv("AFQ_LARGE_BLOCKS")) c->large_blocks = (u32)atoi(s);
  if (const char* s = getenv("AFQ_FORCE_BIN")) c->force_bin = atoi(s);
  if (const char* s = getenv("AFQ_NEED_SHIFT")) c->need_shift = (u32)atoi(s);
  if (const char* s = getenv("AFQ_NO_LANES")) c->no_lanes = atoi(s) != 0;
  if (const char* s = getenv("AFQ_NO_PS")) c->no_ps = atoi(s) != 0;
  if (const char* s = getenv("AFQ_NO_PS_GLOBAL")) c->no_ps_global = atoi(s) != 0;
  if (const char* s = getenv("AFQ_NO_PS_SPLIT")) c->no_ps_split = atoi(s) != 0;
  if (const char* s = getenv("AFQ_BACK_MAX_TIER")) c->back_max_tier = (u32)atoi(s);
  if (const char* s = getenv("AFQ_POOL_MB")) c->pool_default_bytes = (u64)std::max(1, atoi(s)) << 20;
  c->work_dev.sync_sizing = true;
  for (auto& w : c->work_host) {
    w.sync_sizing = false;                // afq_submit never waits for the device
    if (const char* s = getenv("AFQ_SYNC_SIZING")) w.sync_sizing = atoi(s) != 0;   // (A/B: the read-back of round 1)
  }
  if (const char* s = getenv("AFQ_HOST_PIPES")) c->host_pipes = std::max(1, std::min((int)afq_ctx::NPIPE, atoi(s)));
  {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) c->pool_budget_bytes = std::max<u64>(1ull << 30, (u64)total_b / 16);
  }
  if (const char* s = getenv("AFQ_NO_EM_SPLIT")) c->no_em_split = atoi(s) != 0;
  if (const char* s = getenv("AFQ_PS_LIMIT_WORDS")) c->ps_limit_words = (u32)atoi(s);
  if (c->large_cap_log2 < 10) c->large_cap_log2 = 10;
  if (c->large_cap_log2 > 30) c->large_cap_log2 = 30;
  if (c->large_blocks < 1) c->large_blocks = (u32)c->num_sms;
  auto fail = [&](int code) { g_create_err = c->err; afq_destroy(c); return code; };
#define CREATE_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { c->err = std::string(#expr) + ": " + cudaGetErrorString(_e); return fail(AFQ_ERR_CUDA); } } while (0)
  CREATE_TRY(cudaSetDevice(c->device));
  CREATE_TRY(cudaMalloc((void**)&c->d_t2g, n_refs * sizeof(u32)));
  CREATE_TRY(cudaMemcpy(c->d_t2g, tid_to_gid, n_refs * sizeof(u32), cudaMemcpyHostToDevice));
  CREATE_TRY(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
  CREATE_TRY(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  CREATE_TRY(cudaHostAlloc((void**)&c->h_ctl_dev, sizeof(Ctl), cudaHostAllocDefault));
  CREATE_TRY(c->work_dev.create_streams(false));
  for (auto& w : c->work_host) CREATE_TRY(w.create_streams(true));
  // (the giant-cell arenas are allocated by a pipeline's first batch: Work::ensure_large)
  for (auto& s : c->slots) {
    CREATE_TRY(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming));
  }
#undef CREATE_TRY
  int rc;
  if ((rc = setup_bin<0>(c)) || (rc = setup_bin<1>(c)) || (rc = setup_bin<2>(c)) ||
      (rc = setup_bin<3>(c)) || (rc = setup_bin<4>(c)) || (rc = setup_bin<5>(c)) ||
      (rc = setup_ps<0>(c)) || (rc = setup_ps<1>(c)) || (rc = setup_ps<2>(c)) || (rc = setup_ps<3>(c)))
    return fail(rc);
  {   // k_pug_count: presence bitmap over the output slots + counters in dynamic shared memory
    const size_t csm = pc_count_smem_bytes(cfg->num_rows);
    int occ = 0;
    if (csm <= 200 * 1024 &&
        cudaFuncSetAttribute(k_pug_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pug_count, (int)PC_THREADS, csm) == cudaSuccess && occ >= 1)
      c->grid_count = occ * c->num_sms;
    else cudaGetLastError();
  }
  {   // k_pug_back<tier>: the back end's arrays in dynamic shared memory (48 / 100 / 224 KB) or per-CTA global arenas
    int o0 = 0, o1 = 0, o2 = 0, o3 = 0;
    if (cudaFuncSetAttribute(k_pug_back<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pb_arena_words(0) * 4)) == cudaSuccess &&
        cudaFuncSetAttribute(k_pug_back<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pb_arena_words(1) * 4)) == cudaSuccess &&
        cudaFuncSetAttribute(k_pug_back<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pb_arena_words(2) * 4)) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o0, k_pug_back<0>, (int)PB_THREADS, pb_arena_words(0) * 4) == cudaSuccess && o0 >= 1 &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_pug_back<1>, (int)PB_THREADS, pb_arena_words(1) * 4) == cudaSuccess && o1 >= 1 &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_pug_back<2>, (int)PB_THREADS, pb_arena_words(2) * 4) == cudaSuccess && o2 >= 1 &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, k_pug_back<3>, (int)PB_THREADS, 0) == cudaSuccess && o3 >= 1) {
      c->grid_back[0] = o0 * c->num_sms; c->grid_back[1] = o1 * c->num_sms; c->grid_back[2] = o2 * c->num_sms;
      c->grid_back[3] = (o3 > 4 ? 4 : o3) * c->num_sms;     // (tier 3: up to four global arenas per SM)
    } else cudaGetLastError();
  }
  {   // k_em_cells<tier>: slot bitmap + per-cell arrays in dynamic shared memory, or (tier 3) per-CTA global arenas
    const u32 nr = cfg->num_rows;
    const size_t s0 = ec_smem_bytes(0, nr), s1 = ec_smem_bytes(1, nr), s2 = ec_smem_bytes(2, nr), s3 = ec_smem_bytes(3, nr) - 4ull * ec_arena_words(2);
    int o0 = 0, o1 = 0, o2 = 0, o3 = 0;
    if (s2 <= 227 * 1024 &&
        cudaFuncSetAttribute(k_em_cells<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s0) == cudaSuccess &&
        cudaFuncSetAttribute(k_em_cells<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1) == cudaSuccess &&
        cudaFuncSetAttribute(k_em_cells<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2) == cudaSuccess &&
        cudaFuncSetAttribute(k_em_cells<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o0, k_em_cells<0>, (int)EC_THREADS, s0) == cudaSuccess && o0 >= 1 &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_em_cells<1>, (int)EC_THREADS, s1) == cudaSuccess && o1 >= 1 &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_em_cells<2>, (int)EC_THREADS, s2) == cudaSuccess && o2 >= 1 &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, k_em_cells<3>, (int)EC_THREADS, s3) == cudaSuccess && o3 >= 1) {
      c->grid_em[0] = o0 * c->num_sms; c->grid_em[1] = o1 * c->num_sms; c->grid_em[2] = o2 * c->num_sms;
      c->grid_em[3] = (o3 > 4 ? 4 : o3) * c->num_sms;
    } else cudaGetLastError();
  }
  {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gene_eqc, (int)GE_THREADS, 0) != cudaSuccess || occ < 1) occ = 1;
    int cap = 4;   // CTAs per SM (AFQ_GE_OCC overrides for experiments)
    if (const char* e = getenv("AFQ_GE_OCC")) cap = atoi(e) > 0 ? atoi(e) : cap;
    if (occ > cap) occ = cap;
    c->ge_grid = occ * c->num_sms;
  }
  *out = c;
  return AFQ_OK;
}

void afq_destroy(afq_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& p : c->prof) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  c->work_dev.release();
  for (auto& w : c->work_host) w.release();
  for (auto& s : c->slots) s.release();
  if (c->d_t2g) cudaFree(c->d_t2g);
  if (c->h_ctl_dev) cudaFreeHost(c->h_ctl_dev);
  if (c->s_copy) cudaStreamDestroy(c->s_copy);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  delete c;
}

int afq_quant_device(afq_ctx* c, const afq_batch* b, const afq_device_out* o, void* stream) {
  if (!c || !b || !o) return AFQ_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  return run_pipeline(c, c->work_dev, *b, *o, (cudaStream_t)stream);
}

int afq_device_finish(afq_ctx* c, void* stream, uint64_t* nnz, const uint64_t* dev_row_ptr,
                      uint64_t n_cells) {
  if (!c) return AFQ_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (c->work_dev.ctl.p)
    CUDA_TRY(c, cudaMemcpyAsync(c->h_ctl_dev, c->work_dev.ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
  else
    memset(c->h_ctl_dev, 0, sizeof(Ctl));
  u64 h_nnz = 0;
  if (nnz && dev_row_ptr)
    CUDA_TRY(c, cudaMemcpyAsync(&h_nnz, dev_row_ptr + n_cells, sizeof(u64), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  if (nnz) *nnz = h_nnz;
  return check_device_error(c, *c->h_ctl_dev);
}

int afq_submit(afq_ctx* c, const afq_batch* hb, uint64_t* ticket) {
  if (!c || !hb || !ticket) return AFQ_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  if (hb->n_cells >= 0xFFFFFFF0ull || hb->n_refs_total >= 0xFFFFFFF0ull || hb->n_records >= 0xFFFFFFF0ull) {
    c->err = "batch too large: n_cells, n_records and n_refs_total must be < 2^32 (split the batch)";
    return AFQ_ERR_INVALID;
  }
  const u64 t = c->next_ticket;
  Slot& s = c->slots[t % afq_ctx::NSLOT];
  if (s.busy) {
    c->err = "too many batches in flight (at most " + std::to_string(afq_ctx::NSLOT) + "): call afq_wait first";
    return AFQ_ERR_INVALID;
  }
  const u64 nc = hb->n_cells, nr = hb->n_records, nf = hb->n_refs_total;
  CUDA_TRY(c, s.cell_rec_off.ensure(nc + 1));
  const bool use_umi24 = hb->rec_umi24 && !hb->rec_umi32;
  const bool use_refs24 = hb->refs24 && !hb->refs;
  if (use_umi24) CUDA_TRY(c, s.umi24.ensure(3 * nr + 16)); else CUDA_TRY(c, s.umi.ensure(nr + 1));
  CUDA_TRY(c, s.ref_off.ensure(nr + 2));
  if (hb->rec_na8 && !hb->rec_ref_offsets) CUDA_TRY(c, s.na8.ensure(nr + 1));
  if (use_refs24) CUDA_TRY(c, s.refs24.ensure(3 * nf + 16)); else CUDA_TRY(c, s.refs.ensure(nf + 1));
  CUDA_TRY(c, s.row_ptr.ensure(nc + 1));
  CUDA_TRY(c, s.col.ensure(nf + 1));
  CUDA_TRY(c, s.val.ensure(nf + 1));
  CUDA_TRY(c, s.sum_umi.ensure(nc + 1));
  CUDA_TRY(c, s.max_umi.ensure(nc + 1));
  CUDA_TRY(c, s.num_expr.ensure(nc + 1));
  CUDA_TRY(c, s.num_over_mean.ensure(nc + 1));
  CUDA_TRY(c, s.flags.ensure(nc + 1));
  CUDA_TRY(c, s.h_row_ptr.ensure(nc + 1));
  CUDA_TRY(c, s.h_sum.ensure(nc + 1));
  CUDA_TRY(c, s.h_max.ensure(nc + 1));
  CUDA_TRY(c, s.h_num_expr.ensure(nc + 1));
  CUDA_TRY(c, s.h_num_over_mean.ensure(nc + 1));
  CUDA_TRY(c, s.h_flags.ensure(nc + 1));
  CUDA_TRY(c, s.h_ctl.ensure(1));
  // H2D on the copy stream (overlaps the previous batch's kernels)
  CUDA_TRY(c, cudaMemcpyAsync(s.cell_rec_off.p, hb->cell_rec_offsets, (nc + 1) * sizeof(u64), cudaMemcpyHostToDevice, c->s_copy));
  if (!use_umi24 && !hb->rec_umi32 && nr) { c->err = "afq_batch needs rec_umi32 or rec_umi24"; return AFQ_ERR_INVALID; }
  if (!use_refs24 && !hb->refs && nf) { c->err = "afq_batch needs refs or refs24"; return AFQ_ERR_INVALID; }
  if (nr) {
    if (use_umi24) CUDA_TRY(c, cudaMemcpyAsync(s.umi24.p, hb->rec_umi24, 3 * nr, cudaMemcpyHostToDevice, c->s_copy));
    else CUDA_TRY(c, cudaMemcpyAsync(s.umi.p, hb->rec_umi32, nr * sizeof(u32), cudaMemcpyHostToDevice, c->s_copy));
  }
  const bool use_na8 = hb->rec_na8 && !hb->rec_ref_offsets;
  if (!use_na8 && !hb->rec_ref_offsets) { c->err = "afq_batch needs rec_ref_offsets or rec_na8"; return AFQ_ERR_INVALID; }
  if (use_na8) { if (nr) CUDA_TRY(c, cudaMemcpyAsync(s.na8.p, hb->rec_na8, nr * sizeof(u8), cudaMemcpyHostToDevice, c->s_copy)); }
  else CUDA_TRY(c, cudaMemcpyAsync(s.ref_off.p, hb->rec_ref_offsets, (nr + 1) * sizeof(u32), cudaMemcpyHostToDevice, c->s_copy));
  if (nf) {
    if (use_refs24) CUDA_TRY(c, cudaMemcpyAsync(s.refs24.p, hb->refs24, 3 * nf, cudaMemcpyHostToDevice, c->s_copy));
    else CUDA_TRY(c, cudaMemcpyAsync(s.refs.p, hb->refs, nf * sizeof(u32), cudaMemcpyHostToDevice, c->s_copy));
  }
  CUDA_TRY(c, cudaEventRecord(s.ev_h2d, c->s_copy));
  s.pipe = (int)(t % (u64)c->host_pipes);
  CUDA_TRY(c, cudaStreamWaitEvent(c->work_host[s.pipe].st, s.ev_h2d, 0));
  afq_batch db = *hb;
  db.cell_rec_offsets = s.cell_rec_off.p;
  db.rec_umi32 = use_umi24 ? nullptr : s.umi.p;
  db.rec_umi24 = use_umi24 ? s.umi24.p : nullptr;
  db.rec_ref_offsets = use_na8 ? nullptr : s.ref_off.p;
  db.rec_na8 = use_na8 ? s.na8.p : nullptr;
  db.refs = use_refs24 ? nullptr : s.refs.p;
  db.refs24 = use_refs24 ? s.refs24.p : nullptr;
  afq_device_out o{};
  o.row_ptr = s.row_ptr.p; o.cap_cells = nc + 1;
  o.col = s.col.p; o.val = s.val.p; o.cap_nnz = nf + 1;
  o.sum_umi = s.sum_umi.p; o.max_umi = s.max_umi.p;
  o.num_expr = s.num_expr.p; o.num_over_mean = s.num_over_mean.p; o.flags = s.flags.p;
  s.db = db; s.dout = o; s.n_cells = nc; s.n_refs = nf; s.n_records = nr; s.retried = 0;
  int rc = enqueue_slot(c, s);
  if (rc != AFQ_OK) return rc;
  CUDA_TRY(c, cudaEventRecord(s.ev_done, c->work_host[s.pipe].st));
  s.n_cells = nc; s.n_refs = nf; s.ticket = t; s.busy = true;
  c->next_ticket++;
  *ticket = t;
  return AFQ_OK;
}

int afq_wait(afq_ctx* c, uint64_t ticket, afq_result* out) {
  if (!c || !out) return AFQ_ERR_INVALID;
  // The context lock is held for the slot look-up and the bookkeeping only, never across a CUDA
  // synchronisation: a producer thread's afq_submit (H2D of the next batch) must not wait for this
  // batch's kernels. The slot's busy flag protects its buffers in between.
  std::unique_lock<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  Slot& s = c->slots[ticket % afq_ctx::NSLOT];
  if (!s.busy || s.ticket != ticket) { c->err = "afq_wait: unknown or already-collected ticket"; return AFQ_ERR_INVALID; }
  lk.unlock();
  cudaError_t e = cudaEventSynchronize(s.ev_done);
  lk.lock();
  if (e != cudaSuccess) { c->err = std::string("cudaEventSynchronize(ev_done): ") + cudaGetErrorString(e); s.busy = false; return AFQ_ERR_CUDA; }
  for (int attempt = 0; attempt < 2; ++attempt) {      // (each cause at most once)
    const u32 flags = s.h_ctl.p->error;
    int rr = AFQ_OK;
    if ((flags & DEV_ERR_CELL_TOO_LARGE) && !(s.retried & 1)) { s.retried |= 1; rr = grow_large_arena_and_rerun(c, s, s.h_ctl.p->max_cell_refs); }
    else if ((flags & DEV_ERR_POOL) && !(s.retried & 2)) { s.retried |= 2; rr = grow_pools_and_rerun(c, s); }
    else break;
    if (rr != AFQ_OK) { s.busy = false; return rr; }
    c->reruns++;
  }
  int rc = check_device_error(c, *s.h_ctl.p);
  if (rc != AFQ_OK) { s.busy = false; return rc; }
  const u64 nnz = s.h_row_ptr.p[s.n_cells];
  CUDA_TRY(c, s.h_col.ensure(nnz + 1));
  CUDA_TRY(c, s.h_val.ensure(nnz + 1));
  const u64 dq_ncls = s.has_dump ? s.h_cls_ptr.p[s.n_cells] : 0, dq_nlab = s.has_dump ? s.h_lab_base.p[s.n_cells] : 0;
  if (s.has_dump) {
    CUDA_TRY(c, s.h_cls_lab_ptr.ensure(dq_ncls + 2)); CUDA_TRY(c, s.h_dq_counts.ensure(dq_ncls + 1)); CUDA_TRY(c, s.h_dq_labels.ensure(dq_nlab + 1));
    s.h_cls_lab_ptr.p[dq_ncls] = dq_nlab;
    if (dq_ncls) {
      CUDA_TRY(c, cudaMemcpyAsync(s.h_cls_lab_ptr.p, s.dq_cls_lab_ptr.p, dq_ncls * sizeof(u64), cudaMemcpyDeviceToHost, c->s_d2h));
      CUDA_TRY(c, cudaMemcpyAsync(s.h_dq_counts.p, s.dq_counts.p, dq_ncls * sizeof(u32), cudaMemcpyDeviceToHost, c->s_d2h));
    }
    if (dq_nlab) CUDA_TRY(c, cudaMemcpyAsync(s.h_dq_labels.p, s.dq_labels.p, dq_nlab * sizeof(u32), cudaMemcpyDeviceToHost, c->s_d2h));
  }
  if (nnz || dq_ncls) {
    if (nnz) CUDA_TRY(c, cudaMemcpyAsync(s.h_col.p, s.col.p, nnz * sizeof(u32), cudaMemcpyDeviceToHost, c->s_d2h));
    if (nnz) CUDA_TRY(c, cudaMemcpyAsync(s.h_val.p, s.val.p, nnz * sizeof(float), cudaMemcpyDeviceToHost, c->s_d2h));
    CUDA_TRY(c, cudaEventRecord(s.ev_d2h, c->s_d2h));
    lk.unlock();
    e = cudaEventSynchronize(s.ev_d2h);
    lk.lock();
    if (e != cudaSuccess) { c->err = std::string("cudaEventSynchronize(ev_d2h): ") + cudaGetErrorString(e); s.busy = false; return AFQ_ERR_CUDA; }
  }
  out->n_cells = s.n_cells;
  out->nnz = nnz;
  out->row_ptr = s.h_row_ptr.p;
  out->col = s.h_col.p;
  out->val = s.h_val.p;
  out->sum_umi = s.h_sum.p;
  out->max_umi = s.h_max.p;
  out->num_expr = s.h_num_expr.p;
  out->num_over_mean = s.h_num_over_mean.p;
  out->flags = s.h_flags.p;
  return AFQ_OK;
}

int afq_result_eqclasses(afq_ctx* c, const afq_result* res, afq_eqc_dump* out) {
  if (!c || !res || !out) return AFQ_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  for (auto& s : c->slots)
    if (s.busy && s.h_row_ptr.p == res->row_ptr) {
      if (!c->cfg.dump_eq) { c->err = "afq_result_eqclasses: the context was created without dump_eq"; return AFQ_ERR_INVALID; }
      memset(out, 0, sizeof(*out));
      out->n_cells = s.n_cells;
      static const uint64_t zero2[2] = {0, 0};
      if (!s.has_dump) {       // `trivial` / an empty batch: no classes
        out->cell_cls_ptr = s.n_cells ? nullptr : zero2; out->cls_lab_ptr = zero2;
        if (s.n_cells) { c->err = "afq_result_eqclasses: no eq-classes for this resolution"; return AFQ_ERR_UNSUPPORTED; }
        return AFQ_OK;
      }
      out->n_classes = s.h_cls_ptr.p[s.n_cells]; out->n_labels = s.h_lab_base.p[s.n_cells];
      out->cell_cls_ptr = s.h_cls_ptr.p; out->cls_lab_ptr = s.h_cls_lab_ptr.p; out->labels = s.h_dq_labels.p; out->counts = s.h_dq_counts.p;
      return AFQ_OK;
    }
  c->err = "afq_result_eqclasses: unknown or released result";
  return AFQ_ERR_INVALID;
}

void afq_result_release(afq_ctx* c, afq_result* res) {
  if (!c || !res) return;
  std::lock_guard<std::mutex> lk(c->mu);
  for (auto& s : c->slots)
    if (s.busy && s.h_row_ptr.p == res->row_ptr) s.busy = false;
  memset(res, 0, sizeof(*res));
}

int afq_infer(afq_ctx* c, const afq_eqc_table* t, uint64_t n_cells, const uint64_t* cell_off, const uint32_t* cell_eq,
              const uint32_t* cell_cnt, afq_result* out) {
  if (!c || !t || !out || (n_cells && (!cell_off || !cell_eq || !cell_cnt)) || (t->n_classes && (!t->label_offsets || !t->labels))) return AFQ_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  const u32 num_alphas = c->cfg.num_rows;
  const bool usa = c->cfg.usa_mode != 0;
  if (num_alphas == 0 || (usa && num_alphas % 3 != 0)) { c->err = "afq_infer: bad num_rows"; return AFQ_ERR_INVALID; }
  if (inf_smem_bytes(num_alphas) > 220 * 1024) { c->err = "afq_infer: the gene axis does not fit the shared-memory bitmap"; return AFQ_ERR_UNSUPPORTED; }
  const u64 n_classes = t->n_classes, L = n_classes ? t->label_offsets[n_classes] : 0, nnz_in = n_cells ? cell_off[n_cells] : 0;
  if (n_cells >= 0xFFFFFFF0ull || L >= 0xFFFFFFF0ull || n_classes >= 0xFFFFFFF0ull) { c->err = "afq_infer: table too large"; return AFQ_ERR_INVALID; }
  for (u64 i = 0; i < n_classes; ++i) if (t->label_offsets[i + 1] < t->label_offsets[i]) { c->err = "afq_infer: label_offsets must ascend"; return AFQ_ERR_INVALID; }
  for (u64 i = 0; i < L; ++i) if (t->labels[i] >= num_alphas) { c->err = "afq_infer: a class label is >= num_rows"; return AFQ_ERR_INVALID; }
  // staging bounds and arena needs, from the class lengths
  std::vector<u64> stage_off(n_cells + 1, 0);
  u64 garena_words = 0;
  for (u64 cc = 0; cc < n_cells; ++cc) {
    if (cell_off[cc + 1] < cell_off[cc]) { c->err = "afq_infer: cell_offsets must ascend"; return AFQ_ERR_INVALID; }
    u64 E = 0;
    for (u64 k = cell_off[cc]; k < cell_off[cc + 1]; ++k) {
      if (cell_eq[k] >= n_classes) { c->err = "afq_infer: eq-class id out of range"; return AFQ_ERR_INVALID; }
      E += t->label_offsets[cell_eq[k] + 1] - t->label_offsets[cell_eq[k]];
    }
    u64 Sb = E * (usa ? 3 : 1);
    if (Sb > num_alphas) Sb = num_alphas;
    stage_off[cc + 1] = stage_off[cc] + Sb;
    const u64 need = inf_need_words(cell_off[cc + 1] - cell_off[cc], E, Sb, usa);
    if (need > INF_ARENA_WORDS && need > garena_words) garena_words = need;
    if (need >= 0xFFFFFFF0ull) { c->err = "afq_infer: a cell is too large"; return AFQ_ERR_UNSUPPORTED; }
  }
  const u64 n_stage = stage_off[n_cells];
  if (c->grid_infer == 0) {
    const size_t sm = inf_smem_bytes(num_alphas);
    CUDA_TRY(c, cudaFuncSetAttribute(k_em_subset, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int occ = 0;
    CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_em_subset, (int)INF_THREADS, sm));
    if (occ < 1) { c->err = "k_em_subset does not fit an SM"; return AFQ_ERR_CUDA; }
    c->grid_infer = occ * c->num_sms;
  }
  unsigned grid = (unsigned)c->grid_infer;
  if (grid > n_cells) grid = (unsigned)(n_cells ? n_cells : 1);
  DBuf<u64> d_cell_off, d_stage_off, d_row_ptr, d_tiles;
  DBuf<u32> d_eq, d_cnt, d_lab_off, d_labels, d_stage_col, d_num_expr, d_over, d_cursor, d_garena, d_col;
  DBuf<float> d_stage_val, d_sum, d_max, d_val;
  DBuf<u8> d_flags;
  struct Rel { std::vector<std::function<void()>> f; ~Rel() { for (auto& g : f) g(); } } rel;
#define INF_BUF(b, n) do { CUDA_TRY(c, (b).ensure(n)); rel.f.push_back([&] { (b).release(); }); } while (0)
  INF_BUF(d_cell_off, n_cells + 2); INF_BUF(d_stage_off, n_cells + 2); INF_BUF(d_row_ptr, n_cells + 2); INF_BUF(d_tiles, n_cells / SCAN_TILE + 4);
  INF_BUF(d_eq, nnz_in + 8); INF_BUF(d_cnt, nnz_in + 8); INF_BUF(d_lab_off, n_classes + 2); INF_BUF(d_labels, L + 8);
  INF_BUF(d_stage_col, n_stage + 8); INF_BUF(d_stage_val, n_stage + 8); INF_BUF(d_num_expr, n_cells + 2); INF_BUF(d_over, n_cells + 2);
  INF_BUF(d_cursor, 4); INF_BUF(d_sum, n_cells + 2); INF_BUF(d_max, n_cells + 2); INF_BUF(d_flags, n_cells + 2);
  INF_BUF(d_garena, garena_words ? garena_words * grid + 16 : 4);
#undef INF_BUF
  cudaStream_t st = c->work_host[0].st;
  CUDA_TRY(c, cudaMemsetAsync(d_eq.p, 0, (nnz_in + 8) * 4, st));        // (the bulk copies read up to 3 words past a row)
  CUDA_TRY(c, cudaMemsetAsync(d_cnt.p, 0, (nnz_in + 8) * 4, st));
  CUDA_TRY(c, cudaMemsetAsync(d_cursor.p, 0, 16, st));
  if (n_cells) {
    CUDA_TRY(c, cudaMemcpyAsync(d_cell_off.p, cell_off, (n_cells + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(c, cudaMemcpyAsync(d_stage_off.p, stage_off.data(), (n_cells + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nnz_in) {
      CUDA_TRY(c, cudaMemcpyAsync(d_eq.p, cell_eq, nnz_in * 4, cudaMemcpyHostToDevice, st));
      CUDA_TRY(c, cudaMemcpyAsync(d_cnt.p, cell_cnt, nnz_in * 4, cudaMemcpyHostToDevice, st));
    }
  }
  if (n_classes) {
    CUDA_TRY(c, cudaMemcpyAsync(d_lab_off.p, t->label_offsets, (n_classes + 1) * 4, cudaMemcpyHostToDevice, st));
    if (L) CUDA_TRY(c, cudaMemcpyAsync(d_labels.p, t->labels, L * 4, cudaMemcpyHostToDevice, st));
  }
  u64 nnz = 0;
  if (n_cells) {
    InferArgs p{};
    p.n_cells = n_cells; p.cell_off = d_cell_off.p; p.cell_eq = d_eq.p; p.cell_cnt = d_cnt.p; p.lab_off = d_lab_off.p; p.labels = d_labels.p;
    p.num_alphas = num_alphas; p.usa = usa ? 1u : 0u; p.uo = usa ? num_alphas / 3 : 0; p.ao = 2 * p.uo;
    p.init_uniform = c->cfg.em_init_uniform ? 1u : 0u; p.only_unique = 0;
    p.stage_off = d_stage_off.p; p.stage_col = d_stage_col.p; p.stage_val = d_stage_val.p;
    p.sum_umi = d_sum.p; p.max_umi = d_max.p; p.num_expr = d_num_expr.p; p.num_over_mean = d_over.p; p.flags = d_flags.p;
    p.cursor = d_cursor.p; p.garena = d_garena.p; p.garena_words = garena_words;
    c->launches += 5;
    k_em_subset<<<grid, INF_THREADS, inf_smem_bytes(num_alphas), st>>>(p);
    const u32 n_tiles = (u32)((n_cells + SCAN_TILE - 1) / SCAN_TILE);
    k_scan_tile_sums<<<n_tiles, 1024, 0, st>>>(d_num_expr.p, n_cells, d_tiles.p);
    k_scan_tiles<<<1, 1024, 0, st>>>(d_tiles.p, n_tiles, d_row_ptr.p, n_cells);
    k_scan_rows<<<n_tiles, 1024, 0, st>>>(d_num_expr.p, n_cells, d_tiles.p, d_row_ptr.p);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(&nnz, d_row_ptr.p + n_cells, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    CUDA_TRY(c, d_col.ensure(nnz + 4)); rel.f.push_back([&] { d_col.release(); });
    CUDA_TRY(c, d_val.ensure(nnz + 4)); rel.f.push_back([&] { d_val.release(); });
    k_gather_rows_at<<<(unsigned)((n_cells * 32 + 255) / 256), 256, 0, st>>>(n_cells, d_stage_off.p, d_num_expr.p, d_stage_col.p, d_stage_val.p,
                                                                               d_row_ptr.p, d_col.p, d_val.p);
    CUDA_TRY(c, cudaGetLastError());
  }
  c->inf_row_ptr.assign(n_cells + 1, 0); c->inf_col.resize(nnz); c->inf_val.resize(nnz);
  c->inf_sum.resize(n_cells); c->inf_max.resize(n_cells); c->inf_num_expr.resize(n_cells); c->inf_num_over_mean.resize(n_cells); c->inf_flags.resize(n_cells);
  if (n_cells) {
    CUDA_TRY(c, cudaMemcpyAsync(c->inf_row_ptr.data(), d_row_ptr.p, (n_cells + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (nnz) {
      CUDA_TRY(c, cudaMemcpyAsync(c->inf_col.data(), d_col.p, nnz * 4, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(c, cudaMemcpyAsync(c->inf_val.data(), d_val.p, nnz * 4, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->inf_sum.data(), d_sum.p, n_cells * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(c->inf_max.data(), d_max.p, n_cells * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(c->inf_num_expr.data(), d_num_expr.p, n_cells * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(c->inf_num_over_mean.data(), d_over.p, n_cells * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(c->inf_flags.data(), d_flags.p, n_cells, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
  }
  out->n_cells = n_cells; out->nnz = nnz;
  out->row_ptr = c->inf_row_ptr.data(); out->col = c->inf_col.data(); out->val = c->inf_val.data();
  out->sum_umi = c->inf_sum.data(); out->max_umi = c->inf_max.data(); out->num_expr = c->inf_num_expr.data();
  out->num_over_mean = c->inf_num_over_mean.data(); out->flags = c->inf_flags.data();
  return AFQ_OK;
}

int afq_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return AFQ_ERR_INVALID;
  cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
  return e == cudaSuccess ? AFQ_OK : AFQ_ERR_CUDA;
}
int afq_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
void afq_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

uint64_t afq_launch_count(const afq_ctx* c) { return c ? c->launches : 0; }
uint64_t afq_rerun_count(const afq_ctx* c) { return c ? c->reruns : 0; }

int afq_set_profiling(afq_ctx* c, int enable) {
  if (!c) return AFQ_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  c->profiling = enable != 0;
  return AFQ_OK;
}

// Drain the recorded CUDA-event pairs into per-kernel totals (synchronises the device).
int afq_profile_collect(afq_ctx* c) {
  if (!c) return AFQ_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  CUDA_TRY(c, cudaDeviceSynchronize());
  for (auto& p : c->prof) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) c->kid_ms[p.kid] += ms;
    cudaEventDestroy(p.a); cudaEventDestroy(p.b);
  }
  c->prof.clear();
  return AFQ_OK;
}

int afq_profile_reset(afq_ctx* c) {
  if (!c) return AFQ_ERR_INVALID;
  int rc = afq_profile_collect(c);
  std::lock_guard<std::mutex> lk(c->mu);
  for (int i = 0; i < NUM_KID; ++i) { c->kid_ms[i] = 0; c->kid_launches[i] = 0; }
  return rc;
}

// idx-th kernel's accumulated device time; returns nonzero past the end.
int afq_profile_get(afq_ctx* c, int idx, const char** name, double* ms, uint64_t* launches) {
  if (!c || idx < 0 || idx >= NUM_KID) return AFQ_ERR_INVALID;
  if (name) *name = KID_NAMES[idx];
  if (ms) *ms = c->kid_ms[idx];
  if (launches) *launches = c->kid_launches[idx];
  return AFQ_OK;
}

}  // extern "C"
