// afq_synth.cpp — deterministic synthetic collated-cell generator.
//
// Bench/test infrastructure: produces the SoA batch layout of include/afq.h directly
// (what a host would obtain by parsing a collated RAD file). Every draw comes from a
// splitmix64 stream keyed by (seed, cell index), so any cell range can be generated
// independently, on any number of threads, with identical bytes.
//
// Model (per cell): reads-per-cell ~ LogNormal(sigma) scaled to `reads_mean` (or fixed);
// molecules drawn until the read budget is met: gene ~ Zipf(s) over a global ranking,
// reads/UMI ~ 1 + Geometric (mean `reads_per_umi`), UMI uniform over 4^umi_len; mapping
// ambiguity: 85 % one gene (a non-empty subset of its transcripts, re-drawn per read with
// probability 0.3 so a molecule spans several transcript-level eq-classes), `p_multi2` two
// paralogs (g, g+1), `p_multi3` 3-6 consecutive genes; each read's UMI gets one random
// substitution with probability `umi_err`. Records inside a cell are shuffled (record
// order is not part of the RAD contract).
// USA mode: 3 spliced + 1 unspliced transcript per gene (tid = 4g+k); per molecule 60 % S
// only, 25 % U only, 15 % S+U (each read S, U or both); independently 10 % also hit a
// spliced transcript of gene g+1.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

extern "C" {
typedef struct afq_synth_spec {
  uint64_t seed;
  uint32_t n_genes;        /* G */
  int32_t usa_mode;        /* 0: 3 tx/gene, ids = gene; 1: 4 tx/gene, ids 2g / 2g+1 */
  int32_t umi_len;         /* bases (<= 16) */
  int32_t fixed_reads;     /* > 0: exactly this many reads per cell */
  double reads_mean;       /* mean reads per cell (lognormal) */
  double lognorm_sigma;
  double reads_per_umi;    /* mean reads per molecule (>= 1) */
  double zipf_s;
  double p_multi2, p_multi3;
  double umi_err;
} afq_synth_spec;
}

namespace {
using u32 = uint32_t;
using u64 = uint64_t;

struct Rng {
  u64 s;
  explicit Rng(u64 seed) : s(seed) {}
  u64 next() {
    u64 z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  u32 below(u32 n) { return (u32)(((next() >> 32) * (u64)n) >> 32); }
};

struct ZipfTable {
  u32 n = 0;
  double s = 0;
  std::vector<double> cdf;
};
std::mutex g_zipf_mu;
std::vector<ZipfTable*> g_zipf;
const ZipfTable* zipf_table(u32 n, double s) {
  std::lock_guard<std::mutex> lk(g_zipf_mu);
  for (auto* t : g_zipf)
    if (t->n == n && t->s == s) return t;
  auto* t = new ZipfTable();
  t->n = n; t->s = s; t->cdf.resize(n);
  double acc = 0;
  for (u32 i = 0; i < n; ++i) { acc += 1.0 / std::pow((double)(i + 1), s); t->cdf[i] = acc; }
  for (u32 i = 0; i < n; ++i) t->cdf[i] /= acc;
  g_zipf.push_back(t);
  return t;
}
inline u32 zipf_draw(const ZipfTable* t, double u) {
  u32 k = (u32)(std::lower_bound(t->cdf.begin(), t->cdf.end(), u) - t->cdf.begin());
  // scatter ranks over gene ids so that hot genes are not id-adjacent
  u32 r = k < t->n ? k : t->n - 1;
  return (u32)(((u64)r * 2654435761ull) % t->n);
}

struct Read { u32 umi; u32 na; u32 refs[8]; };

inline u64 cell_seed(u64 seed, u64 cell) {
  Rng r(seed ^ (cell * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull));
  return r.next();
}

u32 reads_for_cell(const afq_synth_spec& sp, Rng& rng) {
  if (sp.fixed_reads > 0) return (u32)sp.fixed_reads;
  // Box-Muller
  double u1 = rng.uni(), u2 = rng.uni();
  if (u1 < 1e-300) u1 = 1e-300;
  double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  double mu = std::log(sp.reads_mean) - 0.5 * sp.lognorm_sigma * sp.lognorm_sigma;
  double v = std::exp(mu + sp.lognorm_sigma * z);
  if (v < 1.0) v = 1.0;
  if (v > 4.0e6) v = 4.0e6;
  return (u32)v;
}

inline void sort_small(u32* a, u32 n) {
  for (u32 i = 1; i < n; ++i) {
    u32 x = a[i]; u32 j = i;
    while (j > 0 && a[j - 1] > x) { a[j] = a[j - 1]; --j; }
    a[j] = x;
  }
}

// Generate one cell into `out` (cleared first).
void gen_cell(const afq_synth_spec& sp, const ZipfTable* zt, u64 cell, std::vector<Read>& out) {
  out.clear();
  Rng rng(cell_seed(sp.seed, cell));
  const u32 target = reads_for_cell(sp, rng);
  const u32 G = sp.n_genes;
  const u32 umi_bits = 2 * (u32)sp.umi_len;
  const u64 umi_mask = umi_bits >= 32 ? 0xFFFFFFFFull : ((1ull << umi_bits) - 1);
  const double p_geo = 1.0 / std::max(1.0, sp.reads_per_umi);
  const double log1mp = std::log(1.0 - std::min(p_geo, 0.999999));
  out.reserve(target);
  while (out.size() < target) {
    const u32 g = zipf_draw(zt, rng.uni());
    u32 c = 1;
    if (p_geo < 0.999999) {
      double u = rng.uni();
      if (u < 1e-300) u = 1e-300;
      c += (u32)(std::log(u) / log1mp);
    }
    const u32 umi = (u32)(rng.next() & umi_mask);
    const double kind = rng.uni();
    const u32 base_mask = 1 + rng.below(7);
    const u32 tx_pick = rng.below(3);
    u32 ngenes = 1;
    if (kind >= 1.0 - sp.p_multi3) ngenes = 3 + rng.below(4);
    else if (kind >= 1.0 - sp.p_multi3 - sp.p_multi2) ngenes = 2;
    // USA per-molecule status
    int status = 0;  // 0 S, 1 U, 2 S+U
    bool second_gene = false;
    if (sp.usa_mode) {
      double us = rng.uni();
      status = us < 0.60 ? 0 : (us < 0.85 ? 1 : 2);
      second_gene = rng.uni() < 0.10;
    }
    for (u32 k = 0; k < c && out.size() < target; ++k) {
      Read rd;
      rd.umi = umi;
      rd.na = 0;
      if (!sp.usa_mode) {
        if (ngenes == 1) {
          u32 mask = base_mask;
          if (rng.uni() < 0.3) mask = 1 + rng.below(7);
          for (u32 t = 0; t < 3; ++t)
            if (mask & (1u << t)) rd.refs[rd.na++] = 3 * g + t;
        } else {
          for (u32 j = 0; j < ngenes; ++j) rd.refs[rd.na++] = 3 * ((g + j) % G) + tx_pick;
          sort_small(rd.refs, rd.na);
        }
      } else {
        int rs = status;
        if (status == 2) rs = (int)rng.below(3);  // this read: S, U or both
        if (rs == 0 || rs == 2) {
          u32 mask = base_mask;
          if (rng.uni() < 0.3) mask = 1 + rng.below(7);
          for (u32 t = 0; t < 3; ++t)
            if (mask & (1u << t)) rd.refs[rd.na++] = 4 * g + t;
        }
        if (rs == 1 || rs == 2) rd.refs[rd.na++] = 4 * g + 3;
        if (second_gene) rd.refs[rd.na++] = 4 * ((g + 1) % G) + tx_pick;
        sort_small(rd.refs, rd.na);
      }
      if (sp.umi_err > 0 && rng.uni() < sp.umi_err) {
        u32 pos = rng.below((u32)sp.umi_len);
        u32 delta = 1 + rng.below(3);
        u32 base = (rd.umi >> (2 * pos)) & 3u;
        u32 nb = (base + delta) & 3u;
        rd.umi = (rd.umi & ~(3u << (2 * pos))) | (nb << (2 * pos));
      }
      out.push_back(rd);
    }
  }
  // Fisher-Yates shuffle of the records
  for (size_t i = out.size(); i > 1; --i) {
    size_t j = rng.below((u32)i);
    std::swap(out[i - 1], out[j]);
  }
}

template <class F>
void parallel_cells(u64 n, int n_threads, F f) {
  if (n_threads < 1) n_threads = 1;
  std::atomic<u64> next{0};
  auto work = [&]() {
    std::vector<Read> buf;
    for (;;) {
      u64 c0 = next.fetch_add(64);
      if (c0 >= n) break;
      u64 c1 = std::min(n, c0 + 64);
      for (u64 c = c0; c < c1; ++c) f(c, buf);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
}
}  // namespace

extern "C" {

/* number of transcripts / gene ids / output rows implied by a spec */
uint64_t afq_synth_num_refs(const afq_synth_spec* sp) { return (uint64_t)sp->n_genes * (sp->usa_mode ? 4 : 3); }
uint32_t afq_synth_num_gene_ids(const afq_synth_spec* sp) { return sp->usa_mode ? 2 * sp->n_genes : sp->n_genes; }
uint32_t afq_synth_num_rows(const afq_synth_spec* sp) { return sp->usa_mode ? 3 * sp->n_genes : sp->n_genes; }

/* tid_to_gid as parse_tg_map would build it (src/utils.rs:487-662) for the synthetic t2g */
void afq_synth_t2g(const afq_synth_spec* sp, uint32_t* tid_to_gid) {
  const u32 G = sp->n_genes;
  if (!sp->usa_mode) {
    for (u32 g = 0; g < G; ++g) for (u32 k = 0; k < 3; ++k) tid_to_gid[3 * g + k] = g;
  } else {
    for (u32 g = 0; g < G; ++g) {
      for (u32 k = 0; k < 3; ++k) tid_to_gid[4 * g + k] = 2 * g;
      tid_to_gid[4 * g + 3] = 2 * g + 1;
    }
  }
}

/* pass 1: per-cell record and ref counts for cells [first_cell, first_cell+n_cells) */
void afq_synth_sizes(const afq_synth_spec* sp, uint64_t first_cell, uint64_t n_cells,
                     uint64_t* cell_nrec, uint64_t* cell_nrefs, int n_threads) {
  const ZipfTable* zt = zipf_table(sp->n_genes, sp->zipf_s);
  parallel_cells(n_cells, n_threads, [&](u64 c, std::vector<Read>& buf) {
    gen_cell(*sp, zt, first_cell + c, buf);
    u64 nr = 0;
    for (auto& r : buf) nr += r.na;
    cell_nrec[c] = buf.size();
    cell_nrefs[c] = nr;
  });
}

/* pass 2: fill the SoA arrays. cell_rec_offsets[n_cells+1] and cell_ref_offsets[n_cells+1]
 * are the exclusive prefix sums of pass 1's counts. rec_ref_offsets has n_records+1. */
void afq_synth_fill(const afq_synth_spec* sp, uint64_t first_cell, uint64_t n_cells,
                    const uint64_t* cell_rec_offsets, const uint64_t* cell_ref_offsets,
                    uint32_t* rec_umi32, uint32_t* rec_ref_offsets, uint32_t* refs, int n_threads) {
  const ZipfTable* zt = zipf_table(sp->n_genes, sp->zipf_s);
  parallel_cells(n_cells, n_threads, [&](u64 c, std::vector<Read>& buf) {
    gen_cell(*sp, zt, first_cell + c, buf);
    u64 ri = cell_rec_offsets[c], fi = cell_ref_offsets[c];
    for (auto& r : buf) {
      rec_umi32[ri] = r.umi;
      rec_ref_offsets[ri] = (u32)fi;
      for (u32 k = 0; k < r.na; ++k) refs[fi++] = r.refs[k];
      ++ri;
    }
  });
  rec_ref_offsets[cell_rec_offsets[n_cells]] = (u32)cell_ref_offsets[n_cells];
}

}  // extern "C"
