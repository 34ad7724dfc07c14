// host_quant.cpp — C++ host side of `alevin-fry quant`: the replacement for
// quant::quantify / do_quantify (reference src/quant.rs:359-396, 1327-1951) with the worker
// pool (src/quant.rs:659-1325) delegated to the CUDA C-ABI (include/afq.h).
//
// Contract honoured: input-dir files, t2g parsing (src/utils.rs:487-662),
// record-type sniffing (src/utils.rs:313-377), output files and their formats
// (src/quant.rs:1588-1613, 1786-1847, 1913-1933). Rows are always emitted in chunk order
// (the reference's order with one worker, `-t 2`).
#include <atomic>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <fcntl.h>
#include <functional>
#include <memory>
#include <mutex>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>
#include <zlib.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/afq.h"
#include "../../include/afq_host.h"
#include "rad.h"

namespace {
using namespace afqh;

struct Fail { std::string msg; };
#define REQUIRE(cond, m) do { if (!(cond)) throw Fail{m}; } while (0)

bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }
void mkdirs(const std::string& p) {
  std::string cur;
  for (size_t i = 0; i <= p.size(); ++i) {
    if (i == p.size() || p[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0755); }
    if (i < p.size()) cur.push_back(p[i]);
  }
}
std::string slurp(const std::string& p) {
  std::ifstream f(p, std::ios::binary);
  std::stringstream ss; ss << f.rdbuf();
  return ss.str();
}
// minimal JSON probe: value of a top-level boolean key
bool json_bool(const std::string& js, const std::string& key, bool& out) {
  size_t k = js.find("\"" + key + "\"");
  if (k == std::string::npos) return false;
  size_t c = js.find(':', k);
  if (c == std::string::npos) return false;
  size_t v = js.find_first_not_of(" \t\r\n", c + 1);
  if (v == std::string::npos) return false;
  if (js.compare(v, 4, "true") == 0) { out = true; return true; }
  if (js.compare(v, 5, "false") == 0) { out = false; return true; }
  return false;
}
std::string json_escape(const std::string& s) {
  std::string o;
  for (unsigned char c : s) {
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; } else o.push_back((char)c);
    }
  }
  return o;
}

// Rust `{}` for f32: shortest round-trip digits, never scientific, "NaN" / "inf" spellings.
void append_f32(std::string& out, float v) {
  if (std::isnan(v)) { out += "NaN"; return; }
  if (std::isinf(v)) { out += v < 0 ? "-inf" : "inf"; return; }
  char buf[128];
  if (v >= 0.0f && v < 16777216.0f && v == (float)(uint32_t)v) {   // whole numbers print as integers
    auto r = std::to_chars(buf, buf + sizeof buf, (uint32_t)v);
    out.append(buf, r.ptr);
    return;
  }
  auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
  out.append(buf, r.ptr);
}
// ---- the matrix body is ~13 bytes x nnz of text: formatted with a two-digit table into uninitialised, geometrically grown
// buffers (std::string appends + std::to_chars cost ~200 ns per entry and a memset per reserve; this is ~15 ns) ----------------
static const char DIG2[] =
    "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
    "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
inline unsigned dec_len_u32(uint32_t v) {
  return v < 10u ? 1u : v < 100u ? 2u : v < 1000u ? 3u : v < 10000u ? 4u : v < 100000u ? 5u : v < 1000000u ? 6u : v < 10000000u ? 7u
       : v < 100000000u ? 8u : v < 1000000000u ? 9u : 10u;
}
inline char* put_u32(char* d, uint32_t v) {        // decimal digits of v at d; returns the end
  const unsigned n = dec_len_u32(v);
  char* e = d + n;
  char* q = e;
  while (v >= 100u) { const uint32_t r = v % 100u; v /= 100u; q -= 2; memcpy(q, DIG2 + 2 * r, 2); }
  if (v >= 10u) memcpy(q - 2, DIG2 + 2 * v, 2); else q[-1] = (char)('0' + v);
  return e;
}
inline char* put_u64(char* d, uint64_t v) {
  if (v <= 0xFFFFFFFFull) return put_u32(d, (uint32_t)v);
  char tmp[24];
  auto r = std::to_chars(tmp, tmp + sizeof tmp, v);
  memcpy(d, tmp, (size_t)(r.ptr - tmp));
  return d + (r.ptr - tmp);
}
struct TextBuf {   // a growable char buffer that never zero-fills
  char* p = nullptr; size_t n = 0, cap = 0;
  TextBuf() = default;
  TextBuf(TextBuf&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
  TextBuf& operator=(TextBuf&& o) noexcept { if (this != &o) { free(p); p = o.p; n = o.n; cap = o.cap; o.p = nullptr; o.n = o.cap = 0; } return *this; }
  TextBuf(const TextBuf&) = delete;
  TextBuf& operator=(const TextBuf&) = delete;
  ~TextBuf() { free(p); }
  void room(size_t extra) {             // at least `extra` writable bytes behind n
    if (n + extra <= cap) return;
    size_t nc = cap ? cap + cap / 2 : 4096;
    if (nc < n + extra) nc = n + extra;
    char* q = (char*)realloc(p, nc);
    if (!q) throw Fail{"out of memory while formatting the matrix"};
    p = q; cap = nc;
  }
  const char* data() const { return p; }
  size_t size() const { return n; }
  char operator[](size_t i) const { return p[i]; }
};
// one matrix entry "row col val\n" (row text preformatted); whole numbers below 2^24 take the integer path (Rust prints 3, not 3.0)
inline void put_entry(TextBuf& b, const char* row_txt, unsigned row_len, uint32_t col1, float v) {
  b.room(row_len + 1 + 10 + 1 + 64 + 1);
  char* d = b.p + b.n;
  memcpy(d, row_txt, row_len); d += row_len;
  *d++ = ' ';
  d = put_u32(d, col1);
  *d++ = ' ';
  if (v >= 0.0f && v < 16777216.0f && v == (float)(uint32_t)v) d = put_u32(d, (uint32_t)v);
  else if (std::isnan(v)) { memcpy(d, "NaN", 3); d += 3; }
  else if (std::isinf(v)) { const char* t = v < 0 ? "-inf" : "inf"; const size_t l = strlen(t); memcpy(d, t, l); d += l; }
  else {
    char tmp[128];
    auto r = std::to_chars(tmp, tmp + sizeof tmp, v, std::chars_format::fixed);
    const size_t l = (size_t)(r.ptr - tmp);
    b.n = (size_t)(d - b.p);
    b.room(l + 2);
    d = b.p + b.n;
    memcpy(d, tmp, l); d += l;
  }
  *d++ = '\n';
  b.n = (size_t)(d - b.p);
}

void append_u64(std::string& out, uint64_t v) {
  char buf[24];
  auto r = std::to_chars(buf, buf + sizeof buf, v);
  out.append(buf, r.ptr);
}

// needletail bitmer_to_bytes: A=0 C=1 G=2 T=3, first base in the most significant used bits
std::string decode_barcode(uint64_t bc, unsigned len) {
  static const char N[4] = {'A', 'C', 'G', 'T'};
  std::string s(len, 'A');
  for (unsigned i = 0; i < len; ++i) s[len - 1 - i] = N[(bc >> (2 * i)) & 3];
  return s;
}
bool encode_barcode(const std::string& s, uint64_t& out) {
  out = 0;
  for (char c : s) {
    uint64_t v;
    switch (c) { case 'A': case 'a': case 'N': v = 0; break; case 'C': case 'c': v = 1; break;
                 case 'G': case 'g': v = 2; break; case 'T': case 't': v = 3; break; default: return false; }
    out = (out << 2) | v;
  }
  return true;
}

struct T2G {
  std::vector<uint32_t> tid_to_gid;
  std::vector<std::string> gene_names;
  bool usa = false;
  uint32_t num_gene_ids = 0, num_rows = 0;
};

// parse_tg_map (src/utils.rs:487-662): 2 columns => gene ids in first-seen order; 3 columns =>
// spliced id 2k / unspliced 2k+1. Every RAD reference must be mapped.
T2G parse_t2g(const std::string& path, const std::vector<std::string>& ref_names) {
  std::unordered_map<std::string, uint32_t> rname;
  rname.reserve(ref_names.size() * 2);
  for (uint32_t i = 0; i < ref_names.size(); ++i) rname.emplace(ref_names[i], i);
  std::ifstream f(path);
  REQUIRE(f.good(), "couldn't open file " + path);
  T2G t;
  t.tid_to_gid.assign(ref_names.size(), UINT32_MAX);
  std::unordered_map<std::string, uint32_t> gid;
  std::string line;
  int ncol = 0;
  size_t found = 0;
  uint32_t next_gid = 0;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    std::vector<std::string> col;
    size_t a = 0;
    for (;;) { size_t b = line.find('\t', a); col.push_back(line.substr(a, b == std::string::npos ? b : b - a)); if (b == std::string::npos) break; a = b + 1; }
    if (ncol == 0) {
      ncol = (int)col.size();
      REQUIRE(ncol == 2 || ncol == 3, "Transcript-gene mapping must have either 2 or 3 columns.");
      t.usa = ncol == 3;
    }
    REQUIRE((int)col.size() == ncol, "failed to parse the transcript-to-gene map : inconsistent column count.");
    uint32_t g;
    auto it = gid.find(col[1]);
    if (it == gid.end()) {
      g = t.usa ? next_gid : (uint32_t)gid.size();
      next_gid += 2;
      gid.emplace(col[1], g);
      t.gene_names.push_back(col[1]);
    } else g = it->second;
    auto rt = rname.find(col[0]);
    if (rt != rname.end()) {
      ++found;
      if (!t.usa) t.tid_to_gid[rt->second] = g;
      else if (col[2] == "U" || col[2] == "u") t.tid_to_gid[rt->second] = g + 1;
      else if (col[2] == "S" || col[2] == "s") t.tid_to_gid[rt->second] = g;
      else throw Fail{"Third column in 3 column txp-to-gene file must be S or U"};
    }
  }
  REQUIRE(found == ref_names.size(), "The tg-map must contain a gene mapping for all transcripts in the header");
  for (uint32_t v : t.tid_to_gid) REQUIRE(v != UINT32_MAX, "The tg-map must contain a gene mapping for all transcripts in the header");
  const uint32_t G = (uint32_t)t.gene_names.size();
  if (t.usa) { t.num_gene_ids = 2 * G; t.num_rows = 3 * G; }  // mid = max id + 2 = 2G; mid + mid/2 (src/quant.rs:1627-1645)
  else { t.num_gene_ids = G; t.num_rows = G; }
  return t;
}

const char* resolution_debug_name(const std::string& r) {  // ResolutionStrategy Debug names (src/quant.rs:81-97)
  if (r == "trivial") return "Trivial";
  if (r == "cr-like") return "CellRangerLike";
  if (r == "cr-like-em") return "CellRangerLikeEm";
  if (r == "parsimony-em") return "ParsimonyEm";
  if (r == "parsimony") return "Parsimony";
  if (r == "parsimony-gene-em") return "ParsimonyGeneEm";
  if (r == "parsimony-gene") return "ParsimonyGene";
  return nullptr;
}
int resolution_code(const std::string& r) {
  if (r == "trivial") return AFQ_RES_TRIVIAL;
  if (r == "cr-like") return AFQ_RES_CR_LIKE;
  if (r == "cr-like-em") return AFQ_RES_CR_LIKE_EM;
  if (r == "parsimony-em") return AFQ_RES_PARSIMONY_EM;
  if (r == "parsimony") return AFQ_RES_PARSIMONY;
  if (r == "parsimony-gene-em") return AFQ_RES_PARSIMONY_GENE_EM;
  if (r == "parsimony-gene") return AFQ_RES_PARSIMONY_GENE;
  return -1;
}

// Page tables of the mapped input, filled ahead of the chunk walk by a helper thread (MADV_POPULATE_READ, Linux >= 5.14) instead
// of one minor fault per chunk header on the walking thread (1.5 us per chunk: 0.16 s of the 0.58 s pipeline on 100 k cells).
// Best effort: where the call is not supported the walk faults the pages in as before.
struct Prefault {
  std::thread th;
  std::atomic<bool> stop{false};
  void start(const unsigned char* base, uint64_t size) {
#ifdef MADV_POPULATE_READ
    if (!base || !size || getenv("AFQ_NO_PREFAULT")) return;
    th = std::thread([this, base, size] {
      const uint64_t step = 32ull << 20;
      const uintptr_t a0 = (uintptr_t)base & ~(uintptr_t)4095;
      const uintptr_t a1 = (uintptr_t)base + size;
      for (uintptr_t a = a0; a < a1 && !stop.load(std::memory_order_relaxed); a += step)
        if (madvise((void*)a, (size_t)std::min<uint64_t>(step, a1 - a), MADV_POPULATE_READ) != 0) return;
    });
#else
    (void)base; (void)size;
#endif
  }
  ~Prefault() { stop = true; if (th.joinable()) th.join(); }
};

// afqh_host_stage_bench (no GPU: the host stages alone) keeps the batches in ordinary memory
static bool g_plain_host_memory = false;

template <class T>
struct Pinned {  // growable pinned array (afq_host_alloc)
  T* p = nullptr; size_t n = 0, cap = 0;
  bool plain = false;
  ~Pinned() { release(p); }
  void release(T* q) { if (q) { if (plain) free(q); else afq_host_free(q); } }
  void reserve(size_t want) {
    if (want <= cap) return;
    // batches close one cell past the record budget, so their sizes differ by a few per cent: the first allocation carries
    // 1/8 of head-room and a second one (pinning 100+ MB costs ~50 ms) is normally never needed
    size_t nc = cap ? cap : want + want / 8 + 1024;
    while (nc < want) nc = nc + nc / 2 + 1024;
    void* q = nullptr;
    const bool was_plain = plain;
    if (g_plain_host_memory) { q = malloc(nc * sizeof(T)); if (!q) throw Fail{"host allocation failed"}; }
    else if (afq_host_alloc(&q, nc * sizeof(T)) != AFQ_OK) throw Fail{"pinned host allocation failed"};
    if (n) memcpy(q, p, n * sizeof(T));
    T* old = p;
    plain = was_plain; release(old);
    plain = g_plain_host_memory;
    p = (T*)q; cap = nc;
  }
  void push(T v) { if (n == cap) reserve(n + 1); p[n++] = v; }
  void clear() { n = 0; }
};

// fork-join helper over persistent threads: run(f) calls f(tid) on every worker and on the caller
class Pool {
 public:
  explicit Pool(unsigned n) : n_(n < 1 ? 1 : n) {
    for (unsigned i = 1; i < n_; ++i) th_.emplace_back([this, i] { loop(i); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++gen_; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  unsigned size() const { return n_; }
  void run(const std::function<void(unsigned)>& f) {
    { std::lock_guard<std::mutex> lk(m_); job_ = &f; pending_ = n_ - 1; ++gen_; }
    cv_.notify_all();
    f(0);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
    job_ = nullptr;
  }
 private:
  void loop(unsigned tid) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(unsigned)>* j;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        j = job_;
      }
      (*j)(tid);
      { std::lock_guard<std::mutex> lk(m_); if (--pending_ == 0) done_.notify_one(); }
    }
  }
  unsigned n_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(unsigned)>* job_ = nullptr;
  unsigned pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// one collated chunk (= one cell), located by the header walk; everything a parser thread needs
// to write the chunk's records straight into their final SoA positions
struct ChunkInfo {
  const unsigned char* body;     // first record
  uint32_t body_bytes, nrec;
  uint32_t n_aln;                // alignments in the chunk (derived from nbytes: fixed-size records)
  uint64_t bc;
  uint64_t rec_off, ref_off;     // positions inside the batch
};

struct HostBatch {   // pinned SoA arrays of one device batch, sized exactly by the header walk
  Pinned<uint64_t> cell_rec_off;
  Pinned<uint32_t> umi, ref_off, refs;
  Pinned<uint8_t> na8;           // alignment count per record (sent instead of ref_off when all <= 255)
  Pinned<uint8_t> umi24, refs24; // 3-byte UMIs / transcript ids instead of umi / refs (afq_batch.rec_umi24 / refs24)
  std::vector<ChunkInfo> chunks;
  uint64_t first_cell = 0, n_rec = 0, n_ref = 0;
  bool pack24 = false;
  std::atomic<bool> wide_na{false};   // some record has > 255 alignments
  uint64_t n_cells() const { return chunks.size(); }
};

// unmapped_bc_count_collated.bin (src/quant.rs:1484-1494): corrected barcode -> number of unmapped reads.
// libradicl 0.18's CollatedUnmappedCounts layout is not available here (the crate is not vendored); what IS
// stated in the reference tree is the bincode form of a HashMap<u64, u32> (src/atac/collate.rs:270-283: u64
// entry count, then (u64 key, u32 value) pairs, little-endian), which is also what alevin-fry wrote before the
// libradicl type existed. That form is read exactly (size-checked); anything else is reported and treated like
// the reference treats an unreadable file (empty map), so that the run never silently pretends to know it.
struct UnmappedCounts {
  std::unordered_map<uint64_t, uint32_t> map;
  bool present = false, understood = false;
  uint32_t get(uint64_t bc) const { auto it = map.find(bc); return it == map.end() ? 0u : it->second; }
};
UnmappedCounts load_unmapped(const std::string& path) {
  UnmappedCounts u;
  if (!file_exists(path)) return u;
  u.present = true;
  const std::string raw = slurp(path);
  if (raw.size() >= 8) {
    uint64_t n; memcpy(&n, raw.data(), 8);
    if (n <= raw.size() / 12 && raw.size() == 8 + n * 12) {
      u.map.reserve(n * 2);
      for (uint64_t i = 0; i < n; ++i) {
        uint64_t k; uint32_t v;
        memcpy(&k, raw.data() + 8 + i * 12, 8); memcpy(&v, raw.data() + 16 + i * 12, 4);
        u.map[k] += v;
      }
      u.understood = true;
    }
  } else if (raw.empty()) u.understood = true;   // "no unmapped reads"
  return u;
}

// --dump-eqclasses: the GLOBAL gene-level eq-class table (EqcMap, src/quant.rs:218-229). Ids are handed out in first-seen
// order — cells in row order, a cell's classes in canonical label order (the reference: hash order under one mutex,
// src/quant.rs:1282-1307; the id NUMBERING is arbitrary there too, infer only needs it to be consistent).
struct LabelVecHash {
  size_t operator()(const std::vector<uint32_t>& v) const {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (v.size() * 0xD6E8FEB86659FD93ull);
    for (uint32_t x : v) { h ^= x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); h *= 0xFF51AFD7ED558CCDull; h ^= h >> 32; }
    return (size_t)h;
  }
};
struct EqcMap {
  std::unordered_map<std::vector<uint32_t>, uint64_t, LabelVecHash> global_eqc;
  std::vector<const std::vector<uint32_t>*> by_id;          // id -> label (stable: unordered_map nodes do not move)
  std::vector<std::pair<uint64_t, uint32_t>> cell_level_count;
  std::vector<std::pair<uint64_t, uint64_t>> cell_offset;   // (row index, classes of the cell)
};

struct Outputs {
  FILE* rows = nullptr;
  FILE* feat = nullptr;
  std::vector<TextBuf> mtx_chunks;
  uint64_t nnz = 0, row_index = 0;
  std::vector<uint64_t> alt, empty, tiny;
  EqcMap eqc;
};

// parse the chunks of a batch in parallel (ingest must not be the bottleneck)
void parse_batch(HostBatch& b, const RecordLayout& lay, Pool& pool, std::string& failure) {
  const size_t nc = b.chunks.size();
  b.cell_rec_off.reserve(nc + 1); b.cell_rec_off.n = nc + 1;
  b.na8.reserve(b.n_rec + 16); b.na8.n = b.n_rec;
  if (b.pack24) { b.umi24.reserve(3 * b.n_rec + 16); b.refs24.reserve(3 * b.n_ref + 16); }
  else { b.umi.reserve(b.n_rec + 4); b.refs.reserve(b.n_ref + 4); }
  b.wide_na = false;
  std::atomic<size_t> next{0};
  std::atomic<bool> bad{false};
  std::mutex fm;
  pool.run([&](unsigned) {
    for (;;) {
      const size_t c0 = next.fetch_add(8);
      if (c0 >= nc || bad.load(std::memory_order_relaxed)) return;
      const size_t c1 = c0 + 8 < nc ? c0 + 8 : nc;
      for (size_t c = c0; c < c1; ++c) {
        const ChunkInfo& ci = b.chunks[c];
        b.cell_rec_off.p[c] = ci.rec_off;
        const unsigned char* p = ci.body;
        const unsigned char* end = p + ci.body_bytes;
        uint64_t rec = ci.rec_off, ref = ci.ref_off;
        const uint64_t ref_end = ci.ref_off + ci.n_aln;
        bool ok = true;
        // the usual layout (a 4-byte alignment = the compressed reference id, a UMI of at most 4 bytes) with 24-bit wire arrays:
        // fixed strides the compiler can see, and 4-byte stores for the 3-byte values — the 4th byte lands on the next value's
        // first byte, which is written afterwards; only a chunk's LAST value is stored byte by byte (the next one belongs to
        // another thread)
        if (b.pack24 && lay.aln_bytes == 4 && lay.refid_off == 0 && lay.umi_size <= 4) {
          const size_t rb = lay.read_bytes, uo = lay.umi_off, us = lay.umi_size;
          const uint64_t rec_last = ci.rec_off + ci.nrec - 1;
          uint8_t* const u24 = b.umi24.p;
          uint8_t* const r24 = b.refs24.p;
          uint8_t* const na8 = b.na8.p;
          bool wide = false;
          for (uint32_t r = 0; r < ci.nrec; ++r) {
            if (p + 4 + rb > end) { ok = false; break; }
            uint32_t na; memcpy(&na, p, 4);
            uint32_t rumi = 0; memcpy(&rumi, p + 4 + uo, us);
            p += 4 + rb;
            if (na == 0 || (uint64_t)na > ref_end - ref || p + (size_t)na * 4 > end) { ok = false; break; }
            if (rec != rec_last) memcpy(u24 + 3 * rec, &rumi, 4);
            else { uint8_t* d = u24 + 3 * rec; d[0] = (uint8_t)rumi; d[1] = (uint8_t)(rumi >> 8); d[2] = (uint8_t)(rumi >> 16); }
            na8[rec] = (uint8_t)na;
            wide |= na > 255;
            uint32_t a = 0;
            if (na <= 4 && ref + 5 <= ref_end && p + 16 <= end) {
              // short record well inside its chunk: four unconditional copies instead of a loop whose trip count the branch
              // predictor cannot guess (mean 1.9 alignments); what is written past `na` is overwritten by the next records
              for (int q = 0; q < 4; ++q) {
                uint32_t id; memcpy(&id, p + 4 * q, 4);
                id &= 0x7FFFFFFFu;
                memcpy(r24 + 3 * (ref + q), &id, 4);
              }
              a = na;
            }
            const uint32_t na_fast = (ref + na == ref_end) ? na - 1 : na;     // (all but the chunk's last alignment)
            for (; a < na_fast; ++a) {
              uint32_t id; memcpy(&id, p + 4 * (size_t)a, 4);
              id &= 0x7FFFFFFFu;   // bit 31 = orientation (src/convert.rs:442-445)
              memcpy(r24 + 3 * (ref + a), &id, 4);
            }
            for (; a < na; ++a) {
              uint32_t id; memcpy(&id, p + 4 * (size_t)a, 4);
              id &= 0x7FFFFFFFu;
              uint8_t* d = r24 + 3 * (ref + a); d[0] = (uint8_t)id; d[1] = (uint8_t)(id >> 8); d[2] = (uint8_t)(id >> 16);
            }
            p += (size_t)na * 4;
            ref += na;
            ++rec;
          }
          if (wide) b.wide_na.store(true, std::memory_order_relaxed);
        } else
        for (uint32_t r = 0; r < ci.nrec && ok; ++r) {
          if (p + 4 + lay.read_bytes > end) { ok = false; break; }
          uint32_t na; memcpy(&na, p, 4); p += 4;
          uint64_t rumi = 0;
          memcpy(&rumi, p + lay.umi_off, lay.umi_size);
          p += lay.read_bytes;
          // (a record without alignments cannot come out of a mapper / collate: the reference's writer asserts it,
          // src/convert.rs:134, and its resolvers index label[0])
          if (na == 0 || (uint64_t)na > ref_end - ref || p + (size_t)na * lay.aln_bytes > end) { ok = false; break; }
          if (b.pack24) { uint8_t* d = b.umi24.p + 3 * rec; d[0] = (uint8_t)rumi; d[1] = (uint8_t)(rumi >> 8); d[2] = (uint8_t)(rumi >> 16); }
          else b.umi.p[rec] = (uint32_t)rumi;
          b.na8.p[rec] = (uint8_t)na;
          if (na > 255) b.wide_na.store(true, std::memory_order_relaxed);
          for (uint32_t a = 0; a < na; ++a) {
            uint32_t id; memcpy(&id, p + lay.refid_off, 4); p += lay.aln_bytes;
            id &= 0x7FFFFFFFu;   // bit 31 = orientation (src/convert.rs:442-445)
            if (b.pack24) { uint8_t* d = b.refs24.p + 3 * ref; d[0] = (uint8_t)id; d[1] = (uint8_t)(id >> 8); d[2] = (uint8_t)(id >> 16); }
            else b.refs.p[ref] = id;
            ++ref;
          }
          ++rec;
        }
        if (!ok || ref != ref_end || p != end) {
          bad = true;
          std::lock_guard<std::mutex> lk(fm);
          if (failure.empty()) failure = "record without alignments or overrunning its chunk (corrupt collated RAD, chunk " + std::to_string(b.first_cell + c) + ")";
          return;
        }
      }
    }
  });
  b.cell_rec_off.p[nc] = b.n_rec;
  if (!bad && b.wide_na) {   // rare: CSR offsets instead of 1-byte counts (second walk over the record headers)
    b.ref_off.reserve(b.n_rec + 2); b.ref_off.n = b.n_rec + 1;
    next = 0;
    pool.run([&](unsigned) {
      for (;;) {
        const size_t c = next.fetch_add(1);
        if (c >= nc) return;
        const ChunkInfo& ci = b.chunks[c];
        const unsigned char* p = ci.body;
        uint64_t ref = ci.ref_off;
        for (uint32_t r = 0; r < ci.nrec; ++r) {
          uint32_t na; memcpy(&na, p, 4);
          b.ref_off.p[ci.rec_off + r] = (uint32_t)ref;
          ref += na;
          p += 4 + lay.read_bytes + (size_t)na * lay.aln_bytes;
        }
      }
    });
    b.ref_off.p[b.n_rec] = (uint32_t)b.n_ref;
  }
}

// text for rows / featureDump / mtx of one finished batch, formatted in parallel (N2)
void consume(const HostBatch& hb, const afq_result& r, unsigned bc_len, const UnmappedCounts& unmapped, Outputs& o, Pool& pool) {
  constexpr uint64_t BLK = 256;
  const uint64_t nblk = (r.n_cells + BLK - 1) / BLK;
  std::vector<std::string> rows(nblk), feat(nblk);
  std::vector<TextBuf> mtx(nblk);
  std::atomic<uint64_t> next{0};
  const uint64_t row0 = o.row_index;
  pool.run([&](unsigned) {
    for (;;) {
      const uint64_t k = next.fetch_add(1);
      if (k >= nblk) return;
      std::string& rs = rows[k]; std::string& fs = feat[k]; TextBuf& ms = mtx[k];
      const uint64_t c0 = k * BLK, c1 = c0 + BLK < r.n_cells ? c0 + BLK : r.n_cells;
      ms.room((r.row_ptr[c1] - r.row_ptr[c0]) * 14 + 128);
      for (uint64_t c = c0; c < c1; ++c) {
        const std::string bc = decode_barcode(hb.chunks[c].bc, bc_len);
        rs += bc; rs.push_back('\n');
        // featureDump (src/quant.rs:1181-1196, 1248-1260)
        const uint32_t num_mapped = hb.chunks[c].nrec, num_unmapped = unmapped.get(hb.chunks[c].bc);
        const float sum_umi = r.sum_umi[c], max_umi = r.max_umi[c];
        const float dedup_rate = sum_umi / (float)num_mapped;
        const float mapping_rate = (float)num_mapped / (float)(num_mapped + num_unmapped);
        const float mean_expr = sum_umi / (float)r.num_expr[c];
        const float mean_by_max = mean_expr / max_umi;
        fs += bc; fs.push_back('\t');
        append_u64(fs, (uint64_t)num_mapped + num_unmapped); fs.push_back('\t');
        append_u64(fs, num_mapped); fs.push_back('\t');
        append_f32(fs, sum_umi); fs.push_back('\t');
        append_f32(fs, mapping_rate); fs.push_back('\t');
        append_f32(fs, dedup_rate); fs.push_back('\t');
        append_f32(fs, mean_by_max); fs.push_back('\t');
        append_u64(fs, r.num_expr[c]); fs.push_back('\t');
        append_u64(fs, r.num_over_mean[c]); fs.push_back('\n');
        char row_txt[24];
        const unsigned row_len = (unsigned)(put_u64(row_txt, row0 + c + 1) - row_txt);
        for (uint64_t k2 = r.row_ptr[c]; k2 < r.row_ptr[c + 1]; ++k2) put_entry(ms, row_txt, row_len, r.col[k2] + 1u, r.val[k2]);
      }
    }
  });
  for (uint64_t c = 0; c < r.n_cells; ++c) {
    const uint64_t cell_num = hb.first_cell + c;
    if (r.flags[c] & AFQ_FLAG_ALT) o.alt.push_back(cell_num);
    if (r.flags[c] & AFQ_FLAG_TINY) o.tiny.push_back(cell_num);
    if (r.flags[c] & AFQ_FLAG_EMPTY) o.empty.push_back(cell_num);
  }
  o.row_index += r.n_cells;
  o.nnz += r.nnz;
  for (uint64_t k = 0; k < nblk; ++k) {
    fwrite(rows[k].data(), 1, rows[k].size(), o.rows);
    fwrite(feat[k].data(), 1, feat[k].size(), o.feat);
    o.mtx_chunks.push_back(std::move(mtx[k]));
  }
}

std::string lower(std::string s) { for (auto& c : s) c = (char)tolower((unsigned char)c); return s; }

void json_u64_list(std::string& js, const std::vector<uint64_t>& v, const std::string& ind) {
  if (v.empty()) { js += "[]"; return; }
  js += "[\n";
  for (size_t i = 0; i < v.size(); ++i) { js += ind + "  "; append_u64(js, v[i]); js += i + 1 < v.size() ? ",\n" : "\n"; }
  js += ind + "]";
}

int quantify_impl(const afqh_quant_opts& o) {
  const auto t_entry = std::chrono::steady_clock::now();
  REQUIRE(o.input_dir && o.tg_map && o.output_dir && o.resolution, "input_dir, tg_map, output_dir and resolution are required");
  const std::string in = o.input_dir, out = o.output_dir;
  const std::string res = lower(o.resolution);
  REQUIRE(resolution_code(res) >= 0, "invalid value '" + std::string(o.resolution) + "' for '--resolution <RESOLUTION>'");
  const std::string sa = o.sa_model ? lower(o.sa_model) : "winner-take-all";
  REQUIRE(sa == "winner-take-all" || sa == "prefer-ambig", "invalid value for '--sa-model'");
  REQUIRE(!(o.dump_eq && res == "trivial"), "Gene equivalence classes are not meaningful in case of Trivial resolution.");   // src/main.rs:705-711
  REQUIRE(o.num_bootstraps == 0, "bootstrapping (-b) is not implemented on the CUDA path (the reference's RNG is unseeded; see SURVEY.md §8(f) N4)");
  // src/main.rs:733-734, 759, 812-820
  REQUIRE(file_exists(in + "/generate_permit_list.json"), "The input directory " + in + " did not contain a generate_permit_list.json file; please run generate-permit-list and collate first.");
  bool velo = false;
  json_bool(slurp(in + "/generate_permit_list.json"), "velo_mode", velo);
  REQUIRE(!velo, "velo_mode quantification is not implemented (reference: unimplemented!, src/quant.rs:2033)");
  // src/quant.rs:364-371
  REQUIRE(file_exists(in + "/collate.json"), "could not open the collate.json file.");
  bool compressed = false;
  REQUIRE(json_bool(slurp(in + "/collate.json"), "compressed_output", compressed), "could not read compressed_output field from collate metadata.");
  // src/quant.rs:373-395: map.collated.rad, or the snappy-framed map.collated.rad.sz of `collate --compress`
  const std::string rad_path = in + (compressed ? "/map.collated.rad.sz" : "/map.collated.rad");
  std::vector<unsigned char> zimage;   // the decompressed RAD image (compressed input only)
  std::vector<char> iobuf(8 << 20);
  FILE* f = nullptr;
  if (compressed) {
    const int zfd = open(rad_path.c_str(), O_RDONLY);
    struct stat zst {};
    REQUIRE(zfd >= 0 && fstat(zfd, &zst) == 0, "run collate before quant (could not open " + rad_path + ")");
    const size_t zsize = (size_t)zst.st_size;
    const unsigned char* zmap = zsize ? (const unsigned char*)mmap(nullptr, zsize, PROT_READ, MAP_PRIVATE, zfd, 0) : nullptr;
    close(zfd);
    REQUIRE(!zsize || zmap != MAP_FAILED, "mmap failed for " + rad_path);
    std::string zerr;
    const bool zok = snappy_framed_decompress(zmap, zsize, zimage, std::max(1u, std::min(o.num_threads, 64u)), zerr);
    if (zmap) munmap((void*)zmap, zsize);
    REQUIRE(zok, rad_path + ": " + zerr);
    f = fmemopen(zimage.data(), zimage.size() ? zimage.size() : 1, "rb");
    REQUIRE(f, "fmemopen failed");
  } else {
    f = fopen(rad_path.c_str(), "rb");
    REQUIRE(f, "run collate before quant (could not open " + rad_path + ")");
    setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
  }
  Reader rd(f);
  RadPrelude pre;
  // Hardware work queues: libafq asks for 32 when it is loaded (a host that runs batches ahead of the GPU needs them, afq_cuda.cu),
  // but a context with 32 queues takes 1-2 s longer to create (r2af: 0.58 s -> 1.65 s of set-up on one box) and this tool is
  // bound by its own parse / format stages, not by the GPU: it keeps CUDA's default of 8 unless AFQ_HW_QUEUES says otherwise.
  // (Must happen before the first CUDA call, i.e. before the warm-up thread below.)
  {
    const char* q = getenv("AFQ_HW_QUEUES");
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", (q && atoi(q) > 0) ? q : "8", 1);
  }
  // the CUDA context comes up (~0.3-0.5 s) in the background while the prelude and the t2g map are parsed
  std::thread cuda_warm([] { void* w = nullptr; if (afq_host_alloc(&w, 4096) == AFQ_OK) afq_host_free(w); });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } warm_join{cuda_warm};
  std::string perr;
  if (!parse_prelude(rd, pre, perr)) { fclose(f); throw Fail{"RAD prelude: " + perr}; }
  // record-type sniffing (src/utils.rs:313-377)
  if (auto nb = pre.file_tag("num_barcodes")) { if (nb->u > 1) { fclose(f); throw Fail{"multi-barcode RAD files are not supported by this build (SURVEY.md §8(f) N4)"}; } }
  if (RadPrelude::has(pre.aln_tags, "as") && RadPrelude::has(pre.aln_tags, "start") && RadPrelude::has(pre.aln_tags, "end")) { fclose(f); throw Fail{"long-read RAD files are not supported by this build (SURVEY.md §8(f) N4)"}; }
  if (RadPrelude::has(pre.aln_tags, "type") && RadPrelude::has(pre.aln_tags, "start_pos") && RadPrelude::has(pre.aln_tags, "frag_len")) { fclose(f); throw Fail{"To process atac-seq data, you should use the \"atac\" sub-command"}; }
  const TagValue* cbl = pre.file_tag("cblen");
  if (!cbl) cbl = pre.file_tag("b1len");
  if (!cbl) cbl = pre.file_tag("b0len");
  if (!cbl) { fclose(f); throw Fail{"tag map must contain cblen or bNlen for barcode length"}; }
  const unsigned bc_len = (unsigned)cbl->u;
  const TagValue* ul = pre.file_tag("ulen");
  RecordLayout lay;
  if (!make_layout(pre, lay, perr)) { fclose(f); throw Fail{perr}; }
  // the reference's quant never reads `ulen` (it compares packed UMIs); without the tag the widest UMI the
  // stored integer can hold is assumed, which is exact for the 1-edit enumeration (absent bases are zero bits)
  const unsigned umi_len = ul ? (unsigned)ul->u : (unsigned)(4 * lay.umi_size);
  if (umi_len > 16 || lay.umi_size > 8) { fclose(f); throw Fail{"UMIs longer than 16 bases are not supported on the CUDA path"}; }

  const UnmappedCounts unmapped = load_unmapped(in + "/unmapped_bc_count_collated.bin");
  if (unmapped.present && !unmapped.understood)
    fprintf(stderr, "warning: %s/unmapped_bc_count_collated.bin is not in the bincode HashMap<u64,u32> form this build reads "
                    "(libradicl's CollatedUnmappedCounts layout is unavailable); CorrectedReads / MappingRate in featureDump.txt "
                    "are computed with 0 unmapped reads per barcode, as the reference does when it cannot read the file\n", in.c_str());
  const auto t_t2g0 = std::chrono::steady_clock::now();
  T2G t2g = parse_t2g(o.tg_map, pre.ref_names);
  const auto t_t2g1 = std::chrono::steady_clock::now();

  // --quant-subset (src/utils.rs:1074-1095): one barcode per line
  bool filtering = false;
  std::unordered_set<uint64_t> keep;
  uint64_t num_cells = pre.num_chunks;
  if (o.filter_list && *o.filter_list) {
    std::ifstream fl(o.filter_list);
    if (!fl.good()) { fclose(f); throw Fail{"couldn't open file " + std::string(o.filter_list)}; }
    std::string line;
    while (std::getline(fl, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.empty()) continue;
      uint64_t v;
      if (!encode_barcode(line.substr(0, bc_len), v)) { fclose(f); throw Fail{"bad barcode in --quant-subset file: " + line}; }
      keep.insert(v);
    }
    filtering = true;
    num_cells = keep.size();
  }

  afq_config cfg{};
  cfg.resolution = resolution_code(res);
  cfg.usa_mode = t2g.usa;
  cfg.em_init_uniform = o.init_uniform;
  cfg.pug_exact_umi = o.pug_exact_umi;
  // src/quant.rs:1457-1467: prefer-ambig only makes sense in USA mode; otherwise it is ignored (with a note)
  bool prefer_ambig = sa == "prefer-ambig";
  if (prefer_ambig && !t2g.usa) {
    fprintf(stderr, "When not operating in USA-mode (all-in-one unspliced/spliced/ambiguous), the SplicedAmbiguityModel will be ignored.\n");
    prefer_ambig = false;
  }
  cfg.sa_model = prefer_ambig ? AFQ_SA_PREFER_AMBIG : AFQ_SA_WINNER_TAKE_ALL;
  cfg.num_gene_ids = t2g.num_gene_ids;
  cfg.num_rows = t2g.num_rows;
  cfg.small_thresh = o.small_thresh;
  cfg.large_graph_thresh = o.large_graph_thresh;
  cfg.barcode_len = (uint16_t)bc_len;
  cfg.umi_len = (uint16_t)umi_len;
  cfg.device = o.device;
  cfg.dump_eq = o.dump_eq ? 1 : 0;
  // one context per GPU of the device list (default: the single --device)
  std::vector<int> devs;
  if (o.devices && *o.devices) {
    const std::string dl = lower(o.devices);
    if (dl == "all") {
      int n = afq_device_count();
      for (int d = 0; d < n; ++d) devs.push_back(d);
    } else {
      size_t a = 0;
      for (;;) {
        const size_t b = dl.find(',', a);
        const std::string tok = dl.substr(a, b == std::string::npos ? b : b - a);
        if (tok.empty() || tok.find_first_not_of("0123456789") != std::string::npos) { fclose(f); throw Fail{"bad --devices list '" + dl + "' (expected e.g. 0,1,2,3 or all)"}; }
        devs.push_back(atoi(tok.c_str()));
        if (b == std::string::npos) break;
        a = b + 1;
      }
    }
    // (an ordinal may repeat: two contexts on one GPU are valid and let a single-GPU box exercise this path)
  }
  if (devs.empty()) devs.push_back(o.device);
  const int D = (int)devs.size();
  std::vector<afq_ctx*> ctxs((size_t)D, nullptr);
  auto destroy_all = [&] { for (auto& c : ctxs) if (c) { afq_destroy(c); c = nullptr; } };
  cuda_warm.join();
  const auto t_create0 = std::chrono::steady_clock::now();
  for (int d = 0; d < D; ++d) {
    cfg.device = devs[d];
    if (afq_create(&cfg, t2g.tid_to_gid.data(), t2g.tid_to_gid.size(), &ctxs[d]) != AFQ_OK) {
      std::string m = afq_last_error(nullptr);
      destroy_all();
      fclose(f);
      throw Fail{"afq_create: " + m};
    }
  }

  mkdirs(out);
  mkdirs(out + "/alevin");
  Outputs outs;
  outs.rows = fopen((out + "/alevin/quants_mat_rows.txt").c_str(), "wb");
  outs.feat = fopen((out + "/featureDump.txt").c_str(), "wb");
  if (!outs.rows || !outs.feat) { destroy_all(); fclose(f); throw Fail{"could not create output files in " + out}; }
  fputs("CB\tCorrectedReads\tMappedReads\tDeduplicatedReads\tMappingRate\tDedupRate\tMeanByMax\tNumGenesExpressed\tNumGenesOverMean\n", outs.feat);

  using clk = std::chrono::steady_clock;
  auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  const auto t_begin = clk::now();
  double t_wait = 0, t_format = 0, t_submit = 0, t_parse = 0, t_walk = 0;
  uint64_t cells_seen = 0, total_records = 0;
  std::string failure;

  // ---- map the file and index the chunks (cells): every chunk header is read once, sequentially;
  // record sizes are fixed, so a chunk's alignment count follows from its byte size and the SoA
  // position of every chunk is known before any record is parsed
  const uint64_t body_start = rd.pos();
  fclose(f);
  uint64_t fsize = 0;
  const unsigned char* fmap = nullptr;
  const bool mapped = !compressed;
  if (compressed) { fmap = zimage.data(); fsize = zimage.size(); }
  else {
    const int fd = open(rad_path.c_str(), O_RDONLY);
    struct stat st {};
    if (fd < 0 || fstat(fd, &st) != 0) { if (fd >= 0) close(fd); destroy_all(); fclose(outs.rows); fclose(outs.feat); throw Fail{"couldn't open " + rad_path}; }
    fsize = (uint64_t)st.st_size;
    fmap = fsize ? (const unsigned char*)mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (fsize && fmap == MAP_FAILED) { destroy_all(); fclose(outs.rows); fclose(outs.feat); throw Fail{"mmap failed for " + rad_path}; }
    madvise((void*)fmap, fsize, MADV_SEQUENTIAL);
  }
  Prefault prefault;
  if (mapped) prefault.start(fmap + body_start, fsize - body_start);

  // half of the threads parse the next batch while the other half format the previous result
  const unsigned n_threads = std::max(2u, std::min(o.num_threads < 2 ? 2u : o.num_threads, 128u));
  Pool pool(n_threads / 2), fmt_pool(n_threads - n_threads / 2);
  const uint64_t batch_records = o.batch_records ? o.batch_records : (16ull << 20);
  // three host batches in flight per GPU (the context's slot count); batch k goes to context k mod D and the
  // consumer collects in submission order, so the rows of the one output matrix stay in chunk order
  const int NB = 3 * D;
  auto hb_holder = std::make_unique<std::vector<HostBatch>>((size_t)NB);   // (on the heap: left behind when the process exits anyway)
  std::vector<HostBatch>& hb = *hb_holder;
  std::vector<int> dev_of((size_t)NB, 0);
  uint64_t batch_seq = 0;
  // 24-bit wire arrays whenever the chemistry allows: UMI <= 12 bases and fewer than 2^24 targets
  const bool pack24 = ul && umi_len <= 12 && lay.umi_size <= 4 && pre.ref_names.size() < (1u << 24) && !getenv("AFQ_NO_PACK24");
  for (auto& b : hb) b.pack24 = pack24;
  std::vector<uint64_t> tickets((size_t)NB, 0);
  std::vector<char> inflight((size_t)NB, 0);
  auto finish = [&](int i) {
    afq_result r{};
    const auto ta = clk::now();
    afq_ctx* ctx = ctxs[dev_of[i]];
    if (afq_wait(ctx, tickets[i], &r) != AFQ_OK) throw Fail{std::string("afq_wait: ") + afq_last_error(ctx)};
    const auto tb = clk::now();
    if (o.dump_eq) {   // record the cells' gene eq-classes before the rows are numbered by consume()
      afq_eqc_dump d{};
      if (afq_result_eqclasses(ctx, &r, &d) != AFQ_OK) throw Fail{std::string("afq_result_eqclasses: ") + afq_last_error(ctx)};
      std::vector<uint32_t> key;
      for (uint64_t c = 0; c < d.n_cells; ++c) {
        const uint64_t k0 = d.cell_cls_ptr[c], k1 = d.cell_cls_ptr[c + 1];
        for (uint64_t k = k0; k < k1; ++k) {
          key.assign(d.labels + d.cls_lab_ptr[k], d.labels + d.cls_lab_ptr[k + 1]);
          auto it = outs.eqc.global_eqc.find(key);
          if (it == outs.eqc.global_eqc.end()) {
            it = outs.eqc.global_eqc.emplace(key, (uint64_t)outs.eqc.by_id.size()).first;
            outs.eqc.by_id.push_back(&it->first);
          }
          outs.eqc.cell_level_count.push_back({it->second, d.counts[k]});
        }
        outs.eqc.cell_offset.push_back({outs.row_index + c, k1 - k0});
      }
    }
    consume(hb[i], r, bc_len, unmapped, outs, fmt_pool);
    t_wait += secs(ta, tb); t_format += secs(tb, clk::now());
    afq_result_release(ctx, &r);
  };
  // results are collected, formatted and written by a consumer thread, in submission order
  std::mutex qm;
  std::condition_variable qcv;
  std::deque<int> queue;
  bool producer_done = false;
  std::string consumer_failure;
  std::thread consumer([&] {
    for (;;) {
      int i;
      {
        std::unique_lock<std::mutex> lk(qm);
        qcv.wait(lk, [&] { return !queue.empty() || producer_done; });
        if (queue.empty()) return;
        i = queue.front(); queue.pop_front();
      }
      std::string msg;
      try {
        if (consumer_failure.empty()) finish(i);
        else { afq_result r{}; afq_ctx* ctx = ctxs[dev_of[i]]; if (afq_wait(ctx, tickets[i], &r) == AFQ_OK) afq_result_release(ctx, &r); }
      } catch (const Fail& e) { msg = e.msg; }
      { std::lock_guard<std::mutex> lk(qm); inflight[i] = false; if (!msg.empty() && consumer_failure.empty()) consumer_failure = msg; }
      qcv.notify_all();
    }
  });
  auto wait_free = [&](int i) {
    std::unique_lock<std::mutex> lk(qm);
    qcv.wait(lk, [&] { return !inflight[i]; });
    if (!consumer_failure.empty()) throw Fail{consumer_failure};
  };
  auto submit = [&](int i) {
    HostBatch& b = hb[i];
    if (b.n_cells() == 0) return;
    const auto t0 = clk::now();
    std::string perr2;
    parse_batch(b, lay, pool, perr2);
    t_parse += secs(t0, clk::now());
    if (!perr2.empty()) throw Fail{perr2};
    afq_batch ab{};
    ab.first_cell_index = b.first_cell;
    ab.n_cells = b.n_cells();
    ab.n_records = b.n_rec;
    ab.n_refs_total = b.n_ref;
    ab.cell_rec_offsets = b.cell_rec_off.p;
    if (b.pack24) { ab.rec_umi24 = b.umi24.p; ab.refs24 = b.refs24.p; }     // 3 bytes instead of 4 per UMI / id over PCIe
    else { ab.rec_umi32 = b.umi.p; ab.refs = b.refs.p; }
    if (!b.wide_na) { ab.rec_ref_offsets = nullptr; ab.rec_na8 = b.na8.p; }   // 1 byte instead of 4 per record over PCIe
    else { ab.rec_ref_offsets = b.ref_off.p; ab.rec_na8 = nullptr; }
    const auto ta = clk::now();
    dev_of[i] = (int)(batch_seq++ % (uint64_t)D);
    afq_ctx* ctx = ctxs[dev_of[i]];
    if (afq_submit(ctx, &ab, &tickets[i]) != AFQ_OK) throw Fail{std::string("afq_submit: ") + afq_last_error(ctx)};
    t_submit += secs(ta, clk::now());
    { std::lock_guard<std::mutex> lk(qm); inflight[i] = true; queue.push_back(i); }
    qcv.notify_all();
  };
  try {
    int cur = 0;
    auto reset = [&](HostBatch& b, uint64_t first) { b.chunks.clear(); b.first_cell = first; b.n_rec = b.n_ref = 0; };
    reset(hb[cur], 0);
    const unsigned char* p = fmap + body_start;
    const unsigned char* fend = fmap + fsize;
    const size_t rec_fixed = 4 + lay.read_bytes;
    for (uint64_t ch = 0; ch < pre.num_chunks; ++ch) {
      const auto tw = clk::now();
      REQUIRE(p + 8 <= fend, "truncated chunk header in " + rad_path);
      uint32_t nbytes, nrec;
      memcpy(&nbytes, p, 4); memcpy(&nrec, p + 4, 4);
      REQUIRE(nbytes >= 8, "corrupt chunk header");
      REQUIRE(p + nbytes <= fend, "truncated chunk body");
      REQUIRE(nrec > 0, "Discovered empty chunk; should not happen!");
      const uint64_t payload = (uint64_t)nbytes - 8;
      REQUIRE(payload >= (uint64_t)nrec * rec_fixed && lay.aln_bytes > 0 && (payload - (uint64_t)nrec * rec_fixed) % lay.aln_bytes == 0,
              "record overruns its chunk");
      ChunkInfo ci;
      ci.body = p + 8; ci.body_bytes = (uint32_t)payload; ci.nrec = nrec;
      ci.n_aln = (uint32_t)((payload - (uint64_t)nrec * rec_fixed) / lay.aln_bytes);
      ci.bc = 0;
      memcpy(&ci.bc, ci.body + 4 + lay.bc_off, lay.bc_size);   // collate key = barcode of the first record
      p += nbytes;
      t_walk += secs(tw, clk::now());
      if (filtering && !keep.count(ci.bc)) continue;
      HostBatch& b = hb[cur];
      ci.rec_off = b.n_rec; ci.ref_off = b.n_ref;
      b.n_rec += nrec; b.n_ref += ci.n_aln;
      b.chunks.push_back(ci);
      total_records += nrec;
      ++cells_seen;
      if (b.n_rec >= batch_records || b.n_ref >= (3ull << 30)) {
        submit(cur);
        const int nxt = (cur + 1) % NB;
        wait_free(nxt);
        cur = nxt;
        reset(hb[cur], cells_seen);
      }
    }
    submit(cur);
  } catch (const Fail& e) { failure = e.msg; }
  { std::lock_guard<std::mutex> lk(qm); producer_done = true; }
  qcv.notify_all();
  consumer.join();
  if (failure.empty()) failure = consumer_failure;
  if (!failure.empty()) {
    destroy_all();
    if (fmap && mapped) munmap((void*)fmap, fsize);
    fclose(outs.rows); fclose(outs.feat);
    throw Fail{failure};
  }
  const auto t_td0 = clk::now();
  // the GPU contexts, the pinned batches and the input mapping go away on a helper thread while the output files are written
  // (0.13-0.34 s of cudaFree / cudaFreeHost that nothing below depends on)
  std::thread teardown([&] {
    if (o.process_exits) { (void)hb_holder.release(); return; }      // the operating system reclaims everything in a moment
    destroy_all();
    hb.clear();
    if (fmap && mapped) munmap((void*)fmap, fsize);
  });
  struct TeardownJoin { std::thread& t; ~TeardownJoin() { if (t.joinable()) t.join(); } } teardown_join{teardown};
  fclose(outs.rows);
  fclose(outs.feat);

  // quants_mat_cols.txt (src/quant.rs:1786-1809)
  {
    std::string cols;
    for (auto& g : t2g.gene_names) { cols += g; cols.push_back('\n'); }
    if (t2g.usa) {
      for (auto& g : t2g.gene_names) { cols += g; cols += "-U\n"; }
      for (auto& g : t2g.gene_names) { cols += g; cols += "-A\n"; }
    }
    FILE* fc = fopen((out + "/alevin/quants_mat_cols.txt").c_str(), "wb");
    REQUIRE(fc, "couldn't create gene name file.");
    fwrite(cols.data(), 1, cols.size(), fc);
    fclose(fc);
  }
  // quants_mat.mtx — sprs::io::write_matrix_market layout
  {
    FILE* fm = fopen((out + "/alevin/quants_mat.mtx").c_str(), "wb");
    REQUIRE(fm, "couldn't create quants_mat.mtx");
    std::string hdr = "%%MatrixMarket matrix coordinate real general\n% written by sprs\n";
    append_u64(hdr, num_cells); hdr.push_back(' ');
    append_u64(hdr, t2g.num_rows); hdr.push_back(' ');
    append_u64(hdr, outs.nnz); hdr.push_back('\n');
    fwrite(hdr.data(), 1, hdr.size(), fm);
    fflush(fm);
    // the body chunks go out in parallel at their final offsets
    std::vector<uint64_t> offs(outs.mtx_chunks.size() + 1, hdr.size());
    for (size_t i = 0; i < outs.mtx_chunks.size(); ++i) offs[i + 1] = offs[i] + outs.mtx_chunks[i].size();
    const int mfd = fileno(fm);
    bool wr_ok = ftruncate(mfd, (off_t)offs.back()) == 0;
    std::atomic<size_t> nextc{0};
    std::atomic<bool> wr_bad{false};
    if (wr_ok) fmt_pool.run([&](unsigned) {
      for (;;) {
        const size_t i = nextc.fetch_add(16);
        if (i >= outs.mtx_chunks.size()) return;
        const size_t e = std::min(i + 16, outs.mtx_chunks.size());
        for (size_t j = i; j < e; ++j) {
          const TextBuf& c = outs.mtx_chunks[j];
          size_t done = 0;
          while (done < c.size()) {
            const ssize_t w = pwrite(mfd, c.data() + done, c.size() - done, (off_t)(offs[j] + done));
            if (w <= 0) { wr_bad = true; return; }
            done += (size_t)w;
          }
        }
      }
    });
    fclose(fm);
    REQUIRE(wr_ok && !wr_bad, "writing quants_mat.mtx failed");
  }
  // --dump-eqclasses: geqc_counts.mtx (cells x classes) + gene_eqclass.txt.gz (write_eqc_counts, src/quant.rs:231-355)
  if (o.dump_eq) {
    const EqcMap& em = outs.eqc;
    std::string body = "%%MatrixMarket matrix coordinate real general\n% written by sprs\n";
    append_u64(body, em.cell_offset.size()); body.push_back(' ');
    append_u64(body, em.by_id.size()); body.push_back(' ');
    append_u64(body, em.cell_level_count.size()); body.push_back('\n');
    uint64_t goff = 0;
    for (auto& co : em.cell_offset) {
      for (uint64_t k = goff; k < goff + co.second; ++k) {
        append_u64(body, co.first + 1); body.push_back(' ');
        append_u64(body, em.cell_level_count[k].first + 1); body.push_back(' ');
        append_u64(body, em.cell_level_count[k].second); body.push_back('\n');
      }
      goff += co.second;
    }
    FILE* fm = fopen((out + "/alevin/geqc_counts.mtx").c_str(), "wb");
    REQUIRE(fm, "could not write geqc_counts.mtx");
    fwrite(body.data(), 1, body.size(), fm);
    fclose(fm);
    std::string txt;
    append_u64(txt, t2g.num_rows); txt.push_back('\n');
    append_u64(txt, em.by_id.size()); txt.push_back('\n');
    const uint32_t uo = t2g.num_rows / 3, ao = 2 * uo;
    for (uint64_t id = 0; id < em.by_id.size(); ++id) {
      const std::vector<uint32_t>& gl = *em.by_id[id];
      if (t2g.usa) {   // S -> k, U -> G + k, adjacent S,U of one gene -> 2G + k (src/quant.rs:288-335)
        for (size_t i = 0; i < gl.size(); ++i) {
          const uint32_t cg = gl[i];
          if (i + 1 < gl.size() && ((cg | 1u) == (gl[i + 1] | 1u))) { append_u64(txt, (cg >> 1) + ao); txt.push_back('\t'); ++i; continue; }
          append_u64(txt, (cg & 1u) == 0 ? (cg >> 1) : (cg >> 1) + uo); txt.push_back('\t');
        }
      } else {
        for (uint32_t g : gl) { append_u64(txt, g); txt.push_back('\t'); }
      }
      append_u64(txt, id); txt.push_back('\n');
    }
    gzFile gz = gzopen((out + "/alevin/gene_eqclass.txt.gz").c_str(), "wb");
    REQUIRE(gz, "could not write to gene_eqclass.txt.gz");
    size_t done = 0;
    while (done < txt.size()) {
      const int w = gzwrite(gz, txt.data() + done, (unsigned)std::min<size_t>(txt.size() - done, 1u << 30));
      if (w <= 0) { gzclose(gz); throw Fail{"could not write to gene_eqclass.txt.gz"}; }
      done += (size_t)w;
    }
    gzclose(gz);
  }
  // quant.json (src/quant.rs:1913-1933); keys in sorted order (serde_json Map without preserve_order)
  {
    std::string js = "{\n";
    js += "  \"alt_resolved_cell_numbers\": "; json_u64_list(js, outs.alt, "  "); js += ",\n";
    js += "  \"cmd\": \"" + json_escape(o.cmdline ? o.cmdline : "") + "\",\n";
    js += std::string("  \"dump_eq\": ") + (o.dump_eq ? "true" : "false") + ",\n";
    js += "  \"empty_resolved_cell_numbers\": "; json_u64_list(js, outs.empty, "  "); js += ",\n";
    js += "  \"num_genes\": "; append_u64(js, t2g.num_rows); js += ",\n";
    js += "  \"num_quantified_cells\": "; append_u64(js, num_cells); js += ",\n";
    js += "  \"num_tiny_cell_resolved\": "; append_u64(js, outs.tiny.size()); js += ",\n";
    js += "  \"quant_options\": {\n";
    js += "    \"cmdline\": \"" + json_escape(o.cmdline ? o.cmdline : "") + "\",\n";
    js += std::string("    \"dump_eq\": ") + (o.dump_eq ? "true" : "false") + ",\n";
    js += std::string("    \"filter_list\": ") + ((o.filter_list && *o.filter_list) ? "\"" + json_escape(o.filter_list) + "\"" : std::string("null")) + ",\n";
    js += std::string("    \"init_uniform\": ") + (o.init_uniform ? "true" : "false") + ",\n";
    js += "    \"input_dir\": \"" + json_escape(in) + "\",\n";
    js += "    \"large_graph_thresh\": "; append_u64(js, o.large_graph_thresh); js += ",\n";
    js += "    \"num_bootstraps\": "; append_u64(js, o.num_bootstraps); js += ",\n";
    js += "    \"num_threads\": "; append_u64(js, o.num_threads < 2 ? 2 : o.num_threads); js += ",\n";
    js += "    \"output_dir\": \"" + json_escape(out) + "\",\n";
    js += std::string("    \"pug_exact_umi\": ") + (o.pug_exact_umi ? "true" : "false") + ",\n";
    js += std::string("    \"resolution\": \"") + resolution_debug_name(res) + "\",\n";
    js += std::string("    \"sa_model\": \"") + (sa == "prefer-ambig" ? "PreferAmbiguity" : "WinnerTakeAll") + "\",\n";
    js += "    \"small_thresh\": "; append_u64(js, o.small_thresh); js += ",\n";
    js += std::string("    \"summary_stat\": ") + (o.summary_stat ? "true" : "false") + ",\n";
    js += "    \"tg_map\": \"" + json_escape(o.tg_map) + "\",\n";
    js += "    \"version\": \"" + json_escape(o.version ? o.version : "") + "\"\n";
    js += "  },\n";
    js += std::string("  \"resolution_strategy\": \"") + resolution_debug_name(res) + "\",\n";
    js += "  \"tiny_cell_resolved_cell_numbers\": "; json_u64_list(js, outs.tiny, "  "); js += ",\n";
    js += std::string("  \"usa_mode\": ") + (t2g.usa ? "true" : "false") + ",\n";
    js += "  \"version_str\": \"" + json_escape(o.version ? o.version : "") + "\"\n";
    js += "}";
    FILE* fj = fopen((out + "/quant.json").c_str(), "wb");
    REQUIRE(fj, "couldn't create quant.json file.");
    fwrite(js.data(), 1, js.size(), fj);
    fclose(fj);
  }
  const auto t_written = clk::now();
  teardown.join();
  if (getenv("AFQ_TIMING")) {
    fprintf(stderr, "[afq timing] cells %llu records %llu threads %u | setup %.3f s (t2g %.3f, afq_create %.3f) | pipeline %.3f s (chunk index %.3f, parse %.3f, afq_submit %.3f, afq_wait %.3f, text formatting %.3f) | final writes %.3f s | teardown (in the background of the writes) %.3f s beyond them\n",
            (unsigned long long)cells_seen, (unsigned long long)total_records, n_threads, secs(t_entry, t_begin), secs(t_t2g0, t_t2g1), secs(t_create0, t_begin),
            secs(t_begin, t_td0), t_walk, t_parse, t_submit, t_wait, t_format, secs(t_td0, t_written), secs(t_written, clk::now()));
  }
  return 0;
}

// ---- infer (src/infer.rs) ---------------------------------------------------------------------------------------
std::string dirname_of(const std::string& p) {
  const size_t k = p.find_last_of('/');
  return k == std::string::npos ? std::string(".") : (k == 0 ? std::string("/") : p.substr(0, k));
}

int infer_impl(const afqh_infer_opts& o) {
  REQUIRE(o.count_mat && o.eq_labels && o.output_dir, "count_mat, eq_labels and output_dir are required");
  // ---- the count matrix: MatrixMarket coordinate, `real` (what quant writes) or `integer` (src/infer.rs:55-85) ----
  const std::string mtx = slurp(o.count_mat);
  REQUIRE(!mtx.empty(), std::string("error reading mtx format matrix : cannot read ") + o.count_mat);
  size_t pos = 0;
  auto next_line = [&](std::string& line) { if (pos >= mtx.size()) return false; size_t e = mtx.find('\n', pos); if (e == std::string::npos) e = mtx.size(); line.assign(mtx, pos, e - pos); pos = e + 1; if (!line.empty() && line.back() == '\r') line.pop_back(); return true; };
  std::string line;
  REQUIRE(next_line(line) && line.rfind("%%MatrixMarket", 0) == 0, "error reading mtx format matrix : missing MatrixMarket banner");
  {
    const std::string l = lower(line);
    REQUIRE(l.find("coordinate") != std::string::npos && (l.find(" real") != std::string::npos || l.find(" integer") != std::string::npos) &&
            l.find("general") != std::string::npos, "error reading mtx format matrix : expected `matrix coordinate real|integer general`");
  }
  while (next_line(line) && (line.empty() || line[0] == '%')) {}
  uint64_t n_rows = 0, n_cols = 0, nnz = 0;
  REQUIRE(sscanf(line.c_str(), "%lu %lu %lu", &n_rows, &n_cols, &nnz) == 3, "error reading mtx format matrix : bad size line");
  struct Trip { uint64_t r; uint32_t c, v; };
  std::vector<Trip> trips;
  trips.reserve(nnz);
  {
    const char* p = mtx.data() + pos;
    const char* end = mtx.data() + mtx.size();
    for (uint64_t k = 0; k < nnz; ++k) {
      while (p < end && (*p == '\n' || *p == '\r' || *p == ' ')) ++p;
      REQUIRE(p < end, "error reading mtx format matrix : fewer entries than the size line says");
      char* q;
      const uint64_t r = strtoull(p, &q, 10); p = q;
      const uint64_t c = strtoull(p, &q, 10); p = q;
      const double v = strtod(p, &q);
      REQUIRE(q != p && r >= 1 && r <= n_rows && c >= 1 && c <= n_cols, "error reading mtx format matrix : bad entry");
      p = q;
      // whole UMI counts held in a float: round, do not truncate (src/infer.rs:367-376)
      trips.push_back({r - 1, (uint32_t)(c - 1), (uint32_t)std::llround(v < 0 ? 0.0 : v)});
    }
  }
  std::stable_sort(trips.begin(), trips.end(), [](const Trip& a, const Trip& b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });   // to_csr
  // ---- the global eq-class table (IndexedEqList::init_from_eqc_file, src/eq_class.rs:249-298) ----------------------
  uint64_t num_genes = 0, num_eqc = 0;
  std::vector<std::vector<uint32_t>> eqid_map;
  {
    gzFile gz = gzopen(o.eq_labels, "rb");
    REQUIRE(gz, std::string("cannot open ") + o.eq_labels);
    std::string text;
    char buf[1 << 16];
    int got;
    while ((got = gzread(gz, buf, sizeof buf)) > 0) text.append(buf, (size_t)got);
    gzclose(gz);
    size_t tp = 0;
    auto tline = [&](std::string& l) { if (tp >= text.size()) return false; size_t e = text.find('\n', tp); if (e == std::string::npos) e = text.size(); l.assign(text, tp, e - tp); tp = e + 1; return true; };
    std::string l;
    REQUIRE(tline(l), "gene_eqclass file is empty"); num_genes = strtoull(l.c_str(), nullptr, 10);
    REQUIRE(tline(l), "gene_eqclass file is truncated"); num_eqc = strtoull(l.c_str(), nullptr, 10);
    REQUIRE(num_genes > 0 && num_genes < (1ull << 32), "bad gene count in the gene_eqclass file");
    eqid_map.assign(num_eqc, {});
    while (tline(l)) {
      std::vector<uint32_t> v;
      const char* p = l.c_str();
      for (;;) { while (*p == ' ' || *p == '\t') ++p; if (!*p) break; char* q; v.push_back((uint32_t)strtoull(p, &q, 10)); if (q == p) break; p = q; }
      if (v.empty()) continue;
      const uint32_t id = v.back(); v.pop_back();
      REQUIRE(id < num_eqc, "eq-class id out of range in the gene_eqclass file");
      eqid_map[id] = v;
    }
  }
  REQUIRE(n_cols <= num_eqc, "the count matrix has more columns than the gene_eqclass file has classes");
  std::vector<uint32_t> lab_off(num_eqc + 1, 0), labels;
  for (uint64_t i = 0; i < num_eqc; ++i) { labels.insert(labels.end(), eqid_map[i].begin(), eqid_map[i].end()); lab_off[i + 1] = (uint32_t)labels.size(); }
  // ---- barcodes (rows) + optional filter (src/infer.rs:108-150, 340-352) -------------------------------------------------
  const std::string parent = dirname_of(o.count_mat);
  std::vector<std::string> bcs;
  {
    std::ifstream bf(parent + "/quants_mat_rows.txt");
    REQUIRE(bf.good(), "Unable to read first barcode from " + parent + "/quants_mat_rows.txt");
    std::string l;
    while (std::getline(bf, l)) { if (!l.empty() && l.back() == '\r') l.pop_back(); if (!l.empty()) bcs.push_back(l); }
  }
  REQUIRE(bcs.size() >= n_rows, "quants_mat_rows.txt has fewer barcodes than the count matrix has rows");
  std::unordered_set<std::string> keep;
  bool filtering = false;
  if (o.filter_list && *o.filter_list) {
    std::ifstream fl(o.filter_list);
    REQUIRE(fl.good(), std::string("couldn't open file ") + o.filter_list);
    std::string l;
    while (std::getline(fl, l)) { if (!l.empty() && l.back() == '\r') l.pop_back(); if (!l.empty()) keep.insert(l); }
    filtering = true;
  }
  std::vector<uint64_t> cell_off(1, 0);
  std::vector<uint32_t> cell_eq, cell_cnt;
  std::vector<uint64_t> kept_rows;
  {
    size_t t = 0;
    for (uint64_t r = 0; r < n_rows; ++r) {
      const size_t t0 = t;
      while (t < trips.size() && trips[t].r == r) ++t;
      if (filtering && !keep.count(bcs[r])) continue;
      for (size_t k = t0; k < t; ++k) { cell_eq.push_back(trips[k].c); cell_cnt.push_back(trips[k].v); }
      cell_off.push_back(cell_eq.size());
      kept_rows.push_back(r);
    }
  }
  // ---- the EM on the GPU ---------------------------------------------------------------------------------------------------
  afq_config cfg{};
  cfg.resolution = AFQ_RES_CR_LIKE_EM;
  cfg.usa_mode = o.usa_mode ? 1 : 0;
  cfg.em_init_uniform = 0;                       // infer always runs EmInitType::Informative (src/infer.rs:208)
  cfg.num_rows = (uint32_t)num_genes;
  cfg.num_gene_ids = o.usa_mode ? (uint32_t)(2 * (num_genes / 3)) : (uint32_t)num_genes;
  cfg.small_thresh = 0; cfg.large_graph_thresh = 0; cfg.barcode_len = 16; cfg.umi_len = 12; cfg.device = o.device;
  REQUIRE(!o.usa_mode || num_genes % 3 == 0, "--usa needs a gene count that is a multiple of 3");
  const uint32_t dummy_t2g = 0;
  afq_ctx* ctx = nullptr;
  if (afq_create(&cfg, &dummy_t2g, 1, &ctx) != AFQ_OK) throw Fail{std::string("afq_create: ") + afq_last_error(nullptr)};
  afq_eqc_table tab{num_eqc, lab_off.data(), labels.data()};
  afq_result res{};
  if (afq_infer(ctx, &tab, kept_rows.size(), cell_off.data(), cell_eq.data(), cell_cnt.data(), &res) != AFQ_OK) {
    const std::string m = afq_last_error(ctx);
    afq_destroy(ctx);
    throw Fail{"afq_infer: " + m};
  }
  // ---- outputs ---------------------------------------------------------------------------------------------------------------
  const std::string out = o.output_dir;
  mkdirs(out);
  {
    const std::string cols = slurp(parent + "/quants_mat_cols.txt");
    FILE* fc = fopen((out + "/quants_mat_cols.txt").c_str(), "wb");
    if (!fc) { afq_destroy(ctx); throw Fail{"could not copy column (gene) names to output"}; }
    fwrite(cols.data(), 1, cols.size(), fc);
    fclose(fc);
    std::string rows;
    for (uint64_t r : kept_rows) { rows += bcs[r]; rows.push_back('\n'); }
    FILE* fr = fopen((out + "/quants_mat_rows.txt").c_str(), "wb");
    if (!fr) { afq_destroy(ctx); throw Fail{"couldn't create output barcode file"}; }
    fwrite(rows.data(), 1, rows.size(), fr);
    fclose(fr);
    std::string body = "%%MatrixMarket matrix coordinate real general\n% written by sprs\n";
    append_u64(body, res.n_cells); body.push_back(' '); append_u64(body, num_genes); body.push_back(' '); append_u64(body, res.nnz); body.push_back('\n');
    for (uint64_t c = 0; c < res.n_cells; ++c)
      for (uint64_t k = res.row_ptr[c]; k < res.row_ptr[c + 1]; ++k) {
        append_u64(body, c + 1); body.push_back(' '); append_u64(body, (uint64_t)res.col[k] + 1); body.push_back(' '); append_f32(body, res.val[k]); body.push_back('\n');
      }
    FILE* fm = fopen((out + "/quants_mat.mtx").c_str(), "wb");
    if (!fm) { afq_destroy(ctx); throw Fail{"couldn't create quants_mat.mtx"}; }
    fwrite(body.data(), 1, body.size(), fm);
    fclose(fm);
  }
  afq_destroy(ctx);
  return 0;
}

template <class T> void put(FILE* f, T v) { fwrite(&v, sizeof(T), 1, f); }
void put_str16(FILE* f, const std::string& s) { put<uint16_t>(f, (uint16_t)s.size()); fwrite(s.data(), 1, s.size(), f); }
void put_tag(FILE* f, const std::string& name, uint8_t type) { put_str16(f, name); put<uint8_t>(f, type); }

}  // namespace

extern "C" {

int afqh_quantify(const afqh_quant_opts* opts, char* err, size_t errlen) {
  if (!opts) return 1;
  try {
    return quantify_impl(*opts);
  } catch (const Fail& e) {
    if (err && errlen) { strncpy(err, e.msg.c_str(), errlen - 1); err[errlen - 1] = 0; }
    return 1;
  } catch (const std::exception& e) {
    if (err && errlen) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return 1;
  }
}

int afqh_infer(const afqh_infer_opts* opts, char* err, size_t errlen) {
  if (!opts) return 1;
  try {
    return infer_impl(*opts);
  } catch (const Fail& e) {
    if (err && errlen) { strncpy(err, e.msg.c_str(), errlen - 1); err[errlen - 1] = 0; }
    return 1;
  } catch (const std::exception& e) {
    if (err && errlen) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return 1;
  }
}

int afqh_write_collated_rad(const char* dir, uint64_t n_cells, const uint64_t* cell_rec_offsets,
                            const uint64_t* cell_barcodes, const uint32_t* rec_umi32,
                            const uint32_t* rec_ref_offsets, const uint32_t* refs,
                            const char* const* ref_names, uint64_t n_refs, uint16_t bc_len,
                            uint16_t umi_len, char* err, size_t errlen) {
  auto fail = [&](const std::string& m) { if (err && errlen) { strncpy(err, m.c_str(), errlen - 1); err[errlen - 1] = 0; } return 1; };
  if (bc_len == 0 || bc_len > 32 || umi_len == 0 || umi_len > 16) return fail("bc_len must be 1..32 and umi_len 1..16");
  const std::string d = dir;
  mkdirs(d);
  FILE* f = fopen((d + "/map.collated.rad").c_str(), "wb");
  if (!f) return fail("cannot create " + d + "/map.collated.rad");
  auto width_type = [](unsigned len) -> uint8_t { return len <= 4 ? T_U8 : len <= 8 ? T_U16 : len <= 16 ? T_U32 : T_U64; };  // src/convert.rs:323-343
  const uint8_t bt = width_type(bc_len), ut = width_type(umi_len);
  put<uint8_t>(f, 0);
  put<uint64_t>(f, n_refs);
  for (uint64_t i = 0; i < n_refs; ++i) put_str16(f, ref_names[i]);
  put<uint64_t>(f, n_cells);
  put<uint16_t>(f, 2); put_tag(f, "cblen", T_U16); put_tag(f, "ulen", T_U16);
  put<uint16_t>(f, 2); put_tag(f, "b", bt); put_tag(f, "u", ut);
  put<uint16_t>(f, 1); put_tag(f, "compressed_ori_refid", T_U32);
  put<uint16_t>(f, bc_len); put<uint16_t>(f, umi_len);
  std::vector<unsigned char> buf;
  for (uint64_t c = 0; c < n_cells; ++c) {
    buf.clear();
    auto app = [&](const void* p, size_t n) { const unsigned char* q = (const unsigned char*)p; buf.insert(buf.end(), q, q + n); };
    const uint64_t r0 = cell_rec_offsets[c], r1 = cell_rec_offsets[c + 1];
    for (uint64_t r = r0; r < r1; ++r) {
      const uint32_t na = rec_ref_offsets[r + 1] - rec_ref_offsets[r];
      app(&na, 4);
      const uint64_t bc = cell_barcodes[c], um = rec_umi32[r];
      app(&bc, type_size(bt));
      app(&um, type_size(ut));
      for (uint32_t k = 0; k < na; ++k) { const uint32_t v = refs[rec_ref_offsets[r] + k] | 0x80000000u; app(&v, 4); }
    }
    put<uint32_t>(f, (uint32_t)(buf.size() + 8));
    put<uint32_t>(f, (uint32_t)(r1 - r0));
    fwrite(buf.data(), 1, buf.size(), f);
  }
  fclose(f);
  FILE* j = fopen((d + "/collate.json").c_str(), "wb");
  if (!j) return fail("cannot create collate.json");
  fputs("{\n  \"compressed_output\": false\n}\n", j); fclose(j);
  j = fopen((d + "/generate_permit_list.json").c_str(), "wb");
  if (!j) return fail("cannot create generate_permit_list.json");
  fputs("{\n  \"velo_mode\": false,\n  \"max-ambig-record\": 8\n}\n", j); fclose(j);
  return 0;
}

// The two host stages of `quant` WITHOUT a GPU (tests / scripts/host_stage_bench.py): the chunk index walk + parse_batch
// (the product's parser, into ordinary memory) over every batch of the file, and consume() — the text formatting of the
// rows / featureDump / matrix body — on a synthetic result of the same shape (per cell ~1/8 of its records as non-zeros,
// ascending columns, counts 1..7; every `frac_every`-th value a non-integer). Reports wall seconds of both stages and
// FNV checksums of the parsed arrays in file order (the same sums afqh_rad_summary computes with its own loop).
int afqh_host_stage_bench(const char* rad_path, uint32_t n_threads, uint32_t frac_every, afqh_stage_info* out, char* err, size_t errlen) {
  auto fail = [&](const std::string& m) { if (err && errlen) { strncpy(err, m.c_str(), errlen - 1); err[errlen - 1] = 0; } return 1; };
  if (!rad_path || !out) return fail("null argument");
  memset(out, 0, sizeof(*out));
  g_plain_host_memory = true;
  struct Restore { ~Restore() { g_plain_host_memory = false; } } restore;
  try {
    FILE* f = fopen(rad_path, "rb");
    if (!f) return fail(std::string("cannot open ") + rad_path);
    Reader rd(f);
    RadPrelude pre;
    std::string perr;
    if (!parse_prelude(rd, pre, perr)) { fclose(f); return fail("RAD prelude: " + perr); }
    RecordLayout lay;
    if (!make_layout(pre, lay, perr)) { fclose(f); return fail(perr); }
    const uint64_t body_start = rd.pos();
    fclose(f);
    const TagValue* cbl = pre.file_tag("cblen");
    const TagValue* ul = pre.file_tag("ulen");
    const unsigned bc_len = cbl ? (unsigned)cbl->u : 16;
    const unsigned umi_len = ul ? (unsigned)ul->u : (unsigned)(4 * lay.umi_size);
    const int fd = open(rad_path, O_RDONLY);
    struct stat st {};
    if (fd < 0 || fstat(fd, &st) != 0) { if (fd >= 0) close(fd); return fail("cannot open the file"); }
    const uint64_t fsize = (uint64_t)st.st_size;
    const unsigned char* fmap = fsize ? (const unsigned char*)mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (fsize && fmap == MAP_FAILED) return fail("mmap failed");
    struct Unmap { const unsigned char* p; uint64_t n; ~Unmap() { if (p) munmap((void*)p, n); } } unmap{fmap, fsize};
    Prefault prefault;
    if (fmap) prefault.start(fmap + body_start, fsize - body_start);
    const unsigned nt = std::max(2u, std::min(n_threads < 2 ? 2u : n_threads, 128u));
    Pool pool(nt / 2), fmt_pool(nt - nt / 2);
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    HostBatch b;
    b.pack24 = ul && umi_len <= 12 && lay.umi_size <= 4 && pre.ref_names.size() < (1u << 24) && !getenv("AFQ_NO_PACK24");
    out->pack24 = b.pack24 ? 1 : 0;
    Outputs outs;
    outs.rows = fopen("/dev/null", "wb"); outs.feat = fopen("/dev/null", "wb");
    UnmappedCounts unmapped;
    uint64_t hu = 0xCBF29CE484222325ull, hr = hu, hn = hu;
    auto mix = [](uint64_t& h, uint64_t v) { h = (h ^ v) * 0x100000001B3ull; };
    const uint32_t n_cols = (uint32_t)std::max<size_t>(1, pre.ref_names.size());
    std::vector<uint64_t> row_ptr;
    std::vector<uint32_t> col, num_expr, num_over;
    std::vector<float> val, sum_umi, max_umi;
    std::vector<uint8_t> flags;
    auto flush = [&]() {
      if (b.n_cells() == 0) return;
      std::string pf;
      auto t0 = clk::now();
      parse_batch(b, lay, pool, pf);
      out->parse_s += secs(t0, clk::now());
      if (!pf.empty()) throw Fail{pf};
      // (the product parses into pinned buffers it re-uses and a file it has walked: a second pass over the same batch, with
      // input and output pages mapped, is the steady-state cost of the loop itself)
      t0 = clk::now();
      parse_batch(b, lay, pool, pf);
      out->parse_warm_s += secs(t0, clk::now());
      for (uint64_t r = 0; r < b.n_rec; ++r) {
        uint32_t u;
        if (b.pack24) { const uint8_t* d = b.umi24.p + 3 * r; u = d[0] | (d[1] << 8) | (d[2] << 16); } else u = b.umi.p[r];
        mix(hu, u);
        mix(hn, b.wide_na ? b.ref_off.p[r + 1] - b.ref_off.p[r] : b.na8.p[r]);
      }
      for (uint64_t a = 0; a < b.n_ref; ++a) {
        uint32_t id;
        if (b.pack24) { const uint8_t* d = b.refs24.p + 3 * a; id = d[0] | (d[1] << 8) | (d[2] << 16); } else id = b.refs.p[a];
        mix(hr, id);
      }
      out->n_records += b.n_rec; out->n_alignments += b.n_ref; out->n_cells += b.n_cells();
      // the synthetic result of this batch
      const uint64_t nc = b.n_cells();
      row_ptr.assign(nc + 1, 0); num_expr.assign(nc, 0); num_over.assign(nc, 0); sum_umi.assign(nc, 0); max_umi.assign(nc, 0); flags.assign(nc, 0);
      for (uint64_t c = 0; c < nc; ++c) {
        const uint32_t k = std::min<uint32_t>(b.chunks[c].nrec / 8 + 1, n_cols);
        num_expr[c] = k; row_ptr[c + 1] = row_ptr[c] + k;
      }
      col.resize(row_ptr[nc]); val.resize(row_ptr[nc]);
      for (uint64_t c = 0; c < nc; ++c) {
        const uint32_t k = num_expr[c];
        const uint32_t stride = std::max(1u, n_cols / k);
        float s = 0, m = 0;
        for (uint32_t j = 0; j < k; ++j) {
          const uint64_t e = row_ptr[c] + j;
          col[e] = j * stride;
          float v = (float)(1 + (e * 2654435761ull >> 7) % 7);
          if (frac_every && e % frac_every == 0) v += 1.0f / (float)(2 + e % 5);
          val[e] = v; s += v; m = v > m ? v : m;
        }
        sum_umi[c] = s; max_umi[c] = m;
      }
      afq_result r{};
      r.n_cells = nc; r.nnz = row_ptr[nc]; r.row_ptr = row_ptr.data(); r.col = col.data(); r.val = val.data();
      r.sum_umi = sum_umi.data(); r.max_umi = max_umi.data(); r.num_expr = num_expr.data(); r.num_over_mean = num_over.data(); r.flags = flags.data();
      t0 = clk::now();
      consume(b, r, bc_len, unmapped, outs, fmt_pool);
      out->format_s += secs(t0, clk::now());
      for (auto& c : outs.mtx_chunks) {     // FNV-1a over the whole matrix body (outside the timed stages)
        out->mtx_bytes += c.size();
        uint64_t h = out->mtx_sum ? out->mtx_sum : 0xCBF29CE484222325ull;
        for (size_t i = 0; i < c.size(); ++i) h = (h ^ (unsigned char)c[i]) * 0x100000001B3ull;
        out->mtx_sum = h;
      }
      outs.mtx_chunks.clear();
      out->nnz += r.nnz;
    };
    const unsigned char* p = fmap + body_start;
    const unsigned char* fend = fmap + fsize;
    const size_t rec_fixed = 4 + lay.read_bytes;
    const uint64_t batch_records = 16ull << 20;
    for (uint64_t ch = 0; ch < pre.num_chunks; ++ch) {
      const auto tw = clk::now();
      if (p + 8 > fend) throw Fail{"truncated chunk header"};
      uint32_t nbytes, nrec;
      memcpy(&nbytes, p, 4); memcpy(&nrec, p + 4, 4);
      if (nbytes < 8 || p + nbytes > fend || nrec == 0) throw Fail{"corrupt chunk header"};
      const uint64_t payload = (uint64_t)nbytes - 8;
      if (payload < (uint64_t)nrec * rec_fixed || lay.aln_bytes == 0 || (payload - (uint64_t)nrec * rec_fixed) % lay.aln_bytes != 0) throw Fail{"record overruns its chunk"};
      ChunkInfo ci;
      ci.body = p + 8; ci.body_bytes = (uint32_t)payload; ci.nrec = nrec;
      ci.n_aln = (uint32_t)((payload - (uint64_t)nrec * rec_fixed) / lay.aln_bytes);
      ci.bc = 0;
      memcpy(&ci.bc, ci.body + 4 + lay.bc_off, lay.bc_size);
      p += nbytes;
      out->walk_s += secs(tw, clk::now());
      ci.rec_off = b.n_rec; ci.ref_off = b.n_ref;
      b.n_rec += nrec; b.n_ref += ci.n_aln;
      b.chunks.push_back(ci);
      if (b.n_rec >= batch_records || b.n_ref >= (3ull << 30)) { flush(); b.chunks.clear(); b.first_cell = out->n_cells; b.n_rec = b.n_ref = 0; }
    }
    flush();
    fclose(outs.rows); fclose(outs.feat);
    out->sum_umi = hu; out->sum_refs = hr; out->sum_na = hn;
    out->threads = nt;
  } catch (const Fail& e) { return fail(e.msg); }
  return 0;
}

int afqh_rad_summary(const char* rad_path, afqh_rad_info* info, char* err, size_t errlen) {
  auto fail = [&](const std::string& m) { if (err && errlen) { strncpy(err, m.c_str(), errlen - 1); err[errlen - 1] = 0; } return 1; };
  if (!rad_path || !info) return fail("null argument");
  memset(info, 0, sizeof(*info));
  FILE* f = fopen(rad_path, "rb");
  if (!f) return fail(std::string("cannot open ") + rad_path);
  Reader rd(f);
  RadPrelude pre;
  std::string perr;
  if (!parse_prelude(rd, pre, perr)) { fclose(f); return fail("RAD prelude: " + perr); }
  RecordLayout lay;
  if (!make_layout(pre, lay, perr)) { fclose(f); return fail(perr); }
  info->n_refs = pre.ref_names.size(); info->num_chunks = pre.num_chunks;
  const TagValue* cbl = pre.file_tag("cblen");
  const TagValue* ul = pre.file_tag("ulen");
  info->bc_len = cbl ? (uint32_t)cbl->u : 0; info->umi_len = ul ? (uint32_t)ul->u : 0;
  info->read_bytes = (uint32_t)lay.read_bytes; info->aln_bytes = (uint32_t)lay.aln_bytes;
  info->bc_size = (uint32_t)lay.bc_size; info->umi_size = (uint32_t)lay.umi_size;
  info->bc_off = (uint32_t)lay.bc_off; info->umi_off = (uint32_t)lay.umi_off; info->refid_off = (uint32_t)lay.refid_off;
  info->n_file_tags = (uint32_t)pre.file_tags.size(); info->n_read_tags = (uint32_t)pre.read_tags.size();
  info->n_aln_tags = (uint32_t)pre.aln_tags.size();
  uint64_t hb = 0xCBF29CE484222325ull, hu = hb, hr = hb;
  auto mix = [](uint64_t& h, uint64_t v) { h = (h ^ v) * 0x100000001B3ull; };
  std::vector<unsigned char> buf;
  for (uint64_t ch = 0; ch < pre.num_chunks; ++ch) {
    uint32_t nbytes, nrec;
    if (!rd.get(nbytes) || !rd.get(nrec) || nbytes < 8) { fclose(f); return fail("truncated or corrupt chunk header"); }
    buf.resize(nbytes - 8);
    if (!rd.read(buf.data(), buf.size())) { fclose(f); return fail("truncated chunk body"); }
    const unsigned char* p = buf.data();
    const unsigned char* end = p + buf.size();
    for (uint32_t r = 0; r < nrec; ++r) {
      if (p + 4 + lay.read_bytes > end) { fclose(f); return fail("record overruns its chunk"); }
      uint32_t na; memcpy(&na, p, 4); p += 4;
      uint64_t bc = 0, um = 0;
      memcpy(&bc, p + lay.bc_off, lay.bc_size); memcpy(&um, p + lay.umi_off, lay.umi_size);
      p += lay.read_bytes;
      if (p + (size_t)na * lay.aln_bytes > end) { fclose(f); return fail("record overruns its chunk"); }
      mix(hb, bc); mix(hu, um);
      for (uint32_t a = 0; a < na; ++a) { uint32_t id; memcpy(&id, p + lay.refid_off, 4); p += lay.aln_bytes; mix(hr, id & 0x7FFFFFFFu); }
      info->n_alignments += na;
    }
    if (p != end) { fclose(f); return fail("chunk size does not match its records"); }
    info->n_records += nrec;
  }
  fclose(f);
  info->sum_bc = hb; info->sum_umi = hu; info->sum_refs = hr;
  return 0;
}

int afqh_snappy_framed_decompress(const uint8_t* src, size_t n, uint8_t** out, size_t* out_len, uint32_t n_threads,
                                  char* err, size_t errlen) {
  std::vector<unsigned char> buf;
  std::string e;
  if (!out || !out_len || !snappy_framed_decompress(src, n, buf, n_threads ? n_threads : 1, e)) {
    if (err && errlen) { strncpy(err, e.empty() ? "null argument" : e.c_str(), errlen - 1); err[errlen - 1] = 0; }
    return 1;
  }
  *out = (uint8_t*)malloc(buf.size() ? buf.size() : 1);
  if (!*out) return 1;
  memcpy(*out, buf.data(), buf.size());
  *out_len = buf.size();
  return 0;
}

void afqh_free(void* p) { free(p); }

}  // extern "C"
