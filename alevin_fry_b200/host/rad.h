// rad.h — collated RAD reader/writer (wire format: SURVEY.md §8(b); evidence in the reference:
// src/convert.rs:92-144, 254, 280-383, 472-491, 585-590, 621-626; tests/multi_barcode_integration.rs:57-115).
// libradicl's source is not available; the tag-section layout is reconstructed from those call
// sites and the public RAD specification.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace afqh {

enum RadType : uint8_t { T_BOOL = 0, T_U8 = 1, T_U16 = 2, T_U32 = 3, T_U64 = 4, T_F32 = 5, T_F64 = 6, T_ARRAY = 7, T_STRING = 8 };

struct TagDesc {
  std::string name;
  uint8_t type = 0;
  uint8_t arr_len_type = 0, arr_elem_type = 0;  // for T_ARRAY
};

struct TagValue {
  std::string name;
  uint8_t type = 0;
  uint64_t u = 0;
  double f = 0;
  std::string s;
};

struct RadPrelude {
  uint8_t is_paired = 0;
  uint64_t ref_count = 0;
  std::vector<std::string> ref_names;
  uint64_t num_chunks = 0;
  std::vector<TagDesc> file_tags, read_tags, aln_tags;
  std::vector<TagValue> file_tag_values;
  const TagValue* file_tag(const std::string& n) const {
    for (auto& v : file_tag_values) if (v.name == n) return &v;
    return nullptr;
  }
  static bool has(const std::vector<TagDesc>& v, const std::string& n) {
    for (auto& t : v) if (t.name == n) return true;
    return false;
  }
};

// buffered little-endian file reader
class Reader {
 public:
  explicit Reader(FILE* f) : f_(f) {}
  bool read(void* dst, size_t n);
  template <class T> bool get(T& v) { return read(&v, sizeof(T)); }
  bool skip(size_t n);
  uint64_t pos() const { return pos_; }
 private:
  FILE* f_;
  uint64_t pos_ = 0;
};

size_t type_size(uint8_t t);  // 0 for variable-size types
bool parse_prelude(Reader& r, RadPrelude& p, std::string& err);

// Per-record layout derived from the read / alignment tag sections of a single-barcode
// short-read RNA file (AlevinFryReadRecord[WithPosition]).
struct RecordLayout {
  size_t read_bytes = 0;       // bytes of read-level tags per record
  size_t bc_off = 0, bc_size = 0, umi_off = 0, umi_size = 0;
  size_t aln_bytes = 0;        // bytes per alignment
  size_t refid_off = 0;        // offset of compressed_ori_refid inside an alignment
};
bool make_layout(const RadPrelude& p, RecordLayout& l, std::string& err);

// Snappy framing format (map.collated.rad.sz, `collate --compress`; reference reads it with
// snap::read::FrameDecoder, src/quant.rs:373-395) -> the plain RAD image. Frames decode in parallel.
bool snappy_framed_decompress(const unsigned char* src, size_t n, std::vector<unsigned char>& out, unsigned n_threads,
                              std::string& err);

}  // namespace afqh
