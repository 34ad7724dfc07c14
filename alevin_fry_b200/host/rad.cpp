#include "rad.h"

#include <cstring>

namespace afqh {

bool Reader::read(void* dst, size_t n) {
  if (n == 0) return true;
  size_t got = fread(dst, 1, n, f_);
  pos_ += got;
  return got == n;
}
bool Reader::skip(size_t n) {
  char buf[4096];
  while (n) {
    size_t k = n < sizeof(buf) ? n : sizeof(buf);
    if (!read(buf, k)) return false;
    n -= k;
  }
  return true;
}

size_t type_size(uint8_t t) {
  switch (t) {
    case T_BOOL: case T_U8: return 1;
    case T_U16: return 2;
    case T_U32: case T_F32: return 4;
    case T_U64: case T_F64: return 8;
    default: return 0;
  }
}

static bool read_string16(Reader& r, std::string& s) {
  uint16_t n;
  if (!r.get(n)) return false;
  s.resize(n);
  return r.read(s.data(), n);
}

static bool parse_tag_section(Reader& r, std::vector<TagDesc>& out, std::string& err) {
  uint16_t n;
  if (!r.get(n)) { err = "truncated tag section"; return false; }
  out.resize(n);
  for (auto& t : out) {
    if (!read_string16(r, t.name) || !r.get(t.type)) { err = "truncated tag description"; return false; }
    if (t.type == T_ARRAY) {
      if (!r.get(t.arr_len_type) || !r.get(t.arr_elem_type)) { err = "truncated array tag description"; return false; }
    } else if (t.type > T_STRING) {
      err = "unknown RAD tag type id " + std::to_string(t.type) + " for tag '" + t.name + "'";
      return false;
    }
  }
  return true;
}

static bool read_uint(Reader& r, uint8_t type, uint64_t& v) {
  v = 0;
  size_t n = type_size(type);
  if (n == 0 || type == T_F32 || type == T_F64) return false;
  return r.read(&v, n);
}

bool parse_prelude(Reader& r, RadPrelude& p, std::string& err) {
  if (!r.get(p.is_paired) || !r.get(p.ref_count)) { err = "truncated RAD header"; return false; }
  if (p.ref_count > (1ull << 31)) { err = "implausible ref_count in RAD header"; return false; }
  p.ref_names.resize(p.ref_count);
  for (auto& n : p.ref_names)
    if (!read_string16(r, n)) { err = "truncated reference names"; return false; }
  if (!r.get(p.num_chunks)) { err = "truncated RAD header (num_chunks)"; return false; }
  if (!parse_tag_section(r, p.file_tags, err) || !parse_tag_section(r, p.read_tags, err) ||
      !parse_tag_section(r, p.aln_tags, err))
    return false;
  // file-level tag values follow the three sections (src/convert.rs:364-369)
  for (auto& t : p.file_tags) {
    TagValue v;
    v.name = t.name;
    v.type = t.type;
    if (t.type == T_STRING) {
      if (!read_string16(r, v.s)) { err = "truncated file tag string"; return false; }
    } else if (t.type == T_F32) {
      float f; if (!r.get(f)) { err = "truncated file tag"; return false; } v.f = f;
    } else if (t.type == T_F64) {
      if (!r.get(v.f)) { err = "truncated file tag"; return false; }
    } else if (t.type == T_ARRAY) {
      uint64_t len;
      if (!read_uint(r, t.arr_len_type, len)) { err = "bad array length type in file tag"; return false; }
      size_t es = type_size(t.arr_elem_type);
      if (es == 0) { err = "unsupported array element type in file tag"; return false; }
      if (!r.skip(len * es)) { err = "truncated file tag array"; return false; }
    } else {
      if (!read_uint(r, t.type, v.u)) { err = "truncated file tag"; return false; }
    }
    p.file_tag_values.push_back(v);
  }
  return true;
}

bool make_layout(const RadPrelude& p, RecordLayout& l, std::string& err) {
  l = RecordLayout();
  bool have_b = false, have_u = false, have_ref = false;
  size_t off = 0;
  for (auto& t : p.read_tags) {
    size_t s = type_size(t.type);
    if (s == 0) { err = "variable-size read-level tag '" + t.name + "' is not supported"; return false; }
    if (t.name == "b") { l.bc_off = off; l.bc_size = s; have_b = true; }
    if (t.name == "u") { l.umi_off = off; l.umi_size = s; have_u = true; }
    off += s;
  }
  l.read_bytes = off;
  off = 0;
  for (auto& t : p.aln_tags) {
    size_t s = type_size(t.type);
    if (s == 0) { err = "variable-size alignment-level tag '" + t.name + "' is not supported"; return false; }
    if (t.name == "compressed_ori_refid") { l.refid_off = off; have_ref = (s == 4); }
    off += s;
  }
  l.aln_bytes = off;
  if (!have_b || !have_u) { err = "RAD read-level tags must contain 'b' and 'u'"; return false; }
  if (!have_ref) { err = "RAD alignment-level tags must contain compressed_ori_refid:u32"; return false; }
  return true;
}

}  // namespace afqh
