// snappy_frame.cpp — decoder for the Snappy FRAMING format, the container `alevin-fry collate
// --compress` writes as map.collated.rad.sz and `quant` reads back through snap::read::FrameDecoder
// (reference: src/quant.rs:373-395, src/collate.rs). Written from the published format
// descriptions (framing_format.txt and format_description.txt of google/snappy); no Snappy library
// exists in this image.
//
// A framed stream is a sequence of chunks `type:u8, length:u24le, data`; data chunks carry a masked
// CRC-32C of the UNCOMPRESSED bytes and at most 65536 uncompressed bytes, so they decode
// independently: the chunk headers are walked once (uncompressed sizes are in the block preambles),
// output offsets are prefix sums, and the blocks are decoded in parallel straight into place.
#include <atomic>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "rad.h"

namespace afqh {
namespace {

uint32_t g_crc_table[8][256];
bool g_crc_init = [] {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;   // CRC-32C (Castagnoli), reflected
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xFF];
  return true;
}();

uint32_t crc32c(const unsigned char* p, size_t n) {
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {   // slicing-by-8
    uint32_t lo, hi;
    memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = g_crc_table[7][lo & 0xFF] ^ g_crc_table[6][(lo >> 8) & 0xFF] ^ g_crc_table[5][(lo >> 16) & 0xFF] ^ g_crc_table[4][lo >> 24] ^
        g_crc_table[3][hi & 0xFF] ^ g_crc_table[2][(hi >> 8) & 0xFF] ^ g_crc_table[1][(hi >> 16) & 0xFF] ^ g_crc_table[0][hi >> 24];
    p += 8; n -= 8;
  }
  while (n--) c = (c >> 8) ^ g_crc_table[0][(c ^ *p++) & 0xFF];
  return c ^ 0xFFFFFFFFu;
}
inline uint32_t mask_crc(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }

// varint32 preamble of a raw snappy block; returns bytes consumed (0 on error)
size_t read_varint(const unsigned char* p, size_t n, uint32_t& v) {
  v = 0;
  for (size_t i = 0; i < n && i < 5; ++i) {
    v |= (uint32_t)(p[i] & 0x7F) << (7 * i);
    if (!(p[i] & 0x80)) return i + 1;
  }
  return 0;
}

// one raw snappy block -> exactly `out_len` bytes at `out`
bool decode_block(const unsigned char* p, size_t n, unsigned char* out, size_t out_len) {
  uint32_t ulen;
  const size_t h = read_varint(p, n, ulen);
  if (!h || ulen != out_len) return false;
  size_t ip = h, op = 0;
  while (ip < n) {
    const unsigned tag = p[ip++];
    size_t len, off;
    switch (tag & 3) {
      case 0: {   // literal
        len = (tag >> 2) + 1;
        if (len > 60) {
          const size_t nb = len - 60;
          if (ip + nb > n) return false;
          len = 0;
          for (size_t i = 0; i < nb; ++i) len |= (size_t)p[ip + i] << (8 * i);
          len += 1;
          ip += nb;
        }
        if (ip + len > n || op + len > out_len) return false;
        memcpy(out + op, p + ip, len);
        ip += len; op += len;
        continue;
      }
      case 1:
        if (ip + 1 > n) return false;
        len = ((tag >> 2) & 7) + 4;
        off = ((size_t)(tag >> 5) << 8) | p[ip];
        ip += 1;
        break;
      case 2:
        if (ip + 2 > n) return false;
        len = (tag >> 2) + 1;
        off = (size_t)p[ip] | ((size_t)p[ip + 1] << 8);
        ip += 2;
        break;
      default:
        if (ip + 4 > n) return false;
        len = (tag >> 2) + 1;
        off = (size_t)p[ip] | ((size_t)p[ip + 1] << 8) | ((size_t)p[ip + 2] << 16) | ((size_t)p[ip + 3] << 24);
        ip += 4;
        break;
    }
    if (off == 0 || off > op || op + len > out_len) return false;
    if (off >= len) memcpy(out + op, out + op - off, len);
    else for (size_t i = 0; i < len; ++i) out[op + i] = out[op - off + i];   // overlapping copy = run
    op += len;
  }
  return op == out_len;
}

struct Frame { size_t in_off, in_len, out_off, out_len; uint32_t crc; bool compressed; };

}  // namespace

bool snappy_framed_decompress(const unsigned char* src, size_t n, std::vector<unsigned char>& out, unsigned n_threads,
                              std::string& err) {
  std::vector<Frame> frames;
  size_t ip = 0, total = 0;
  bool seen_id = false;
  while (ip < n) {
    if (ip + 4 > n) { err = "truncated snappy frame header"; return false; }
    const unsigned type = src[ip];
    const size_t len = (size_t)src[ip + 1] | ((size_t)src[ip + 2] << 8) | ((size_t)src[ip + 3] << 16);
    ip += 4;
    if (ip + len > n) { err = "truncated snappy frame"; return false; }
    if (type == 0xFF) {
      if (len != 6 || memcmp(src + ip, "sNaPpY", 6) != 0) { err = "bad snappy stream identifier"; return false; }
      seen_id = true;
    } else if (type == 0x00 || type == 0x01) {
      if (!seen_id) { err = "snappy data chunk before the stream identifier"; return false; }
      if (len < 4) { err = "snappy data chunk without checksum"; return false; }
      Frame f{};
      memcpy(&f.crc, src + ip, 4);
      f.in_off = ip + 4; f.in_len = len - 4; f.out_off = total; f.compressed = type == 0x00;
      if (f.compressed) {
        uint32_t ulen;
        if (!read_varint(src + f.in_off, f.in_len, ulen)) { err = "bad snappy block preamble"; return false; }
        f.out_len = ulen;
      } else {
        f.out_len = f.in_len;
      }
      if (f.out_len > 65536) { err = "snappy frame larger than 65536 bytes"; return false; }
      total += f.out_len;
      frames.push_back(f);
    } else if (type >= 0x02 && type <= 0x7F) {
      err = "reserved unskippable snappy chunk type " + std::to_string(type);
      return false;
    }   // 0x80..0xFE: skippable / padding
    ip += len;
  }
  out.resize(total);
  std::atomic<size_t> next{0};
  std::atomic<int> bad{0};
  auto work = [&] {
    for (;;) {
      const size_t b = next.fetch_add(64);
      if (b >= frames.size() || bad.load()) return;
      const size_t e = b + 64 < frames.size() ? b + 64 : frames.size();
      for (size_t i = b; i < e; ++i) {
        const Frame& f = frames[i];
        unsigned char* dst = out.data() + f.out_off;
        if (f.compressed) { if (!decode_block(src + f.in_off, f.in_len, dst, f.out_len)) { bad = 1; return; } }
        else memcpy(dst, src + f.in_off, f.out_len);
        if (mask_crc(crc32c(dst, f.out_len)) != f.crc) { bad = 2; return; }
      }
    }
  };
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> th;
  for (unsigned t = 1; t < n_threads && t < frames.size() / 64 + 1; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (bad == 1) { err = "corrupt snappy block"; return false; }
  if (bad == 2) { err = "snappy frame checksum mismatch"; return false; }
  return true;
}

}  // namespace afqh
