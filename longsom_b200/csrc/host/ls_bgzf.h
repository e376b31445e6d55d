// BGZF helpers shared by the host-side BAM reader and splitter (SAM/BAM spec v1, section 4.1):
// a BGZF file is a series of gzip members, each with a 'BC' extra sub-field holding its total
// size minus one; the uncompressed payload of a member is at most 64 KiB.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <zlib.h>

#include "ls_inflate.h"

#include <atomic>
#include <string>
#include <thread>
#include <vector>

namespace lsbgzf {

static inline uint32_t rd32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline void wr16(uint8_t *p, uint32_t v) {
  p[0] = (uint8_t)v;
  p[1] = (uint8_t)(v >> 8);
}
static inline void wr32(uint8_t *p, uint32_t v) {
  p[0] = (uint8_t)v;
  p[1] = (uint8_t)(v >> 8);
  p[2] = (uint8_t)(v >> 16);
  p[3] = (uint8_t)(v >> 24);
}

constexpr uint32_t kBlockPayload = 0xff00;  // uncompressed bytes per written member (htslib's BGZF_BLOCK_SIZE)

static const uint8_t kEofBlock[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                      0x02, 0,    0x1b, 0,    0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// Deflate `n` bytes into BGZF members of kBlockPayload bytes each, `threads` members at a time.
// out = the concatenated members (no EOF marker); block_coff[i] = file offset of member i, so the
// virtual offset of uncompressed position u is (block_coff[u / kBlockPayload] << 16) | (u % kBlockPayload).
static inline bool deflate_stream(const uint8_t *data, size_t n, int threads, int level, std::vector<uint8_t> &out,
                                  std::vector<uint64_t> &block_coff, std::string &err) {
  const size_t nb = (n + kBlockPayload - 1) / kBlockPayload;
  std::vector<std::vector<uint8_t>> comp(nb);
  std::atomic<size_t> next(0);
  std::atomic<int> bad(0);
  auto worker = [&]() {
    z_stream zs;
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= nb) break;
      const size_t lo = i * kBlockPayload;
      const uint32_t len = (uint32_t)((n - lo) < kBlockPayload ? (n - lo) : kBlockPayload);
      std::vector<uint8_t> &c = comp[i];
      c.resize(18 + compressBound(len) + 8);
      memset(&zs, 0, sizeof zs);
      if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
        bad = 1;
        break;
      }
      zs.next_in = const_cast<Bytef *>(data + lo);
      zs.avail_in = len;
      zs.next_out = c.data() + 18;
      zs.avail_out = (uInt)(c.size() - 26);
      const int rc = deflate(&zs, Z_FINISH);
      const size_t dlen = zs.total_out;
      deflateEnd(&zs);
      if (rc != Z_STREAM_END || 18 + dlen + 8 > 65536) {  // incompressible payloads still fit: 0xff00 + overhead < 64 KiB
        bad = 1;
        break;
      }
      static const uint8_t head[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0};
      memcpy(c.data(), head, 16);
      wr16(c.data() + 16, (uint32_t)(18 + dlen + 8 - 1));
      wr32(c.data() + 18 + dlen, (uint32_t)crc32(crc32(0L, Z_NULL, 0), data + lo, len));
      wr32(c.data() + 18 + dlen + 4, len);
      c.resize(18 + dlen + 8);
    }
  };
  {
    if (threads < 1) threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
  }
  if (bad) {
    err = "BGZF deflate failed";
    return false;
  }
  block_coff.resize(nb + 1);
  size_t total = 0;
  for (size_t i = 0; i < nb; ++i) {
    block_coff[i] = total;
    total += comp[i].size();
  }
  block_coff[nb] = total;
  out.resize(total);
  for (size_t i = 0; i < nb; ++i) memcpy(out.data() + block_coff[i], comp[i].data(), comp[i].size());
  return true;
}

static inline bool inflate_member(const uint8_t *src, uint32_t csize, uint8_t *dst, uint32_t usize) {
  if (csize < 26) return false;
  const uint32_t xlen = rd16(src + 10);
  if (12u + xlen + 8u > csize) return false;
  const uint8_t *def = src + 12 + xlen;
  const uint32_t dlen = csize - 12 - xlen - 8;
  {
    static thread_local lsinf::Tables tabs;  // own decoder first; zlib judges whatever it does not accept
    if (!getenv("LS_ZLIB_INFLATE") && lsinf::inflate_raw(def, dlen, dst, usize, tabs)) return true;
  }
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, -15) != Z_OK) return false;
  zs.next_in = const_cast<Bytef *>(def);
  zs.avail_in = dlen;
  zs.next_out = dst;
  zs.avail_out = usize;
  const int rc = inflate(&zs, Z_FINISH);
  inflateEnd(&zs);
  return rc == Z_STREAM_END && zs.total_out == usize;
}

// Read a whole BGZF file and inflate it, members in parallel.
static inline bool inflate_file(const char *path, int threads, std::vector<uint8_t> &raw, std::string &err) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    err = std::string("cannot open ") + path;
    return false;
  }
  fseek(f, 0, SEEK_END);
  const uint64_t fsize = (uint64_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> comp(fsize);
  if (fsize && fread(comp.data(), 1, fsize, f) != fsize) {
    fclose(f);
    err = "short read";
    return false;
  }
  fclose(f);
  struct Member {
    uint64_t coff, uoff;
    uint32_t csize, usize;
  };
  std::vector<Member> members;
  uint64_t off = 0, uoff = 0;
  while (off + 18 <= fsize) {
    const uint8_t *p = comp.data() + off;
    if (p[0] != 0x1f || p[1] != 0x8b || !(p[3] & 4)) {
      err = "not a BGZF file (bad gzip member header)";
      return false;
    }
    const uint32_t xlen = rd16(p + 10);
    uint32_t bsize = 0;
    const uint8_t *x = p + 12, *xe = p + 12 + xlen;
    while (x + 4 <= xe) {
      const uint32_t slen = rd16(x + 2);
      if (x[0] == 'B' && x[1] == 'C' && slen == 2) bsize = (uint32_t)rd16(x + 4) + 1;
      x += 4 + slen;
    }
    if (bsize == 0 || off + bsize > fsize) {
      err = "corrupt BGZF block";
      return false;
    }
    Member m;
    m.coff = off;
    m.csize = bsize;
    m.usize = rd32(p + bsize - 4);
    m.uoff = uoff;
    members.push_back(m);
    uoff += m.usize;
    off += bsize;
  }
  raw.resize(uoff);
  if (threads < 1) threads = 1;
  std::atomic<size_t> next(0);
  std::atomic<int> bad(0);
  auto worker = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(16);
      if (i >= members.size()) break;
      for (size_t j = i; j < i + 16 && j < members.size(); ++j) {
        const Member &m = members[j];
        if (m.usize && !inflate_member(comp.data() + m.coff, m.csize, raw.data() + m.uoff, m.usize)) bad = 1;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back(worker);
  for (auto &t : th) t.join();
  if (bad) {
    err = "inflate failed";
    return false;
  }
  return true;
}

}  // namespace lsbgzf
