// Streaming BAM -> structure-of-arrays decoder (host, no CUDA): the file is consumed a chunk of BGZF members at a
// time, so the memory in flight is bounded by the chunk size whatever the size of the BAM.
//
// Replaces, for the drop-in BaseCellCounter, the per-window `pysam.AlignmentFile.pileup` fetches of
// workflow/scripts/SNVCalling/BaseCellCounter.py:190-191,344-409 (the reference streams windows through a process
// pool; here the reads stream through chunks and the windows are completed on the fly, see pipeline.StreamCounter).
//
//   h = ls_bams_open(path, threads)                    header + contigs
//   n = ls_bams_next(h, target_bytes, &n_cigar, &n_bases)   inflate ~target_bytes (members in parallel), index the
//                                                      complete records; a record cut by the chunk end is carried over
//   ls_bams_fill(h, ...)                               parallel copy of the chunk into CALLER buffers (pinned staging
//                                                      memory of the CUDA library: the decoder writes where the H2D
//                                                      copy reads)
// Barcodes (CB:Z) are interned across chunks: ids are stable for the whole file.
// Every length field is checked against the bytes that are actually there: a truncated or corrupt file is an error
// string, never an out-of-bounds read.
#include <zlib.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

struct Member {
  size_t coff;     // offset in the compressed chunk buffer
  uint32_t csize, usize;
  size_t uoff;     // offset in the inflated buffer (after the carried bytes)
};

struct Stream {
  std::string err;
  FILE *f = nullptr;
  int threads = 1;
  bool header_done = false, eof = false;
  std::vector<std::string> contig_names;
  std::vector<int32_t> contig_lens;
  std::vector<std::string> barcodes;
  std::unordered_map<std::string, int32_t> bmap;
  std::vector<uint8_t> comp;        // compressed bytes read but not yet consumed
  size_t comp_lo = 0;               // first unconsumed byte of comp
  std::vector<uint8_t> raw;         // carried tail of the previous chunk + this chunk's inflated bytes
  std::vector<const uint8_t *> recs;  // records of the current chunk (pointers into raw, past block_size)
  std::vector<uint32_t> cig_off;
  std::vector<uint64_t> base_off;
  std::vector<int32_t> cb;
  size_t tail_lo = 0;               // first byte of raw not covered by a complete record
};

bool inflate_block(z_stream &zs, const uint8_t *src, uint32_t csize, uint8_t *dst, uint32_t usize) {
  if (csize < 26) return false;
  const uint32_t xlen = rd16(src + 10);
  if (12u + xlen + 8u > csize) return false;
  const uint8_t *def = src + 12 + xlen;
  const uint32_t dlen = csize - 12 - xlen - 8;
  if (inflateReset(&zs) != Z_OK) return false;
  zs.next_in = const_cast<Bytef *>(def);
  zs.avail_in = dlen;
  zs.next_out = dst;
  zs.avail_out = usize;
  const int rc = inflate(&zs, Z_FINISH);
  return rc == Z_STREAM_END && zs.total_out == usize;
}

// "CB" aux tag of type Z: pointer to its NUL-terminated text, or nullptr
const char *find_cb(const uint8_t *p, const uint8_t *end) {
  while (p + 3 <= end) {
    const uint8_t t0 = p[0], t1 = p[1], ty = p[2];
    p += 3;
    const bool is_cb = (t0 == 'C' && t1 == 'B');
    switch (ty) {
      case 'A': case 'c': case 'C': p += 1; break;
      case 's': case 'S': p += 2; break;
      case 'i': case 'I': case 'f': p += 4; break;
      case 'Z': case 'H': {
        const uint8_t *s = p;
        while (p < end && *p) ++p;
        if (p >= end) return nullptr;
        ++p;
        if (is_cb && ty == 'Z') return reinterpret_cast<const char *>(s);
        break;
      }
      case 'B': {
        if (p + 5 > end) return nullptr;
        const uint8_t sub = p[0];
        const uint32_t cnt = rd32(p + 1);
        p += 5;
        const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        if ((size_t)(end - p) < es * (size_t)cnt) return nullptr;
        p += es * (size_t)cnt;
        break;
      }
      default: return nullptr;
    }
  }
  return nullptr;
}

// Make sure at least `want` unconsumed compressed bytes are buffered (or the file is exhausted).  Bytes below `pin`
// (the first member of the call in progress, still to be inflated) are kept; returns how far the buffer was shifted.
size_t refill(Stream *s, size_t want, size_t pin) {
  if (s->comp.size() - s->comp_lo >= want || s->eof) return 0;
  size_t shift = 0;
  if (pin) {
    s->comp.erase(s->comp.begin(), s->comp.begin() + (ptrdiff_t)pin);
    s->comp_lo -= pin;
    shift = pin;
  }
  const size_t have = s->comp.size();
  const size_t grab = want > (size_t)(8u << 20) ? want : (size_t)(8u << 20);
  s->comp.resize(have + grab);
  const size_t got = fread(s->comp.data() + have, 1, grab, s->f);
  s->comp.resize(have + got);
  if (got < grab) s->eof = true;
  return shift;
}

// Inflate members until the inflated size of this call reaches target (at least one member); appended to raw.
bool inflate_some(Stream *s, size_t target) {
  std::vector<Member> mem;
  size_t usum = 0;
  const size_t raw0 = s->raw.size();
  size_t pin = s->comp_lo;  // everything from here on is needed until the members below are inflated
  auto need = [&](size_t want) {
    const size_t shift = refill(s, want, pin);
    if (shift) {
      for (auto &m : mem) m.coff -= shift;
      pin = 0;
    }
    return s->comp.size() - s->comp_lo >= want;
  };
  for (;;) {
    if (!need(18)) {
      if (s->comp.size() - s->comp_lo == 0) break;
      s->err = "truncated BGZF member header";
      return false;
    }
    const uint8_t *p = s->comp.data() + s->comp_lo;
    if (p[0] != 0x1f || p[1] != 0x8b || !(p[3] & 4)) {
      s->err = "not a BGZF file (bad gzip member header)";
      return false;
    }
    const uint32_t xlen = rd16(p + 10);
    if (!need(12 + (size_t)xlen)) {
      s->err = "truncated BGZF extra field";
      return false;
    }
    p = s->comp.data() + s->comp_lo;
    uint32_t bsize = 0;
    const uint8_t *x = p + 12, *xe = p + 12 + xlen;
    while (x + 4 <= xe) {
      const uint32_t slen = rd16(x + 2);
      if (x[0] == 'B' && x[1] == 'C' && slen == 2 && x + 6 <= xe) bsize = (uint32_t)rd16(x + 4) + 1;
      x += 4 + slen;
    }
    if (bsize < 12 + xlen + 8) {
      s->err = "corrupt BGZF block (no BC field or bad size)";
      return false;
    }
    if (!need(bsize)) {
      s->err = "truncated BGZF member";
      return false;
    }
    p = s->comp.data() + s->comp_lo;
    Member m;
    m.coff = s->comp_lo;
    m.csize = bsize;
    m.usize = rd32(p + bsize - 4);
    if (m.usize > 65536u) {
      s->err = "corrupt BGZF block (ISIZE > 64 KiB)";
      return false;
    }
    m.uoff = raw0 + usum;
    mem.push_back(m);
    usum += m.usize;
    s->comp_lo += bsize;
    if (usum >= target) break;
  }
  if (mem.empty()) return true;
  s->raw.resize(raw0 + usum);
  std::atomic<size_t> next(0);
  std::atomic<int> bad(0);
  auto worker = [&]() {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) {
      bad = 1;
      return;
    }
    for (;;) {
      const size_t i = next.fetch_add(8);
      if (i >= mem.size()) break;
      for (size_t j = i; j < i + 8 && j < mem.size(); ++j)
        if (mem[j].usize && !inflate_block(zs, s->comp.data() + mem[j].coff, mem[j].csize, s->raw.data() + mem[j].uoff, mem[j].usize))
          bad = 1;
    }
    inflateEnd(&zs);
  };
  std::vector<std::thread> th;
  for (int t = 0; t < s->threads; ++t) th.emplace_back(worker);
  for (auto &t : th) t.join();
  if (bad) {
    s->err = "inflate failed";
    return false;
  }
  return true;
}

// BAM header + reference table at the start of raw; returns false if more bytes are needed
bool parse_header(Stream *s, size_t &consumed, bool &need_more) {
  need_more = false;
  const uint8_t *p = s->raw.data(), *end = p + s->raw.size();
  if (s->raw.size() < 12) {
    need_more = true;
    return true;
  }
  if (memcmp(p, "BAM\1", 4) != 0) {
    s->err = "bad BAM magic";
    return false;
  }
  const uint64_t l_text = rd32(p + 4);
  if ((uint64_t)(end - p) < 12 + l_text) {
    need_more = true;
    return true;
  }
  const uint8_t *q = p + 8 + l_text;
  const uint32_t n_ref = rd32(q);
  q += 4;
  std::vector<std::string> names;
  std::vector<int32_t> lens;
  for (uint32_t i = 0; i < n_ref; ++i) {
    if (end - q < 4) {
      need_more = true;
      return true;
    }
    const uint64_t l_name = rd32(q);
    if ((uint64_t)(end - q) < 8 + l_name) {
      need_more = true;
      return true;
    }
    names.emplace_back(reinterpret_cast<const char *>(q + 4), l_name ? (size_t)l_name - 1 : 0);
    lens.push_back((int32_t)rd32(q + 4 + l_name));
    q += 8 + l_name;
  }
  s->contig_names.swap(names);
  s->contig_lens.swap(lens);
  consumed = (size_t)(q - p);
  return true;
}

}  // namespace

extern "C" {

void *ls_bams_open(const char *path, int threads) {
  Stream *s = new Stream();
  s->threads = threads < 1 ? 1 : threads;
  s->f = fopen(path, "rb");
  if (!s->f) {
    s->err = std::string("cannot open ") + path;
    return s;
  }
  // the header may span several members
  size_t consumed = 0;
  for (;;) {
    const size_t before = s->raw.size();
    if (!inflate_some(s, 1u << 20)) return s;
    bool need_more = false;
    if (!parse_header(s, consumed, need_more)) return s;
    if (!need_more) break;
    if (s->raw.size() == before) {
      s->err = "truncated BAM header";
      return s;
    }
  }
  s->raw.erase(s->raw.begin(), s->raw.begin() + (ptrdiff_t)consumed);
  s->header_done = true;
  return s;
}

const char *ls_bams_error(void *h) {
  Stream *s = (Stream *)h;
  return s->err.empty() ? nullptr : s->err.c_str();
}
void ls_bams_close(void *h) {
  Stream *s = (Stream *)h;
  if (s->f) fclose(s->f);
  delete s;
}
int32_t ls_bams_n_contigs(void *h) { return (int32_t)((Stream *)h)->contig_names.size(); }
const char *ls_bams_contig_name(void *h, int i) { return ((Stream *)h)->contig_names[i].c_str(); }
int32_t ls_bams_contig_len(void *h, int i) { return ((Stream *)h)->contig_lens[i]; }
int32_t ls_bams_n_barcodes(void *h) { return (int32_t)((Stream *)h)->barcodes.size(); }
const char *ls_bams_barcode(void *h, int i) { return ((Stream *)h)->barcodes[i].c_str(); }

// Next chunk: number of complete records (0 = end of file, -1 = error).  The records stay valid until the next call.
int64_t ls_bams_next(void *h, int64_t target_bytes, int64_t *n_cigar, int64_t *n_bases) {
  Stream *s = (Stream *)h;
  if (!s->err.empty() || !s->header_done) return -1;
  // drop what the previous chunk consumed, keep its partial tail
  if (s->tail_lo) {
    s->raw.erase(s->raw.begin(), s->raw.begin() + (ptrdiff_t)s->tail_lo);
    s->tail_lo = 0;
  }
  s->recs.clear();
  const size_t before = s->raw.size();
  if (!inflate_some(s, target_bytes > 65536 ? (size_t)target_bytes : 65536)) return -1;
  const uint8_t *p = s->raw.data(), *end = p + s->raw.size();
  while (end - p >= 4) {
    const uint64_t bs = rd32(p);
    if (bs < 32) {
      s->err = "corrupt BAM record (block_size < 32)";
      return -1;
    }
    if ((uint64_t)(end - p) < 4 + bs) break;  // cut by the chunk end: carried over
    s->recs.push_back(p + 4);
    p += 4 + bs;
  }
  s->tail_lo = (size_t)(p - s->raw.data());
  const int64_t n = (int64_t)s->recs.size();
  if (n == 0) {
    if (s->raw.size() == before && s->raw.size() - s->tail_lo > 0) {
      s->err = "truncated BAM record at end of file";
      return -1;
    }
    if (s->raw.size() != before) return ls_bams_next(h, target_bytes, n_cigar, n_bases);  // one record larger than the chunk
    *n_cigar = 0;
    *n_bases = 0;
    return 0;
  }
  s->cig_off.resize((size_t)n + 1);
  s->base_off.resize((size_t)n + 1);
  s->cb.resize((size_t)n);
  uint64_t co = 0, bo = 0;
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t *r = s->recs[(size_t)i];
    const uint64_t bs = rd32(r - 4);
    const uint64_t l_name = r[8], n_cig = rd16(r + 12), l_seq = rd32(r + 16);
    const uint64_t fixed = 32 + l_name + 4 * n_cig + (l_seq + 1) / 2 + l_seq;
    if (fixed > bs) {
      s->err = "corrupt BAM record (fields exceed block_size)";
      return -1;
    }
    s->cig_off[(size_t)i] = (uint32_t)co;
    s->base_off[(size_t)i] = bo;
    co += n_cig;
    bo += (l_seq + 15u) & ~(uint64_t)15u;
    if (co >= 0xffffffffull) {
      s->err = "more than 2^32 CIGAR operations in one chunk";
      return -1;
    }
    const char *cbs = find_cb(r + fixed, r + bs);
    if (!cbs) {
      s->cb[(size_t)i] = -1;
    } else {
      auto it = s->bmap.find(cbs);
      if (it == s->bmap.end()) {
        const int32_t id = (int32_t)s->barcodes.size();
        s->barcodes.emplace_back(cbs);
        s->bmap.emplace(s->barcodes.back(), id);
        s->cb[(size_t)i] = id;
      } else {
        s->cb[(size_t)i] = it->second;
      }
    }
  }
  s->cig_off[(size_t)n] = (uint32_t)co;
  s->base_off[(size_t)n] = bo;
  *n_cigar = (int64_t)co;
  *n_bases = (int64_t)bo;
  return n;
}

// Copy the current chunk into caller buffers: per-read arrays [n] (cigar_off / base_off: [n + 1]), cigar [n_cigar],
// seq4 [n_bases / 2], qual [n_bases].  Reads start at multiples of 16 bases, the padding is zeroed.
int ls_bams_fill(void *h, int32_t *tid, int32_t *pos, uint16_t *flag, uint8_t *mapq, int32_t *cb, int32_t *lq,
                 uint32_t *cigar_off, uint64_t *base_off, uint32_t *cigar, uint8_t *seq4, uint8_t *qual) {
  Stream *s = (Stream *)h;
  if (!s->err.empty()) return -1;
  const int64_t n = (int64_t)s->recs.size();
  memcpy(cigar_off, s->cig_off.data(), (size_t)(n + 1) * 4);
  memcpy(base_off, s->base_off.data(), (size_t)(n + 1) * 8);
  memcpy(cb, s->cb.data(), (size_t)n * 4);
  std::atomic<int64_t> nx(0);
  auto filler = [&]() {
    for (;;) {
      const int64_t i0 = nx.fetch_add(2048);
      if (i0 >= n) break;
      for (int64_t i = i0; i < i0 + 2048 && i < n; ++i) {
        const uint8_t *r = s->recs[(size_t)i];
        const uint32_t l_name = r[8], n_cig = rd16(r + 12), l_seq = rd32(r + 16);
        tid[i] = (int32_t)rd32(r);
        pos[i] = (int32_t)rd32(r + 4);
        mapq[i] = r[9];
        flag[i] = rd16(r + 14);
        lq[i] = (int32_t)l_seq;
        const uint8_t *cg = r + 32 + l_name;
        memcpy(cigar + s->cig_off[(size_t)i], cg, 4 * (size_t)n_cig);
        const uint8_t *sq = cg + 4 * (size_t)n_cig;
        const uint64_t bo = s->base_off[(size_t)i], pad = s->base_off[(size_t)i + 1] - bo;
        const uint32_t sb = (l_seq + 1) / 2;
        memcpy(seq4 + bo / 2, sq, sb);
        memset(seq4 + bo / 2 + sb, 0, (size_t)(pad / 2 - sb));
        memcpy(qual + bo, sq + sb, l_seq);
        memset(qual + bo + l_seq, 0, (size_t)(pad - l_seq));
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < s->threads; ++t) th.emplace_back(filler);
  for (auto &t : th) t.join();
  return 0;
}

}  // extern "C"
