// Host-side BAM decoder: BGZF inflate (zlib, multi-threaded) + record parse into the
// structure-of-arrays batch of include/longsom_b200.h.
//
// Takes the place of pysam/htslib record decoding on the hot path (reference call sites:
// pysam.AlignmentFile(BAM) at BaseCellCounter.py:190, SingleCellGenotype.py:123).  The
// reference re-opens and re-decodes the BAM once per 50 kb window and touches every read
// through Python objects; here the file is inflated once, block-parallel, and every record
// becomes one row of the SoA arrays that are handed to the GPU.
//
// Format facts used (SAM/BAM spec v1, sections 4.1-4.2): BGZF = concatenated gzip members with
// a BC extra sub-field holding BSIZE; BAM record = block_size, refID, pos, l_read_name, mapq,
// bin, n_cigar_op, flag, l_seq, next_refID, next_pos, tlen, read_name, cigar, seq (4-bit),
// qual, aux.  The CB:Z aux tag is interned (raw text) into dense ids.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include "ls_inflate.h"

#include <algorithm>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Block {
  uint64_t coff;   // offset of the gzip member in the file
  uint32_t csize;  // compressed member size
  uint32_t usize;  // ISIZE
  uint64_t uoff;   // offset in the inflated stream
};

struct Bam {
  std::string err;
  std::vector<std::string> contig_names;
  std::vector<int32_t> contig_lens;
  std::string header_text;
  // SoA
  std::vector<int32_t> tid, pos, cb, lq;
  std::vector<uint16_t> flag;
  std::vector<uint8_t> mapq;
  std::vector<uint32_t> cigar_off;
  // large payload arrays: allocated uninitialised and first touched by the parallel fill (a value-initialised
  // std::vector would zero ~3 GB per million long reads on one thread before the copy even starts)
  std::unique_ptr<uint32_t[]> cigar;
  std::unique_ptr<uint8_t[]> seq4, qual;
  size_t n_cigar = 0, n_qual = 0;
  std::vector<uint64_t> base_off;
  std::vector<std::string> barcodes;
  int64_t n_reads = 0;
};

static inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// zs: a raw-deflate stream the calling thread initialised once (inflateInit2(-15)); reset per member, which
// saves the ~40 KB state allocation that inflateInit2 / inflateEnd would do for every 64 KB member
static bool inflate_block(z_stream &zs, const uint8_t *src, uint32_t csize, uint8_t *dst, uint32_t usize) {
  // gzip member: 10-byte header + XLEN extra, deflate stream, CRC32, ISIZE
  if (csize < 26) return false;
  uint32_t xlen = rd16(src + 10);
  if (12u + xlen + 8u > csize) return false;  // the extra field must leave room for the CRC32 / ISIZE trailer
  const uint8_t *def = src + 12 + xlen;
  uint32_t dlen = csize - 12 - xlen - 8;
  {
    static thread_local lsinf::Tables tabs;  // own decoder first; zlib judges whatever it does not accept
    if (!getenv("LS_ZLIB_INFLATE") && lsinf::inflate_raw(def, dlen, dst, usize, tabs)) return true;
  }
  if (inflateReset(&zs) != Z_OK) return false;
  zs.next_in = const_cast<Bytef *>(def);
  zs.avail_in = dlen;
  zs.next_out = dst;
  zs.avail_out = usize;
  int rc = inflate(&zs, Z_FINISH);
  return rc == Z_STREAM_END && zs.total_out == usize;
}

// find "CB" aux tag of type Z; returns pointer to the NUL-terminated text or nullptr
static const char *find_cb(const uint8_t *p, const uint8_t *end) {
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
        size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        p += es * (size_t)cnt;
        break;
      }
      default: return nullptr;  // unknown type: stop scanning
    }
  }
  return nullptr;
}

}  // namespace

extern "C" {

static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void *ls_bam_read(const char *path, int threads) {
  Bam *b = new Bam();
  const bool timing = getenv("LS_BAM_TIMING") != nullptr;
  double t_prev = now_s();
  auto lap = [&](const char *what) {
    if (!timing) return;
    const double t = now_s();
    fprintf(stderr, "[ls_bam_read] %-28s %.3f s\n", what, t - t_prev);
    t_prev = t;
  };
  FILE *f = fopen(path, "rb");
  if (!f) {
    b->err = std::string("cannot open ") + path;
    return b;
  }
  fseek(f, 0, SEEK_END);
  const uint64_t fsize = (uint64_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  std::unique_ptr<uint8_t[]> comp(new uint8_t[fsize ? fsize : 1]);
  if (fsize && fread(comp.get(), 1, fsize, f) != fsize) {
    fclose(f);
    b->err = "short read";
    return b;
  }
  fclose(f);
  lap("read file");
  // pass 1: BGZF block table
  std::vector<Block> blocks;
  uint64_t off = 0, uoff = 0;
  while (off + 18 <= fsize) {
    const uint8_t *p = comp.get() + off;
    if (p[0] != 0x1f || p[1] != 0x8b || !(p[3] & 4)) {
      b->err = "not a BGZF file (bad gzip member header)";
      return b;
    }
    uint32_t xlen = rd16(p + 10);
    uint32_t bsize = 0;
    const uint8_t *x = p + 12, *xe = p + 12 + xlen;
    while (x + 4 <= xe) {
      uint32_t slen = rd16(x + 2);
      if (x[0] == 'B' && x[1] == 'C' && slen == 2) bsize = (uint32_t)rd16(x + 4) + 1;
      x += 4 + slen;
    }
    if (bsize == 0 || off + bsize > fsize) {
      b->err = "corrupt BGZF block";
      return b;
    }
    Block bl;
    bl.coff = off;
    bl.csize = bsize;
    bl.usize = rd32(p + bsize - 4);
    bl.uoff = uoff;
    blocks.push_back(bl);
    uoff += bl.usize;
    off += bsize;
  }
  lap("member table");
  // pass 2: inflate in parallel
  std::unique_ptr<uint8_t[]> raw(new uint8_t[uoff ? uoff : 1]);
  const size_t raw_size = (size_t)uoff;
  if (threads < 1) threads = 1;
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
      size_t i = next.fetch_add(16);
      if (i >= blocks.size()) break;
      for (size_t j = i; j < i + 16 && j < blocks.size(); ++j) {
        const Block &bl = blocks[j];
        if (bl.usize && !inflate_block(zs, comp.get() + bl.coff, bl.csize, raw.get() + bl.uoff, bl.usize)) bad = 1;
      }
    }
    inflateEnd(&zs);
  };
  {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
  }
  lap("alloc + inflate");
  comp.reset();
  if (bad) {
    b->err = "inflate failed";
    return b;
  }
  // header
  const uint8_t *p = raw.get(), *end = raw.get() + raw_size;
  if (raw_size < 12 || memcmp(p, "BAM\1", 4) != 0) {
    b->err = "bad BAM magic";
    return b;
  }
  // every length field is checked against the bytes that are there: a truncated or corrupt file is an error string
  const uint64_t l_text = rd32(p + 4);
  if (12 + l_text > raw_size) {
    b->err = "truncated BAM header (l_text)";
    return b;
  }
  b->header_text.assign(reinterpret_cast<const char *>(p + 8), l_text);
  p += 8 + l_text;
  uint32_t n_ref = rd32(p);
  p += 4;
  for (uint32_t i = 0; i < n_ref; ++i) {
    if (end - p < 4) {
      b->err = "truncated BAM header (reference table)";
      return b;
    }
    const uint64_t l_name = rd32(p);
    if ((uint64_t)(end - p) < 8 + l_name) {
      b->err = "truncated BAM header (reference name)";
      return b;
    }
    b->contig_names.emplace_back(reinterpret_cast<const char *>(p + 4), l_name ? l_name - 1 : 0);
    b->contig_lens.push_back((int32_t)rd32(p + 4 + l_name));
    p += 8 + l_name;
  }
  // pass 3: record offsets (sequential walk over block_size fields)
  std::vector<const uint8_t *> recs;
  while (p + 4 <= end) {
    const uint64_t bs = rd32(p);
    if (bs < 32 || (uint64_t)(end - p) < 4 + bs) {
      b->err = "truncated or corrupt BAM record";
      return b;
    }
    {
      const uint8_t *r = p + 4;
      const uint64_t need = 32 + (uint64_t)r[8] + 4 * (uint64_t)rd16(r + 12) + ((uint64_t)rd32(r + 16) + 1) / 2 + rd32(r + 16);
      if (need > bs) {
        b->err = "corrupt BAM record (fields exceed block_size)";
        return b;
      }
    }
    recs.push_back(p + 4);
    p += 4 + bs;
  }
  lap("record offsets");
  const int64_t n = (int64_t)recs.size();
  b->n_reads = n;
  b->tid.resize(n);
  b->pos.resize(n);
  b->cb.resize(n);
  b->lq.resize(n);
  b->flag.resize(n);
  b->mapq.resize(n);
  b->cigar_off.resize(n + 1);
  b->base_off.resize(n + 1);
  // pass 4: sizes + barcode interning (sequential: the map is shared)
  std::unordered_map<std::string, int32_t> bmap;
  uint32_t co = 0;
  uint64_t bo = 0;
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t *r = recs[i];
    const uint32_t l_name = r[8];
    const uint32_t n_cig = rd16(r + 12);
    const uint32_t l_seq = rd32(r + 16);
    b->tid[i] = (int32_t)rd32(r);
    b->pos[i] = (int32_t)rd32(r + 4);
    b->mapq[i] = r[9];
    b->flag[i] = rd16(r + 14);
    b->lq[i] = (int32_t)l_seq;
    b->cigar_off[i] = co;
    b->base_off[i] = bo;
    co += n_cig;
    bo += ((uint64_t)l_seq + 15u) & ~(uint64_t)15u;
    const uint8_t *aux = r + 32 + l_name + 4 * n_cig + (l_seq + 1) / 2 + l_seq;
    const uint8_t *rend = r + rd32(r - 4);
    const char *cbs = aux <= rend ? find_cb(aux, rend) : nullptr;
    if (!cbs) {
      b->cb[i] = -1;
    } else {
      auto it = bmap.find(cbs);
      if (it == bmap.end()) {
        int32_t id = (int32_t)b->barcodes.size();
        b->barcodes.emplace_back(cbs);
        bmap.emplace(b->barcodes.back(), id);
        b->cb[i] = id;
      } else {
        b->cb[i] = it->second;
      }
    }
  }
  lap("fixed fields + barcodes");
  b->cigar_off[n] = co;
  b->base_off[n] = bo;
  b->n_cigar = co;
  b->n_qual = (size_t)bo;
  b->cigar.reset(new uint32_t[co ? co : 1]);
  b->seq4.reset(new uint8_t[bo / 2 ? bo / 2 : 1]);
  b->qual.reset(new uint8_t[bo ? bo : 1]);
  lap("alloc seq/qual");
  // pass 5: fill cigar / seq / qual in parallel
  std::atomic<int64_t> nx(0);
  auto filler = [&]() {
    for (;;) {
      int64_t i0 = nx.fetch_add(4096);
      if (i0 >= n) break;
      for (int64_t i = i0; i < i0 + 4096 && i < n; ++i) {
        const uint8_t *r = recs[i];
        const uint32_t l_name = r[8];
        const uint32_t n_cig = rd16(r + 12);
        const uint32_t l_seq = rd32(r + 16);
        const uint8_t *cg = r + 32 + l_name;
        memcpy(b->cigar.get() + b->cigar_off[i], cg, 4 * (size_t)n_cig);
        const uint8_t *sq = cg + 4 * n_cig;
        const uint64_t bo_i = b->base_off[i], pad = b->base_off[i + 1] - bo_i;  // padded to a multiple of 16 bases
        const uint32_t sb = (l_seq + 1) / 2;
        memcpy(b->seq4.get() + bo_i / 2, sq, sb);
        memset(b->seq4.get() + bo_i / 2 + sb, 0, (size_t)(pad / 2 - sb));
        memcpy(b->qual.get() + bo_i, sq + sb, l_seq);
        memset(b->qual.get() + bo_i + l_seq, 0, (size_t)(pad - l_seq));
      }
    }
  };
  {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(filler);
    for (auto &t : th) t.join();
  }
  lap("fill cigar/seq/qual");
  return b;
}

const char *ls_bam_error(void *h) { Bam *b = (Bam *)h; return b->err.empty() ? nullptr : b->err.c_str(); }
void ls_bam_free(void *h) { delete (Bam *)h; }
int64_t ls_bam_n_reads(void *h) { return ((Bam *)h)->n_reads; }
int64_t ls_bam_n_cigar(void *h) { return (int64_t)((Bam *)h)->n_cigar; }
int64_t ls_bam_n_bases(void *h) { return (int64_t)((Bam *)h)->n_qual; }
int32_t ls_bam_n_contigs(void *h) { return (int32_t)((Bam *)h)->contig_names.size(); }
const char *ls_bam_contig_name(void *h, int i) { return ((Bam *)h)->contig_names[i].c_str(); }
int32_t ls_bam_contig_len(void *h, int i) { return ((Bam *)h)->contig_lens[i]; }
int32_t ls_bam_n_barcodes(void *h) { return (int32_t)((Bam *)h)->barcodes.size(); }
const char *ls_bam_barcode(void *h, int i) { return ((Bam *)h)->barcodes[i].c_str(); }
const char *ls_bam_header_text(void *h) { return ((Bam *)h)->header_text.c_str(); }
// array getters: 0 tid 1 pos 2 flag 3 mapq 4 cb 5 cigar_off 6 cigar 7 base_off 8 l_qseq 9 seq4 10 qual
const void *ls_bam_array(void *h, int which) {
  Bam *b = (Bam *)h;
  switch (which) {
    case 0: return b->tid.data();
    case 1: return b->pos.data();
    case 2: return b->flag.data();
    case 3: return b->mapq.data();
    case 4: return b->cb.data();
    case 5: return b->cigar_off.data();
    case 6: return b->cigar.get();
    case 7: return b->base_off.data();
    case 8: return b->lq.data();
    case 9: return b->seq4.get();
    case 10: return b->qual.get();
  }
  return nullptr;
}

}  // extern "C"

// ---- BaseCellCounter TSV writer --------------------------------------------------------------
// Formats the per-site table of ls_pileup_count as the reference prints it
// (BaseCellCounter.py:297-309): chrom, pos+1, REF, "DP|NC|CC|BC|BQ|BCf|BCr", and the seven
// '|'-joined fields with ':'-joined allele vectors (BC = BCf + BCr).  Threads format disjoint row
// ranges into private buffers which are then written in order.
namespace {
inline char *put_u32(char *p, uint32_t v) {
  char tmp[12];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}
inline char *put_vec6(char *p, const uint32_t *v) {
  for (int i = 0; i < 6; ++i) {
    p = put_u32(p, v[i]);
    *p++ = i == 5 ? '|' : ':';
  }
  return p;
}
}  // namespace

extern "C" int ls_write_counter_rows(const char *path, const char *chrom, const int32_t *pos, const uint8_t *ref,
                                     const uint32_t *counts, int64_t n, int threads, int append) {
  // open(2) + pwrite(2): the threads that format a batch of rows also write it, each at its own offset (a single
  // writer is bound by the page-cache allocation of the file, ~0.5 GB/s; the formatting itself does 1.7e7 rows/s)
  const int fd = open(path, O_WRONLY | O_CREAT | (append ? 0 : O_TRUNC), 0666);
  if (fd < 0) return -1;
  off_t file_off = append ? lseek(fd, 0, SEEK_END) : 0;
  if (file_off < 0) {
    close(fd);
    return -1;
  }
  if (threads < 1) threads = 1;
  const size_t clen = strlen(chrom);
  const int64_t CH = 1 << 16;
  const int64_t nch = (n + CH - 1) / CH;
  int rc = 0;
  for (int64_t c0 = 0; c0 < nch; c0 += threads) {
    const int nt = (int)std::min<int64_t>(threads, nch - c0);
    // (plain arrays: a std::string would zero-fill its ~420 reserved bytes per row before the ~115 that get written)
    std::vector<std::unique_ptr<char[]>> bufs((size_t)nt);
    std::vector<size_t> lens((size_t)nt, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
      th.emplace_back([&, t]() {
        const int64_t lo = (c0 + t) * CH, hi = std::min(n, lo + CH);
        bufs[(size_t)t].reset(new char[(size_t)(hi - lo) * (clen + 420) + 16]);
        char *const base = bufs[(size_t)t].get();
        char *p = base;
        for (int64_t i = lo; i < hi; ++i) {
          const uint32_t *r = counts + i * 26;
          memcpy(p, chrom, clen);
          p += clen;
          *p++ = '\t';
          p = put_u32(p, (uint32_t)(pos[i] + 1));
          *p++ = '\t';
          *p++ = (char)ref[i];
          memcpy(p, "\tDP|NC|CC|BC|BQ|BCf|BCr\t", 24);
          p += 24;
          p = put_u32(p, r[0]);
          *p++ = '|';
          p = put_u32(p, r[1]);
          *p++ = '|';
          p = put_vec6(p, r + 2);  // CC
          uint32_t bc[6];
          for (int k = 0; k < 6; ++k) bc[k] = r[8 + k] + r[14 + k];
          p = put_vec6(p, bc);      // BC
          p = put_vec6(p, r + 20);  // BQ
          p = put_vec6(p, r + 8);   // BCf
          p = put_vec6(p, r + 14);  // BCr
          p[-1] = '\n';
        }
        lens[(size_t)t] = (size_t)(p - base);
      });
    }
    for (auto &t : th) t.join();
    th.clear();
    std::vector<off_t> offs((size_t)nt);
    for (int t = 0; t < nt; ++t) {
      offs[(size_t)t] = file_off;
      file_off += (off_t)lens[(size_t)t];
    }
    std::atomic<int> bad(0);
    for (int t = 0; t < nt; ++t) {
      th.emplace_back([&, t]() {
        const char *b = bufs[(size_t)t].get();
        size_t left = lens[(size_t)t];
        off_t o = offs[(size_t)t];
        while (left > 0) {
          const ssize_t w = pwrite(fd, b, left, o);
          if (w <= 0) {
            bad = 1;
            return;
          }
          b += w;
          o += w;
          left -= (size_t)w;
        }
      });
    }
    for (auto &t : th) t.join();
    if (bad) rc = -2;
  }
  if (close(fd) != 0 && rc == 0) rc = -2;
  return rc;
}
