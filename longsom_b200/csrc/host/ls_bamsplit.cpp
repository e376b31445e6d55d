// Host-side BAM splitter: the native half of the SplitBamCellTypes drop-in.
//
// Reference behaviour restated (workflow/scripts/PreProcessing/SplitBamCellTypes.py:39-192; the
// pysam/htslib calls it makes are AlignmentFile.fetch / .write / pysam.index):
//   * every record with a reference id (fetch() without a region visits the placed records only) is
//     counted in Total_reads; no CB:Z tag -> CB_not_found; barcode (text before the first '-') not in
//     the table -> CB_not_matched;
//   * optional filters, evaluated together and reported as a ';'-joined reason: nM > max_nM
//     (nM_not_found when the tag is absent), NH > max_NH (NH_not_found), mapq < min_MQ (only when min_MQ > 0);
//   * optional end trimming: base qualities of the first trim_start and last trim_end query bases are
//     set to 0; a leading / trailing soft clip of length L extends the trim to L + n_trim, except that
//     20 <= L < 30 is treated as 30 (SplitBamCellTypes.py:140-160);
//   * the record is appended, otherwise byte for byte, to the BAM of its cell type; outputs get the
//     input's header, a BGZF EOF marker and a .bai index.
// Differences by construction: members are deflated in parallel (so the compressed bytes differ from
// htslib's, the records do not) and the whole input is inflated in memory, like ls_bamread.cpp.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "ls_bgzf.h"

using namespace lsbgzf;

namespace {

struct AuxHit {
  bool found = false;
  bool is_int = false, is_float = false, is_text = false;
  int64_t ival = 0;
  double fval = 0;
  const char *text = nullptr;
};

// look one two-letter tag up in the aux block [p, end)
static AuxHit find_aux(const uint8_t *p, const uint8_t *end, char t0, char t1) {
  AuxHit h;
  while (p + 3 <= end) {
    const bool mine = (p[0] == (uint8_t)t0 && p[1] == (uint8_t)t1);
    const uint8_t ty = p[2];
    p += 3;
    switch (ty) {
      case 'A': case 'c': case 'C':
        if (mine) { h.found = h.is_int = true; h.ival = ty == 'c' ? (int8_t)p[0] : p[0]; return h; }
        p += 1; break;
      case 's': case 'S':
        if (mine) { h.found = h.is_int = true; h.ival = ty == 's' ? (int16_t)rd16(p) : rd16(p); return h; }
        p += 2; break;
      case 'i': case 'I':
        if (mine) { h.found = h.is_int = true; h.ival = ty == 'i' ? (int64_t)(int32_t)rd32(p) : (int64_t)rd32(p); return h; }
        p += 4; break;
      case 'f':
        if (mine) { h.found = h.is_float = true; uint32_t u = rd32(p); float f; memcpy(&f, &u, 4); h.fval = f; return h; }
        p += 4; break;
      case 'Z': case 'H': {
        const uint8_t *s = p;
        while (p < end && *p) ++p;
        if (p >= end) return h;
        ++p;
        if (mine) { h.found = h.is_text = true; h.text = reinterpret_cast<const char *>(s); return h; }
        break;
      }
      case 'B': {
        if (p + 5 > end) return h;
        const uint8_t sub = p[0];
        const uint32_t cnt = rd32(p + 1);
        p += 5;
        const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        if (mine) { h.found = true; return h; }
        p += es * (size_t)cnt;
        break;
      }
      default: return h;
    }
  }
  return h;
}

static int reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

struct OutRec {
  int32_t tid, beg, end;
  uint64_t u0, u1;  // uncompressed stream offsets of the record
  bool mapped;
};

static void put32(std::vector<uint8_t> &v, uint32_t x) {
  uint8_t b[4];
  wr32(b, x);
  v.insert(v.end(), b, b + 4);
}
static void put64(std::vector<uint8_t> &v, uint64_t x) {
  put32(v, (uint32_t)x);
  put32(v, (uint32_t)(x >> 32));
}

// .bai of a coordinate-sorted record list (SAM spec 5.2), including the 37450 metadata bin samtools writes
static bool build_bai(const std::vector<OutRec> &recs, int n_ref, const std::vector<uint64_t> &coff, std::vector<uint8_t> &bai,
                      std::string &err) {
  auto voff = [&](uint64_t u) { return (coff[u / kBlockPayload] << 16) | (u % kBlockPayload); };
  bai.clear();
  bai.insert(bai.end(), {'B', 'A', 'I', 1});
  put32(bai, (uint32_t)n_ref);
  size_t i = 0;
  int32_t prev_tid = -1, prev_beg = -1;
  for (const OutRec &r : recs) {
    if (r.tid < prev_tid || (r.tid == prev_tid && r.beg < prev_beg)) {
      err = "records are not sorted by coordinate: cannot index";
      return false;
    }
    prev_tid = r.tid;
    prev_beg = r.beg;
  }
  for (int t = 0; t < n_ref; ++t) {
    std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
    std::vector<uint64_t> lin;
    uint64_t off_beg = 0, off_end = 0, n_map = 0, n_unmap = 0;
    bool any = false;
    int last_bin = -1;
    while (i < recs.size() && recs[i].tid == t) {
      const OutRec &r = recs[i];
      const uint64_t v0 = voff(r.u0), v1 = voff(r.u1);
      const int b = reg2bin(r.beg, r.end);
      auto &ch = bins[(uint32_t)b];
      if (b == last_bin && !ch.empty() && ch.back().second == v0)
        ch.back().second = v1;  // consecutive records of one bin form one chunk
      else
        ch.emplace_back(v0, v1);
      last_bin = b;
      const size_t w0 = (size_t)(r.beg >> 14), w1 = (size_t)((r.end - 1) >> 14);
      if (lin.size() <= w1) lin.resize(w1 + 1, 0);
      for (size_t w = w0; w <= w1; ++w)
        if (lin[w] == 0) lin[w] = v0;
      if (!any) off_beg = v0;
      off_end = v1;
      any = true;
      (r.mapped ? n_map : n_unmap)++;
      ++i;
    }
    for (size_t w = 1; w < lin.size(); ++w)
      if (lin[w] == 0) lin[w] = lin[w - 1];
    put32(bai, (uint32_t)(bins.size() + (any ? 1 : 0)));
    for (auto &kv : bins) {
      put32(bai, kv.first);
      put32(bai, (uint32_t)kv.second.size());
      for (auto &c : kv.second) {
        put64(bai, c.first);
        put64(bai, c.second);
      }
    }
    if (any) {
      put32(bai, 37450u);
      put32(bai, 2u);
      put64(bai, off_beg);
      put64(bai, off_end);
      put64(bai, n_map);
      put64(bai, n_unmap);
    }
    put32(bai, (uint32_t)lin.size());
    for (uint64_t v : lin) put64(bai, v);
  }
  put64(bai, 0);  // records without coordinates: none are written
  return true;
}

static bool write_file(const std::string &path, const std::vector<uint8_t> &a, const uint8_t *tail, size_t ntail, std::string &err) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) {
    err = "cannot create " + path;
    return false;
  }
  bool ok = a.empty() || fwrite(a.data(), 1, a.size(), f) == a.size();
  if (ok && ntail) ok = fwrite(tail, 1, ntail, f) == ntail;
  ok = (fclose(f) == 0) && ok;
  if (!ok) err = "short write to " + path;
  return ok;
}

}  // namespace

// Reason bits of a filtered read (reported by the caller as 'nM;NH;MAPQ'-style keys)
enum { LS_SPLIT_NM = 1, LS_SPLIT_NM_MISSING = 2, LS_SPLIT_NH = 4, LS_SPLIT_NH_MISSING = 8, LS_SPLIT_MAPQ = 16 };

extern "C" {

// counters: [0] Total_reads, [1] Pass_reads, [2] CB_not_found, [3] CB_not_matched, [4 + mask] reads filtered
// for reason set `mask` (1..31); first_seen[mask] = ordinal (1-based, among visited reads) of the first read
// filtered for that set, 0 if none -- the reference's report lists reasons in first-occurrence order.
// max_nm / max_nh < 0 switch the filter off.  Returns 0, or -1 with a message in err.
int ls_bam_split(const char *in_path, int n_types, const char *const *out_paths, const char *bc_blob,
                 const uint32_t *bc_off, const int32_t *bc_type, int64_t n_bc, int min_mapq, int max_nm, int max_nh,
                 int n_trim, int threads, int level, int64_t *counters, int64_t *first_seen, char *err, int errlen) {
  std::string e;
  auto fail = [&](const std::string &m) {
    snprintf(err, (size_t)errlen, "%s", m.c_str());
    return -1;
  };
  for (int k = 0; k < 36; ++k) counters[k] = 0;
  for (int k = 0; k < 32; ++k) first_seen[k] = 0;
  std::vector<uint8_t> raw;
  if (!inflate_file(in_path, threads, raw, e)) return fail(e);
  const uint8_t *p = raw.data(), *end = raw.data() + raw.size();
  if (raw.size() < 12 || memcmp(p, "BAM\1", 4) != 0) return fail("bad BAM magic");
  const uint32_t l_text = rd32(p + 4);
  if ((uint64_t)12 + l_text > raw.size()) return fail("truncated BAM header");
  p += 8 + l_text;
  const uint32_t n_ref = rd32(p);
  p += 4;
  for (uint32_t i = 0; i < n_ref; ++i) {
    if (p + 8 > end) return fail("truncated BAM header");
    p += 8 + rd32(p);
  }
  if (p > end) return fail("truncated BAM header");
  const size_t header_len = (size_t)(p - raw.data());

  std::unordered_map<std::string, int32_t> table;
  table.reserve((size_t)n_bc * 2 + 1);
  for (int64_t i = 0; i < n_bc; ++i)
    table[std::string(bc_blob + bc_off[i], bc_off[i + 1] - bc_off[i])] = bc_type[i];  // later rows win, like dict()

  std::vector<std::vector<uint8_t>> streams((size_t)n_types);
  std::vector<std::vector<OutRec>> recs((size_t)n_types);
  for (auto &s : streams) s.assign(raw.data(), raw.data() + header_len);

  int64_t visited = 0;
  std::string key;
  while (p + 4 <= end) {
    const uint32_t bs = rd32(p);
    const uint8_t *r = p + 4;
    if (bs < 32 || r + bs > end) return fail("truncated BAM record");
    p = r + bs;
    const int32_t tid = (int32_t)rd32(r);
    if (tid < 0) continue;  // fetch() without a region does not visit reads without a reference
    ++visited;
    ++counters[0];
    const int32_t pos = (int32_t)rd32(r + 4);
    const uint32_t l_name = r[8], mapq = r[9];
    const uint32_t n_cig = rd16(r + 12), flag = rd16(r + 14), l_seq = rd32(r + 16);
    const uint8_t *cig = r + 32 + l_name;
    const uint8_t *seq = cig + 4 * (size_t)n_cig;
    const uint8_t *qual = seq + (l_seq + 1) / 2;
    const uint8_t *aux = qual + l_seq;
    if (aux > r + bs) return fail("corrupt BAM record");
    const AuxHit cb = find_aux(aux, r + bs, 'C', 'B');
    if (!cb.found) {
      ++counters[2];
      continue;
    }
    if (!cb.is_text) return fail("CB tag is not a string (the reference fails on barcode.split here)");
    const char *dash = strchr(cb.text, '-');
    key.assign(cb.text, dash ? (size_t)(dash - cb.text) : strlen(cb.text));
    auto it = table.find(key);
    if (it == table.end()) {
      ++counters[3];
      continue;
    }
    int mask = 0;
    if (max_nm >= 0) {
      const AuxHit h = find_aux(aux, r + bs, 'n', 'M');
      if (!h.found)
        mask |= LS_SPLIT_NM_MISSING;
      else if (!(h.is_int || h.is_float))
        return fail("nM tag is not numeric");
      else if ((h.is_int ? (double)h.ival : h.fval) > (double)max_nm)
        mask |= LS_SPLIT_NM;
    }
    if (max_nh >= 0) {
      const AuxHit h = find_aux(aux, r + bs, 'N', 'H');
      if (!h.found)
        mask |= LS_SPLIT_NH_MISSING;
      else if (!(h.is_int || h.is_float))
        return fail("NH tag is not numeric");
      else if ((h.is_int ? (double)h.ival : h.fval) > (double)max_nh)
        mask |= LS_SPLIT_NH;
    }
    if (min_mapq > 0 && (int)mapq < min_mapq) mask |= LS_SPLIT_MAPQ;
    if (mask) {
      ++counters[4 + mask];
      if (!first_seen[mask]) first_seen[mask] = visited;
      continue;
    }
    ++counters[1];
    std::vector<uint8_t> &s = streams[(size_t)it->second];
    const uint64_t u0 = s.size();
    s.insert(s.end(), r - 4, r + bs);
    if (n_trim > 0) {
      uint32_t trim_start = (uint32_t)n_trim, trim_end = (uint32_t)n_trim;
      if (n_cig > 1) {
        const uint32_t c0 = rd32(cig), c1 = rd32(cig + 4 * (size_t)(n_cig - 1));
        if ((c0 & 15u) == 4u) trim_start = (((c0 >> 4) >= 20 && (c0 >> 4) < 30) ? 30u : (c0 >> 4)) + (uint32_t)n_trim;
        if ((c1 & 15u) == 4u) trim_end = (((c1 >> 4) >= 20 && (c1 >> 4) < 30) ? 30u : (c1 >> 4)) + (uint32_t)n_trim;
      }
      if (l_seq == 0 || qual[0] == 0xff) return fail("read without base qualities cannot be trimmed (TypeError in the reference)");
      if (trim_start > l_seq || trim_end > l_seq)
        return fail("IndexError: read shorter than the trimmed ends (the reference fails here as well)");
      uint8_t *q = s.data() + u0 + 4 + (size_t)(qual - r);
      memset(q, 0, trim_start);
      memset(q + l_seq - trim_end, 0, trim_end);
    }
    // reference span for the index: M/D/N/=/X lengths (bam_endpos; 1 for records without one)
    int64_t span = 0;
    for (uint32_t k = 0; k < n_cig; ++k) {
      const uint32_t c = rd32(cig + 4 * (size_t)k), op = c & 15u;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += c >> 4;
    }
    const bool mapped = !(flag & 4u);
    OutRec o;
    o.tid = tid;
    o.beg = pos;
    o.end = (int32_t)(pos + ((mapped && span > 0) ? span : 1));
    o.u0 = u0;
    o.u1 = s.size();
    o.mapped = mapped;
    recs[(size_t)it->second].push_back(o);
  }
  raw.clear();
  raw.shrink_to_fit();

  for (int t = 0; t < n_types; ++t) {
    std::vector<uint8_t> comp, bai;
    std::vector<uint64_t> coff;
    if (!deflate_stream(streams[(size_t)t].data(), streams[(size_t)t].size(), threads, level, comp, coff, e)) return fail(e);
    if (!write_file(out_paths[t], comp, kEofBlock, sizeof kEofBlock, e)) return fail(e);
    if (!build_bai(recs[(size_t)t], (int)n_ref, coff, bai, e)) return fail(e + " (" + out_paths[t] + ")");
    if (!write_file(std::string(out_paths[t]) + ".bai", bai, nullptr, 0, e)) return fail(e);
    std::vector<uint8_t>().swap(streams[(size_t)t]);
  }
  return 0;
}

}  // extern "C"
