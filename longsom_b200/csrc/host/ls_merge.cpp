// MergeBaseCellCounts: positional merge of the per-cell-type BaseCellCounter tables, native (no CUDA).
//
// Replaces the cursor loop of workflow/scripts/SNVCalling/MergeBaseCellCounts.py:116-204 (restated in Python in
// longsom_b200/cli/merge.py): N file cursors advanced in lock-step over (chrom, pos)-sorted tables, one row per table
// in memory, chromosomes in lexicographic order, 'NA' where a cell type lacks the site, rows whose position does not
// increase inside their chromosome skipped, a blank line ends its table.  The parser is strict: a row that is not
// "chrom <tab> integer <tab> ref <tab> info <tab> counts" makes the call return 1 and the caller runs the Python
// restatement, whose exceptions are the reference's.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

namespace {

struct Cursor {
  FILE *f = nullptr;
  std::string chrom = "x", ref = "0", info, bc = "0";
  bool has_info = false;  // the reference keeps the integer 0 until a row has been taken
  int64_t pos = 0;
  bool done = false;
  std::vector<char> buf, iobuf;
  std::string line;
};

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// next line, str.strip()-ped; false at end of file
bool read_stripped(Cursor &c, std::string &out) {
  out.clear();
  bool any = false;
  for (;;) {
    if (!fgets(c.buf.data(), (int)c.buf.size(), c.f)) break;
    any = true;
    const size_t n = strlen(c.buf.data());
    out.append(c.buf.data(), n);
    if (n && c.buf[n - 1] == '\n') break;
  }
  if (!any) return false;
  size_t a = 0, b = out.size();
  while (a < b && is_space(out[a])) ++a;
  while (b > a && is_space(out[b - 1])) --b;
  out.resize(b);
  if (a) out.erase(0, a);
  return true;  // (a lone '\r' inside a line ends a line in the reference's text mode: take() refuses such rows)
}

// 0 taken or skipped, 1 malformed
int take(Cursor &c, const std::string &line) {
  if (line.find('\r') != std::string::npos) return 1;
  size_t t[4];
  size_t from = 0;
  for (int i = 0; i < 4; ++i) {
    t[i] = line.find('\t', from);
    if (t[i] == std::string::npos) return 1;
    from = t[i] + 1;
  }
  if (line.find('\t', from) != std::string::npos) return 1;  // more than five fields
  if (t[0] == 0) return 1;                                    // an empty chromosome name is the end marker of the loop
  // plain (optionally negative) integer
  size_t p = t[0] + 1;
  bool neg = false;
  if (p < t[1] && line[p] == '-') {
    neg = true;
    ++p;
  }
  if (p >= t[1] || t[1] - p > 18) return 1;
  int64_t v = 0;
  for (; p < t[1]; ++p) {
    if (line[p] < '0' || line[p] > '9') return 1;
    v = v * 10 + (line[p] - '0');
  }
  if (neg) v = -v;
  if (line.compare(0, t[0], c.chrom) == 0 && c.chrom.size() == t[0] && c.pos >= v) return 0;  // not increasing: skipped
  c.chrom.assign(line, 0, t[0]);
  c.pos = v;
  c.ref.assign(line, t[1] + 1, t[2] - t[1] - 1);
  c.info.assign(line, t[2] + 1, t[3] - t[2] - 1);
  c.has_info = true;
  c.bc.assign(line, t[3] + 1, std::string::npos);
  return 0;
}

// sort_set of the reference (:48-57): drop 'NA', drop '.' when something else is present, order by frequency
// (ties keep first-seen order), join with '|'
void most_common_join(std::vector<const std::string *> &vals, std::string &out) {
  static const std::string NA = "NA", DOT = ".";
  {  // the usual row: every table that has the site agrees -> that value (or nothing when all are 'NA')
    const std::string *first = nullptr;
    bool same = true;
    for (auto *s : vals) {
      if (*s == NA) continue;
      if (!first)
        first = s;
      else if (*s != *first)
        same = false;
    }
    if (same) {
      out.clear();
      if (first) out = *first;
      return;
    }
  }
  std::vector<const std::string *> v;
  for (auto *s : vals)
    if (*s != NA) v.push_back(s);
  for (;;) {  // while len > 1 and '.' in vals: remove the first '.'
    if (v.size() <= 1) break;
    size_t i = 0;
    while (i < v.size() && *v[i] != DOT) ++i;
    if (i == v.size()) break;
    v.erase(v.begin() + (ptrdiff_t)i);
  }
  std::vector<const std::string *> keys;
  std::vector<int> cnt;
  for (auto *s : v) {
    size_t i = 0;
    while (i < keys.size() && *keys[i] != *s) ++i;
    if (i == keys.size()) {
      keys.push_back(s);
      cnt.push_back(1);
    } else {
      ++cnt[i];
    }
  }
  // stable sort by count, descending
  std::vector<size_t> ord(keys.size());
  for (size_t i = 0; i < ord.size(); ++i) ord[i] = i;
  for (size_t i = 1; i < ord.size(); ++i) {
    const size_t x = ord[i];
    size_t j = i;
    while (j > 0 && cnt[ord[j - 1]] < cnt[x]) {
      ord[j] = ord[j - 1];
      --j;
    }
    ord[j] = x;
  }
  out.clear();
  for (size_t i = 0; i < ord.size(); ++i) {
    if (i) out.push_back('|');
    out += *keys[ord[i]];
  }
}

}  // namespace

extern "C" {

// header: everything that goes in front of the rows (date line, ##INFO lines, column names).  n_header_lines: lines to
// skip at the top of every input (9).  Returns 0, 1 = a row the strict parser refuses (run the Python restatement), -1 =
// I/O error (err).
int ls_merge_tables(int32_t n, const char *const *paths, const char *out_path, const char *header, int32_t n_header_lines,
                    char *err, int32_t errlen) {
  std::vector<Cursor> cur((size_t)n);
  auto fail = [&](int rc, const char *msg) {
    if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", msg);
    for (auto &c : cur)
      if (c.f) fclose(c.f);
    return rc;
  };
  for (int i = 0; i < n; ++i) {
    Cursor &c = cur[(size_t)i];
    c.buf.resize(1 << 16);
    c.f = fopen(paths[i], "rb");
    if (!c.f) return fail(-1, "cannot open an input table");
    c.iobuf.resize(1 << 20);
    setvbuf(c.f, c.iobuf.data(), _IOFBF, c.iobuf.size());
    std::string l;
    for (int k = 0; k < n_header_lines; ++k) {  // readline() x 9: whole lines, whatever they hold
      for (;;) {
        if (!fgets(c.buf.data(), (int)c.buf.size(), c.f)) break;
        const size_t m = strlen(c.buf.data());
        if (m && c.buf[m - 1] == '\n') break;
      }
    }
    if (!read_stripped(c, l) || l.empty()) {
      c.done = true;
      c.chrom.clear();
      c.pos = -1;
    } else if (take(c, l)) {
      return fail(1, "malformed row");
    }
  }
  FILE *out = fopen(out_path, "wb");
  if (!out) return fail(-1, "cannot open the output table");
  std::vector<char> obuf(1 << 20);
  setvbuf(out, obuf.data(), _IOFBF, obuf.size());
  fputs(header, out);
  static const std::string NA = "NA";
  std::string cur_chr, refs_joined, fmt_joined, row, l;
  bool have_chr = false;
  int64_t cur_pos = 0;
  std::vector<const std::string *> tmp;
  auto all_done = [&]() {
    for (auto &c : cur)
      if (!c.done) return false;
    return true;
  };
  while (!all_done()) {
    for (auto &c : cur) {
      while (have_chr && c.chrom == cur_chr && c.pos <= cur_pos) {
        if (!read_stripped(c, l) || l.empty()) {  // end of file or a blank line: the table ends here
          c.done = true;
          c.chrom.clear();
          c.pos = -1;
          break;
        }
        if (take(c, l)) {
          fclose(out);
          return fail(1, "malformed row");
        }
      }
    }
    const bool go = !all_done();
    bool on_chr = false;
    if (have_chr)
      for (auto &c : cur) on_chr = on_chr || c.chrom == cur_chr;
    if (on_chr) {
      if (!go) continue;
      bool any = false;
      int64_t low = 0;
      for (auto &c : cur)
        if (c.chrom == cur_chr && c.pos > cur_pos && (!any || c.pos < low)) {
          low = c.pos;
          any = true;
        }
      if (!any) {  // min() of an empty sequence in the reference
        fclose(out);
        return fail(1, "no position to advance to");
      }
      cur_pos = low;
      tmp.clear();
      for (auto &c : cur) tmp.push_back((c.chrom == cur_chr && c.pos == cur_pos) ? &c.ref : &NA);
      most_common_join(tmp, refs_joined);
      tmp.clear();
      for (auto &c : cur) tmp.push_back(c.has_info ? &c.info : &NA);
      most_common_join(tmp, fmt_joined);
      row.clear();
      row += cur_chr;
      row.push_back('\t');
      char nb[24];
      const int nn = snprintf(nb, sizeof nb, "%lld", (long long)cur_pos);
      row.append(nb, (size_t)nn);
      row.push_back('\t');
      row.append(nb, (size_t)nn);
      row.push_back('\t');
      row += refs_joined;
      row.push_back('\t');
      row += fmt_joined;
      for (auto &c : cur) {
        row.push_back('\t');
        row += (c.chrom == cur_chr && c.pos == cur_pos) ? c.bc : NA;
      }
      row.push_back('\n');
      fwrite(row.data(), 1, row.size(), out);
    } else if (go) {
      // the smallest chromosome name any cursor points at
      bool any = false;
      for (auto &c : cur)
        if (!c.chrom.empty() && (!any || c.chrom < cur_chr)) {
          cur_chr = c.chrom;
          any = true;
        }
      have_chr = any;
      cur_pos = 0;
    }
  }
  const bool bad = ferror(out) != 0;
  if (fclose(out) != 0 || bad) return fail(-1, "write error");
  for (auto &c : cur)
    if (c.f) {
      fclose(c.f);
      c.f = nullptr;
    }
  return 0;
}

}  // extern "C"
