// BaseCellCalling.step1, the per-row host work in native code (no CUDA).
//
// Replaces the two per-row Python passes of longsom_b200/cli/step1.py, which restate
// workflow/scripts/SNVCalling/BaseCellCalling.step1.py:78-467 (the beta-binomial tails in between stay on the GPU:
// K2, ls_betabinom_sf):
//   ls_s1_parse   one pass over a byte range of the merged table: per site and cell type the candidate alternative
//                 alleles, and every (k, n) query of the range gathered into four arrays;
//   ls_s1_format  the label cascade on the rounded p-values and the output lines, byte for byte what the reference
//                 prints (allele order, '1' vs '1.0' of the noise test, Q5-Q8 of SURVEY.md).
// The parser is strict: any row the reference would stumble over (missing column, non-numeric field, a candidate allele
// with reads but no cells, ...) makes ls_s1_parse fail, and the caller hands the whole range to the Python restatement,
// whose exceptions are the reference's.  The strand-bias column (--fisher_cutoff != 1) stays in Python as well.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

const char ALLELE[8] = {'A', 'C', 'T', 'G', 'I', 'D', 'N', 'O'};  // step1.py:20
const int SORTED4[4] = {0, 1, 3, 2};                               // A, C, G, T: sorted(alt_bc) of the reference

struct TypeCall {
  int32_t ct;            // index into the caller's cell-type list
  int32_t DP, NC;
  uint8_t has_bc, has_cc;  // bit x: allele x (A, C, T, G) is in alt_bc / alt_cc
  int32_t bc[4], cc[4];
  int32_t q_bc[4], q_cc[4];  // query index of allele x
};

struct Row {
  int64_t lo, hi;        // the line in data, without its terminator
  int32_t tab4;          // offset (relative to lo) of the 4th tab, where the INFO block is inserted; -1: no such tab
  bool verbatim;
  int32_t chrom, n_qual;
  int64_t pos;
  int64_t sum_alts_bc, sum_alts_cc, sum_dp, sum_nc;
  int64_t rest1, rest2;  // query indices of the noise test, -1: none
  int32_t call0, ncall;
};

struct S1 {
  const char *data;
  int64_t len;
  std::vector<Row> rows;
  std::vector<TypeCall> calls;
  std::vector<int32_t> q1k, q1n, q2k, q2n;
  std::vector<std::string> chroms;
  std::unordered_map<std::string, int32_t> chrom_id;
  std::string out;
  int64_t n_data = 0;
};

bool parse_uint(const char *p, const char *e, int64_t &v) {
  if (p >= e || e - p > 18) return false;
  int64_t x = 0;
  for (; p < e; ++p) {
    if (*p < '0' || *p > '9') return false;
    x = x * 10 + (*p - '0');
  }
  v = x;
  return true;
}

// "a:b:c..." -> up to 8 non-negative ints
bool parse_colon(const char *p, const char *e, int64_t *v, int &n) {
  n = 0;
  for (;;) {
    const char *q = p;
    while (q < e && *q != ':') ++q;
    if (n >= 8 || !parse_uint(p, q, v[n])) return false;
    ++n;
    if (q >= e) return true;
    p = q + 1;
  }
}

void fail(char *err, int errlen, const char *msg, int64_t row) {
  if (err && errlen > 0) snprintf(err, (size_t)errlen, "row %lld: %s", (long long)row, msg);
}

void put_int(std::string &o, int64_t v) {
  char b[24];
  int n = 0;
  uint64_t u = v < 0 ? (uint64_t)(-(v + 1)) + 1u : (uint64_t)v;
  do {
    b[n++] = (char)('0' + u % 10);
    u /= 10;
  } while (u);
  if (v < 0) b[n++] = '-';
  while (n) o.push_back(b[--n]);
}

// str(round(v, 4)) of Python / str() of a numpy float64 that is already rounded: the decimal number with four places
// nearest to v, trailing zeros stripped down to one decimal.  Fast path for 0 <= v < 1e6 when v * 1e4 is not within 1e-6 of a
// tie (then the rounding is unambiguous in double arithmetic); everything else goes through the correctly rounded "%.4f".
void put_round4(std::string &o, double v) {
  if (std::isnan(v)) {  // the noise test of a site whose rest depth went negative: scipy answers nan, printed as is
    o += "nan";
    return;
  }
  if (v >= 0.0 && v < 1e6) {
    const double sc = v * 1e4;
    const double fl = std::floor(sc);
    const double frac = sc - fl;
    if (std::fabs(frac - 0.5) > 1e-6) {
      const uint64_t m = (uint64_t)fl + (frac > 0.5 ? 1u : 0u);
      put_int(o, (int64_t)(m / 10000u));
      o.push_back('.');
      const unsigned f = (unsigned)(m % 10000u);
      char d[4] = {(char)('0' + f / 1000), (char)('0' + f / 100 % 10), (char)('0' + f / 10 % 10), (char)('0' + f % 10)};
      int n = 4;
      while (n > 1 && d[n - 1] == '0') --n;
      o.append(d, (size_t)n);
      return;
    }
  }
  char b[48];
  const int n = snprintf(b, sizeof b, "%.4f", v);
  int m = n;
  while (m > 0 && b[m - 1] == '0') --m;
  if (m > 0 && b[m - 1] == '.') ++m;
  o.append(b, (size_t)m);
}

// longest run of equal characters in a + b (step1.py:478-483)
int longest_run2(const char *a, size_t na, const char *b, size_t nb) {
  int best = 0, cur = 0;
  char prev = 0;
  bool first = true;
  for (size_t i = 0; i < na + nb; ++i) {
    const char ch = i < na ? a[i] : b[i - na];
    cur = (!first && ch == prev) ? cur + 1 : 1;
    first = false;
    prev = ch;
    if (cur > best) best = cur;
  }
  return best;
}

}  // namespace

extern "C" {

void *ls_s1_parse(const char *data, int64_t len, const int32_t *ct_cols, int32_t n_ct, int32_t min_reads, int32_t min_cells,
                  char *err, int32_t errlen) {
  S1 *s = new S1();
  s->data = data;
  s->len = len;
  int32_t max_col = 3;
  for (int i = 0; i < n_ct; ++i)
    if (ct_cols[i] > max_col) max_col = ct_cols[i];
  std::vector<int64_t> tab;  // start offsets of the columns of the current line
  int32_t last_chrom = -1;
  int64_t p = 0;
  while (p < len) {
    // one line, universal newlines: "\r\n", "\n" and a lone "\r" all end it
    int64_t q = p;
    while (q < len && data[q] != '\n' && data[q] != '\r') ++q;
    int64_t next = q;
    if (q < len) next = (data[q] == '\r' && q + 1 < len && data[q + 1] == '\n') ? q + 2 : q + 1;
    Row r;
    memset(&r, 0, sizeof r);
    r.lo = p;
    r.hi = q;
    r.rest1 = r.rest2 = -1;
    r.tab4 = -1;
    const int64_t row_no = (int64_t)s->rows.size();
    if (q - p >= 2 && data[p] == '#' && data[p + 1] == '#') {
      r.verbatim = true;
      s->rows.push_back(r);
      p = next;
      continue;
    }
    tab.clear();
    tab.push_back(p);
    for (int64_t i = p; i < q; ++i)
      if (data[i] == '\t') tab.push_back(i + 1);
    const int ncol = (int)tab.size();
    auto col_end = [&](int c) { return c + 1 < ncol ? tab[(size_t)c + 1] - 1 : q; };
    if (ncol <= max_col) {
      fail(err, errlen, "fewer columns than the header names", row_no);
      delete s;
      return nullptr;
    }
    if (ncol > 4) r.tab4 = (int32_t)(tab[4] - 1 - p);
    if (!parse_uint(data + tab[1], data + col_end(1), r.pos)) {
      fail(err, errlen, "POS is not a plain integer", row_no);
      delete s;
      return nullptr;
    }
    const size_t clen0 = (size_t)(col_end(0) - tab[0]);
    if (!s->rows.empty() && last_chrom >= 0 && s->chroms[(size_t)last_chrom].size() == clen0 &&
        memcmp(s->chroms[(size_t)last_chrom].data(), data + tab[0], clen0) == 0) {
      r.chrom = last_chrom;  // tables are sorted: almost every row repeats the contig of the row before
    } else {
      std::string c(data + tab[0], clen0);
      auto it = s->chrom_id.find(c);
      if (it == s->chrom_id.end()) {
        r.chrom = (int32_t)s->chroms.size();
        s->chrom_id.emplace(c, r.chrom);
        s->chroms.push_back(c);
      } else {
        r.chrom = it->second;
      }
      last_chrom = r.chrom;
    }
    int ref_idx = -1;  // REF equals one of the allele letters
    if (col_end(3) - tab[3] == 1)
      for (int x = 0; x < 8; ++x)
        if (data[tab[3]] == ALLELE[x]) ref_idx = x;
    r.call0 = (int32_t)s->calls.size();
    for (int i = 0; i < n_ct; ++i) {
      const char *f = data + tab[(size_t)ct_cols[i]], *fe = data + col_end(ct_cols[i]);
      if (fe - f >= 2 && f[0] == 'N' && f[1] == 'A') continue;
      // DP|NC|CC|BC|BQ|BCf|BCr
      const char *part[8];
      int np = 0;
      part[np++] = f;
      for (const char *c = f; c < fe; ++c)
        if (*c == '|') {
          if (np >= 8) {
            np = 99;
            break;
          }
          part[np++] = c + 1;
        }
      if (np != 7) {
        fail(err, errlen, "a cell-type column does not have seven '|' fields", row_no);
        delete s;
        return nullptr;
      }
      part[7] = fe + 1;
      int64_t DP, NC;
      if (!parse_uint(part[0], part[1] - 1, DP) || !parse_uint(part[1], part[2] - 1, NC) || DP > 0x7fffffff || NC > 0x7fffffff) {
        fail(err, errlen, "DP / NC is not a plain integer", row_no);
        delete s;
        return nullptr;
      }
      if (!(DP >= min_reads && NC >= min_cells)) continue;
      ++r.n_qual;
      int64_t cc[8], bc[8];
      int ncc, nbc;
      if (!parse_colon(part[2], part[3] - 1, cc, ncc) || !parse_colon(part[3], part[4] - 1, bc, nbc)) {
        fail(err, errlen, "CC / BC is not a ':' list of at most eight integers", row_no);
        delete s;
        return nullptr;
      }
      for (int x = 0; x < nbc; ++x)
        if (x != ref_idx && x != 7) r.sum_alts_bc += bc[x];
      for (int x = 0; x < ncc; ++x)
        if (x != ref_idx && x != 7) r.sum_alts_cc += cc[x];
      r.sum_dp += DP;
      r.sum_nc += NC;
      TypeCall t;
      memset(&t, 0, sizeof t);
      t.ct = i;
      t.DP = (int32_t)DP;
      t.NC = (int32_t)NC;
      for (int x = 0; x < 4 && x < nbc; ++x)
        if (x != ref_idx && bc[x] > 0) {
          if (bc[x] > 0x7fffffff) {
            fail(err, errlen, "count above 2^31", row_no);
            delete s;
            return nullptr;
          }
          t.has_bc |= (uint8_t)(1u << x);
          t.bc[x] = (int32_t)bc[x];
        }
      for (int x = 0; x < 4 && x < ncc; ++x)
        if (x != ref_idx && cc[x] > 0) {
          if (cc[x] > 0x7fffffff) {
            fail(err, errlen, "count above 2^31", row_no);
            delete s;
            return nullptr;
          }
          t.has_cc |= (uint8_t)(1u << x);
          t.cc[x] = (int32_t)cc[x];
        }
      if (!t.has_bc) continue;
      if (t.has_bc & ~t.has_cc) {  // the reference dies with a KeyError here
        fail(err, errlen, "an allele has reads but no cells", row_no);
        delete s;
        return nullptr;
      }
      // queries in dict order (A, C, T, G)
      int64_t b0 = 0, c0 = 0;
      for (int x = 0; x < 4; ++x)
        if (t.has_bc & (1u << x)) {
          t.q_bc[x] = (int32_t)s->q1k.size();
          s->q1k.push_back(t.bc[x]);
          s->q1n.push_back(t.DP);
          b0 += t.bc[x];
        }
      for (int x = 0; x < 4; ++x)
        if (t.has_cc & (1u << x)) {
          t.q_cc[x] = (int32_t)s->q2k.size();
          s->q2k.push_back(t.cc[x]);
          s->q2n.push_back(t.NC);
          if (t.has_bc & (1u << x)) c0 += t.cc[x];
        }
      r.sum_dp -= b0;
      r.sum_nc -= c0;
      r.sum_alts_bc -= b0;
      r.sum_alts_cc -= c0;
      s->calls.push_back(t);
    }
    r.ncall = (int32_t)s->calls.size() - r.call0;
    if (r.sum_alts_bc > 0) {  // noise test
      if (r.sum_alts_bc > 0x7fffffff || r.sum_dp > 0x7fffffff || r.sum_dp < -0x7fffffff || r.sum_alts_cc > 0x7fffffff ||
          r.sum_alts_cc < -0x7fffffff || r.sum_nc > 0x7fffffff || r.sum_nc < -0x7fffffff) {
        fail(err, errlen, "count above 2^31", row_no);
        delete s;
        return nullptr;
      }
      r.rest1 = (int64_t)s->q1k.size();
      r.rest2 = (int64_t)s->q2k.size();
      s->q1k.push_back((int32_t)r.sum_alts_bc);
      s->q1n.push_back((int32_t)r.sum_dp);
      s->q2k.push_back((int32_t)r.sum_alts_cc);
      s->q2n.push_back((int32_t)r.sum_nc);
    }
    ++s->n_data;
    s->rows.push_back(r);
    p = next;
  }
  return s;
}

void ls_s1_free(void *h) { delete (S1 *)h; }
int64_t ls_s1_n_rows(void *h) { return (int64_t)((S1 *)h)->rows.size(); }
int64_t ls_s1_n_data_rows(void *h) { return ((S1 *)h)->n_data; }
int64_t ls_s1_n_q1(void *h) { return (int64_t)((S1 *)h)->q1k.size(); }
int64_t ls_s1_n_q2(void *h) { return (int64_t)((S1 *)h)->q2k.size(); }
void ls_s1_queries(void *h, int32_t *q1k, int32_t *q1n, int32_t *q2k, int32_t *q2n) {
  S1 *s = (S1 *)h;
  if (!s->q1k.empty()) {
    memcpy(q1k, s->q1k.data(), s->q1k.size() * 4);
    memcpy(q1n, s->q1n.data(), s->q1n.size() * 4);
  }
  if (!s->q2k.empty()) {
    memcpy(q2k, s->q2k.data(), s->q2k.size() * 4);
    memcpy(q2n, s->q2n.data(), s->q2n.size() * 4);
  }
}
int32_t ls_s1_n_chroms(void *h) { return (int32_t)((S1 *)h)->chroms.size(); }
const char *ls_s1_chrom(void *h, int32_t i) { return ((S1 *)h)->chroms[(size_t)i].c_str(); }
// per row (verbatim rows: chrom -1): contig id and POS, for the caller's reference-context lookup
void ls_s1_sites(void *h, int32_t *chrom, int64_t *pos) {
  S1 *s = (S1 *)h;
  for (size_t i = 0; i < s->rows.size(); ++i) {
    chrom[i] = s->rows[i].verbatim ? -1 : s->rows[i].chrom;
    pos[i] = s->rows[i].pos;
  }
}

// ctx: [n_rows][11] reference bases around the site (fetch(CHROM, POS - 6, POS + 5), upper case), ctx_len[row] = how many
// of them exist (-1: no context, printed as '.').  r1 / r2: the ROUNDED tails of the range's queries.
// Returns the length of the text (owned by the handle, valid until ls_s1_free), -1 on error.
int64_t ls_s1_format(void *h, const double *r1, const double *r2, const uint8_t *ctx, const int8_t *ctx_len,
                     const char *const *ct_names, int32_t min_ac_cells, int32_t min_ac_reads, int32_t min_cell_types,
                     int32_t max_cell_types, const char **text) {
  S1 *s = (S1 *)h;
  std::string &o = s->out;
  o.clear();
  o.reserve((size_t)s->len + s->rows.size() * 224);
  char alts[64][8];          // "A|C|G|T" at most; one per call of the row
  uint8_t alts_len[64];
  const char *filt[64];
  std::string up, down, FILTER;
  for (size_t ri = 0; ri < s->rows.size(); ++ri) {
    const Row &r = s->rows[ri];
    const char *line = s->data + r.lo;
    const int64_t ll = r.hi - r.lo;
    if (r.verbatim) {
      o.append(line, (size_t)ll);
      if (r.hi < s->len) o.push_back('\n');  // (a last line without a line end is copied as it is)
      continue;
    }
    // columns 0-3, then the twenty INFO fields, then the rest of the line
    if (r.tab4 >= 0)
      o.append(line, (size_t)r.tab4);
    else
      o.append(line, (size_t)ll);
    o.push_back('\t');
    const int cl = ctx_len ? ctx_len[ri] : -1;
    if (cl < 0) {
      up = ".";
      down = ".";
    } else {
      const char *c = reinterpret_cast<const char *>(ctx + ri * 11);
      up.assign(c, (size_t)(cl < 5 ? cl : 5));
      down.assign(cl > 6 ? c + 6 : c, (size_t)(cl > 6 ? cl - 6 : 0));
    }
    const bool has_rest = r.rest1 >= 0;
    const double bc_noise = has_rest ? r1[r.rest1] : 1.0, cc_noise = has_rest ? r2[r.rest2] : 1.0;
    auto put_rest = [&](int64_t a, int64_t b, double p) {
      put_int(o, a);
      o.push_back(';');
      put_int(o, b);
      o.push_back(';');
      if (has_rest)
        put_round4(o, p);
      else
        o.push_back('1');  // a plain int in the reference
    };
    if (r.ncall > 64) return -2;  // more cell types than the fixed per-row tables hold: the Python passes take the table
    if (r.ncall > 0) {
      const TypeCall *tc = s->calls.data() + r.call0;
      int n_pass = 0, n_nonsig = 0;
      bool any_multi = false;
      for (int k = 0; k < r.ncall; ++k) {
        const TypeCall &t = tc[k];
        int ncand = 0, al = 0;
        double mb = INFINITY, mc = INFINITY;
        int64_t bsum_single = 0, csum_single = 0;
        for (int j = 0; j < 4; ++j) {
          const int x = SORTED4[j];
          if (!(t.has_bc & (1u << x))) continue;
          if (ncand) alts[k][al++] = '|';
          alts[k][al++] = ALLELE[x];
          ++ncand;
          bsum_single = t.bc[x];
          csum_single = t.cc[x];
        }
        alts_len[k] = (uint8_t)al;
        for (int x = 0; x < 4; ++x) {
          if (t.has_bc & (1u << x)) {
            if (std::isnan(r1[t.q_bc[x]])) return -1;  // min() over a NaN is order dependent in the reference
            mb = r1[t.q_bc[x]] < mb ? r1[t.q_bc[x]] : mb;
          }
          if (t.has_cc & (1u << x)) {
            if (std::isnan(r2[t.q_cc[x]])) return -1;
            mc = r2[t.q_cc[x]] < mc ? r2[t.q_cc[x]] : mc;
          }
        }
        const char *lab = nullptr;
        if (mb >= 0.05 || mc >= 0.05) {
          lab = "Non-Significant";
          ++n_nonsig;
        } else if ((0.001 < mb && mb < 0.05) || (0.001 < mc && mc < 0.05)) {
          lab = "Low-Significance";
        } else if (ncand > 1) {
          lab = "Multi-allelic";
          any_multi = true;
        } else if (csum_single < min_ac_cells) {
          lab = "Low_cells";
        } else if (bsum_single < min_ac_reads) {
          lab = "Low_reads";
        } else {
          lab = "PASS";
          ++n_pass;
        }
        filt[k] = lab;
      }
      const int nfilt = r.ncall;
      int len_alts = 0;
      for (int i = 0; i < r.ncall; ++i) {
        bool seen = false;
        for (int j = 0; j < i; ++j) seen = seen || (alts_len[j] == alts_len[i] && memcmp(alts[j], alts[i], alts_len[i]) == 0);
        if (!seen) ++len_alts;
      }
      // site-level FILTER
      FILTER.clear();
      auto add = [&](const char *x) {
        if (!FILTER.empty()) FILTER.push_back(',');
        FILTER += x;
      };
      if (n_pass > max_cell_types) add("Multiple_cell_types");
      if (len_alts > 1 || any_multi) add("Multi-allelic");
      if (r.n_qual < min_cell_types) add("Min_cell_types");
      if (nfilt - n_pass - n_nonsig > 0) add("Cell_type_noise");
      if (bc_noise < 0.05 || cc_noise < 0.05) add("Noisy_site");
      auto homopolymer = [&](const std::string &context, bool upstream) {
        if (context == ".") return false;
        int m = 0;
        for (int i = 0; i < r.ncall; ++i) {
          const int v = upstream ? longest_run2(context.data(), context.size(), alts[i], alts_len[i])
                                 : longest_run2(alts[i], alts_len[i], context.data(), context.size());
          if (v > m) m = v;
        }
        return m >= 4;
      };
      if (homopolymer(up, true)) add("LC_Upstream");
      if (homopolymer(down, false)) add("LC_Downstream");
      if (FILTER.empty()) {
        if (n_pass > 0)
          FILTER = "PASS";
        else
          for (int i = 0; i < nfilt; ++i) {
            if (i) FILTER.push_back(',');
            FILTER += filt[i];
          }
      }
      auto join_calls = [&](auto &&f) {
        for (int k = 0; k < r.ncall; ++k) {
          if (k) o.push_back(',');
          f(tc[k]);
        }
      };
      auto per_cand = [&](const TypeCall &t, auto &&f) {
        bool first = true;
        for (int j = 0; j < 4; ++j) {
          const int x = SORTED4[j];
          if (!(t.has_bc & (1u << x))) continue;
          if (!first) o.push_back('|');
          first = false;
          f(x);
        }
      };
      for (int i = 0; i < r.ncall; ++i) {  // ALT
        if (i) o.push_back(',');
        o.append(alts[i], alts_len[i]);
      }
      o.push_back('\t');
      o += FILTER;
      o.push_back('\t');
      join_calls([&](const TypeCall &t) { o += ct_names[t.ct]; });
      o.push_back('\t');
      o += up;
      o.push_back('\t');
      o += down;
      o.push_back('\t');
      put_int(o, len_alts);
      o.push_back('\t');
      join_calls([&](const TypeCall &t) { put_int(o, t.DP); });
      o.push_back('\t');
      join_calls([&](const TypeCall &t) { put_int(o, t.NC); });
      o.push_back('\t');
      join_calls([&](const TypeCall &t) { per_cand(t, [&](int x) { put_int(o, t.bc[x]); }); });
      o.push_back('\t');
      join_calls([&](const TypeCall &t) { per_cand(t, [&](int x) { put_int(o, t.cc[x]); }); });
      o.push_back('\t');
      join_calls([&](const TypeCall &t) { per_cand(t, [&](int x) { put_round4(o, (double)t.bc[x] / (double)t.DP); }); });
      o.push_back('\t');
      join_calls([&](const TypeCall &t) { per_cand(t, [&](int x) { put_round4(o, (double)t.cc[x] / (double)t.NC); }); });
      o.push_back('\t');
      join_calls([&](const TypeCall &t) { per_cand(t, [&](int x) { put_round4(o, r1[t.q_bc[x]]); }); });
      o.push_back('\t');
      join_calls([&](const TypeCall &t) { per_cand(t, [&](int x) { put_round4(o, r2[t.q_cc[x]]); }); });
      o.push_back('\t');
      put_int(o, r.n_qual);
      o.push_back('\t');
      put_int(o, r.n_qual);
      o.push_back('\t');
      put_rest(r.sum_alts_bc, r.sum_dp, bc_noise);
      o.push_back('\t');
      put_rest(r.sum_alts_cc, r.sum_nc, cc_noise);
      o += "\t.\t";  // Fisher_p: '.' when --fisher_cutoff is 1
      for (int i = 0; i < nfilt; ++i) {
        if (i) o.push_back(',');
        o += filt[i];
      }
    } else {
      const bool noisy = bc_noise < 0.001 || cc_noise < 0.001;
      o += ".\t";
      o += noisy ? "Noisy_site" : ".";
      o += "\t.\t";
      o += up;
      o.push_back('\t');
      o += down;
      o += "\t.\t.\t.\t.\t.\t.\t.\t.\t.\t";
      put_int(o, r.n_qual);
      o.push_back('\t');
      put_int(o, r.n_qual);
      o.push_back('\t');
      put_rest(r.sum_alts_bc, r.sum_dp, bc_noise);
      o.push_back('\t');
      put_rest(r.sum_alts_cc, r.sum_nc, cc_noise);
      o += "\t.\t.";
    }
    if (r.tab4 >= 0) o.append(line + r.tab4, (size_t)(ll - r.tab4));  // "\t" + the columns from the fifth on
    o.push_back('\n');
  }
  *text = o.data();
  return (int64_t)o.size();
}

}  // extern "C"
