// SingleCellGenotype / HCCVSingleCellGenotype: the dense long table, written natively (no CUDA).
//
// The reference prints one row per (candidate site, barcode of the metadata) -- SingleCellGenotype.py:128-214,
// HCCVSingleCellGenotype.py:126-212 -- i.e. n_sites x n_cells rows, almost all of them "NoCoverage".  The GPU returns
// only the touched (site, cell) pairs (ls_genotype_sparse_*); this expands them to the reference's rows: per site a
// fixed prefix (built by the caller), per cell "barcode <tab> cell type", and the six or eight value columns with the
// reference's text (str(round(ALT / DP, 4)), str(numpy.float64(p)), the MutationStatus labels).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <string>

namespace {

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

// str(round(v, 4)) / str(numpy.float64 rounded to 4 places), see ls_step1.cpp
double round4(double v, bool &exact_path) {
  const double sc = v * 1e4, fl = std::floor(sc), frac = sc - fl;
  exact_path = v >= 0.0 && v < 1e6 && std::fabs(frac - 0.5) > 1e-6;
  return exact_path ? (fl + (frac > 0.5 ? 1.0 : 0.0)) : 0.0;
}

void put_round4(std::string &o, double v, double *rounded) {
  if (std::isnan(v)) {
    o += "nan";
    if (rounded) *rounded = v;
    return;
  }
  bool fast;
  const double m = round4(v, fast);
  if (fast) {
    const uint64_t u = (uint64_t)m;
    put_int(o, (int64_t)(u / 10000u));
    o.push_back('.');
    const unsigned f = (unsigned)(u % 10000u);
    char d[4] = {(char)('0' + f / 1000), (char)('0' + f / 100 % 10), (char)('0' + f / 10 % 10), (char)('0' + f % 10)};
    int n = 4;
    while (n > 1 && d[n - 1] == '0') --n;
    o.append(d, (size_t)n);
    if (rounded) *rounded = (double)u / 1e4;  // the double nearest to the decimal, like Python's round()
    return;
  }
  char b[48];
  const int n = snprintf(b, sizeof b, "%.4f", v);
  if (rounded) *rounded = strtod(b, nullptr);
  int k = n;
  while (k > 0 && b[k - 1] == '0') --k;
  if (k > 0 && b[k - 1] == '.') ++k;
  o.append(b, (size_t)k);
}

}  // namespace

extern "C" {

// n_sites sites, each with: prefix[s] (the seven leading columns, tab-joined), index[s] (the INDEX column, NULL when
// hccv), chrm[s] (the chrM shortcut applies to this site), hits hit_lo[s] .. hit_hi[s] of the tuple arrays (sorted by
// cell; p already rounded to four places).  cell_text[c] = "barcode\tcell type".  Returns the text length; *text is
// malloc'd (ls_geno_rows_free).
int64_t ls_geno_rows(int32_t n_sites, const char *const *prefix, const char *const *index, const uint8_t *chrm,
                     const int64_t *hit_lo, const int64_t *hit_hi, const int32_t *t_cell, const int32_t *t_dp,
                     const int32_t *t_alt, const double *t_p, int32_t n_cells, const char *const *cell_text, int32_t hccv,
                     double pvalue, char **text) {
  std::string o;
  size_t per_cell = 0;
  for (int32_t c = 0; c < n_cells; ++c) per_cell += strlen(cell_text[c]) + 40;
  size_t total = 0;
  for (int32_t s = 0; s < n_sites; ++s) total += per_cell + (size_t)n_cells * (strlen(prefix[s]) + (index ? strlen(index[s]) : 0) + 4);
  o.reserve(total + 64);
  for (int32_t s = 0; s < n_sites; ++s) {
    const size_t pl = strlen(prefix[s]);
    const size_t il = index ? strlen(index[s]) : 0;
    int64_t h = hit_lo[s];
    const int64_t he = hit_hi[s];
    for (int32_t c = 0; c < n_cells; ++c) {
      int32_t dp = 0, alt = 0;
      double pv = 0.0;
      if (h < he && t_cell[h] == c) {
        dp = t_dp[h];
        alt = t_alt[h];
        pv = t_p[h];
        ++h;
      }
      o.append(prefix[s], pl);
      o.push_back('\t');
      o += cell_text[c];
      o.push_back('\t');
      put_int(o, dp);
      o.push_back('\t');
      put_int(o, alt);
      o.push_back('\t');
      const char *status = "NoCoverage";
      if (dp > 0) {
        double vaf = 0.0;
        if (alt > 0 || !hccv) {
          put_round4(o, (double)alt / (double)dp, &vaf);   // VAF
        } else {
          o += "0.0";                                       // HCCV prints float(0) for a covered pair without ALT reads
        }
        o.push_back('\t');
        if (alt > 0) {
          if (chrm[s]) {
            o.push_back('.');
            status = vaf < 0.3 ? "LowVAFChrM" : "PASS";
          } else {
            put_round4(o, pv, nullptr);                      // BetaBin
            status = pv < pvalue ? "PASS" : "BetaBin_problem";
          }
        } else {
          o.push_back('.');
          status = "NoAltReads";
        }
      } else {
        o += ".\t.";
      }
      o.push_back('\t');
      o += status;
      if (!hccv) {
        o.push_back('\t');
        o.push_back(status[0] == 'P' ? '1' : (status[2] == 'C' ? '3' : '0'));  // PASS 1, NoCoverage 3, everything else 0
        o.push_back('\t');
        o.append(index[s], il);
      }
      o.push_back('\n');
    }
    if (h != he) return -1;  // a hit of a cell outside [0, n_cells) or unsorted hits
  }
  char *buf = (char *)malloc(o.size() + 1);
  if (!buf) return -2;
  memcpy(buf, o.data(), o.size());
  buf[o.size()] = 0;
  *text = buf;
  return (int64_t)o.size();
}

void ls_geno_rows_free(char *text) { free(text); }

}  // extern "C"
