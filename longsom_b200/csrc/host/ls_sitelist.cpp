// BaseCellCalling.step2: the RNA-editing / panel-of-normals site lists, read natively (no CUDA).
//
// Replaces the line loop of build_dict (workflow/scripts/SNVCalling/BaseCellCalling.step2.py:197-221) as restated in
// longsom_b200/cli/step2.py:read_site_list: tab separated, '#' comments, columns chrom, pos.  Millions of rows; the
// membership test itself is K3 on the GPU.  Strict: a line the Python loop would not parse the same way (fewer than two
// columns, a position that is not a plain integer, non-ASCII bytes -- a gzip file looks like that) makes the call return 1
// and the caller runs the Python loop, which then applies the reference's "any failure empties the list" rule.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

namespace {
struct SiteList {
  std::vector<int32_t> chrom;
  std::vector<int64_t> pos;
  std::vector<std::string> names;
};
}  // namespace

extern "C" {

// 0 = ok (*h set), 1 = refused, -1 = cannot read the file
int ls_sitelist_read(const char *path, void **h) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  std::vector<char> data;
  {
    std::vector<char> buf(1 << 22);
    size_t n;
    while ((n = fread(buf.data(), 1, buf.size(), f)) > 0) data.insert(data.end(), buf.begin(), buf.begin() + (ptrdiff_t)n);
    fclose(f);
  }
  SiteList *s = new SiteList();
  std::unordered_map<std::string, int32_t> ids;
  int32_t last = -1;
  const char *p = data.data(), *end = p + data.size();
  while (p < end) {
    const char *q = (const char *)memchr(p, '\n', (size_t)(end - p));
    const char *le = q ? q : end;  // line without its '\n'
    if (*p != '#') {
      for (const char *c = p; c + 1 < le; ++c)  // a '\r' inside a line ends a line in the reference's text mode
        if (*c == '\r') {
          delete s;
          return 1;
        }
      const char *t1 = (const char *)memchr(p, '\t', (size_t)(le - p));
      if (!t1) {
        delete s;
        return 1;
      }
      const char *t2 = (const char *)memchr(t1 + 1, '\t', (size_t)(le - t1 - 1));
      const char *ne = t2 ? t2 : le;
      // int(): optional sign, digits; surrounding blanks (a '\r' before the line end) are accepted by Python as well
      const char *a = t1 + 1;
      while (a < ne && (*a == ' ' || *a == '\r')) ++a;
      const char *b = ne;
      while (b > a && (b[-1] == ' ' || b[-1] == '\r')) --b;
      bool neg = false;
      if (a < b && (*a == '-' || *a == '+')) {
        neg = *a == '-';
        ++a;
      }
      if (a >= b || b - a > 18) {
        delete s;
        return 1;
      }
      int64_t v = 0;
      for (const char *c = a; c < b; ++c) {
        if (*c < '0' || *c > '9') {
          delete s;
          return 1;
        }
        v = v * 10 + (*c - '0');
      }
      for (const char *c = p; c < t1; ++c)
        if ((unsigned char)*c >= 0x80 || *c == '\r') {
          delete s;
          return 1;
        }
      const size_t cl = (size_t)(t1 - p);
      int32_t id;
      if (last >= 0 && s->names[(size_t)last].size() == cl && memcmp(s->names[(size_t)last].data(), p, cl) == 0) {
        id = last;
      } else {
        std::string name(p, cl);
        auto it = ids.find(name);
        if (it == ids.end()) {
          id = (int32_t)s->names.size();
          ids.emplace(name, id);
          s->names.push_back(name);
        } else {
          id = it->second;
        }
        last = id;
      }
      s->chrom.push_back(id);
      s->pos.push_back(neg ? -v : v);
    }
    if (!q) break;
    p = q + 1;
  }
  *h = s;
  return 0;
}
int64_t ls_sitelist_n(void *h) { return (int64_t)((SiteList *)h)->pos.size(); }
int32_t ls_sitelist_n_chroms(void *h) { return (int32_t)((SiteList *)h)->names.size(); }
const char *ls_sitelist_chrom(void *h, int32_t i) { return ((SiteList *)h)->names[(size_t)i].c_str(); }
void ls_sitelist_fill(void *h, int32_t *chrom, int64_t *pos) {
  SiteList *s = (SiteList *)h;
  if (!s->pos.empty()) {
    memcpy(chrom, s->chrom.data(), s->chrom.size() * 4);
    memcpy(pos, s->pos.data(), s->pos.size() * 8);
  }
}
void ls_sitelist_free(void *h) { delete (SiteList *)h; }

}  // extern "C"
