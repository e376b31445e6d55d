/*
 * pileup_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C CPU restatement of the reference's pileup counting, used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product path (longsom_b200/) never links, imports or calls this file.
 *
 * Parity status: the reference ships no golden vectors for this path.
 * This restatement is pinned against outputs of the reference scripts themselves,
 * run in the build container over oracle/shims (a literal re-statement of the htslib
 * pileup engine) -- see oracle/make_golden.py and tests/golden/.
 *
 * What it follows, column for column:
 *   oracle_pileup_count   : BaseCellCounter.run_interval (workflow/scripts/SNVCalling/
 *                           BaseCellCounter.py:182-320), EasyReadPileup (:152-180), and the
 *                           pileup semantics of SURVEY.md Appendix A (htslib bam_plp / pysam).
 *   oracle_genotype_count : the pileup loop of SingleCellGenotype.run_interval
 *                           (workflow/scripts/CellClustering/SingleCellGenotype.py:114-178).
 *
 * The structure deliberately differs from the CUDA path: per window, dense per-column
 * accumulators; the window's reads are visited cell by cell, and the NC / CC set
 * cardinalities (the reference's len(set(...)), :283,292) count a cell the first time a
 * per-site / per-(site, class) stamp sees it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/longsom_b200.h"

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

static int is_match(uint32_t op) { return op == OP_M || op == OP_EQ || op == OP_X; }

/* pysam __advance_samtools: flag_filter 0x704, min_mapping_quality, ignore_orphans */
static int engine_keeps(uint32_t flag, uint32_t mapq, int min_mq) {
  if (flag & 0x704u) return 0;
  if ((int)mapq < min_mq) return 0;
  if ((flag & 0x1u) && !(flag & 0x2u)) return 0;
  return 1;
}

/* htslib resolve_cigar2: sign of p->indel at the last reference column of op k */
static int indel_sign(const uint32_t *cig, uint32_t k, uint32_t kend, uint32_t op) {
  if (k + 1 >= kend) return 0;
  uint32_t op2 = cig[k + 1] & 15u;
  if (op2 == OP_D && op != OP_D) return -1;
  if (op2 == OP_I) return 1;
  if (op2 == OP_P && k + 2 < kend) {
    uint32_t l3 = 0;
    for (uint32_t kk = k + 2; kk < kend; ++kk) {
      uint32_t o = cig[kk] & 15u;
      if (o == OP_I)
        l3 += cig[kk] >> 4;
      else if (o == OP_D || o == OP_M || o == OP_N || o == OP_EQ || o == OP_X)
        break;
    }
    if (l3) return 1;
  }
  return 0;
}

static const char CLASS_LETTER[9] = {'A', 'C', 'T', 'G', 'I', 'D', 'N', 'O', '?'};

/* EasyReadPileup on the string pysam would build for this entry */
static int entry_class(uint32_t op, int ind, uint32_t code) {
  if (ind < 0) return LS_CLASS_D; /* x[1] == '-' */
  if (ind > 0) return LS_CLASS_I; /* x[1] == '+' */
  if (is_match(op)) {
    switch (code) {
      case 1: return LS_CLASS_A;
      case 2: return LS_CLASS_C;
      case 4: return LS_CLASS_G;
      case 8: return LS_CLASS_T;
      case 15: return LS_CLASS_N;
      default: return LS_CLASS_NA;
    }
  }
  if (op == OP_D) return LS_CLASS_O; /* '*' */
  return LS_CLASS_NA;                /* '>' '<' */
}

static int32_t read_end(const ls_read_batch *b, int64_t r) {
  int32_t x = b->pos[r];
  for (uint32_t k = b->cigar_off[r]; k < b->cigar_off[r + 1]; ++k) {
    uint32_t op = b->cigar[k] & 15u;
    if (is_match(op) || op == OP_D || op == OP_N) x += (int32_t)(b->cigar[k] >> 4);
  }
  return x;
}

typedef struct {
  int32_t cell;
  int64_t r;
} cellread;

static int cmp_cellread(const void *pa, const void *pb) {
  const cellread *a = (const cellread *)pa, *b = (const cellread *)pb;
  if (a->cell != b->cell) return a->cell < b->cell ? -1 : 1;
  return a->r < b->r ? -1 : (a->r > b->r);
}

static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return x < y ? -1 : (x > y);
}

/* min-heap of int32 for the depth-cap simulation */
typedef struct {
  int32_t *v;
  int64_t n, cap;
} heap;
static void heap_push(heap *h, int32_t x) {
  if (h->n == h->cap) {
    h->cap = h->cap ? h->cap * 2 : 1024;
    h->v = (int32_t *)realloc(h->v, (size_t)h->cap * 4);
  }
  int64_t i = h->n++;
  h->v[i] = x;
  while (i > 0) {
    int64_t p = (i - 1) / 2;
    if (h->v[p] <= h->v[i]) break;
    int32_t t = h->v[p];
    h->v[p] = h->v[i];
    h->v[i] = t;
    i = p;
  }
}
static void heap_pop(heap *h) {
  h->v[0] = h->v[--h->n];
  int64_t i = 0;
  for (;;) {
    int64_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < h->n && h->v[l] < h->v[m]) m = l;
    if (r < h->n && h->v[r] < h->v[m]) m = r;
    if (m == i) break;
    int32_t t = h->v[m];
    h->v[m] = h->v[i];
    h->v[i] = t;
    i = m;
  }
}

/* Records one pileup(CHROM, START, END) call feeds to the column engine, in file order:
 * region fetch + engine filter + max_depth rule. */
static int64_t fetch_window(const ls_read_batch *b, const int32_t *rend, const int32_t *pmax, int32_t tid,
                            int32_t ws, int32_t we, int min_mq, int max_depth, int64_t **list, int64_t *cap) {
  /* first read of the contig */
  int64_t lo = 0, hi = b->n_reads;
  while (lo < hi) {
    int64_t m = (lo + hi) / 2;
    if (b->tid[m] < tid) lo = m + 1; else hi = m;
  }
  int64_t r0 = lo;
  /* first read whose running max end exceeds ws (pmax is non-decreasing within a contig) */
  hi = b->n_reads;
  while (lo < hi) {
    int64_t m = (lo + hi) / 2;
    if (b->tid[m] == tid && pmax[m] <= ws) lo = m + 1; else hi = m;
  }
  (void)r0;
  int64_t n = 0;
  heap live = {0, 0, 0};
  int32_t last_p = -1;
  int any_at_p = 0;
  for (int64_t r = lo; r < b->n_reads && b->tid[r] == tid && b->pos[r] < we; ++r) {
    if (!engine_keeps(b->flag[r], b->mapq[r], min_mq)) continue;
    int32_t e = rend[r] > b->pos[r] ? rend[r] : b->pos[r] + 1; /* bam_endpos */
    if (e <= ws) continue;
    int32_t P = b->pos[r];
    if (max_depth > 0) {
      if (P != last_p) { last_p = P; any_at_p = 0; }
      while (live.n && live.v[0] < P) heap_pop(&live);
      if (any_at_p && 1 + live.n > (int64_t)max_depth) continue; /* bam_plp_push overflow drop */
      any_at_p = 1;
      heap_push(&live, rend[r]);
    }
    if (n == *cap) {
      *cap = *cap ? *cap * 2 : 4096;
      *list = (int64_t *)realloc(*list, (size_t)*cap * 8);
    }
    (*list)[n++] = r;
  }
  free(live.v);
  return n;
}

typedef struct {
  int64_t n;
  int32_t *pos;
  uint8_t *ref;
  uint32_t *counts;
} chunk;

static void count_window(const ls_read_batch *b, const ls_windows *w, int64_t wi, const int32_t *rend,
                         const int32_t *pmax, const ls_count_params *P, chunk *out) {
  const int32_t tid = w->tid[wi], ws = w->start[wi], we = w->end[wi];
  const int32_t L = we - ws;
  out->n = 0;
  out->pos = NULL;
  out->ref = NULL;
  out->counts = NULL;
  if (L <= 0) return;
  int64_t *list = NULL, cap = 0;
  int64_t nr = fetch_window(b, rend, pmax, tid, ws, we, P->min_mq, P->max_depth, &list, &cap);
  if (nr == 0) { free(list); return; }
  uint32_t *aligned = (uint32_t *)calloc((size_t)L, 4);  /* get_num_aligned */
  uint32_t *callable = (uint32_t *)calloc((size_t)L, 4); /* entries whose class != NA */
  uint32_t *ac = (uint32_t *)calloc((size_t)L, 4);       /* EasyReadPileup AC */
  uint32_t *cf = (uint32_t *)calloc((size_t)L * 8, 4);
  uint32_t *cr = (uint32_t *)calloc((size_t)L * 8, 4);
  uint32_t *bq = (uint32_t *)calloc((size_t)L * 8, 4);
  /* Distinct cells per site (NC) and per (site, class) (CC): the reads are visited cell by cell (every per-site sum
   * is order independent), so "this cell has already been seen here" is one stamp per site / per (site, class). */
  uint32_t *nc = (uint32_t *)calloc((size_t)L, 4);
  uint32_t *cc = (uint32_t *)calloc((size_t)L * 8, 4);
  uint32_t *stamp_any = (uint32_t *)calloc((size_t)L, 4);
  uint32_t *stamp_cls = (uint32_t *)calloc((size_t)L * 8, 4);
  {
    cellread *cr2 = (cellread *)malloc((size_t)nr * sizeof(cellread));
    for (int64_t li = 0; li < nr; ++li) { cr2[li].cell = b->cell[list[li]]; cr2[li].r = list[li]; }
    qsort(cr2, (size_t)nr, sizeof(cellread), cmp_cellread);
    for (int64_t li = 0; li < nr; ++li) list[li] = cr2[li].r;
    free(cr2);
  }
  const uint8_t *ref = w->ref + w->ref_off[wi];
  for (int64_t li = 0; li < nr; ++li) {
    const int64_t r = list[li];
    const uint32_t k0 = b->cigar_off[r], kend = b->cigar_off[r + 1];
    const uint32_t flag = b->flag[r];
    const int counted = b->cell[r] >= 0 && !(flag & 0x800u);
    const int rev = (flag & 0x10u) != 0;
    const uint8_t *q = b->qual + b->base_off[r];
    const uint8_t *s4 = b->seq4 + (b->base_off[r] >> 1);
    const uint32_t lq = (uint32_t)b->l_qseq[r];
    int32_t x = b->pos[r];
    uint32_t y = 0;
    for (uint32_t k = k0; k < kend; ++k) {
      const uint32_t op = b->cigar[k] & 15u;
      const int32_t len = (int32_t)(b->cigar[k] >> 4);
      if (is_match(op) || op == OP_D || op == OP_N) {
        for (int32_t j = 0; j < len; ++j) {
          const int32_t p = x + j;
          if (p < ws || p >= we) continue;
          const uint32_t qpos = is_match(op) ? y + (uint32_t)j : y;
          const uint32_t qv = qpos < lq ? q[qpos] : 0u;
          if ((int)qv < P->min_bq) continue; /* pileup_base_qual_skip */
          const int s = p - ws;
          aligned[s]++;
          int ind = (j == len - 1) ? indel_sign(b->cigar, k, kend, op) : 0;
          uint32_t code = 15u;
          if (is_match(op) && qpos < lq) code = (qpos & 1u) ? (s4[qpos >> 1] & 15u) : (uint32_t)(s4[qpos >> 1] >> 4);
          const int cls = entry_class(op, ind, code);
          if (cls == LS_CLASS_NA) continue;
          callable[s]++;
          {
            uint8_t rb = ref[s];
            if (rb >= 'a' && rb <= 'z') rb -= 32;
            if (cls == LS_CLASS_D || cls == LS_CLASS_I) ac[s]++;
            else if (cls != LS_CLASS_O && (uint8_t)CLASS_LETTER[cls] != rb) ac[s]++;
          }
          if (!counted) continue;
          if (rev) cr[(size_t)s * 8 + cls]++; else cf[(size_t)s * 8 + cls]++;
          bq[(size_t)s * 8 + cls] += qv;
          {
            const uint32_t tag = (uint32_t)b->cell[r] + 1u;
            if (stamp_any[s] != tag) { stamp_any[s] = tag; nc[s]++; }
            if (stamp_cls[(size_t)s * 8 + cls] != tag) { stamp_cls[(size_t)s * 8 + cls] = tag; cc[(size_t)s * 8 + cls]++; }
          }
        }
      }
      if (is_match(op)) { x += len; y += (uint32_t)len; }
      else if (op == OP_D || op == OP_N) x += len;
      else if (op == OP_I || op == OP_S) y += (uint32_t)len;
    }
  }
  free(list);
  free(stamp_any);
  free(stamp_cls);
  out->pos = (int32_t *)malloc((size_t)L * 4);
  out->ref = (uint8_t *)malloc((size_t)L);
  out->counts = (uint32_t *)malloc((size_t)L * LS_SITE_WORDS * 4);
  int64_t n = 0;
  for (int32_t s = 0; s < L; ++s) {
    uint8_t rb = ref[s];
    if (rb >= 'a' && rb <= 'z') rb -= 32;
    if (aligned[s] == 0) continue;                                 /* no pileup column */
    if (!((int)aligned[s] >= P->min_dp && rb != 'N')) continue;    /* :211 */
    if ((int)callable[s] < P->min_dp || (int)ac[s] < P->min_ac) continue; /* :221 */
    uint32_t count = 0;
    for (int c = 0; c < 8; ++c) count += cf[(size_t)s * 8 + c] + cr[(size_t)s * 8 + c];
    if ((int)count < P->min_dp) continue;                          /* :282 */
    if ((int)nc[s] < P->min_cc) continue;                          /* :294 */
    uint32_t *o = out->counts + (size_t)n * LS_SITE_WORDS;
    o[LS_SITE_DP] = count;
    o[LS_SITE_NC] = nc[s];
    for (int c = 0; c < 6; ++c) {
      o[LS_SITE_CC + c] = cc[(size_t)s * 8 + c];
      o[LS_SITE_BCF + c] = cf[(size_t)s * 8 + c];
      o[LS_SITE_BCR + c] = cr[(size_t)s * 8 + c];
      o[LS_SITE_BQ + c] = bq[(size_t)s * 8 + c];
    }
    out->pos[n] = ws + s;
    out->ref[n] = rb;
    ++n;
  }
  out->n = n;
  free(aligned); free(callable); free(ac); free(cf); free(cr); free(bq); free(nc); free(cc);
}

static void prep_ends(const ls_read_batch *b, int32_t **rend_o, int32_t **pmax_o) {
  int32_t *rend = (int32_t *)malloc((size_t)(b->n_reads + 1) * 4);
  int32_t *pmax = (int32_t *)malloc((size_t)(b->n_reads + 1) * 4);
  for (int64_t r = 0; r < b->n_reads; ++r) {
    rend[r] = read_end(b, r);
    int32_t e = rend[r] > b->pos[r] ? rend[r] : b->pos[r] + 1;
    pmax[r] = (r > 0 && b->tid[r - 1] == b->tid[r] && pmax[r - 1] > e) ? pmax[r - 1] : e;
  }
  *rend_o = rend;
  *pmax_o = pmax;
}

/* Returns the number of passing sites (also when it exceeds out->capacity; then nothing is
 * written beyond capacity and the caller should retry), or a negative value on bad input. */
int64_t oracle_pileup_count(const ls_read_batch *b, const ls_windows *w, const ls_count_params *P,
                            ls_site_counts *out, int threads, int64_t *n_aligned) {
  int32_t *rend, *pmax;
  prep_ends(b, &rend, &pmax);
  if (n_aligned) {
    int64_t al = 0;
    for (int64_t k = 0; k < b->n_cigar; ++k)
      if (is_match(b->cigar[k] & 15u)) al += b->cigar[k] >> 4;
    *n_aligned = al;
  }
  chunk *ch = (chunk *)calloc((size_t)(w->n_windows ? w->n_windows : 1), sizeof(chunk));
  if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int64_t wi = 0; wi < w->n_windows; ++wi) count_window(b, w, wi, rend, pmax, P, &ch[wi]);
  int64_t n = 0;
  for (int64_t wi = 0; wi < w->n_windows; ++wi) {
    for (int64_t i = 0; i < ch[wi].n; ++i, ++n) {
      if (n < out->capacity) {
        out->tid[n] = w->tid[wi];
        out->pos[n] = ch[wi].pos[i];
        out->ref[n] = ch[wi].ref[i];
        memcpy(out->counts + (size_t)n * LS_SITE_WORDS, ch[wi].counts + (size_t)i * LS_SITE_WORDS,
               LS_SITE_WORDS * 4);
      }
    }
    free(ch[wi].pos); free(ch[wi].ref); free(ch[wi].counts);
  }
  free(ch); free(rend); free(pmax);
  out->n_sites = n;
  return n;
}

/* SingleCellGenotype.run_interval pileup loop (:114-178).  The reference makes one pileup()
 * call per 50 kb bin of candidate sites over [min-1, max+1); bin_id[s] names the bin of each
 * site (sites sorted by (tid,pos), bins contiguous).  dp/alt are [n_sites][n_cells]. */
int64_t oracle_genotype_count(const ls_read_batch *b, const int32_t *site_tid, const int32_t *site_pos,
                              const uint8_t *alt_class, const int64_t *bin_id, int64_t n_sites, int32_t n_cells,
                              const ls_geno_params *P, int32_t *dp, int32_t *alt) {
  int32_t *rend, *pmax;
  prep_ends(b, &rend, &pmax);
  memset(dp, 0, (size_t)n_sites * (size_t)n_cells * 4);
  memset(alt, 0, (size_t)n_sites * (size_t)n_cells * 4);
  int64_t events = 0;
  int64_t s0 = 0;
  while (s0 < n_sites) {
    int64_t s1 = s0 + 1;
    while (s1 < n_sites && bin_id[s1] == bin_id[s0]) ++s1;
    const int32_t tid = site_tid[s0];
    const int32_t ws = site_pos[s0] - 1, we = site_pos[s1 - 1] + 1; /* START, END of the pileup call */
    int64_t *list = NULL, cap = 0;
    int64_t nr = fetch_window(b, rend, pmax, tid, ws < 0 ? 0 : ws, we, P->min_mq, P->max_depth, &list, &cap);
    for (int64_t li = 0; li < nr; ++li) {
      const int64_t r = list[li];
      const uint32_t k0 = b->cigar_off[r], kend = b->cigar_off[r + 1];
      const uint32_t flag = b->flag[r];
      const int32_t cell = b->cell[r];
      if (cell < 0 || cell >= n_cells || (flag & 0x800u)) continue;
      const uint8_t *q = b->qual + b->base_off[r];
      const uint8_t *s4 = b->seq4 + (b->base_off[r] >> 1);
      const uint32_t lq = (uint32_t)b->l_qseq[r];
      int32_t x = b->pos[r];
      uint32_t y = 0;
      for (uint32_t k = k0; k < kend; ++k) {
        const uint32_t op = b->cigar[k] & 15u;
        const int32_t len = (int32_t)(b->cigar[k] >> 4);
        if ((is_match(op) || op == OP_D || op == OP_N) && len > 0) {
          /* target sites inside [x, x+len) */
          int64_t lo = s0, hi = s1;
          while (lo < hi) {
            int64_t m = (lo + hi) / 2;
            if (site_pos[m] < x) lo = m + 1; else hi = m;
          }
          for (int64_t s = lo; s < s1 && site_pos[s] < x + len; ++s) {
            const int32_t j = site_pos[s] - x;
            const uint32_t qpos = is_match(op) ? y + (uint32_t)j : y;
            const uint32_t qv = qpos < lq ? q[qpos] : 0u;
            if ((int)qv < P->min_bq) continue;
            int ind = (j == len - 1) ? indel_sign(b->cigar, k, kend, op) : 0;
            uint32_t code = 15u;
            if (is_match(op) && qpos < lq) code = (qpos & 1u) ? (s4[qpos >> 1] & 15u) : (uint32_t)(s4[qpos >> 1] >> 4);
            const int cls = entry_class(op, ind, code);
            int use;
            if (P->alt_only) use = (cls == (int)alt_class[s]) && cls != LS_CLASS_NA;
            else use = cls != LS_CLASS_NA && cls != LS_CLASS_O;
            if (!use) continue;
            dp[(size_t)s * n_cells + cell]++;
            if (cls == (int)alt_class[s]) alt[(size_t)s * n_cells + cell]++;
            ++events;
          }
        }
        if (is_match(op)) { x += len; y += (uint32_t)len; }
        else if (op == OP_D || op == OP_N) x += len;
        else if (op == OP_I || op == OP_S) y += (uint32_t)len;
      }
    }
    free(list);
    s0 = s1;
  }
  free(rend); free(pmax);
  (void)cmp_i32;
  return events;
}
