/*
 * ls_synth.c -- deterministic synthetic long-read scRNA data (test + bench infrastructure).
 *
 * Produces exactly the structure-of-arrays batch the C-ABI consumes (include/longsom_b200.h)
 * plus the reference genome, so the CUDA path, the CPU oracle and (through the BAM writer in
 * longsom_b200/bamio.py) the reference scripts all see identical input.  Shapes follow
 * spliced PacBio-Kinnex-like reads (~1.5 kb, lognormal), 2-12 exon genes,
 * indel / mismatch errors, mixed base qualities, MAPQ / flag noise, log-normal expression,
 * planted somatic / germline / editing-like variants, optional hotspot genes and a chrM.
 *
 * Everything is a pure function of (seed, read uid): the generator runs a read three times
 * (position, size, fill) instead of storing per-read state, so it parallelises with OpenMP
 * and needs no memory beyond the outputs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t seed;
  int32_t n_contigs;
  int32_t n_genes;
  int64_t n_reads;
  int32_t n_cells;        /* cells listed in the barcodes file: ids 0..n_cells-1 */
  int32_t n_extra_cells;  /* barcodes NOT in the barcodes file: ids n_cells..n_cells+n_extra-1 */
  double frac_cancer;     /* cells [0, frac*n_cells) are Cancer */
  int32_t n_hot_genes;    /* first n_hot genes receive hot_fraction of all reads (config 4) */
  double hot_fraction;
  int32_t chrm_tid;       /* contig index of chrM, -1 = none; gets one high-depth gene */
  double chrm_fraction;   /* fraction of reads on chrM */
  double mean_len;        /* 1500 */
  double sigma_len;       /* 0.35 */
  double p_mismatch, p_ins, p_del;
  double p_softclip;
  double p_no_cb, p_extra_cb;
  double p_reverse, p_suppl, p_secondary, p_dup, p_qcfail, p_lowmapq;
  int32_t variants_per_gene;
} synth_params;

typedef struct {
  int32_t tid, n_exons, strand;
  int32_t tlen;          /* transcript length */
  int32_t ex_first;      /* index into exon arrays */
  int32_t var_first, n_var;
  double weight;
} gene_t;

typedef struct {
  int32_t tpos;      /* transcript coordinate */
  uint8_t alt;       /* ASCII */
  uint8_t scope;     /* 0 all cells, 1 cancer-cell subset */
  float read_prob;   /* P(read shows alt | cell carries) */
  float cell_frac;   /* fraction of in-scope cells carrying it */
} var_t;

typedef struct {
  synth_params p;
  int32_t *contig_len;
  uint64_t *contig_off; /* into ref */
  uint8_t *ref;
  uint64_t ref_len;
  gene_t *genes;
  int32_t *ex_start, *ex_len, *ex_tstart; /* genomic start, length, transcript start */
  int32_t n_exons_total;
  var_t *vars;
  int32_t n_vars;
  double *cum_w;  /* cumulative gene weights (non-hot, non-chrM) */
  double *cell_cum; /* cumulative cell expression weights */
  int64_t *order;   /* sorted read uids */
  int32_t *rtid, *rpos;
} synth_plan;

/* ---- RNG ---- */
static inline uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
typedef struct { uint64_t s; } rng_t;
static inline rng_t rng_make(uint64_t seed, uint64_t stream, uint64_t ctr) {
  rng_t r; r.s = mix64(seed ^ mix64(stream * 0x632be59bd9b4e019ull + ctr)); return r;
}
static inline uint64_t rng_u64(rng_t *r) { r->s += 0x9e3779b97f4a7c15ull; uint64_t z = r->s;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
static inline double rng_f(rng_t *r) { return (double)(rng_u64(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline int64_t rng_int(rng_t *r, int64_t n) { return (int64_t)(rng_f(r) * (double)n); }
static inline double rng_normal(rng_t *r) {
  double u1 = rng_f(r), u2 = rng_f(r);
  if (u1 < 1e-300) u1 = 1e-300;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
/* number of failures before the next success with probability p (>=0) */
static inline int64_t rng_geom(rng_t *r, double p) {
  if (p <= 0) return (int64_t)1 << 40;
  double u = rng_f(r);
  if (u < 1e-300) u = 1e-300;
  return (int64_t)(log(u) / log1p(-p));
}

static const char BASES[4] = {'A', 'C', 'G', 'T'};
static inline uint8_t code_of(uint8_t c) {
  switch (c) { case 'A': case 'a': return 1; case 'C': case 'c': return 2; case 'G': case 'g': return 4;
               case 'T': case 't': return 8; default: return 15; }
}

static int bsearch_cum(const double *cum, int n, double x) {
  int lo = 0, hi = n - 1;
  while (lo < hi) { int m = (lo + hi) / 2; if (cum[m] <= x) lo = m + 1; else hi = m; }
  return lo;
}

void synth_free(synth_plan *pl) {
  if (!pl) return;
  free(pl->contig_len); free(pl->contig_off); free(pl->ref); free(pl->genes); free(pl->ex_start);
  free(pl->ex_len); free(pl->ex_tstart); free(pl->vars); free(pl->cum_w); free(pl->cell_cum);
  free(pl->order); free(pl->rtid); free(pl->rpos); free(pl);
}

/* per-read header: everything that does not need the error stream */
typedef struct {
  int32_t gene, tstart, tlen, strand, tid, pos;
  int32_t clip5, clip3;
  int32_t cell;
  uint16_t flag; uint8_t mapq;
  rng_t rng; /* state after the header draws: error stream continues from here */
} rhead;

static void read_header(const synth_plan *pl, int64_t uid, rhead *h) {
  const synth_params *p = &pl->p;
  rng_t r = rng_make(p->seed, 7, (uint64_t)uid);
  int g;
  double u = rng_f(&r);
  int n_special = p->n_hot_genes + (p->chrm_tid >= 0 ? 1 : 0);
  if (p->chrm_tid >= 0 && u < p->chrm_fraction) g = p->n_hot_genes; /* the chrM gene sits right after the hot genes */
  else if (p->n_hot_genes > 0 && rng_f(&r) < p->hot_fraction) g = (int)rng_int(&r, p->n_hot_genes);
  else g = n_special + bsearch_cum(pl->cum_w, p->n_genes - n_special, rng_f(&r) * pl->cum_w[p->n_genes - n_special - 1]);
  const gene_t *G = &pl->genes[g];
  double L = exp(log(p->mean_len) + p->sigma_len * rng_normal(&r));
  if (L < 300) L = 300; if (L > 6000) L = 6000;
  int32_t tlen = (int32_t)L;
  if (tlen > G->tlen) tlen = G->tlen;
  int32_t tstart = (int32_t)rng_int(&r, G->tlen - tlen + 1);
  /* most reads of a 3'-biased protocol end at the transcript end */
  if (rng_f(&r) < 0.5) tstart = G->tlen - tlen;
  h->gene = g; h->tstart = tstart; h->tlen = tlen; h->tid = G->tid;
  /* genomic position of tstart */
  int e = 0;
  while (e + 1 < G->n_exons && pl->ex_tstart[G->ex_first + e + 1] <= tstart) ++e;
  h->pos = pl->ex_start[G->ex_first + e] + (tstart - pl->ex_tstart[G->ex_first + e]);
  h->strand = rng_f(&r) < p->p_reverse ? 1 : 0;
  h->clip5 = rng_f(&r) < p->p_softclip ? 5 + (int32_t)rng_int(&r, 36) : 0;
  h->clip3 = rng_f(&r) < p->p_softclip ? 5 + (int32_t)rng_int(&r, 36) : 0;
  uint16_t flag = h->strand ? 0x10 : 0;
  if (rng_f(&r) < p->p_suppl) flag |= 0x800;
  if (rng_f(&r) < p->p_secondary) flag |= 0x100;
  if (rng_f(&r) < p->p_dup) flag |= 0x400;
  if (rng_f(&r) < p->p_qcfail) flag |= 0x200;
  h->flag = flag;
  h->mapq = rng_f(&r) < p->p_lowmapq ? (uint8_t)rng_int(&r, 60) : 60;
  double uc = rng_f(&r);
  if (uc < p->p_no_cb) h->cell = -1;
  else if (uc < p->p_no_cb + p->p_extra_cb && p->n_extra_cells > 0) h->cell = p->n_cells + (int32_t)rng_int(&r, p->n_extra_cells);
  else h->cell = bsearch_cum(pl->cell_cum, p->n_cells, rng_f(&r) * pl->cell_cum[p->n_cells - 1]);
  h->rng = r;
}

static inline int cell_is_cancer(const synth_params *p, int32_t cell) {
  return cell >= 0 && cell < (int32_t)(p->frac_cancer * p->n_cells);
}

/* Materialise one read.  mode 0: sizes only.  mode 1: write cigar / seq4 / qual.
 * seq4/qual point at the read's first base (base offset even). */
static void read_body(const synth_plan *pl, const rhead *h, int mode, int32_t *n_cigar_o, int32_t *lq_o,
                      uint32_t *cigar, uint8_t *seq4, uint8_t *qual) {
  const synth_params *p = &pl->p;
  const gene_t *G = &pl->genes[h->gene];
  rng_t r = h->rng;
  int32_t nc = 0, y = 0;
#define PUSH_OP(op, len) do { if ((len) > 0) { if (mode) cigar[nc] = ((uint32_t)(len) << 4) | (op); ++nc; } } while (0)
#define PUT_BASE(code, q) do { if (mode) { if (y & 1) seq4[y >> 1] |= (uint8_t)(code); else seq4[y >> 1] = (uint8_t)((code) << 4); \
                                 qual[y] = (uint8_t)(q); } ++y; } while (0)
#define DRAW_Q() (qw = rng_u64(&r), (qw & 0xffffu) < 58982u ? 40 : ((qw & 0xffffu) < 62259u ? 93 : 2 + (int)((qw >> 16) % 28u)))
  uint64_t qw; const uint32_t mm_thr = (uint32_t)(p->p_mismatch * 1048576.0);
  for (int i = 0; i < h->clip5; ++i) { int q = DRAW_Q(); int cd = code_of(BASES[rng_int(&r, 4)]); PUT_BASE(cd, q); }
  PUSH_OP(4, h->clip5);
  /* walk the transcript interval [tstart, tstart+tlen) exon by exon */
  int e = 0;
  while (e + 1 < G->n_exons && pl->ex_tstart[G->ex_first + e + 1] <= h->tstart) ++e;
  int32_t t = h->tstart, tend = h->tstart + h->tlen;
  int32_t run_m = 0; /* pending M length */
  int64_t next_ins = rng_geom(&r, p->p_ins), next_del = rng_geom(&r, p->p_del);
  int vi = 0; /* next variant index (variants sorted by tpos) */
  while (vi < G->n_var && pl->vars[G->var_first + vi].tpos < t) ++vi;
  const uint8_t *ref = pl->ref + pl->contig_off[G->tid];
  int first_base = 1;
  while (t < tend) {
    const int32_t es = pl->ex_start[G->ex_first + e], ets = pl->ex_tstart[G->ex_first + e], el = pl->ex_len[G->ex_first + e];
    int32_t stop = ets + el < tend ? ets + el : tend;
    while (t < stop) {
      /* deletion (never first/last base of an exon block, never adjacent to another indel) */
      if (next_del <= 0 && !first_base && run_m > 0 && t + 3 < stop) {
        int32_t dl = 1 + (int32_t)rng_geom(&r, 0.7); if (dl > stop - t - 2) dl = stop - t - 2;
        PUSH_OP(0, run_m); run_m = 0; PUSH_OP(2, dl);
        t += dl; next_del = rng_geom(&r, p->p_del);
        while (vi < G->n_var && pl->vars[G->var_first + vi].tpos < t) ++vi;
        /* the base after a deletion is always a match */
      } else if (next_ins <= 0 && !first_base && run_m > 0 && t + 2 < stop) {
        int32_t il = 1 + (int32_t)rng_geom(&r, 0.7); if (il > 30) il = 30;
        PUSH_OP(0, run_m); run_m = 0; PUSH_OP(1, il);
        for (int i = 0; i < il; ++i) { int q = DRAW_Q(); int cd = code_of(BASES[rng_int(&r, 4)]); PUT_BASE(cd, q); }
        next_ins = rng_geom(&r, p->p_ins);
      }
      --next_del; --next_ins;
      uint8_t rb = ref[es + (t - ets)];
      uint8_t b = rb;
      if (vi < G->n_var && pl->vars[G->var_first + vi].tpos == t) {
        const var_t *v = &pl->vars[G->var_first + vi];
        int in_scope = v->scope == 0 ? (h->cell >= 0) : cell_is_cancer(p, h->cell);
        if (in_scope) {
          uint64_t hsh = mix64(p->seed ^ mix64(((uint64_t)(G->var_first + vi) << 32) | (uint32_t)h->cell));
          double carry = (double)(hsh >> 11) * (1.0 / 9007199254740992.0);
          if (carry < v->cell_frac && rng_f(&r) < v->read_prob) b = v->alt;
        }
        ++vi;
      }
      int q = DRAW_Q();
      if (((uint32_t)(qw >> 24) & 0xfffffu) < mm_thr) b = BASES[(qw >> 44) & 3u];
      int cd = code_of(b);
      PUT_BASE(cd, q);
      ++run_m; ++t; first_base = 0;
    }
    if (t < tend) { /* intron */
      PUSH_OP(0, run_m); run_m = 0;
      int32_t gap = pl->ex_start[G->ex_first + e + 1] - (es + el);
      PUSH_OP(3, gap);
      ++e; first_base = 1;
    }
  }
  PUSH_OP(0, run_m);
  for (int i = 0; i < h->clip3; ++i) { int q = DRAW_Q(); int cd = code_of(BASES[rng_int(&r, 4)]); PUT_BASE(cd, q); }
  PUSH_OP(4, h->clip3);
  *n_cigar_o = nc; *lq_o = y;
#undef PUSH_OP
#undef PUT_BASE
#undef DRAW_Q
}

typedef struct { int32_t tid, pos; int64_t uid; } skey;
static int cmp_skey(const void *a, const void *b) {
  const skey *x = (const skey *)a, *y = (const skey *)b;
  if (x->tid != y->tid) return x->tid < y->tid ? -1 : 1;
  if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
  return x->uid < y->uid ? -1 : (x->uid > y->uid);
}

/* Build genome, genes, variants and the sorted read order.  Returns NULL on failure. */
synth_plan *synth_plan_create(const synth_params *pp, const int32_t *contig_len) {
  synth_plan *pl = (synth_plan *)calloc(1, sizeof(synth_plan));
  pl->p = *pp;
  const synth_params *p = &pl->p;
  pl->contig_len = (int32_t *)malloc(sizeof(int32_t) * p->n_contigs);
  pl->contig_off = (uint64_t *)malloc(sizeof(uint64_t) * (p->n_contigs + 1));
  uint64_t tot = 0;
  for (int c = 0; c < p->n_contigs; ++c) { pl->contig_len[c] = contig_len[c]; pl->contig_off[c] = tot; tot += (uint64_t)contig_len[c]; }
  pl->contig_off[p->n_contigs] = tot;
  pl->ref_len = tot;
  pl->ref = (uint8_t *)malloc(tot ? tot : 1);
  /* reference: 41% GC, lowercase (soft-masked) stretches, N runs */
#pragma omp parallel for schedule(static)
  for (int64_t blk = 0; blk < (int64_t)((tot + 65535) / 65536); ++blk) {
    rng_t r = rng_make(p->seed, 1, (uint64_t)blk);
    uint64_t lo = (uint64_t)blk * 65536, hi = lo + 65536 < tot ? lo + 65536 : tot;
    for (uint64_t i = lo; i < hi; ++i) {
      double u = rng_f(&r);
      pl->ref[i] = u < 0.295 ? 'A' : (u < 0.59 ? 'T' : (u < 0.795 ? 'G' : 'C'));
    }
  }
  {
    rng_t r = rng_make(p->seed, 2, 0);
    uint64_t n_masks = tot / 20000;
    for (uint64_t m = 0; m < n_masks; ++m) {
      uint64_t s = (uint64_t)rng_int(&r, (int64_t)tot), l = 50 + (uint64_t)rng_int(&r, 400);
      for (uint64_t i = s; i < s + l && i < tot; ++i) pl->ref[i] |= 0x20;
    }
    uint64_t n_nruns = tot / 200000;
    for (uint64_t m = 0; m < n_nruns; ++m) {
      uint64_t s = (uint64_t)rng_int(&r, (int64_t)tot), l = 200 + (uint64_t)rng_int(&r, 1600);
      for (uint64_t i = s; i < s + l && i < tot; ++i) pl->ref[i] = 'N';
    }
  }
  /* genes: laid out left to right on contigs proportionally to contig length */
  pl->genes = (gene_t *)calloc((size_t)p->n_genes, sizeof(gene_t));
  pl->ex_start = (int32_t *)malloc(sizeof(int32_t) * (size_t)p->n_genes * 12);
  pl->ex_len = (int32_t *)malloc(sizeof(int32_t) * (size_t)p->n_genes * 12);
  pl->ex_tstart = (int32_t *)malloc(sizeof(int32_t) * (size_t)p->n_genes * 12);
  pl->vars = (var_t *)malloc(sizeof(var_t) * (size_t)p->n_genes * (size_t)(p->variants_per_gene + 1));
  int n_special = p->n_hot_genes + (p->chrm_tid >= 0 ? 1 : 0);
  if (p->n_genes <= n_special) { synth_free(pl); return NULL; }
  uint64_t tot_nuc = 0; /* length available to ordinary genes */
  for (int c = 0; c < p->n_contigs; ++c) if (c != p->chrm_tid) tot_nuc += (uint64_t)contig_len[c];
  rng_t gr = rng_make(p->seed, 3, 0);
  int nex = 0, nvar = 0;
  /* assign genes to contigs */
  int g = 0;
  int *per_contig = (int *)calloc((size_t)p->n_contigs, sizeof(int));
  {
    int n_ord = p->n_genes - (p->chrm_tid >= 0 ? 1 : 0);
    int assigned = 0;
    for (int c = 0; c < p->n_contigs; ++c) {
      if (c == p->chrm_tid) continue;
      per_contig[c] = (int)((double)n_ord * (double)contig_len[c] / (double)tot_nuc);
      assigned += per_contig[c];
    }
    for (int c = 0; assigned < n_ord; c = (c + 1) % p->n_contigs) if (c != p->chrm_tid) { per_contig[c]++; assigned++; }
  }
  /* gene order in the genes[] array: hot genes first, then chrM gene, then ordinary ones; we create
   * them contig by contig and place hot genes among the first ordinary slots */
  int *slot_of = (int *)malloc(sizeof(int) * (size_t)p->n_genes); /* creation index -> genes[] index */
  {
    int idx_ord = n_special, idx_hot = 0, made = 0;
    int n_ord = p->n_genes - (p->chrm_tid >= 0 ? 1 : 0);
    int hot_every = p->n_hot_genes > 0 ? n_ord / p->n_hot_genes : 0;
    for (int i = 0; i < n_ord; ++i) {
      if (p->n_hot_genes > 0 && idx_hot < p->n_hot_genes && i % hot_every == 0) slot_of[made++] = idx_hot++;
      else slot_of[made++] = idx_ord++;
    }
    if (p->chrm_tid >= 0) slot_of[made++] = p->n_hot_genes;
  }
  int made = 0;
  for (int c = 0; c < p->n_contigs; ++c) {
    int ng = (c == p->chrm_tid) ? 1 : per_contig[c];
    if (ng == 0) continue;
    int64_t span = contig_len[c] / ng;
    for (int k = 0; k < ng; ++k, ++made) {
      gene_t *G = &pl->genes[slot_of[made]];
      G->tid = c; G->ex_first = nex; G->var_first = nvar; G->n_var = 0;
      G->strand = rng_f(&gr) < 0.5;
      int n_ex = 2 + (int)rng_int(&gr, 11);
      int64_t lo = (int64_t)k * span + 1, hi = lo + span - 1; /* keep position 0 free */
      if (hi > contig_len[c]) hi = contig_len[c];
      /* exon / intron lengths, shrunk to fit the slot */
      int32_t el[12], il[12];
      int64_t need = 0;
      for (int e = 0; e < n_ex; ++e) { el[e] = 80 + (int32_t)rng_int(&gr, 521); need += el[e]; }
      for (int e = 0; e + 1 < n_ex; ++e) {
        double u = rng_f(&gr);
        il[e] = (int32_t)(100.0 * pow(200.0, u)); /* log-uniform 100..20000 */
        need += il[e];
      }
      if (c == p->chrm_tid) { n_ex = 2; el[0] = (int32_t)(contig_len[c] * 0.4); el[1] = (int32_t)(contig_len[c] * 0.3); il[0] = 120; need = el[0] + el[1] + il[0]; }
      while (need > hi - lo - 10 && n_ex > 2) { --n_ex; need -= el[n_ex] + il[n_ex - 1]; }
      if (need > hi - lo - 10) { /* slot too small: shrink introns */
        for (int e = 0; e + 1 < n_ex; ++e) { need -= il[e] - 100; il[e] = 100; }
      }
      if (need > hi - lo - 10) { for (int e = 0; e < n_ex; ++e) { need -= el[e] - 80; el[e] = 80; } }
      int64_t start = lo + (hi - lo - need > 0 ? rng_int(&gr, hi - lo - need) : 0);
      int32_t tpos = 0;
      for (int e = 0; e < n_ex; ++e) {
        pl->ex_start[nex] = (int32_t)start; pl->ex_len[nex] = el[e]; pl->ex_tstart[nex] = tpos;
        start += el[e]; tpos += el[e];
        if (e + 1 < n_ex) start += il[e];
        ++nex;
      }
      G->n_exons = n_ex; G->tlen = tpos;
      G->weight = exp(1.2 * rng_normal(&gr));
      /* un-mask and un-N the exons so that expressed sequence is callable */
      for (int e = 0; e < n_ex; ++e) {
        uint8_t *rp = pl->ref + pl->contig_off[c] + pl->ex_start[G->ex_first + e];
        for (int32_t i = 0; i < el[e]; ++i) {
          if (rp[i] == 'N') rp[i] = BASES[rng_int(&gr, 4)];
          else if (rng_f(&gr) < 0.9) rp[i] &= (uint8_t)~0x20;
        }
      }
      /* planted variants, sorted by transcript position */
      int nv = p->variants_per_gene;
      if (nv > 0) {
        int32_t *tp = (int32_t *)malloc(sizeof(int32_t) * (size_t)nv);
        for (int v = 0; v < nv; ++v) tp[v] = 5 + (int32_t)rng_int(&gr, G->tlen - 10);
        for (int a = 1; a < nv; ++a) { int32_t x = tp[a]; int b2 = a - 1; while (b2 >= 0 && tp[b2] > x) { tp[b2 + 1] = tp[b2]; --b2; } tp[b2 + 1] = x; }
        for (int v = 0; v < nv; ++v) {
          if (v > 0 && tp[v] <= tp[v - 1] + 12) continue; /* keep variants apart */
          var_t *V = &pl->vars[nvar];
          V->tpos = tp[v];
          /* reference base at tp */
          int e = 0; while (e + 1 < n_ex && pl->ex_tstart[G->ex_first + e + 1] <= tp[v]) ++e;
          uint8_t rb = pl->ref[pl->contig_off[c] + pl->ex_start[G->ex_first + e] + (tp[v] - pl->ex_tstart[G->ex_first + e])];
          rb &= (uint8_t)~0x20;
          double kind = rng_f(&gr);
          if (kind < 0.45) { V->scope = 1; V->cell_frac = (float)(0.05 + 0.75 * rng_f(&gr)); V->read_prob = 0.5f; }      /* somatic */
          else if (kind < 0.70) { V->scope = 0; V->cell_frac = 1.0f; V->read_prob = 0.5f; }                              /* germline het */
          else if (kind < 0.85) { V->scope = 0; V->cell_frac = 1.0f; V->read_prob = (float)(0.05 + 0.25 * rng_f(&gr)); } /* editing-like */
          else { V->scope = 0; V->cell_frac = 1.0f; V->read_prob = (float)(0.02 + 0.1 * rng_f(&gr)); }                   /* PoN-like noise */
          uint8_t alt;
          if (kind >= 0.70 && kind < 0.85 && rb == 'A') alt = 'G';
          else do { alt = BASES[rng_int(&gr, 4)]; } while (alt == rb);
          V->alt = alt;
          ++nvar; G->n_var++;
        }
        free(tp);
      }
    }
  }
  free(per_contig); free(slot_of);
  pl->n_exons_total = nex; pl->n_vars = nvar;
  /* cumulative weights */
  pl->cum_w = (double *)malloc(sizeof(double) * (size_t)(p->n_genes - n_special));
  { double s = 0; for (int i = n_special; i < p->n_genes; ++i) { s += pl->genes[i].weight; pl->cum_w[i - n_special] = s; } }
  pl->cell_cum = (double *)malloc(sizeof(double) * (size_t)p->n_cells);
  { rng_t cr = rng_make(p->seed, 4, 0); double s = 0; for (int i = 0; i < p->n_cells; ++i) { s += exp(0.8 * rng_normal(&cr)); pl->cell_cum[i] = s; } }
  /* read positions and sorted order */
  skey *keys = (skey *)malloc(sizeof(skey) * (size_t)(p->n_reads ? p->n_reads : 1));
#pragma omp parallel for schedule(static)
  for (int64_t u = 0; u < p->n_reads; ++u) { rhead h; read_header(pl, u, &h); keys[u].tid = h.tid; keys[u].pos = h.pos; keys[u].uid = u; }
  qsort(keys, (size_t)p->n_reads, sizeof(skey), cmp_skey);
  pl->order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(p->n_reads ? p->n_reads : 1));
  for (int64_t i = 0; i < p->n_reads; ++i) pl->order[i] = keys[i].uid;
  free(keys);
  (void)g;
  return pl;
}

uint64_t synth_ref_len(const synth_plan *pl) { return pl->ref_len; }
const uint8_t *synth_ref(const synth_plan *pl) { return pl->ref; }
int32_t synth_n_vars(const synth_plan *pl) { return pl->n_vars; }
int32_t synth_n_exons(const synth_plan *pl) { return pl->n_exons_total; }

/* exon table: tid, start, len per exon (for tests / bed files) */
void synth_exons(const synth_plan *pl, int32_t *tid, int32_t *start, int32_t *len) {
  for (int g = 0; g < pl->p.n_genes; ++g)
    for (int e = 0; e < pl->genes[g].n_exons; ++e) {
      int i = pl->genes[g].ex_first + e;
      tid[i] = pl->genes[g].tid; start[i] = pl->ex_start[i]; len[i] = pl->ex_len[i];
    }
}

/* planted variants in genomic coordinates: tid, pos, alt, scope, read_prob, cell_frac */
void synth_variants(const synth_plan *pl, int32_t *tid, int32_t *pos, uint8_t *alt, uint8_t *scope, float *read_prob,
                    float *cell_frac) {
  for (int g = 0; g < pl->p.n_genes; ++g) {
    const gene_t *G = &pl->genes[g];
    for (int v = 0; v < G->n_var; ++v) {
      const var_t *V = &pl->vars[G->var_first + v];
      int e = 0; while (e + 1 < G->n_exons && pl->ex_tstart[G->ex_first + e + 1] <= V->tpos) ++e;
      int i = G->var_first + v;
      tid[i] = G->tid; pos[i] = pl->ex_start[G->ex_first + e] + (V->tpos - pl->ex_tstart[G->ex_first + e]);
      alt[i] = V->alt; scope[i] = V->scope; read_prob[i] = V->read_prob; cell_frac[i] = V->cell_frac;
    }
  }
}

/* pass 1: per-read sizes in sorted order.  cigar_off[n+1], base_off[n+1] (padded to 16). */
void synth_sizes(const synth_plan *pl, const int64_t *sel, int64_t n, uint32_t *cigar_off, uint64_t *base_off, int32_t *l_qseq) {
  int32_t *nc = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < n; ++i) {
    rhead h; read_header(pl, pl->order[sel ? sel[i] : i], &h);
    int32_t c, l; read_body(pl, &h, 0, &c, &l, NULL, NULL, NULL);
    nc[i] = c; l_qseq[i] = l;
  }
  uint32_t co = 0; uint64_t bo = 0;
  for (int64_t i = 0; i < n; ++i) { cigar_off[i] = co; base_off[i] = bo; co += (uint32_t)nc[i]; bo += ((uint64_t)l_qseq[i] + 15u) & ~(uint64_t)15u; }
  cigar_off[n] = co; base_off[n] = bo;
  free(nc);
}

/* pass 2: fill every array (seq4 must be zero-initialised or is overwritten nibble-wise here) */
void synth_fill(const synth_plan *pl, const int64_t *sel, int64_t n, const uint32_t *cigar_off, const uint64_t *base_off, int32_t *tid, int32_t *pos,
                uint16_t *flag, uint8_t *mapq, int32_t *cell, uint32_t *cigar, uint8_t *seq4, uint8_t *qual,
                int64_t *uid_out) {
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < n; ++i) {
    rhead h; read_header(pl, pl->order[sel ? sel[i] : i], &h);
    tid[i] = h.tid; pos[i] = h.pos; flag[i] = h.flag; mapq[i] = h.mapq; cell[i] = h.cell;
    if (uid_out) uid_out[i] = pl->order[sel ? sel[i] : i];
    int32_t c, l;
    uint64_t bo = base_off[i];
    uint64_t padded = base_off[i + 1] - bo;
    memset(seq4 + (bo >> 1), 0, (size_t)(padded >> 1));
    memset(qual + bo, 0, (size_t)padded);
    read_body(pl, &h, 1, &c, &l, cigar + cigar_off[i], seq4 + (bo >> 1), qual + bo);
  }
}

/* header-only view of every read in sorted order: position, end of its gene (an upper bound of
 * the alignment end) and transcript bases (~ aligned bases); used to shard without materialising */
void synth_headers(const synth_plan *pl, int32_t *tid, int32_t *pos, int32_t *gene_end, int32_t *tlen) {
  const int64_t n = pl->p.n_reads;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    rhead h; read_header(pl, pl->order[i], &h);
    const gene_t *G = &pl->genes[h.gene];
    int last = G->ex_first + G->n_exons - 1;
    tid[i] = h.tid; pos[i] = h.pos; tlen[i] = h.tlen;
    gene_end[i] = pl->ex_start[last] + pl->ex_len[last];
  }
}
