// Shared internals of the longsom_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/longsom_b200.h"

#define LS_NUM_SMS_DEFAULT 148

// BAM CIGAR op codes (SAM spec section 4.2): MIDNSHP=X
enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

// pysam pileup() default flag_filter: UNMAP|SECONDARY|QCFAIL|DUP
#define LS_FLAG_FILTER 0x704u
#define LS_FLAG_PAIRED 0x1u
#define LS_FLAG_PROPER 0x2u
#define LS_FLAG_REVERSE 0x10u
#define LS_FLAG_SUPPL 0x800u

struct DBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

// one (read, tile) work item of the pileup-count kernel (16 B, self-contained: the count kernel never touches the
// per-read arrays)
struct __align__(16) Segment {
  uint32_t p0;      // first piece of the segment in pieces[] (consecutive: a read's CIGAR visits a tile once)
  uint32_t np_nu;   // bits 0-15: number of pieces; bits 16-31: number of 32-base units the pieces expand to
  uint32_t boff16;  // base_off[read] / 16: where the read's qualities / bases start
  uint32_t flags;   // bit 0: reverse strand
};

// One CIGAR op clipped to one tile, produced by the segment builder so that the count kernel does not walk
// CIGARs: a match piece covers query bases [ya, ya + n) at tile columns [col, col + n); a deletion piece covers
// n columns that all carry the quality of query base ya (htslib: qpos of a deletion = the next query base).
// `ind` marks a piece whose last column is the op's last column and is followed by an insertion (1) or a
// deletion (2): that column's class becomes I / D.  A ref-skip followed by an indel is a 1-column deletion piece.
// `virt` marks query positions past the stored sequence (malformed record): quality 0, base 'N'.
struct __align__(8) Piece {
  uint32_t ya;
  uint32_t meta;  // col: bits 0-8, n (1..512): bits 9-18, deletion-like: bit 19, ind: bits 20-21, virt: bit 22
};
__host__ __device__ __forceinline__ uint32_t piece_meta(uint32_t col, uint32_t n, uint32_t del, uint32_t ind,
                                                        uint32_t virt) {
  return col | (n << 9) | (del << 19) | (ind << 20) | (virt << 22);
}
// Units of a piece: the piece cut at the 32-column windows of its tile (a unit never crosses a window, so the
// count kernel can keep one window's per-site state in registers and the padded accumulator rows need no carry).
__host__ __device__ __forceinline__ uint32_t piece_units(uint32_t meta) {
  const uint32_t col = meta & 511u, n = (meta >> 9) & 1023u;
  return ((col + n - 1u) >> 5) - (col >> 5) + 1u;
}

// Device copies of qual[] / seq4[] start LS_QPAD bytes into their allocations (and end with as much slack): a unit
// loads the aligned 40-byte / 24-byte blocks around its bases, which may reach a few bytes past either end.
#define LS_QPAD 64

struct ls_ctx {
  int device = 0;
  int num_sms = LS_NUM_SMS_DEFAULT;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  std::string err;

  // ---- uploaded batch ----
  bool have_batch = false;
  int64_t n_reads = 0, n_cigar = 0, n_bases = 0, n_windows = 0;
  int32_t max_cell = -1;
  DBuf tid, pos, flag, mapq, cell, cigar_off, cigar, base_off, lq, seq4, qual;
  DBuf wtid, wstart, wend, wref_off, ref, wtile_base;
  // seq4 on the device is nibble-swapped at upload (even base in the LOW nibble): base n of the batch then sits at
  // bits 4n .. 4n+3 of the little-endian byte stream, so aligning a unit to its window is one funnel shift
  const uint8_t *qual_d() const { return as_u8(qual) + LS_QPAD; }
  const uint8_t *seq4_d() const { return as_u8(seq4) + LS_QPAD; }
  static const uint8_t *as_u8(const DBuf &b) { return reinterpret_cast<const uint8_t *>(b.p); }
  std::vector<int32_t> h_wtid, h_wstart, h_wend;
  std::vector<int64_t> h_wtile_base;  // [n_windows+1]
  int64_t n_tiles_total = 0;

  // ---- run state ----
  bool have_run = false, compacted = false;
  // sizes of the last completed run on this batch, for the replay mode of ls_pileup_run
  bool cache_valid = false;
  ls_count_params cache_params = {};
  uint64_t cache_seg[3] = {0, 0, 0}, cache_tot[16] = {0};
  int64_t cache_n_sites = 0;
  ls_count_params params = {};
  DBuf segs, pieces, keys_a, keys_b, vals_a, vals_b, rs_hist, scan_tmp, counters;
  DBuf tile_flag, tile_rank, slot_tile, slot_lo, slot_out, slot_mask, slot_npass, slot_off;
  DBuf drop_keys, rend, wcount, part_slot, slot_done, slot_desc, acbuf;  // part_slot: PartDesc records
  DBuf offs_s, offs_m, offs_u, units, goffs, gdir, mrank, mlist;
  int64_t n_drop = 0;
  bool k1_attr_set = false;
  int64_t n_segments = 0, n_pieces = 0, n_slots = 0, n_sites = 0;
  int cell_bits = 0;
  uint64_t *sorted_keys = nullptr;
  uint32_t *sorted_vals = nullptr;
  ls_run_stats stats = {};

  // ---- fetch / misc scratch ----
  DBuf out_tid, out_pos, out_ref, out_counts;
  DBuf l2_scratch;
  DBuf g_a, g_b, g_c, g_d, g_e;  // genotype / betabinom / mask scratch
  // K3: resident sorted site tables (slot LS_SITE_TABLES is the scratch table of the one-shot ls_site_mask)
  DBuf site_tab[LS_SITE_TABLES + 1];
  int64_t site_tab_n[LS_SITE_TABLES + 1] = {};
  bool site_tab_ok[LS_SITE_TABLES + 1] = {};
  DBuf gs_cnt, gs_hits_a, gs_hits_b, gs_flag, gs_tup, gs_p, gs_skip;  // sparse genotyping
  int64_t n_tuples = 0;
  bool have_tuples = false;
  double gs_alpha = 0.0, gs_beta = 0.0;
};

#define LS_CK(call)                                                                        \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      char _b[512];                                                                        \
      snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,                \
               cudaGetErrorString(_e));                                                    \
      ctx->err = _b;                                                                       \
      return LS_E_CUDA;                                                                    \
    }                                                                                      \
  } while (0)

#define LS_FAIL(code, msg) \
  do {                     \
    ctx->err = (msg);      \
    return (code);         \
  } while (0)

static inline int ls_bits_for(uint64_t v) {  // number of bits needed to hold values 0..v
  int b = 0;
  while (v) {
    ++b;
    v >>= 1;
  }
  return b < 1 ? 1 : b;
}

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ bool op_is_match(uint32_t op) { return op == OP_M || op == OP_EQ || op == OP_X; }
__device__ __forceinline__ bool op_consumes_ref(uint32_t op) {
  return op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X;
}

// read-level filter of the pileup engine (pysam __advance_samtools)
__device__ __forceinline__ bool read_passes_engine(uint32_t flag, uint32_t mapq, int min_mq) {
  if (flag & LS_FLAG_FILTER) return false;
  if ((int)mapq < min_mq) return false;
  if ((flag & LS_FLAG_PAIRED) && !(flag & LS_FLAG_PROPER)) return false;  // ignore_orphans
  return true;
}

// Sign of htslib's pileup "indel" field at the LAST reference position of op k
// (resolve_cigar2: peek at the following ops).  Returns +1 (insertion follows),
// -1 (deletion follows), 0.
__device__ __forceinline__ int indel_after(const uint32_t *__restrict__ cigar, uint32_t k, uint32_t kend,
                                           uint32_t op) {
  if (k + 1 >= kend) return 0;
  uint32_t op2 = cigar[k + 1] & 15u;
  if (op2 == OP_D && op != OP_D) return -1;
  if (op2 == OP_I) return +1;
  if (op2 == OP_P && k + 2 < kend) {
    uint32_t l3 = 0;
    for (uint32_t kk = k + 2; kk < kend; ++kk) {
      uint32_t c = cigar[kk];
      uint32_t o = c & 15u;
      if (o == OP_I)
        l3 += c >> 4;
      else if (o == OP_D || o == OP_M || o == OP_N || o == OP_EQ || o == OP_X)
        break;
    }
    if (l3 > 0) return +1;
  }
  return 0;
}

// BAM nibble code -> allele class (EasyReadPileup, BaseCellCounter.py:152-180)
__device__ __forceinline__ int class_of_code(uint32_t code) {
  // =ACMGRSVTWYHKDBN : 1=A 2=C 4=G 8=T 15=N
  switch (code) {
    case 1: return LS_CLASS_A;
    case 2: return LS_CLASS_C;
    case 4: return LS_CLASS_G;
    case 8: return LS_CLASS_T;
    case 15: return LS_CLASS_N;
    default: return LS_CLASS_NA;
  }
}

__device__ __forceinline__ uint8_t class_letter(int cls) {
  // "ACTGIDNO" packed little-endian, one byte per class id
  return (uint8_t)((0x4F4E444947544341ull >> (8 * (cls & 7))) & 0xffu);
}

__device__ __forceinline__ uint8_t upper_ascii(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

int ls_depth_cap_host(ls_ctx *ctx, int min_mq, int max_depth, const std::vector<uint32_t> &wcount);
int ls_tile_size(void);

// ---- utilities implemented in ls_util.cu --------------------------------------------------
// several independent exclusive scans in one launch (ls_util.cu)
constexpr int LS_SCAN_MAX_JOBS = 6;
struct LsScanJob {
  const uint32_t *in;
  uint32_t *out;
  int64_t n;
  uint64_t *total;  // may be null
};
cudaError_t ls_scan_exclusive_u32_multi(const LsScanJob *jobs, int n_jobs, DBuf &tmp, cudaStream_t st);
cudaError_t ls_scan_exclusive_u32(const uint32_t *d_in, uint32_t *d_out, int64_t n, uint64_t *d_total,
                                  DBuf &tmp, cudaStream_t st);
cudaError_t ls_radix_sort_pairs(uint64_t *keys_a, uint64_t *keys_b, uint32_t *vals_a, uint32_t *vals_b,
                                int64_t n, int key_bits, DBuf &hist, uint64_t **sorted_keys,
                                uint32_t **sorted_vals, int num_sms, cudaStream_t st, int *launches);
cudaError_t ls_radix_sort_keys32(uint32_t *keys_a, uint32_t *keys_b, int64_t n, int key_bits, DBuf &hist,
                                 uint32_t **sorted_keys, int num_sms, cudaStream_t st, int *launches);
cudaError_t ls_radix_sort_keys(uint64_t *keys_a, uint64_t *keys_b, int64_t n, int key_bits, DBuf &hist,
                               uint64_t **sorted_keys, int num_sms, cudaStream_t st, int *launches);
