// K1 pileup-count kernel (third generation) + part bookkeeping.  Included by ls_pileup.cu.
//
// Unit of work: a PART = up to K1_PART_SEGS consecutive (tile, cell)-sorted segments of one
// tile.  Most tiles are one part; deep tiles (chrM, hotspot genes: >1e5 reads per locus) are
// split so that no CTA owns more than ~4e5 pileup entries.  Same-cell runs never straddle
// parts (a run belongs to the part / chunk that holds its first segment), so NC / CC stay
// exact and every output word is additive across parts; multi-part tiles add into their HBM
// slot and the last part to finish applies the reference's gates.
//
// Inside a part (CTA of 8 warps, tile accumulators in shared memory):
//   * a warp grabs a chunk of <= 32 segments; lane j loads segment j's record and read
//     metadata (one global-latency for the whole chunk instead of one per segment);
//     and L2-prefetches the query bytes / first piece line that segment will touch;
//   * per segment the lanes fetch its pieces (CIGAR ops clipped to the tile, precomputed by the segment
//     builder in ls_segments.cuh) 32 at a time and go through them warp-uniformly -- no CIGAR walk here;
//   * a match piece is consumed 128 query bases per step: lane L loads one aligned 32-bit word of
//     qualities and one 16-bit word of 4-bit bases (vectorised, coalesced), classifies its 4
//     bases and issues ONE packed shared atomic per visible base (count<<20 | quality).
// (A two-phase variant that flattens op pieces into a per-warp unit queue was measured 1.8x
//  slower on C2 -- register spills and the per-word unit search cost more than the idle lanes.)
// Distinct-cell counts: NC/CC = reads minus same-cell duplicates.  Runs of >1 same-cell
// segments mark (site, class) bits with atomicOr in a per-warp seen[] array, so the order in
// which a run's bases are visited does not matter; single-segment runs skip this entirely.
// The accumulator column of site s is [s mod 4][s / 4]: the 4-bases-per-lane pattern then
// hits 32 distinct banks on each of its 4 atomics.
#pragma once

constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_PART_SEGS = 2048;       // nominal segments per part
constexpr int K1_MAX_RUN_PACKED = 2039;  // packed counters (12-bit) need: part segs + run extension <= 4095
constexpr int K1_ROWS = 18;               // 9 row pairs: 8 classes + the dump pair
constexpr uint32_t K1_CLASS_STRIDE = 2u * LS_TILE * 4u;  // bytes between the row pairs of consecutive classes
constexpr uint32_t K1_DUMP_OFF = 8u * K1_CLASS_STRIDE;
constexpr uint32_t K1_CNT_SHIFT = 20;    // packed word: count << 20 | base-quality sum (<= 4095 * 255 < 2^20)

static_assert(LS_TILE % 128 == 0 && LS_TILE % K1_THREADS == 0 && LS_TILE <= 65536, "tile / block shape");

__host__ __device__ __forceinline__ int swz(int s) { return ((s & 3) * (LS_TILE / 4)) | (s >> 2); }

struct CountArgs {
  const uint16_t *flag;
  const Piece *pieces;
  const uint64_t *base_off;
  const int32_t *lq;
  const uint8_t *seq4, *qual;
  const Segment *segs;
  const uint64_t *keys;
  const uint32_t *vals;
  const int64_t *slot_tile;
  const uint32_t *slot_lo;
  const uint32_t *part_slot, *part_k, *slot_nparts;
  uint32_t *slot_done;
  const uint32_t *n_parts;
  int64_t n_windows;
  const int32_t *wstart, *wend;
  const int64_t *wtile_base;
  const uint64_t *wref_off;
  const uint8_t *ref;
  uint32_t *out;    // [n_slots][LS_SITE_WORDS][LS_TILE]
  uint32_t *acbuf;  // [n_slots][LS_TILE], only when min_ac > 0
  uint32_t *mask;   // [n_slots][LS_TILE/32]
  uint32_t *npass;  // [n_slots]
  unsigned long long *n_events;
  int cell_bits;
  uint32_t uncounted_key;
  int min_bq, min_dp, min_cc, min_ac;
  int prefetch;
};

template <bool PACKED>
struct TileSmemT {
  // rows [cls*2+strand], cls 0..7 real classes, cls 8 = "dump" rows that absorb invisible / ignored bases so that
  // the 4-base step needs no branches.  PACKED: cnt<<20|bq; else rows [0..17] counts, [18..35] quality sums.
  uint32_t hist[PACKED ? K1_ROWS : 2 * K1_ROWS][LS_TILE];
  uint32_t dupcc[3][LS_TILE];                // (cell, class) already seen at the site: class c in half-word c/3 of row c%3
  uint32_t lut2[16];                         // BAM nibble code -> byte offset of the class row pair (dump rows if ignored)
  uint32_t dupnc[LS_TILE];                   // entries whose cell was already seen at the site (any class)
  uint32_t acx[LS_TILE];                     // alt entries of visible-but-uncounted reads (--min_ac > 0)
  uint32_t seen[K1_WARPS][LS_TILE / 4];      // per warp: class bits (1 byte per site) of the current same-cell run
  uint8_t ref[LS_TILE];
  uint32_t next, npass, ticket;
};

__device__ __forceinline__ int64_t window_of_tile(const CountArgs &a, int64_t tile) {
  int64_t lo = 0, hi = a.n_windows;  // last w with wtile_base[w] <= tile
  while (hi - lo > 1) {
    int64_t m = (lo + hi) >> 1;
    if (a.wtile_base[m] <= tile)
      lo = m;
    else
      hi = m;
  }
  return lo;
}

// shared-memory reduction on a 32-bit shared-window address (no generic->shared conversion per base)
__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v) {
  // no "memory" clobber: nothing reads the accumulators before the CTA barrier, so loads may move across
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v));
}

// hs = shared address of hist[strand][0]: the (class, strand) row of class c starts 2*LS_TILE words further per class
template <bool PACKED>
__device__ __forceinline__ void note_seen(TileSmemT<PACKED> &sm, uint32_t *seen, int s, int cls) {
  const int sidx = swz(s);
  const int sh = 8 * (s & 3);
  const uint32_t old = (atomicOr(&seen[s >> 2], (1u << cls) << sh) >> sh) & 255u;
  if (((old >> cls) & 1u) && cls < 6) atomicAdd(&sm.dupcc[cls % 3][sidx], 1u << (16 * (cls / 3)));
  if (old) atomicAdd(&sm.dupnc[sidx], 1u);
}

template <bool PACKED, bool SEEN>
__device__ __forceinline__ void add_entry(TileSmemT<PACKED> &sm, uint32_t *seen, int s, int cls, uint32_t q,
                                          uint32_t hs) {
  const uint32_t addr = hs + (uint32_t)cls * K1_CLASS_STRIDE + (uint32_t)swz(s) * 4u;
  if (PACKED) {
    red_shared_add(addr, (1u << K1_CNT_SHIFT) | q);
  } else {
    red_shared_add(addr, 1u);
    red_shared_add(addr + (uint32_t)K1_ROWS * LS_TILE * 4u, q);
  }
  if (SEEN) note_seen<PACKED>(sm, seen, s, cls);
}

template <bool PACKED>
__device__ __forceinline__ void add_uncounted(TileSmemT<PACKED> &sm, int s, int cls) {
  // AC pre-gate of BaseCellCounter.py:165-174,221 for reads the counts ignore (no CB / supplementary)
  const bool alt = (cls == LS_CLASS_D || cls == LS_CLASS_I) || (cls != LS_CLASS_O && class_letter(cls) != sm.ref[s]);
  if (alt) atomicAdd(&sm.acx[swz(s)], 1u);
}

// nibble code -> class id, 4 bits per entry: 1->A(0) 2->C(1) 4->G(3) 8->T(2) 15->N(6), else NA(8)
#define K1_CLASS_LUT 0x6888888288838108ull

struct SegMeta {
  uint32_t p0, np, lq;
  uint32_t y0, qlen;  // query extent of the segment (prefetch only; not shuffled)
  uint64_t boff;
  int strand;
};

__device__ __forceinline__ SegMeta load_meta(const CountArgs &a, uint32_t i) {
  const Segment sg = a.segs[a.vals[i]];
  SegMeta m;
  m.p0 = sg.p0;
  m.np = sg.np & 0xffffu;
  m.y0 = sg.y0;
  m.qlen = sg.np >> 16;
  m.boff = a.base_off[sg.read];
  m.lq = (uint32_t)a.lq[sg.read];
  m.strand = (a.flag[sg.read] & LS_FLAG_REVERSE) ? 1 : 0;
  return m;
}

__device__ __forceinline__ SegMeta shfl_meta(const SegMeta &m, int j) {
  SegMeta r;
  r.p0 = __shfl_sync(0xffffffffu, m.p0, j);
  r.np = __shfl_sync(0xffffffffu, m.np, j);
  r.lq = __shfl_sync(0xffffffffu, m.lq, j);
  r.boff = __shfl_sync(0xffffffffu, m.boff, j);
  r.strand = __shfl_sync(0xffffffffu, m.strand, j);
  return r;
}

// ---- one segment, warp-cooperative: its pieces (CIGAR ops clipped to the tile, precomputed by the segment
// builder) are fetched 32 at a time; a match piece is consumed 128 query bases per step (4 per lane, one 32-bit
// quality word + one 16-bit base word) ----------------------------------------------------------------------
template <bool PACKED, bool SEEN, bool COUNTED>
__device__ __forceinline__ void process_segment_fast(const CountArgs &a, TileSmemT<PACKED> &sm, uint32_t *seen,
                                                     const SegMeta m, int lane, uint32_t hist_s) {
  const uint8_t *__restrict__ qual = a.qual + m.boff;
  const uint8_t *__restrict__ seq4 = a.seq4 + (m.boff >> 1);
  const uint32_t strand = hist_s + (uint32_t)m.strand * (LS_TILE * 4u);  // shared address of hist[strand][0]
  const uint32_t lq = m.lq;
  for (uint32_t pb = 0; pb < m.np; pb += 32u) {
    Piece pc;
    pc.ya = 0u;
    pc.meta = 0u;
    if (pb + (uint32_t)lane < m.np) pc = a.pieces[m.p0 + pb + (uint32_t)lane];
    const int nloc = (int)((m.np - pb) < 32u ? (m.np - pb) : 32u);
    for (int t = 0; t < nloc; ++t) {
      const uint32_t y0 = __shfl_sync(0xffffffffu, pc.ya, t);
      const uint32_t meta = __shfl_sync(0xffffffffu, pc.meta, t);
      const int sbase = (int)(meta & 511u);
      const uint32_t n = (meta >> 9) & 1023u;
      const uint32_t ind = (meta >> 20) & 3u;
      const int indcls = ind == 2u ? LS_CLASS_D : LS_CLASS_I;
      {
        if ((meta >> 19) & 1u) {  // deletion / ref-skip columns: every column carries the quality of the next query base
          const uint32_t q = y0 < lq ? qual[y0] : 0u;
          if ((int)q >= a.min_bq) {
            for (uint32_t p = (uint32_t)lane; p < n; p += 32u) {
              const int cls = (p == n - 1u && ind != 0u) ? indcls : LS_CLASS_O;
              if (COUNTED)
                add_entry<PACKED, SEEN>(sm, seen, sbase + (int)p, cls, q, strand);
              else
                add_uncounted<PACKED>(sm, sbase + (int)p, cls);
            }
          }
        } else {
          const uint32_t ya = y0, yb = y0 + n;                                       // query range inside the tile
          const uint32_t ybl = yb < lq ? yb : lq;                                    // bases that exist
          const uint32_t ylast = (ind != 0) ? (yb - 1u) : 0xffffffffu;
          // the base that carries a following indel is handled on its own (below); the word loop stops before it
          const uint32_t yw = (ind != 0 && ylast < ybl) ? ylast : ybl;
          if (COUNTED && a.min_bq >= 0 && a.min_bq <= 128) {
            // Branch-free 4-base step.  Class decode = one shared lookup per base byte (2 bases); bases that must
            // not count (quality below min_bq, ignored code, outside [ya, yw)) are steered to the dump rows instead
            // of being branched around.  (site & 3) of byte j is identical for all lanes and steps, so the
            // swizzled column of byte j is colb[j] + 4 * (site_of_byte0 >> 2) with colb[] warp-uniform.
            const int a4 = (sbase - (int)(ya & 3u)) & 3;
            uint32_t colb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) colb[j] = (uint32_t)((((a4 + j) & 3) * (LS_TILE / 4)) + ((a4 + j) >> 2)) * 4u;
            const uint32_t mq4 = (uint32_t)a.min_bq * 0x01010101u;
            for (uint32_t g = (ya & ~3u) + 4u * (uint32_t)lane; g < yw; g += 128u) {
              const uint32_t w = *reinterpret_cast<const uint32_t *>(qual + g);
              const uint32_t h = *reinterpret_cast<const uint16_t *>(seq4 + (g >> 1));
              const int d0 = (int)(g - ya);  // index inside the piece of this word's byte 0 (< 0 in the head word)
              // bit 8j+7 <=> quality of byte j >= min_bq   ((q|0x80) >= 128 >= min_bq: no borrow between bytes)
              uint32_t okm = (((w | 0x80808080u) - mq4) | w) & 0x80808080u;
              const int hi = (int)(yw - g);
              if (d0 < 0 || hi < 4) {  // head / tail word
                uint32_t m = 0xffffffffu;
                if (d0 < 0) m <<= 8 * (-d0);
                if (hi < 4) m &= 0xffffffffu >> (8 * (4 - hi));
                okm &= m;
              }
              const int c4 = (sbase + d0) >> 2;  // site of byte 0, divided by 4 (arithmetic shift: -1 in the head word)
              const uint32_t rowb = strand + (uint32_t)c4 * 4u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // 16-entry table: every lane pair reads the same word (broadcast) or different banks -> conflict-free
                uint32_t off = sm.lut2[(h >> (8 * (j >> 1) + ((j & 1) ? 0 : 4))) & 15u];
                if (!((okm >> (8 * j + 7)) & 1u)) off = K1_DUMP_OFF;
                const uint32_t q = (w >> (8 * j)) & 255u;
                const uint32_t addr = rowb + colb[j] + off;
                if (PACKED) {
                  red_shared_add(addr, (1u << K1_CNT_SHIFT) | q);
                } else {
                  red_shared_add(addr, 1u);
                  red_shared_add(addr + (uint32_t)K1_ROWS * LS_TILE * 4u, q);
                }
                if (SEEN) {
                  if (off != K1_DUMP_OFF) {  // same-cell duplicate marks; (site & 3) and the column are warp-uniform per j
                    const int tj = a4 + j;
                    const int sh = 8 * (tj & 3);
                    const uint32_t cls = off / K1_CLASS_STRIDE;
                    const uint32_t bit = (1u << cls) << sh;
                    const uint32_t old = atomicOr(&seen[c4 + (tj >> 2)], bit);
                    const uint32_t dcol = (uint32_t)(c4 + (int)(colb[j] >> 2));
                    if ((old & bit) && cls < 6u) atomicAdd(&sm.dupcc[0][0] + (cls % 3u) * LS_TILE + dcol, 1u << (16u * (cls / 3u)));
                    if ((old >> sh) & 255u) atomicAdd(&sm.dupnc[dcol], 1u);
                  }
                }
              }
            }
          } else {
            for (uint32_t g = (ya & ~3u) + 4u * (uint32_t)lane; g < yw; g += 128u) {
              const uint32_t w = *reinterpret_cast<const uint32_t *>(qual + g);
              const uint32_t h = *reinterpret_cast<const uint16_t *>(seq4 + (g >> 1));
              const int sg = sbase + (int)(g - ya);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t q = (w >> (8 * j)) & 255u;
                const bool ok = (int)q >= a.min_bq && (g + (uint32_t)j >= ya) && (g + (uint32_t)j < yw);
                const uint32_t code = (h >> (8 * (j >> 1) + ((j & 1) ? 0 : 4))) & 15u;
                const int cls = class_of_code(code);
                if (ok && cls != LS_CLASS_NA) {
                  if (COUNTED)
                    add_entry<PACKED, SEEN>(sm, seen, sg + j, cls, q, strand);
                  else
                    add_uncounted<PACKED>(sm, sg + j, cls);
                }
              }
            }
          }
          if (ind != 0 && ylast < ybl && lane == 0) {  // last base of the op, followed by an insertion / deletion
            const uint32_t q = qual[ylast];
            if ((int)q >= a.min_bq) {
              if (COUNTED)
                add_entry<PACKED, SEEN>(sm, seen, sbase + (int)(ylast - ya), indcls, q, strand);
              else
                add_uncounted<PACKED>(sm, sbase + (int)(ylast - ya), indcls);
            }
          }
          // query positions past the stored sequence (malformed record): base 'N', quality 0
          if (yb > lq && a.min_bq <= 0) {
            for (uint32_t qp = (ya > lq ? ya : lq) + (uint32_t)lane; qp < yb; qp += 32u) {
              const int cls = (qp == ylast) ? indcls : LS_CLASS_N;
              if (COUNTED)
                add_entry<PACKED, SEEN>(sm, seen, sbase + (int)(qp - ya), cls, 0u, strand);
              else
                add_uncounted<PACKED>(sm, sbase + (int)(qp - ya), cls);
            }
          }
        }
      }
    }
  }
}

template <bool PACKED>
__global__ void __launch_bounds__(K1_THREADS, 4) pileup_count_kernel(CountArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmemT<PACKED> &sm = *reinterpret_cast<TileSmemT<PACKED> *>(smem_raw);
  const uint32_t part = blockIdx.x;
  const uint32_t pslot = a.part_slot[part];
  if (pslot == 0xffffffffu) return;  // unused entry between the heavy (front) and light (back) parts
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t slot = pslot;
  const uint32_t pk = a.part_k[part];
  const uint32_t nparts = a.slot_nparts[slot];
  const int64_t tile = a.slot_tile[slot];
  const uint32_t slot_lo = a.slot_lo[slot], slot_hi = a.slot_lo[slot + 1];
  const uint32_t my_lo = slot_lo + pk * K1_PART_SEGS;
  const uint32_t my_hi = (my_lo + K1_PART_SEGS) < slot_hi ? (my_lo + K1_PART_SEGS) : slot_hi;
  const int64_t w = window_of_tile(a, tile);
  const int32_t tile_start = a.wstart[w] + (int32_t)(tile - a.wtile_base[w]) * LS_TILE;
  const int32_t tile_end = (tile_start + LS_TILE) < a.wend[w] ? (tile_start + LS_TILE) : a.wend[w];
  const uint64_t ref_base = a.wref_off[w] + (uint64_t)(tile_start - a.wstart[w]);

  {  // zero the accumulators, stage the reference bases
    uint32_t *z = reinterpret_cast<uint32_t *>(&sm);
    constexpr int NZ = ((PACKED ? K1_ROWS : 2 * K1_ROWS) + 3) * LS_TILE;  // hist + dupcc (dupnc / acx zeroed below)
    for (int i = threadIdx.x; i < NZ; i += K1_THREADS) z[i] = 0u;
    for (int i = threadIdx.x; i < LS_TILE; i += K1_THREADS) {
      sm.ref[i] = (tile_start + i < tile_end) ? upper_ascii(a.ref[ref_base + i]) : (uint8_t)'N';
      sm.dupnc[i] = 0u;
      sm.acx[i] = 0u;
    }
    if (threadIdx.x < 16)  // BAM nibble code -> byte offset of the class row pair; ignored codes -> dump rows (class 8)
      sm.lut2[threadIdx.x] = (uint32_t)class_of_code((uint32_t)threadIdx.x) * K1_CLASS_STRIDE;
    if (threadIdx.x == 0) {
      sm.next = 0;
      sm.npass = 0;
      sm.ticket = 0;
    }
  }
  __syncthreads();

  const uint64_t cmask = (1ull << a.cell_bits) - 1ull;
  const uint64_t unc = (uint64_t)a.uncounted_key;
  uint32_t *seen = sm.seen[warp];
  const uint32_t hist_s = (uint32_t)__cvta_generic_to_shared(&sm.hist[0][0]);
  const uint32_t nmine = my_hi - my_lo;
  for (;;) {
    // guided self-scheduling: grab ~1/(2*warps) of what is left (<= 32 segments, >= 2), so that the
    // last grabs are small and the warps of the CTA reach the barrier together
    uint32_t g = 0, chunk = 0;
    if (lane == 0) {
      const uint32_t seen_next = *(volatile uint32_t *)&sm.next;
      const uint32_t left = seen_next < nmine ? nmine - seen_next : 0u;
      chunk = left / (2u * K1_WARPS);
      chunk = chunk < 2u ? 2u : (chunk > 32u ? 32u : chunk);
      g = atomicAdd(&sm.next, chunk);
    }
    g = __shfl_sync(0xffffffffu, g, 0);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (g >= nmine) break;
    const uint32_t ca = my_lo + g;
    const uint32_t cb = (ca + chunk) < my_hi ? (ca + chunk) : my_hi;
    const int n = (int)(cb - ca);
    // lane j < n owns segment ca + j: key, "starts a run" flag, record + read metadata
    uint64_t ck = ~0ull;
    bool start = false;
    SegMeta mm = {};
    if (lane < n) {
      const uint32_t i = ca + (uint32_t)lane;
      ck = a.keys[i] & cmask;
      start = (i == slot_lo) || ck == unc || (a.keys[i - 1] & cmask) != ck;
      mm = load_meta(a, i);
      // lane-parallel warm-up for the 32 segments of the chunk: the first piece, and the first lines of the
      // qualities / bases it points at, are pulled towards the SM while earlier segments are being processed
      if (a.prefetch) {
        const uint64_t qb = mm.boff + (uint64_t)mm.y0;
        const uint8_t *q = a.qual + qb;
        const uint8_t *sq = a.seq4 + (qb >> 1);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(sq));
        if (a.prefetch & 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pieces + mm.p0));
        if (a.prefetch & 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pieces + mm.p0));
        if ((a.prefetch & 15) >= 2) {
          // the remaining lines of the segment's query bytes (typically ~200 qualities, ~100 base bytes)
          const uint32_t qe = ((uint32_t)(qb & 127u) + mm.qlen) >> 7;          // extra quality lines
          const uint32_t se = ((uint32_t)((qb >> 1) & 127u) + (mm.qlen >> 1)) >> 7;  // extra base lines
          for (uint32_t l = 1; l <= qe && l <= (uint32_t)(a.prefetch & 15); ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 128u * l));
          for (uint32_t l = 1; l <= se && l <= (uint32_t)(a.prefetch & 15); ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(sq + 128u * l));
        }
      }
    }
    const uint32_t startmask = __ballot_sync(0xffffffffu, start);
    uint32_t rem = startmask;
    while (rem) {
      const int j0 = __ffs(rem) - 1;
      rem &= rem - 1;
      const int j1 = rem ? (__ffs(rem) - 1) : n;
      const uint64_t rk = __shfl_sync(0xffffffffu, ck, j0);
      const bool counted = rk != unc;
      // a run that reaches the end of the chunk continues into the following segments of the tile
      uint32_t ext = 0;
      if (j1 == n && counted) {
        while (cb + ext < slot_hi && (a.keys[cb + ext] & cmask) == rk) ++ext;
      }
      const uint32_t runlen = (uint32_t)(j1 - j0) + ext;
      if (runlen == 1) {
        if (counted)
          process_segment_fast<PACKED, false, true>(a, sm, seen, shfl_meta(mm, j0), lane, hist_s);
        else
          process_segment_fast<PACKED, false, false>(a, sm, seen, shfl_meta(mm, j0), lane, hist_s);
      } else {
        for (int q = lane; q < LS_TILE / 4; q += 32) seen[q] = 0u;
        __syncwarp();
        for (int j = j0; j < j1; ++j)
          process_segment_fast<PACKED, true, true>(a, sm, seen, shfl_meta(mm, j), lane, hist_s);
        for (uint32_t e = 0; e < ext; ++e)
          process_segment_fast<PACKED, true, true>(a, sm, seen, load_meta(a, cb + e), lane, hist_s);
      }
    }
  }
  __syncthreads();

  // ---- site epilogue: gates of BaseCellCounter.py:211,220-222,282,294 -------------------
  uint32_t *out = a.out + (size_t)slot * LS_SITE_WORDS * LS_TILE;
  bool last_part = true;
  uint32_t nev = 0;
  for (int pass_no = 0; pass_no < 2; ++pass_no) {
    // pass 0: per-part words (direct store for single-part tiles, atomic merge otherwise)
    // pass 1: only for the last part of a multi-part tile: gates on the merged totals
    if (pass_no == 1) {
      if (nparts == 1) break;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) sm.ticket = atomicAdd(&a.slot_done[slot], 1u);
      __syncthreads();
      last_part = sm.ticket == nparts - 1;
      if (!last_part) break;
      __threadfence();
    }
    for (int s = threadIdx.x; s < LS_TILE; s += K1_THREADS) {
      const int sidx = swz(s);
      const uint8_t rb = sm.ref[s];
      uint32_t dp = 0, nc = 0, ac = 0;
      uint32_t f[8], r[8], bq[6], cc[6];
      if (pass_no == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (PACKED) {
            const uint32_t hf = sm.hist[k * 2][sidx], hr = sm.hist[k * 2 + 1][sidx];
            f[k] = hf >> K1_CNT_SHIFT;
            r[k] = hr >> K1_CNT_SHIFT;
            if (k < 6) bq[k] = (hf & ((1u << K1_CNT_SHIFT) - 1u)) + (hr & ((1u << K1_CNT_SHIFT) - 1u));
          } else {
            f[k] = sm.hist[k * 2][sidx];
            r[k] = sm.hist[k * 2 + 1][sidx];
            if (k < 6) bq[k] = sm.hist[K1_ROWS + k * 2][sidx] + sm.hist[K1_ROWS + k * 2 + 1][sidx];
          }
          dp += f[k] + r[k];
        }
        nev += dp;
        nc = dp - sm.dupnc[sidx];
#pragma unroll
        for (int k = 0; k < 6; ++k) cc[k] = f[k] + r[k] - ((sm.dupcc[k % 3][sidx] >> (16 * (k / 3))) & 0xffffu);
        if (a.min_ac > 0) {
          ac = sm.acx[sidx] + f[LS_CLASS_I] + r[LS_CLASS_I] + f[LS_CLASS_D] + r[LS_CLASS_D];
          const int base_cls[5] = {LS_CLASS_A, LS_CLASS_C, LS_CLASS_T, LS_CLASS_G, LS_CLASS_N};
#pragma unroll
          for (int j = 0; j < 5; ++j)
            if (class_letter(base_cls[j]) != rb) ac += f[base_cls[j]] + r[base_cls[j]];
        }
        if (nparts > 1) {
          if (dp) atomicAdd(&out[LS_SITE_DP * LS_TILE + s], dp);
          if (nc) atomicAdd(&out[LS_SITE_NC * LS_TILE + s], nc);
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            if (cc[k]) atomicAdd(&out[(LS_SITE_CC + k) * LS_TILE + s], cc[k]);
            if (f[k]) atomicAdd(&out[(LS_SITE_BCF + k) * LS_TILE + s], f[k]);
            if (r[k]) atomicAdd(&out[(LS_SITE_BCR + k) * LS_TILE + s], r[k]);
            if (bq[k]) atomicAdd(&out[(LS_SITE_BQ + k) * LS_TILE + s], bq[k]);
          }
          if (a.min_ac > 0 && ac) atomicAdd(&a.acbuf[(size_t)slot * LS_TILE + s], ac);
          continue;  // gates are applied in pass 1 by the last part
        }
      } else {
        dp = __ldcg(&out[LS_SITE_DP * LS_TILE + s]);
        nc = __ldcg(&out[LS_SITE_NC * LS_TILE + s]);
        if (a.min_ac > 0) ac = __ldcg(&a.acbuf[(size_t)slot * LS_TILE + s]);
      }
      bool pass = (tile_start + s < tile_end) && rb != 'N' && dp > 0 && (int)dp >= a.min_dp && (int)nc >= a.min_cc;
      if (pass && a.min_ac > 0) pass = (int)ac >= a.min_ac;
      const uint32_t bal = __ballot_sync(0xffffffffu, pass);
      if (lane == 0) {
        a.mask[(size_t)slot * (LS_TILE / 32) + (s >> 5)] = bal;
        if (bal) atomicAdd(&sm.npass, (uint32_t)__popc(bal));
      }
      if (pass && pass_no == 0) {
        out[LS_SITE_DP * LS_TILE + s] = dp;
        out[LS_SITE_NC * LS_TILE + s] = nc;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          out[(LS_SITE_CC + k) * LS_TILE + s] = cc[k];
          out[(LS_SITE_BCF + k) * LS_TILE + s] = f[k];
          out[(LS_SITE_BCR + k) * LS_TILE + s] = r[k];
          out[(LS_SITE_BQ + k) * LS_TILE + s] = bq[k];
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nev += __shfl_xor_sync(0xffffffffu, nev, o);
  if (lane == 0 && nev) atomicAdd(a.n_events, (unsigned long long)nev);
  __syncthreads();
  if (last_part && threadIdx.x == 0) a.npass[slot] = sm.npass;
}

// One CTA per slot: split the slot into parts, register them (part_slot[] pre-set to 0xffffffff), zero the HBM slot of
// multi-part tiles (their parts merge with atomics).
__global__ void __launch_bounds__(128) part_build_kernel(const uint32_t *__restrict__ slot_lo, int64_t n_slots,
                                                         uint32_t *__restrict__ part_slot, uint32_t *__restrict__ part_k,
                                                         uint32_t *__restrict__ slot_nparts, uint32_t *__restrict__ slot_done,
                                                         uint32_t *__restrict__ n_parts, uint32_t *__restrict__ n_light,
                                                         uint32_t max_parts, uint32_t *__restrict__ out,
                                                         uint32_t *__restrict__ acbuf) {
  const int64_t slot = blockIdx.x;
  if (slot >= n_slots) return;
  const uint32_t n = slot_lo[slot + 1] - slot_lo[slot];
  const uint32_t np = n == 0 ? 1u : (n + K1_PART_SEGS - 1) / K1_PART_SEGS;
  __shared__ uint32_t base;
  if (threadIdx.x == 0) {
    // deep tiles first: their parts sit at the front of the grid, shallow tiles fill in behind (no long tail)
    if (n >= (uint32_t)K1_PART_SEGS / 2)
      base = atomicAdd(n_parts, np);
    else
      base = max_parts - np - atomicAdd(n_light, np);
    slot_nparts[slot] = np;
    slot_done[slot] = 0;
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < np; k += blockDim.x) {
    part_slot[base + k] = (uint32_t)slot;
    part_k[base + k] = k;
  }
  if (np > 1) {
    uint32_t *o = out + (size_t)slot * LS_SITE_WORDS * LS_TILE;
    for (int i = threadIdx.x; i < LS_SITE_WORDS * LS_TILE; i += blockDim.x) o[i] = 0u;
    if (acbuf)
      for (int i = threadIdx.x; i < LS_TILE; i += blockDim.x) acbuf[(size_t)slot * LS_TILE + i] = 0u;
  }
}

// any same-(tile, cell) run longer than K1_MAX_RUN_PACKED segments? (then the 12-bit packed counters could overflow)
__global__ void __launch_bounds__(256) long_run_kernel(const uint64_t *__restrict__ keys, int64_t n, uint64_t cmask,
                                                       uint64_t unc, uint32_t *__restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K1_MAX_RUN_PACKED || i >= n) return;
  const uint64_t k = keys[i];
  if ((k & cmask) != unc && keys[i - K1_MAX_RUN_PACKED] == k) *flag = 1u;
}
