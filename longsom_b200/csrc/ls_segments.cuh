// Segment builder: one thread per read walks its CIGAR and emits, per (read, tile) pair with pileup
// entries, one Segment plus its Pieces (CIGAR ops clipped to the tile).  Included by ls_pileup.cu.
//
// One kernel: each thread walks its read once, counting segments / pieces and staging them in thread-local
// memory; a warp-level prefix sum and one atomicAdd pair per warp reserve the output ranges; the staged
// output is copied out (reads that overflow the staging area are walked a second time).  The block's CIGAR
// ops are first copied to shared memory with coalesced loads.
// Output order across warps follows the atomics and is not deterministic; the (tile, cell) sort
// and the commutative integer accumulation downstream make the results independent of it.
#pragma once

static_assert(LS_TILE == 512, "Piece::meta packs tile columns in 9 bits");

struct SegArgs {
  int64_t n_reads;
  const int32_t *tid, *pos, *cell;
  const uint16_t *flag;
  const uint8_t *mapq;
  const uint32_t *cigar_off, *cigar;
  const uint64_t *base_off;
  const int32_t *lq;
  int64_t n_windows;
  const int32_t *wtid, *wstart, *wend;
  const int64_t *wtile_base;
  int min_mq, cell_bits, emit_uncounted;
  uint32_t uncounted_key;
  const uint64_t *drop_keys;  // sorted (window<<32 | read) pairs removed by the depth cap
  int64_t n_drop;
};

__device__ __forceinline__ int64_t first_window(const SegArgs &a, int32_t tid, int32_t x) {
  // first window w with (wtid, wend) > (tid, x)
  int64_t lo = 0, hi = a.n_windows;
  while (lo < hi) {
    int64_t m = (lo + hi) >> 1;
    int32_t t = a.wtid[m];
    bool le = (t < tid) || (t == tid && a.wend[m] <= x);
    if (le)
      lo = m + 1;
    else
      hi = m;
  }
  return lo;
}

__device__ __forceinline__ bool is_dropped(const SegArgs &a, int64_t w, uint32_t r) {
  if (a.n_drop == 0) return false;
  uint64_t key = ((uint64_t)w << 32) | r;
  int64_t lo = 0, hi = a.n_drop;
  while (lo < hi) {
    int64_t m = (lo + hi) >> 1;
    if (a.drop_keys[m] < key)
      lo = m + 1;
    else
      hi = m;
  }
  return lo < a.n_drop && a.drop_keys[lo] == key;
}

struct WalkOut {
  uint32_t nseg, npiece, nunit;
  uint64_t aligned;
  int32_t end;
  int64_t wfirst;  // first_window(tid, pos) when the walk happened to look it up, else -1
};

struct SegSink {
  Segment *segs;
  uint64_t *keys;
  Piece *pieces;
  uint32_t seg_base, piece_base;
  uint64_t cell_key;
  int cell_bits;
};

// First-walk staging in thread-local memory: most reads fit, and then the second walk is a plain copy.
constexpr int SEG_STAGE_PIECES = 40;
constexpr int SEG_STAGE_SEGS = 12;
constexpr int SEG_CIG_SMEM = 6144;  // CIGAR ops of one block kept in shared memory (24 KB)
struct SegStage {
  Piece piece[SEG_STAGE_PIECES];
  uint64_t tile[SEG_STAGE_SEGS];
  uint32_t p0[SEG_STAGE_SEGS];  // first piece of the segment, relative to the read
  uint32_t nu[SEG_STAGE_SEGS];  // 32-base units of the segment's pieces
};

// A window cached in registers while the read stays inside it.
struct WinCur {
  int64_t w;
  int32_t ws, we;
  int64_t tb;
  bool same_tid, dropped;
};

__device__ __forceinline__ void load_window(const SegArgs &a, WinCur &c, int32_t tid, uint32_t r) {
  c.same_tid = c.w < a.n_windows && a.wtid[c.w] == tid;
  if (c.same_tid) {
    c.ws = a.wstart[c.w];
    c.we = a.wend[c.w];
    c.tb = a.wtile_base[c.w];
    c.dropped = is_dropped(a, c.w, r);
  }
}

// EMIT = false: count, and stage the output in `stage` while it fits; EMIT = true: write to the sink.
template <bool EMIT>
__device__ __forceinline__ WalkOut walk_read(const SegArgs &a, const uint32_t *__restrict__ cig, int64_t r,
                                             const SegSink &sink, SegStage *stage) {
  // cig: the CIGAR array, indexed like a.cigar (the kernel points it at its shared-memory copy when that fits)
  const uint32_t k0 = a.cigar_off[r], kend = a.cigar_off[r + 1];
  const int32_t tid = a.tid[r];
  const uint32_t flag = a.flag[r];
  const uint64_t boff = a.base_off[r];
  const uint32_t lq = (uint32_t)a.lq[r];
  int32_t x = a.pos[r];
  uint32_t y = 0;
  WalkOut o;
  o.nseg = 0;
  o.npiece = 0;
  o.nunit = 0;
  o.aligned = 0;
  o.wfirst = -1;
  const bool engine_ok = read_passes_engine(flag, a.mapq[r], a.min_mq) && tid >= 0;
  const bool counted = a.cell[r] >= 0 && !(flag & LS_FLAG_SUPPL);
  const bool want = engine_ok && (counted || a.emit_uncounted);
  WinCur cur;
  cur.w = -1;
  cur.same_tid = false;
  cur.dropped = false;
  cur.ws = cur.we = 0;
  cur.tb = 0;
  int64_t last_tile = -1;
  uint32_t seg_p0 = 0, seg_nu = 0;
  auto close_segment = [&]() {  // the segment that ends here: o.nseg - 1
    if (EMIT) {
      Segment s;
      s.p0 = sink.piece_base + seg_p0;
      s.np_nu = (o.npiece - seg_p0) | (seg_nu << 16);
      s.boff16 = (uint32_t)(boff >> 4);
      s.flags = (flag & LS_FLAG_REVERSE) ? 1u : 0u;
      sink.segs[sink.seg_base + o.nseg - 1] = s;
    } else if (o.nseg <= (uint32_t)SEG_STAGE_SEGS) {
      stage->nu[o.nseg - 1] = seg_nu;
    }
  };
  auto put_piece = [&](uint32_t ya, uint32_t col, uint32_t n, uint32_t del, uint32_t ind, uint32_t virt) {
    Piece p;
    p.ya = ya;
    p.meta = piece_meta(col, n, del, ind, virt);
    if (EMIT)
      sink.pieces[sink.piece_base + o.npiece] = p;
    else if (o.npiece < (uint32_t)SEG_STAGE_PIECES)
      stage->piece[o.npiece] = p;
    const uint32_t pu = piece_units(p.meta);
    seg_nu += pu;
    o.nunit += pu;
    ++o.npiece;
  };
  uint32_t cnext = k0 < kend ? cig[k0] : 0xfu;
  for (uint32_t k = k0; k < kend; ++k) {
    const uint32_t c = cnext;
    cnext = (k + 1 < kend) ? cig[k + 1] : 0xfu;
    const uint32_t op = c & 15u;
    const int32_t len = (int32_t)(c >> 4);
    const bool match = op_is_match(op);
    if (match) o.aligned += (uint64_t)len;
    if (want && len > 0 && (match || op == OP_D || op == OP_N)) {
      // class of the op's last column when an indel follows (htslib resolve_cigar2)
      const uint32_t op2 = cnext & 15u;
      int ind = 0;
      if (op2 == OP_D && op != OP_D)
        ind = -1;
      else if (op2 == OP_I)
        ind = 1;
      else if (op2 == OP_P)
        ind = indel_after(cig, k, kend, op);
      const uint32_t indcode = ind > 0 ? 1u : (ind < 0 ? 2u : 0u);
      int32_t xa = x, xb = x + len;
      if (op == OP_N) {  // a ref-skip only matters through its last column, and only if an indel follows
        xa = indcode ? x + len - 1 : xb;
      }
      if (xb > xa) {
        if (cur.w < 0) {
          cur.w = first_window(a, tid, xa);
          if (xa == a.pos[r]) o.wfirst = cur.w;
          load_window(a, cur, tid, (uint32_t)r);
        }
        while (cur.same_tid && cur.we <= xa) {
          ++cur.w;
          load_window(a, cur, tid, (uint32_t)r);
        }
        WinCur ww = cur;  // an op that crosses a window border visits the following windows too
        while (ww.same_tid && ww.ws < xb) {
          int32_t lo = xa > ww.ws ? xa : ww.ws;
          const int32_t hi = xb < ww.we ? xb : ww.we;
          if (lo < hi && !ww.dropped) {
            while (lo < hi) {
              const uint32_t trel = (uint32_t)(lo - ww.ws) / (uint32_t)LS_TILE;
              const int32_t tstart = ww.ws + (int32_t)(trel * (uint32_t)LS_TILE);
              const int32_t tend = (tstart + LS_TILE) < ww.we ? (tstart + LS_TILE) : ww.we;
              const int32_t shi = hi < tend ? hi : tend;
              const int64_t tile = ww.tb + (int64_t)trel;
              if (tile != last_tile) {
                if (o.nseg > 0) close_segment();
                if (EMIT) {
                  sink.keys[sink.seg_base + o.nseg] = ((uint64_t)tile << sink.cell_bits) | sink.cell_key;
                } else if (o.nseg < (uint32_t)SEG_STAGE_SEGS) {
                  stage->tile[o.nseg] = (uint64_t)tile;
                  stage->p0[o.nseg] = o.npiece;
                }
                seg_p0 = o.npiece;
                seg_nu = 0;
                ++o.nseg;
                last_tile = tile;
              }
              const uint32_t col = (uint32_t)(lo - tstart), n = (uint32_t)(shi - lo);
              const uint32_t pind = (shi == x + len) ? indcode : 0u;
              if (match) {
                // query bases [ya, ya + n); positions at or past l_qseq do not exist (malformed record): the
                // pileup engine reports them with quality 0 and base 'N' -> a separate "virtual" piece
                const uint32_t ya = y + (uint32_t)(lo - x);
                if (ya + n <= lq) {
                  put_piece(ya, col, n, 0u, pind, 0u);
                } else if (ya >= lq) {
                  put_piece(ya, col, n, 0u, pind, 1u);
                } else {
                  put_piece(ya, col, lq - ya, 0u, 0u, 0u);
                  put_piece(lq, col + (lq - ya), n - (lq - ya), 0u, pind, 1u);
                }
              } else {
                put_piece(y, col, n, 1u, pind, y >= lq ? 1u : 0u);
              }
              lo = shi;
            }
          }
          if (ww.we >= xb) break;
          ++ww.w;
          load_window(a, ww, tid, (uint32_t)r);
        }
      }
    }
    if (match) {
      x += len;
      y += (uint32_t)len;
    } else if (op == OP_D || op == OP_N) {
      x += len;
    } else if (op == OP_I || op == OP_S) {
      y += (uint32_t)len;
    }
  }
  if (o.nseg > 0) close_segment();
  o.end = x;
  return o;
}

// totals[0] = segments, totals[1] = pieces (both keep counting past the capacities, so that the host can
// size the buffers exactly and relaunch), totals[2] = 32-base units the pieces expand to; nothing is written by a
// warp whose range does not fit.
__global__ void __launch_bounds__(256) seg_build_kernel(SegArgs a, Segment *__restrict__ segs, uint64_t *__restrict__ keys,
                                                        Piece *__restrict__ pieces, uint64_t seg_cap, uint64_t piece_cap,
                                                        unsigned long long *__restrict__ totals,
                                                        unsigned long long *__restrict__ n_aligned,
                                                        int32_t *__restrict__ rend, uint32_t *__restrict__ wcount,
                                                        uint32_t max_depth, uint32_t *__restrict__ cap_flag) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = r < a.n_reads;
  // The CIGARs of the block's 256 consecutive reads are one contiguous range: copy it to shared memory with
  // coalesced loads, so that the per-thread serial walks do not each pay global-memory latency per op.
  __shared__ uint32_t scig[SEG_CIG_SMEM];
  const uint32_t *cig = a.cigar;
  {
    const int64_t r0 = (int64_t)blockIdx.x * blockDim.x;
    const int64_t r1 = (r0 + (int64_t)blockDim.x) < a.n_reads ? (r0 + (int64_t)blockDim.x) : a.n_reads;
    const uint32_t c0 = a.cigar_off[r0], c1 = a.cigar_off[r1];
    if (c1 - c0 <= (uint32_t)SEG_CIG_SMEM) {
      for (uint32_t i = threadIdx.x; i < c1 - c0; i += blockDim.x) scig[i] = a.cigar[c0 + i];
      cig = scig - c0;
    }
    __syncthreads();
  }
  SegSink sink;
  sink.segs = segs;
  sink.keys = keys;
  sink.pieces = pieces;
  sink.seg_base = sink.piece_base = 0;
  sink.cell_key = 0;
  sink.cell_bits = a.cell_bits;
  WalkOut o;
  o.nseg = o.npiece = o.nunit = 0;
  o.aligned = 0;
  o.end = 0;
  SegStage stage;
  if (live) {
    o = walk_read<false>(a, cig, r, sink, &stage);
    if (rend) rend[r] = o.end;
  }
  if (wcount) {
    // records each pileup() call (window) would fetch: overlap of [pos, bam_endpos) with the window.  Reads are
    // position-sorted, so the lanes of a warp mostly hit the same window: one atomic per distinct first window;
    // cap_flag is raised by the add that takes a window past max_depth (only then does the host look at wcount).
    int64_t w0 = -1;
    int32_t tid = -1, e = 0;
    if (live) {
      tid = a.tid[r];
      const int32_t p0 = a.pos[r];
      if (tid >= 0 && read_passes_engine(a.flag[r], a.mapq[r], a.min_mq)) {
        e = o.end > p0 ? o.end : p0 + 1;
        w0 = o.wfirst >= 0 ? o.wfirst : first_window(a, tid, p0);
        if (!(w0 < a.n_windows && a.wtid[w0] == tid && a.wstart[w0] < e)) w0 = -1;
      }
    }
    const long long key = w0 >= 0 ? (long long)w0 : -1ll - (long long)lane;
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (w0 >= 0) {
      if (lane == __ffs(peers) - 1) {
        const uint32_t add = (uint32_t)__popc(peers);
        const uint32_t old = atomicAdd(&wcount[w0], add);
        if (old <= max_depth && old + add > max_depth) *cap_flag = 1u;
      }
      for (int64_t w = w0 + 1; w < a.n_windows && a.wtid[w] == tid && a.wstart[w] < e; ++w) {
        const uint32_t old = atomicAdd(&wcount[w], 1u);
        if (old == max_depth) *cap_flag = 1u;
      }
    }
  }
  // warp prefix sums of the two counts, one reservation per warp
  uint32_t ps = o.nseg, pp = o.npiece;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t ts = __shfl_up_sync(0xffffffffu, ps, d);
    const uint32_t tp = __shfl_up_sync(0xffffffffu, pp, d);
    if (lane >= d) {
      ps += ts;
      pp += tp;
    }
  }
  const uint32_t wseg = __shfl_sync(0xffffffffu, ps, 31), wpiece = __shfl_sync(0xffffffffu, pp, 31);
  unsigned long long bs = 0, bp = 0;
  if (lane == 31 && wseg) {
    bs = atomicAdd(&totals[0], (unsigned long long)wseg);
    bp = atomicAdd(&totals[1], (unsigned long long)wpiece);
  }
  bs = __shfl_sync(0xffffffffu, bs, 31);
  bp = __shfl_sync(0xffffffffu, bp, 31);
  if (live && o.nseg && bs + wseg <= seg_cap && bp + wpiece <= piece_cap) {
    sink.seg_base = (uint32_t)bs + (ps - o.nseg);
    sink.piece_base = (uint32_t)bp + (pp - o.npiece);
    const int32_t cell = a.cell[r];
    const bool counted = cell >= 0 && !(a.flag[r] & LS_FLAG_SUPPL);
    sink.cell_key = counted ? (uint64_t)(uint32_t)cell : (uint64_t)a.uncounted_key;
    if (o.nseg <= (uint32_t)SEG_STAGE_SEGS && o.npiece <= (uint32_t)SEG_STAGE_PIECES) {
      for (uint32_t i = 0; i < o.npiece; ++i) pieces[sink.piece_base + i] = stage.piece[i];
      const uint32_t boff16 = (uint32_t)(a.base_off[r] >> 4);
      const uint32_t sflags = (a.flag[r] & LS_FLAG_REVERSE) ? 1u : 0u;
      for (uint32_t i = 0; i < o.nseg; ++i) {
        Segment sg;
        sg.p0 = sink.piece_base + stage.p0[i];
        sg.np_nu = ((i + 1 < o.nseg ? stage.p0[i + 1] : o.npiece) - stage.p0[i]) | (stage.nu[i] << 16);
        sg.boff16 = boff16;
        sg.flags = sflags;
        segs[sink.seg_base + i] = sg;
        keys[sink.seg_base + i] = (stage.tile[i] << a.cell_bits) | sink.cell_key;
      }
    } else {
      walk_read<true>(a, cig, r, sink, nullptr);
    }
  }
  uint64_t al = o.aligned;
  uint32_t nu = o.nunit;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    al += __shfl_xor_sync(0xffffffffu, al, d);
    nu += __shfl_xor_sync(0xffffffffu, nu, d);
  }
  if (lane == 0 && al) atomicAdd(n_aligned, (unsigned long long)al);
  if (lane == 0 && nu) atomicAdd(&totals[2], (unsigned long long)nu);
}
