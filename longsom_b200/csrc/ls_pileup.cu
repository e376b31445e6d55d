// K1: per-barcode base pileup for BaseCellCounter.
//
// Reference behaviour restated (not ported) from
//   workflow/scripts/SNVCalling/BaseCellCounter.py:182-320 (run_interval),
//   :152-180 (EasyReadPileup) and the htslib pileup engine behind
//   pysam.AlignmentFile.pileup.
//
// Design (B200): the reference walks reference columns and, per column, a Python
// list of reads.  Here the unit of work is a TILE of LS_TILE reference positions
// owned by exactly one CTA, so all accumulation is shared-memory atomics and the
// result leaves the SM with plain coalesced stores (no global atomics, no
// zero-initialised global histograms).
//   1. seg_count / seg_fill : one thread per read walks its CIGAR once and emits one
//      Segment per (read, tile) pair that has at least one pileup entry.
//   2. radix sort of segments by (tile, cell)  -> same-cell segments of a tile are adjacent.
//   3. pileup_count_kernel : CTA per non-empty tile; a warp owns a run of same-cell
//      segments, lanes stride consecutive reference positions (coalesced seq4/qual
//      reads, conflict-free shared atomics); distinct-cell counts NC/CC are
//      "reads minus same-cell duplicates", the duplicates found with a per-warp
//      T-byte seen[] mask that is only touched for runs longer than one segment.
//   4. site epilogue applies the reference's gates and writes [slot][field][T] words.
#include "ls_common.cuh"

#ifndef LS_TILE
#define LS_TILE 512
#endif

struct SegArgs {
  int64_t n_reads;
  const int32_t *tid, *pos, *cell;
  const uint16_t *flag;
  const uint8_t *mapq;
  const uint32_t *cigar_off, *cigar;
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

// Walk one read; EMIT(tile, k, xk, yk) is called once per (read, tile) with entries.
template <typename EMIT>
__device__ __forceinline__ uint32_t walk_read(const SegArgs &a, int64_t r, uint64_t *aligned, int32_t *end_out,
                                              EMIT emit) {
  const uint32_t k0 = a.cigar_off[r], kend = a.cigar_off[r + 1];
  const int32_t tid = a.tid[r];
  const uint32_t flag = a.flag[r];
  int32_t x = a.pos[r];
  uint32_t y = 0;
  uint64_t al = 0;
  const bool engine_ok = read_passes_engine(flag, a.mapq[r], a.min_mq) && tid >= 0;
  const bool counted = a.cell[r] >= 0 && !(flag & LS_FLAG_SUPPL);
  const bool want = engine_ok && (counted || a.emit_uncounted);
  int64_t last_tile = -1;
  int64_t w = -1;
  uint32_t n = 0;
  for (uint32_t k = k0; k < kend; ++k) {
    const uint32_t c = a.cigar[k];
    const uint32_t op = c & 15u;
    const int32_t len = (int32_t)(c >> 4);
    if (op_is_match(op)) al += (uint64_t)len;
    if (want && len > 0) {
      int32_t xa = 0, xb = 0;
      if (op_is_match(op) || op == OP_D) {
        xa = x;
        xb = x + len;
      } else if (op == OP_N && indel_after(a.cigar, k, kend, op) != 0) {
        xa = x + len - 1;  // a ref-skip whose last column carries a following indel
        xb = x + len;
      }
      if (xb > xa) {
        if (w < 0) w = first_window(a, tid, xa);
        while (w < a.n_windows && a.wtid[w] == tid && a.wend[w] <= xa) ++w;
        int64_t ww = w;
        while (ww < a.n_windows && a.wtid[ww] == tid && a.wstart[ww] < xb) {
          int32_t lo = xa > a.wstart[ww] ? xa : a.wstart[ww];
          int32_t hi = xb < a.wend[ww] ? xb : a.wend[ww];
          if (lo < hi && !is_dropped(a, ww, (uint32_t)r)) {
            int64_t t0 = a.wtile_base[ww] + (lo - a.wstart[ww]) / LS_TILE;
            int64_t t1 = a.wtile_base[ww] + (hi - 1 - a.wstart[ww]) / LS_TILE;
            for (int64_t t = t0; t <= t1; ++t) {
              if (t != last_tile) {
                emit(n, t, k, x, y);
                ++n;
                last_tile = t;
              }
            }
          }
          if (a.wend[ww] >= xb) break;
          ++ww;
        }
      }
    }
    if (op_is_match(op)) {
      x += len;
      y += (uint32_t)len;
    } else if (op == OP_D || op == OP_N) {
      x += len;
    } else if (op == OP_I || op == OP_S) {
      y += (uint32_t)len;
    }
  }
  *aligned = al;
  *end_out = x;
  return n;
}

__global__ void __launch_bounds__(256) seg_count_kernel(SegArgs a, uint32_t *__restrict__ nseg,
                                                        unsigned long long *__restrict__ n_aligned,
                                                        int32_t *__restrict__ rend, uint32_t *__restrict__ wcount) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t al = 0;
  if (r < a.n_reads) {
    int32_t end;
    uint32_t n = walk_read(a, r, &al, &end, [](uint32_t, int64_t, uint32_t, int32_t, uint32_t) {});
    nseg[r] = n;
    rend[r] = end;
    // records each pileup() call (window) would fetch: overlap of [pos, bam_endpos) with the window
    const int32_t tid = a.tid[r], p0 = a.pos[r];
    if (wcount && tid >= 0 && read_passes_engine(a.flag[r], a.mapq[r], a.min_mq)) {
      const int32_t e = end > p0 ? end : p0 + 1;
      for (int64_t w = first_window(a, tid, p0); w < a.n_windows && a.wtid[w] == tid && a.wstart[w] < e; ++w)
        atomicAdd(&wcount[w], 1u);
    }
  }
  // block reduce of aligned bases
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) al += __shfl_xor_sync(0xffffffffu, al, o);
  if ((threadIdx.x & 31) == 0 && al) atomicAdd(n_aligned, (unsigned long long)al);
}

__global__ void __launch_bounds__(256) seg_fill_kernel(SegArgs a, const uint32_t *__restrict__ seg_off,
                                                       Segment *__restrict__ segs, uint64_t *__restrict__ keys) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n_reads) return;
  const uint32_t off = seg_off[r];
  const int32_t cell = a.cell[r];
  const bool counted = cell >= 0 && !(a.flag[r] & LS_FLAG_SUPPL);
  const uint64_t ck = counted ? (uint64_t)(uint32_t)cell : (uint64_t)a.uncounted_key;
  const int cb = a.cell_bits;
  uint64_t al;
  int32_t end;
  walk_read(a, r, &al, &end, [&](uint32_t i, int64_t t, uint32_t k, int32_t xk, uint32_t yk) {
    Segment s;
    s.read = (uint32_t)r;
    s.cig = k;
    s.x0 = xk;
    s.y0 = yk;
    segs[off + i] = s;
    keys[off + i] = ((uint64_t)t << cb) | ck;
  });
}

// tile boundaries in the sorted key array
__global__ void __launch_bounds__(256) tile_flag_kernel(const uint64_t *__restrict__ keys, int64_t n, int cell_bits,
                                                        uint32_t *__restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t t = keys[i] >> cell_bits;
  flag[i] = (i == 0 || (keys[i - 1] >> cell_bits) != t) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) slot_fill_kernel(const uint64_t *__restrict__ keys, int64_t n, int cell_bits,
                                                        const uint32_t *__restrict__ flag,
                                                        const uint32_t *__restrict__ rank, int64_t n_slots,
                                                        int64_t *__restrict__ slot_tile, uint32_t *__restrict__ slot_lo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) slot_lo[n_slots] = (uint32_t)n;
  if (i >= n) return;
  if (flag[i]) {
    uint32_t s = rank[i];
    slot_tile[s] = (int64_t)(keys[i] >> cell_bits);
    slot_lo[s] = (uint32_t)i;
  }
}

#include "ls_count.cuh"

// ---- compaction of passing sites (fetch path) -----------------------------------------------
struct CompactArgs {
  const int64_t *slot_tile;
  const uint32_t *slot_off;
  const uint32_t *out, *mask;
  int64_t n_windows;
  const int32_t *wtid, *wstart;
  const int64_t *wtile_base;
  const uint64_t *wref_off;
  const uint8_t *ref;
  int32_t *o_tid, *o_pos;
  uint8_t *o_ref;
  uint32_t *o_counts;
};

__global__ void __launch_bounds__(256) compact_sites_kernel(CompactArgs a) {
  __shared__ uint16_t list[LS_TILE];
  __shared__ uint32_t words[LS_TILE / 32];
  __shared__ uint32_t wpre[LS_TILE / 32 + 1];
  const int64_t slot = blockIdx.x;
  if (threadIdx.x < LS_TILE / 32) words[threadIdx.x] = a.mask[(size_t)slot * (LS_TILE / 32) + threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < LS_TILE / 32; ++i) {
      wpre[i] = run;
      run += __popc(words[i]);
    }
    wpre[LS_TILE / 32] = run;
  }
  __syncthreads();
  const uint32_t np = wpre[LS_TILE / 32];
  if (np == 0) return;
  const int64_t tile = a.slot_tile[slot];
  int64_t lo = 0, hi = a.n_windows;
  while (hi - lo > 1) {
    int64_t m = (lo + hi) >> 1;
    if (a.wtile_base[m] <= tile)
      lo = m;
    else
      hi = m;
  }
  const int64_t w = lo;
  const int32_t tile_start = a.wstart[w] + (int32_t)(tile - a.wtile_base[w]) * LS_TILE;
  const uint64_t ref_base = a.wref_off[w] + (uint64_t)(tile_start - a.wstart[w]);
  const uint32_t off = a.slot_off[slot];
  for (int s = threadIdx.x; s < LS_TILE; s += blockDim.x) {
    uint32_t wd = words[s >> 5];
    if ((wd >> (s & 31)) & 1u) {
      uint32_t rk = wpre[s >> 5] + __popc(wd & ((1u << (s & 31)) - 1u));
      list[rk] = (uint16_t)s;
      a.o_tid[off + rk] = a.wtid[w];
      a.o_pos[off + rk] = tile_start + s;
      a.o_ref[off + rk] = upper_ascii(a.ref[ref_base + s]);
    }
  }
  __syncthreads();
  const uint32_t *src = a.out + (size_t)slot * LS_SITE_WORDS * LS_TILE;
  uint32_t *dst = a.o_counts + (size_t)off * LS_SITE_WORDS;
  for (uint32_t e = threadIdx.x; e < np * LS_SITE_WORDS; e += blockDim.x) {
    uint32_t rk = e / LS_SITE_WORDS, f = e - rk * LS_SITE_WORDS;
    dst[e] = src[f * LS_TILE + list[rk]];
  }
}

// ---- host-side orchestration ------------------------------------------------------------------
static SegArgs make_seg_args(ls_ctx *ctx, const ls_count_params &p, bool emit_uncounted) {
  SegArgs a;
  a.n_reads = ctx->n_reads;
  a.tid = ctx->tid.as<int32_t>();
  a.pos = ctx->pos.as<int32_t>();
  a.cell = ctx->cell.as<int32_t>();
  a.flag = ctx->flag.as<uint16_t>();
  a.mapq = ctx->mapq.as<uint8_t>();
  a.cigar_off = ctx->cigar_off.as<uint32_t>();
  a.cigar = ctx->cigar.as<uint32_t>();
  a.n_windows = ctx->n_windows;
  a.wtid = ctx->wtid.as<int32_t>();
  a.wstart = ctx->wstart.as<int32_t>();
  a.wend = ctx->wend.as<int32_t>();
  a.wtile_base = ctx->wtile_base.as<int64_t>();
  a.min_mq = p.min_mq;
  a.cell_bits = ctx->cell_bits;
  a.emit_uncounted = emit_uncounted ? 1 : 0;
  a.uncounted_key = (uint32_t)(ctx->max_cell + 1);
  a.drop_keys = ctx->drop_keys.as<uint64_t>();
  a.n_drop = ctx->n_drop;
  return a;
}

int ls_tile_size(void) { return LS_TILE; }
extern "C" int ls_pileup_tile_size(void) { return LS_TILE; }

extern "C" int ls_pileup_run(ls_ctx *ctx, const ls_count_params *params, int64_t *n_sites, ls_run_stats *stats) {
  if (!ctx) return LS_E_ARG;
  if (!params) LS_FAIL(LS_E_ARG, "ls_pileup_run: params is null");
  if (!ctx->have_batch) LS_FAIL(LS_E_STATE, "ls_pileup_run: no batch uploaded");
  if (params->min_dp < 1) LS_FAIL(LS_E_ARG, "ls_pileup_run: min_dp must be >= 1");
  LS_CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  ctx->params = *params;
  ctx->have_run = false;
  ls_run_stats S;
  memset(&S, 0, sizeof S);
  int launches = 0;

  const int64_t n = ctx->n_reads;
  ctx->cell_bits = ls_bits_for((uint64_t)(ctx->max_cell + 1));
  const int tile_bits = ls_bits_for((uint64_t)(ctx->n_tiles_total > 0 ? ctx->n_tiles_total - 1 : 0));
  SegArgs sa = make_seg_args(ctx, *params, params->min_ac > 0);

  LS_CK(ctx->counters.ensure(64));
  LS_CK(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
  unsigned long long *d_aligned = ctx->counters.as<unsigned long long>();
  unsigned long long *d_events = d_aligned + 1;
  uint64_t *d_nseg_total = reinterpret_cast<uint64_t *>(d_aligned + 2);
  uint64_t *d_nslot_total = reinterpret_cast<uint64_t *>(d_aligned + 3);
  uint32_t *d_nparts = reinterpret_cast<uint32_t *>(d_aligned + 4);
  uint32_t *d_longrun = reinterpret_cast<uint32_t *>(d_aligned + 5);
  uint32_t *d_nlight = reinterpret_cast<uint32_t *>(d_aligned + 6);

  LS_CK(cudaEventRecord(ctx->ev[0], st));
  uint64_t h_tot[6] = {0, 0, 0, 0, 0, 0};
  if (n > 0 && ctx->n_windows > 0) {
    LS_CK(ctx->nseg.ensure((size_t)n * 4));
    LS_CK(ctx->seg_off.ensure((size_t)n * 4));
    LS_CK(ctx->rend.ensure((size_t)n * 4));
    LS_CK(ctx->wcount.ensure((size_t)ctx->n_windows * 4));
    ctx->n_drop = 0;
    sa.n_drop = 0;
    // the depth cap can only fire in a window that fetches more than max_depth records
    const bool cap_possible = params->max_depth > 0 && n > (int64_t)params->max_depth;
    if (cap_possible) LS_CK(cudaMemsetAsync(ctx->wcount.p, 0, (size_t)ctx->n_windows * 4, st));
    seg_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sa, ctx->nseg.as<uint32_t>(), d_aligned,
                                                                   ctx->rend.as<int32_t>(),
                                                                   cap_possible ? ctx->wcount.as<uint32_t>() : nullptr);
    ++launches;
    if (cap_possible) {
      std::vector<uint32_t> wc((size_t)ctx->n_windows);
      LS_CK(cudaMemcpyAsync(wc.data(), ctx->wcount.p, wc.size() * 4, cudaMemcpyDeviceToHost, st));
      LS_CK(cudaStreamSynchronize(st));
      bool any = false;
      for (uint32_t c : wc) any = any || (int64_t)c > (int64_t)params->max_depth;
      if (any) {
        int rc = ls_depth_cap_host(ctx, params->min_mq, params->max_depth, wc);
        if (rc != LS_OK) return rc;
        if (ctx->n_drop > 0) {  // redo the count with the dropped (window, read) pairs
          sa = make_seg_args(ctx, *params, params->min_ac > 0);
          LS_CK(cudaMemsetAsync(d_aligned, 0, 8, st));
          seg_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sa, ctx->nseg.as<uint32_t>(), d_aligned,
                                                                         ctx->rend.as<int32_t>(), nullptr);
          ++launches;
        }
      }
    }
    LS_CK(ls_scan_exclusive_u32(ctx->nseg.as<uint32_t>(), ctx->seg_off.as<uint32_t>(), n, d_nseg_total,
                                ctx->scan_tmp, st));
    launches += 3;
    LS_CK(cudaMemcpyAsync(h_tot, ctx->counters.p, 32, cudaMemcpyDeviceToHost, st));
    LS_CK(cudaStreamSynchronize(st));
  }
  const int64_t nseg = (int64_t)h_tot[2];
  if (nseg >= (int64_t)0xffffffffll) LS_FAIL(LS_E_ARG, "ls_pileup_run: more than 2^32 segments in one batch");
  S.n_aligned = (int64_t)h_tot[0];
  S.n_segments = nseg;
  ctx->n_segments = nseg;
  int64_t n_slots = 0;
  if (nseg > 0) {
    LS_CK(ctx->segs.ensure((size_t)nseg * sizeof(Segment)));
    LS_CK(ctx->keys_a.ensure((size_t)nseg * 8));
    LS_CK(ctx->keys_b.ensure((size_t)nseg * 8));
    LS_CK(ctx->vals_a.ensure((size_t)nseg * 4));
    LS_CK(ctx->vals_b.ensure((size_t)nseg * 4));
    seg_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sa, ctx->seg_off.as<uint32_t>(),
                                                                 ctx->segs.as<Segment>(), ctx->keys_a.as<uint64_t>());
    ++launches;
    LS_CK(cudaEventRecord(ctx->ev[1], st));
    LS_CK(ls_radix_sort_pairs(ctx->keys_a.as<uint64_t>(), ctx->keys_b.as<uint64_t>(), ctx->vals_a.as<uint32_t>(),
                              ctx->vals_b.as<uint32_t>(), nseg, tile_bits + ctx->cell_bits, ctx->rs_hist,
                              &ctx->sorted_keys, &ctx->sorted_vals, ctx->num_sms, st, &launches));
    // non-empty tiles -> slots
    LS_CK(ctx->tile_flag.ensure((size_t)nseg * 4));
    LS_CK(ctx->tile_rank.ensure((size_t)nseg * 4));
    tile_flag_kernel<<<(unsigned)((nseg + 255) / 256), 256, 0, st>>>(ctx->sorted_keys, nseg, ctx->cell_bits,
                                                                     ctx->tile_flag.as<uint32_t>());
    ++launches;
    LS_CK(ls_scan_exclusive_u32(ctx->tile_flag.as<uint32_t>(), ctx->tile_rank.as<uint32_t>(), nseg, d_nslot_total,
                                ctx->scan_tmp, st));
    launches += 3;
    {
      const uint64_t cmask = (1ull << ctx->cell_bits) - 1ull;
      long_run_kernel<<<(unsigned)((nseg + 255) / 256), 256, 0, st>>>(ctx->sorted_keys, nseg, cmask,
                                                                      (uint64_t)(uint32_t)(ctx->max_cell + 1), d_longrun);
      ++launches;
    }
    LS_CK(cudaMemcpyAsync(h_tot, ctx->counters.p, 48, cudaMemcpyDeviceToHost, st));
    LS_CK(cudaStreamSynchronize(st));
    n_slots = (int64_t)h_tot[3];
    const bool packed = (h_tot[5] & 0xffffffffull) == 0;
    LS_CK(ctx->slot_tile.ensure((size_t)n_slots * 8));
    LS_CK(ctx->slot_lo.ensure((size_t)(n_slots + 1) * 4));
    slot_fill_kernel<<<(unsigned)((nseg + 255) / 256), 256, 0, st>>>(
        ctx->sorted_keys, nseg, ctx->cell_bits, ctx->tile_flag.as<uint32_t>(), ctx->tile_rank.as<uint32_t>(), n_slots,
        ctx->slot_tile.as<int64_t>(), ctx->slot_lo.as<uint32_t>());
    ++launches;
    LS_CK(cudaEventRecord(ctx->ev[2], st));

    LS_CK(ctx->slot_out.ensure((size_t)n_slots * LS_SITE_WORDS * LS_TILE * 4));
    LS_CK(ctx->slot_mask.ensure((size_t)n_slots * (LS_TILE / 32) * 4));
    LS_CK(ctx->slot_npass.ensure((size_t)n_slots * 4));
    LS_CK(ctx->slot_off.ensure((size_t)n_slots * 4));
    const int64_t max_parts = n_slots + nseg / K1_PART_SEGS + 1;
    LS_CK(ctx->part_slot.ensure((size_t)max_parts * 4));
    LS_CK(ctx->part_k.ensure((size_t)max_parts * 4));
    LS_CK(ctx->slot_nparts.ensure((size_t)n_slots * 4));
    LS_CK(ctx->slot_done.ensure((size_t)n_slots * 4));
    if (params->min_ac > 0) LS_CK(ctx->acbuf.ensure((size_t)n_slots * LS_TILE * 4));
    LS_CK(cudaMemsetAsync(ctx->part_slot.p, 0xff, (size_t)max_parts * 4, st));
    part_build_kernel<<<(unsigned)n_slots, 128, 0, st>>>(
        ctx->slot_lo.as<uint32_t>(), n_slots, ctx->part_slot.as<uint32_t>(), ctx->part_k.as<uint32_t>(),
        ctx->slot_nparts.as<uint32_t>(), ctx->slot_done.as<uint32_t>(), d_nparts, d_nlight, (uint32_t)max_parts,
        ctx->slot_out.as<uint32_t>(),
        params->min_ac > 0 ? ctx->acbuf.as<uint32_t>() : nullptr);
    ++launches;
    CountArgs ca;
    ca.part_slot = ctx->part_slot.as<uint32_t>();
    ca.part_k = ctx->part_k.as<uint32_t>();
    ca.slot_nparts = ctx->slot_nparts.as<uint32_t>();
    ca.slot_done = ctx->slot_done.as<uint32_t>();
    ca.n_parts = d_nparts;
    ca.acbuf = params->min_ac > 0 ? ctx->acbuf.as<uint32_t>() : nullptr;
    ca.flag = ctx->flag.as<uint16_t>();
    ca.cigar_off = ctx->cigar_off.as<uint32_t>();
    ca.cigar = ctx->cigar.as<uint32_t>();
    ca.base_off = ctx->base_off.as<uint64_t>();
    ca.lq = ctx->lq.as<int32_t>();
    ca.seq4 = ctx->seq4.as<uint8_t>();
    ca.qual = ctx->qual.as<uint8_t>();
    ca.segs = ctx->segs.as<Segment>();
    ca.keys = ctx->sorted_keys;
    ca.vals = ctx->sorted_vals;
    ca.slot_tile = ctx->slot_tile.as<int64_t>();
    ca.slot_lo = ctx->slot_lo.as<uint32_t>();
    ca.n_windows = ctx->n_windows;
    ca.wstart = ctx->wstart.as<int32_t>();
    ca.wend = ctx->wend.as<int32_t>();
    ca.wtile_base = ctx->wtile_base.as<int64_t>();
    ca.wref_off = ctx->wref_off.as<uint64_t>();
    ca.ref = ctx->ref.as<uint8_t>();
    ca.out = ctx->slot_out.as<uint32_t>();
    ca.mask = ctx->slot_mask.as<uint32_t>();
    ca.npass = ctx->slot_npass.as<uint32_t>();
    ca.n_events = d_events;
    ca.cell_bits = ctx->cell_bits;
    ca.uncounted_key = (uint32_t)(ctx->max_cell + 1);
    ca.min_bq = params->min_bq;
    ca.min_dp = params->min_dp;
    ca.min_cc = params->min_cc;
    ca.min_ac = params->min_ac;
    if (!ctx->k1_attr_set) {
      LS_CK(cudaFuncSetAttribute(pileup_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(TileSmemT<true>)));
      LS_CK(cudaFuncSetAttribute(pileup_count_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(TileSmemT<false>)));
      ctx->k1_attr_set = true;
    }
    // 12-bit packed counters unless some cell has > K1_MAX_RUN_PACKED reads in one tile
    if (packed)
      pileup_count_kernel<true><<<(unsigned)max_parts, K1_THREADS, sizeof(TileSmemT<true>), st>>>(ca);
    else
      pileup_count_kernel<false><<<(unsigned)max_parts, K1_THREADS, sizeof(TileSmemT<false>), st>>>(ca);
    ++launches;
    LS_CK(cudaGetLastError());
    LS_CK(cudaEventRecord(ctx->ev[3], st));
    LS_CK(ls_scan_exclusive_u32(ctx->slot_npass.as<uint32_t>(), ctx->slot_off.as<uint32_t>(), n_slots, d_nslot_total,
                                ctx->scan_tmp, st));
    launches += 3;
    LS_CK(cudaEventRecord(ctx->ev[4], st));
    LS_CK(cudaMemcpyAsync(h_tot, ctx->counters.p, 32, cudaMemcpyDeviceToHost, st));
    LS_CK(cudaStreamSynchronize(st));
    ctx->n_sites = (int64_t)h_tot[3];
    S.n_events = (int64_t)h_tot[1];
    LS_CK(cudaEventElapsedTime(&S.ms_segments, ctx->ev[0], ctx->ev[1]));
    LS_CK(cudaEventElapsedTime(&S.ms_sort, ctx->ev[1], ctx->ev[2]));
    LS_CK(cudaEventElapsedTime(&S.ms_count, ctx->ev[2], ctx->ev[3]));
    LS_CK(cudaEventElapsedTime(&S.ms_total, ctx->ev[0], ctx->ev[4]));
  } else {
    ctx->n_sites = 0;
  }
  ctx->n_slots = n_slots;
  S.n_tiles = n_slots;
  S.count_launches = launches;
  ctx->stats = S;
  ctx->have_run = true;
  if (n_sites) *n_sites = ctx->n_sites;
  if (stats) *stats = S;
  return LS_OK;
}

extern "C" int ls_pileup_fetch(ls_ctx *ctx, ls_site_counts *out) {
  if (!ctx) return LS_E_ARG;
  if (!out) LS_FAIL(LS_E_ARG, "ls_pileup_fetch: out is null");
  if (!ctx->have_run) LS_FAIL(LS_E_STATE, "ls_pileup_fetch: ls_pileup_run has not completed");
  out->n_sites = ctx->n_sites;
  if (ctx->n_sites == 0) return LS_OK;
  if (out->capacity < ctx->n_sites) LS_FAIL(LS_E_CAPACITY, "ls_pileup_fetch: capacity < n_sites");
  if (!out->tid || !out->pos || !out->ref || !out->counts) LS_FAIL(LS_E_ARG, "ls_pileup_fetch: null output array");
  LS_CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t ns = ctx->n_sites;
  LS_CK(ctx->out_tid.ensure((size_t)ns * 4));
  LS_CK(ctx->out_pos.ensure((size_t)ns * 4));
  LS_CK(ctx->out_ref.ensure((size_t)ns));
  LS_CK(ctx->out_counts.ensure((size_t)ns * LS_SITE_WORDS * 4));
  CompactArgs a;
  a.slot_tile = ctx->slot_tile.as<int64_t>();
  a.slot_off = ctx->slot_off.as<uint32_t>();
  a.out = ctx->slot_out.as<uint32_t>();
  a.mask = ctx->slot_mask.as<uint32_t>();
  a.n_windows = ctx->n_windows;
  a.wtid = ctx->wtid.as<int32_t>();
  a.wstart = ctx->wstart.as<int32_t>();
  a.wtile_base = ctx->wtile_base.as<int64_t>();
  a.wref_off = ctx->wref_off.as<uint64_t>();
  a.ref = ctx->ref.as<uint8_t>();
  a.o_tid = ctx->out_tid.as<int32_t>();
  a.o_pos = ctx->out_pos.as<int32_t>();
  a.o_ref = ctx->out_ref.as<uint8_t>();
  a.o_counts = ctx->out_counts.as<uint32_t>();
  LS_CK(cudaEventRecord(ctx->ev[5], st));
  compact_sites_kernel<<<(unsigned)ctx->n_slots, 256, 0, st>>>(a);
  LS_CK(cudaGetLastError());
  LS_CK(cudaEventRecord(ctx->ev[6], st));
  LS_CK(cudaMemcpyAsync(out->tid, ctx->out_tid.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
  LS_CK(cudaMemcpyAsync(out->pos, ctx->out_pos.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
  LS_CK(cudaMemcpyAsync(out->ref, ctx->out_ref.p, (size_t)ns, cudaMemcpyDeviceToHost, st));
  LS_CK(cudaMemcpyAsync(out->counts, ctx->out_counts.p, (size_t)ns * LS_SITE_WORDS * 4, cudaMemcpyDeviceToHost, st));
  LS_CK(cudaStreamSynchronize(st));
  LS_CK(cudaEventElapsedTime(&ctx->stats.ms_compact, ctx->ev[5], ctx->ev[6]));
  return LS_OK;
}

extern "C" int ls_pileup_count(ls_ctx *ctx, const ls_read_batch *batch, const ls_windows *windows,
                               const ls_count_params *params, ls_site_counts *out, ls_run_stats *stats) {
  int rc = ls_pileup_upload(ctx, batch, windows);
  if (rc != LS_OK) return rc;
  int64_t ns = 0;
  rc = ls_pileup_run(ctx, params, &ns, stats);
  if (rc != LS_OK) return rc;
  rc = ls_pileup_fetch(ctx, out);
  if (stats) stats->ms_compact = ctx->stats.ms_compact;
  return rc;
}
