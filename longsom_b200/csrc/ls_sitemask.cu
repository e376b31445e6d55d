// K3: (chrom,pos) membership masks.
// Replaces step2.build_dict (BaseCellCalling.step2.py:197-221: {chrom:{pos//20000:set(pos)}})
// and the EDITING / PON_SR / PON_LR lookups of GetExtraFilters (:124-160).  The Python dict
// of sets becomes one sorted uint64 key table (tid<<32 | pos) in HBM: radix sort once,
// then one binary search per candidate site.
#include "ls_common.cuh"

__global__ void __launch_bounds__(256) mask_lookup_kernel(const uint64_t *__restrict__ keys, int64_t nk,
                                                          const uint64_t *__restrict__ query, int64_t m,
                                                          uint8_t *__restrict__ hit) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t q = query[i];
  int64_t lo = 0, hi = nk;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < q)
      lo = mid + 1;
    else
      hi = mid;
  }
  hit[i] = (lo < nk && keys[lo] == q) ? 1 : 0;
}

__global__ void __launch_bounds__(256) key_max_kernel(const uint64_t *__restrict__ keys, int64_t n,
                                                      unsigned long long *__restrict__ out) {
  uint64_t m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = keys[i] > m ? keys[i] : m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, (unsigned long long)m);
}

// Upload a key table into slot `t` and sort it there; it stays resident until the slot is loaded again.
static int site_table_load(ls_ctx *ctx, int t, const uint64_t *keys, int64_t n_keys, float *ms_sort, int *launches) {
  cudaStream_t st = ctx->stream;
  ctx->site_tab_ok[t] = false;
  ctx->site_tab_n[t] = 0;
  LS_CK(ctx->site_tab[t].ensure((size_t)(n_keys ? n_keys : 1) * 8));
  LS_CK(ctx->g_b.ensure((size_t)(n_keys ? n_keys : 1) * 8));
  LS_CK(ctx->g_e.ensure(64));
  if (n_keys) {
    LS_CK(cudaMemcpyAsync(ctx->site_tab[t].p, keys, (size_t)n_keys * 8, cudaMemcpyHostToDevice, st));
    // the number of radix passes follows the largest key, found on the device (8 bytes come back)
    unsigned long long *d_max = ctx->g_e.as<unsigned long long>();
    LS_CK(cudaMemsetAsync(d_max, 0, 8, st));
    const int64_t nb = (n_keys + 255) / 256;
    key_max_kernel<<<(unsigned)(nb < 4096 ? nb : 4096), 256, 0, st>>>(ctx->site_tab[t].as<uint64_t>(), n_keys, d_max);
    ++*launches;
    unsigned long long h_max = 0;
    LS_CK(cudaMemcpyAsync(&h_max, d_max, 8, cudaMemcpyDeviceToHost, st));
    LS_CK(cudaStreamSynchronize(st));
    uint64_t *sorted = ctx->site_tab[t].as<uint64_t>();
    LS_CK(cudaEventRecord(ctx->ev[0], st));
    LS_CK(ls_radix_sort_keys(ctx->site_tab[t].as<uint64_t>(), ctx->g_b.as<uint64_t>(), n_keys, ls_bits_for(h_max),
                             ctx->rs_hist, &sorted, ctx->num_sms, st, launches));
    if (sorted != ctx->site_tab[t].as<uint64_t>())
      LS_CK(cudaMemcpyAsync(ctx->site_tab[t].p, sorted, (size_t)n_keys * 8, cudaMemcpyDeviceToDevice, st));
    LS_CK(cudaEventRecord(ctx->ev[1], st));
    LS_CK(cudaStreamSynchronize(st));
    if (ms_sort) LS_CK(cudaEventElapsedTime(ms_sort, ctx->ev[0], ctx->ev[1]));
  }
  ctx->site_tab_n[t] = n_keys;
  ctx->site_tab_ok[t] = true;
  return LS_OK;
}

static int site_table_lookup(ls_ctx *ctx, int t, const uint64_t *query, int64_t m, uint8_t *hit, float *ms_lookup,
                             int *launches) {
  cudaStream_t st = ctx->stream;
  LS_CK(ctx->g_c.ensure((size_t)m * 8));
  LS_CK(ctx->g_d.ensure((size_t)m));
  LS_CK(cudaMemcpyAsync(ctx->g_c.p, query, (size_t)m * 8, cudaMemcpyHostToDevice, st));
  LS_CK(cudaEventRecord(ctx->ev[1], st));
  mask_lookup_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(ctx->site_tab[t].as<uint64_t>(), ctx->site_tab_n[t],
                                                                  ctx->g_c.as<uint64_t>(), m, ctx->g_d.as<uint8_t>());
  ++*launches;
  LS_CK(cudaGetLastError());
  LS_CK(cudaEventRecord(ctx->ev[2], st));
  LS_CK(cudaMemcpyAsync(hit, ctx->g_d.p, (size_t)m, cudaMemcpyDeviceToHost, st));
  LS_CK(cudaStreamSynchronize(st));
  if (ms_lookup) LS_CK(cudaEventElapsedTime(ms_lookup, ctx->ev[1], ctx->ev[2]));
  return LS_OK;
}

extern "C" int ls_site_table_load(ls_ctx *ctx, int table, const uint64_t *keys, int64_t n_keys, ls_run_stats *stats) {
  if (!ctx) return LS_E_ARG;
  if (table < 0 || table >= LS_SITE_TABLES || n_keys < 0 || (n_keys > 0 && !keys))
    LS_FAIL(LS_E_ARG, "ls_site_table_load: bad arguments");
  LS_CK(cudaSetDevice(ctx->device));
  ls_run_stats S;
  memset(&S, 0, sizeof S);
  int launches = 0;
  const int rc = site_table_load(ctx, table, keys, n_keys, &S.ms_sort, &launches);
  if (rc != LS_OK) return rc;
  S.ms_total = S.ms_sort;
  S.count_launches = launches;
  S.n_events = n_keys;
  if (stats) *stats = S;
  return LS_OK;
}

extern "C" int ls_site_table_lookup(ls_ctx *ctx, int table, const uint64_t *query, int64_t m, uint8_t *hit,
                                    ls_run_stats *stats) {
  if (!ctx) return LS_E_ARG;
  if (table < 0 || table >= LS_SITE_TABLES || m < 0 || (m > 0 && (!query || !hit)))
    LS_FAIL(LS_E_ARG, "ls_site_table_lookup: bad arguments");
  if (!ctx->site_tab_ok[table]) LS_FAIL(LS_E_STATE, "ls_site_table_lookup: no table loaded in this slot");
  LS_CK(cudaSetDevice(ctx->device));
  ls_run_stats S;
  memset(&S, 0, sizeof S);
  int launches = 0;
  if (m > 0) {
    const int rc = site_table_lookup(ctx, table, query, m, hit, &S.ms_count, &launches);
    if (rc != LS_OK) return rc;
  }
  S.ms_total = S.ms_count;
  S.count_launches = launches;
  S.n_events = m;
  if (stats) *stats = S;
  return LS_OK;
}

extern "C" int ls_site_mask(ls_ctx *ctx, const uint64_t *keys, int64_t n_keys, const uint64_t *query, int64_t m,
                            uint8_t *hit, ls_run_stats *stats) {
  if (!ctx) return LS_E_ARG;
  if (n_keys < 0 || m < 0 || (n_keys > 0 && !keys) || (m > 0 && (!query || !hit)))
    LS_FAIL(LS_E_ARG, "ls_site_mask: bad arguments");
  LS_CK(cudaSetDevice(ctx->device));
  ls_run_stats S;
  memset(&S, 0, sizeof S);
  if (m == 0) {
    if (stats) *stats = S;
    return LS_OK;
  }
  int launches = 0;
  int rc = site_table_load(ctx, LS_SITE_TABLES, keys, n_keys, &S.ms_sort, &launches);
  if (rc != LS_OK) return rc;
  rc = site_table_lookup(ctx, LS_SITE_TABLES, query, m, hit, &S.ms_count, &launches);
  if (rc != LS_OK) return rc;
  S.ms_total = S.ms_sort + S.ms_count;
  S.count_launches = launches;
  S.n_events = m;
  if (stats) *stats = S;
  return LS_OK;
}
