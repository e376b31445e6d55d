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

extern "C" int ls_site_mask(ls_ctx *ctx, const uint64_t *keys, int64_t n_keys, const uint64_t *query, int64_t m,
                            uint8_t *hit, ls_run_stats *stats) {
  if (!ctx) return LS_E_ARG;
  if (n_keys < 0 || m < 0 || (n_keys > 0 && !keys) || (m > 0 && (!query || !hit)))
    LS_FAIL(LS_E_ARG, "ls_site_mask: bad arguments");
  LS_CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  ls_run_stats S;
  memset(&S, 0, sizeof S);
  if (m == 0) {
    if (stats) *stats = S;
    return LS_OK;
  }
  LS_CK(ctx->g_a.ensure((size_t)(n_keys ? n_keys : 1) * 8));
  LS_CK(ctx->g_b.ensure((size_t)(n_keys ? n_keys : 1) * 8));
  LS_CK(ctx->g_c.ensure((size_t)m * 8));
  LS_CK(ctx->g_d.ensure((size_t)m));
  uint64_t maxkey = 0;
  for (int64_t i = 0; i < n_keys; ++i) maxkey = keys[i] > maxkey ? keys[i] : maxkey;
  if (n_keys) LS_CK(cudaMemcpyAsync(ctx->g_a.p, keys, (size_t)n_keys * 8, cudaMemcpyHostToDevice, st));
  LS_CK(cudaMemcpyAsync(ctx->g_c.p, query, (size_t)m * 8, cudaMemcpyHostToDevice, st));
  int launches = 0;
  uint64_t *sorted = ctx->g_a.as<uint64_t>();
  LS_CK(cudaEventRecord(ctx->ev[0], st));
  LS_CK(ls_radix_sort_keys(ctx->g_a.as<uint64_t>(), ctx->g_b.as<uint64_t>(), n_keys, ls_bits_for(maxkey),
                           ctx->rs_hist, &sorted, ctx->num_sms, st, &launches));
  LS_CK(cudaEventRecord(ctx->ev[1], st));
  mask_lookup_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(sorted, n_keys, ctx->g_c.as<uint64_t>(), m,
                                                                  ctx->g_d.as<uint8_t>());
  ++launches;
  LS_CK(cudaGetLastError());
  LS_CK(cudaEventRecord(ctx->ev[2], st));
  LS_CK(cudaMemcpyAsync(hit, ctx->g_d.p, (size_t)m, cudaMemcpyDeviceToHost, st));
  LS_CK(cudaStreamSynchronize(st));
  LS_CK(cudaEventElapsedTime(&S.ms_sort, ctx->ev[0], ctx->ev[1]));
  LS_CK(cudaEventElapsedTime(&S.ms_count, ctx->ev[1], ctx->ev[2]));
  S.ms_total = S.ms_sort + S.ms_count;
  S.count_launches = launches;
  S.n_events = m;
  if (stats) *stats = S;
  return LS_OK;
}
