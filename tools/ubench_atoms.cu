// Micro-benchmark: shared-memory atomic / RMW throughput on sm_100a (informs the K1 design).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned *out, int iters, int spread) {
  extern __shared__ unsigned h[];
  for (int i = threadIdx.x; i < 16384; i += 256) h[i] = 0;
  __syncthreads();
  unsigned tid = threadIdx.x, acc = 0;
  unsigned a = tid;
  for (int i = 0; i < iters; ++i) {
    unsigned idx = (a + i * spread) & 16383u;
    if (MODE == 0) atomicAdd(&h[idx], 1u);                       // 1 atomic, conflict-free within warp
    if (MODE == 1) { atomicAdd(&h[idx], 1u); atomicAdd(&h[(idx + 8192) & 16383u], 3u); }
    if (MODE == 2) { unsigned v = h[idx]; h[idx] = v + 1; }       // plain RMW (racy across warps; throughput only)
    if (MODE == 3) acc += atomicAdd(&h[idx], 1u);                 // atomic with return
    if (MODE == 4) atomicAdd(&h[(idx * 17u) & 16383u], 1u);       // random-ish bank pattern
  }
  __syncthreads();
  unsigned s = acc;
  for (int i = threadIdx.x; i < 16384; i += 256) s += h[i];
  if (s == 0xdeadbeef) out[0] = s;
}
template <int MODE> void run(const char *name, int blocks_per_sm, int spread) {
  int sms = 148, iters = 20000;
  unsigned *d; cudaMalloc(&d, 4);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * blocks_per_sm, 256, 65536>>>(d, 100, spread);
  cudaEventRecord(e0);
  k<MODE><<<sms * blocks_per_sm, 256, 65536>>>(d, iters, spread);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)sms * blocks_per_sm * 256 * iters * (MODE == 1 ? 2 : 1);
  printf("%-28s blocks/SM=%d spread=%d: %.2f ms, %.1f Gop/s, %.2f lane-ops/clk/SM @1.965GHz\n", name, blocks_per_sm, spread, ms,
         ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  for (int b = 1; b <= 3; ++b) {
    run<0>("atoms.add x1", b, 32); run<1>("atoms.add x2", b, 32); run<2>("lds+sts rmw", b, 32);
    run<3>("atoms.add ret", b, 32); run<4>("atoms.add scattered", b, 32);
  }
  return 0;
}
