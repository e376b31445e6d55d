// Micro-benchmark 2: shared-memory atomic throughput on sm_100a as a function of the ADDRESS PATTERN and of the
// number of active lanes (informs the K1 design: lane-per-unit scatters 32 lanes over 32 unrelated addresses).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_atoms2 ubench_atoms2.cu && ./ubench_atoms2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void red(uint32_t addr, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ void red64(uint32_t addr, unsigned long long v) { asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(addr), "l"(v)); }
__device__ __forceinline__ uint32_t atom_or(uint32_t addr, uint32_t v) { uint32_t o; asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v)); return o; }

// MODE 0: consecutive words (lane l -> word base+l), base random per warp-op
// MODE 1: distinct banks, random rows: word = row(l,i)*512 + ((l + i) & 511) with row random per lane
// MODE 2: random bank, random row (fully scattered)
// MODE 3: all lanes one address
// MODE 4: MODE 2 with returning atomic (atom.or)
// MODE 5: MODE 2 done as plain lds + add + sts (racy; throughput only)
// MODE 6: 4 lanes share a bank (systematic 4-way), others distinct
// MODE 7: MODE 1 with value 0 (does a zero add cost the same?)
// MODE 8: 64-bit red, fully scattered 8-byte words; MODE 9: 64-bit red, consecutive 8-byte words
// MODE 10: MODE 2 + one 16-entry table lookup (ld.shared) per op, as in the K1 inner loop
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, int iters, uint32_t lanes_mask) {
  extern __shared__ uint32_t h[];
  for (int i = threadIdx.x; i < 16384; i += 256) h[i] = 0;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(h);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = threadIdx.x * 2654435761u + 12345u, acc = 0;
  const bool active = (lanes_mask >> lane) & 1u;
  if (active) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        x = x * 1664525u + 1013904223u;
        const uint32_t r = x >> 8;
        uint32_t w;
        if (MODE == 0) w = ((uint32_t)(i * 8 + u) * 97u + warp * 1031u) * 32u + lane;
        else if (MODE == 1 || MODE == 7) w = (r & 15u) * 512u + ((lane + (uint32_t)(i * 8 + u)) & 511u);
        else if (MODE == 2 || MODE == 4 || MODE == 5) w = r;
        else if (MODE == 3) w = (uint32_t)(i * 8 + u) * 33u;
        else w = (r & 15u) * 512u + ((((lane >> 2) << 2) + (uint32_t)(i * 8 + u)) & 31u) + 32u * (lane & 3u);  // MODE 6
        if (MODE == 8) { red64(base + ((r & 4095u) << 3), ((unsigned long long)(r & 3u) << 32) | (r & 63u)); continue; }
        if (MODE == 9) { red64(base + (((((uint32_t)(i * 8 + u) * 97u + warp * 1031u) * 32u + lane) & 4095u) << 3), 1ull << 20); continue; }
        if (MODE == 10) { uint32_t t; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(base + 32768u + ((r >> 13) & 0x3cu))); red(base + (((r + t) & 8191u) << 2), 1u); continue; }
        const uint32_t a = base + ((w & 8191u) << 2);
        if (MODE == 4) acc += atom_or(a, 1u << (r & 7u));
        else if (MODE == 5) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v + 1u)); }
        else red(a, MODE == 7 ? 0u : ((1u << 20) | (r & 63u)));
      }
    }
  }
  __syncthreads();
  uint32_t s = acc;
  for (int i = threadIdx.x; i < 8192; i += 256) s += h[i];
  if (s == 0xdeadbeefu) out[0] = s;
}

template <int MODE>
void run(const char *name, int blocks_per_sm, uint32_t mask) {
  const int sms = 148, iters = 4000;
  uint32_t *d;
  cudaMalloc(&d, 4);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<sms * blocks_per_sm, 256, 65536>>>(d, 50, mask);
  cudaEventRecord(e0);
  k<MODE><<<sms * blocks_per_sm, 256, 65536>>>(d, iters, mask);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const int nl = __builtin_popcount(mask);
  const double warp_ops = (double)sms * blocks_per_sm * 8 * iters * 8;
  printf("%-34s CTAs/SM=%d lanes=%2d: %7.2f ms  %6.2f lane-ops/clk/SM  %5.3f warp-ops/clk/SM\n", name, blocks_per_sm, nl, ms,
         warp_ops * nl / (ms * 1e-3) / sms / 1.965e9, warp_ops / (ms * 1e-3) / sms / 1.965e9);
  cudaFree(d);
}

int main() {
  for (int b = 1; b <= 3; b += 2) {
    run<0>("red consecutive", b, 0xffffffffu);
    run<1>("red distinct banks, random rows", b, 0xffffffffu);
    run<7>("red distinct banks, value 0", b, 0xffffffffu);
    run<2>("red fully scattered", b, 0xffffffffu);
    run<2>("red fully scattered", b, 0x0000ffffu);
    run<2>("red fully scattered", b, 0x000000ffu);
    run<2>("red fully scattered", b, 0x55555555u);
    run<1>("red distinct banks", b, 0x0000ffffu);
    run<1>("red distinct banks", b, 0x000000ffu);
    run<3>("red one address", b, 0xffffffffu);
    run<6>("red 4-way systematic", b, 0xffffffffu);
    run<4>("atom.or scattered", b, 0xffffffffu);
    run<5>("lds+add+sts scattered", b, 0xffffffffu);
    run<8>("red.u64 scattered", b, 0xffffffffu);
    run<9>("red.u64 consecutive", b, 0xffffffffu);
    run<10>("lds table + red scattered", b, 0xffffffffu);
  }
  return 0;
}
