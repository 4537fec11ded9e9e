// microbench.cu -- integer-pipe calibration for the IMAD roofline: measured issue rates of
// IMAD / IMAD.HI / IMAD.WIDE / IADD3 on this GPU and the throughput of the engine's Fq Montgomery product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench.bin tools/microbench.cu
// Output: one JSON line (ops per clock per SM, and Gops/s at the observed clock).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define BN254_INLINE_MUL 1  // calibration measures the inlined product
#include "../bn254_b200/csrc/tower.cuh"
using namespace bn;

#define ITERS 4096
#define NACC 8

template <int MODE>
__global__ void __launch_bounds__(256) k_rate(uint32_t* out, uint32_t seed, long long* cycles) {
  uint32_t a[NACC], b = seed | 1, c = seed * 7 + 3;
  uint64_t w[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { a[i] = threadIdx.x + i * 977 + seed; w[i] = a[i]; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      else if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      else if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
      else if (MODE == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      else if (MODE == 4) {  // 1 wide + 1 add (different pipes)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      } else if (MODE == 5) {  // wide with carry chain: lo/hi pair fused by ptxas into IMAD.WIDE.U32.X
        uint32_t lo = (uint32_t)w[i], hi = (uint32_t)(w[i] >> 32);
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(b), "r"(c));
        w[i] = ((uint64_t)hi << 32) | lo;
      }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) acc ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// chained Fq products: NCH independent chains per thread
template <int NCH>
__global__ void __launch_bounds__(256) k_fqmul(uint32_t* out, uint32_t seed, int iters, long long* cycles) {
  fq x[NCH], y;
  for (int j = 0; j < NCH; j++)
    for (int i = 0; i < 8; i++) x[j].l[i] = (threadIdx.x * 2654435761u + i * 40503u + j + seed) & (i == 7 ? 0x0fffffffu : 0xffffffffu);
  for (int i = 0; i < 8; i++) y.l[i] = (seed * 97 + i * 7919u + blockIdx.x) & (i == 7 ? 0x0fffffffu : 0xffffffffu);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < NCH; j++) x[j] = fq_mul(x[j], y);
  }
  long long t1 = clock64();
  uint32_t acc = 0;
  for (int j = 0; j < NCH; j++)
    for (int i = 0; i < 8; i++) acc ^= x[j].l[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
// Fq2 products through memory (the engine's out-of-line routine), operands in local memory
__global__ void __launch_bounds__(256) k_fq2mul(uint32_t* out, uint32_t seed, int iters) {
  fq2 x, y;
  for (int i = 0; i < 8; i++) {
    x.c0.l[i] = (threadIdx.x * 2654435761u + i * 40503u + seed) & (i == 7 ? 0x0fffffffu : 0xffffffffu);
    x.c1.l[i] = (threadIdx.x * 40503u + i * 2654435761u + seed) & (i == 7 ? 0x0fffffffu : 0xffffffffu);
    y.c0.l[i] = (seed * 97 + i * 7919u + blockIdx.x) & (i == 7 ? 0x0fffffffu : 0xffffffffu);
    y.c1.l[i] = (seed * 31 + i * 104729u + blockIdx.x) & (i == 7 ? 0x0fffffffu : 0xffffffffu);
  }
  for (int it = 0; it < iters; it++) fq2_mul(&x, &x, &y);
  uint32_t acc = 0;
  for (int i = 0; i < 8; i++) acc ^= x.c0.l[i] ^ x.c1.l[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class L>
static double time_ms(L launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5.0;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, (size_t)sms * 16 * 256 * 4); cudaMalloc(&cyc, 8);
  // ("imad_wide": ptxas hoists the loop-invariant product of this pattern and emits IADD3 pairs, so the figure is an add rate;
  //  "wide_x_carry" is the one that issues one IMAD.WIDE.U32 per iteration -- SASS checked, see profiles/r02_tuning_log.md)
  const char* names[6] = {"imad_lo", "imad_hi", "imad_wide_hoisted_adds", "iadd", "wide_plus_add", "imad_wide_acc64"};
  double per[6];
  printf("{\"sms\": %d, \"clock_khz_max\": %d", sms, prop.clockRate);
  for (int mode = 0; mode < 6; mode++) {
    int blocks = sms * 8;  // 8 x 256 threads = 2048 threads / SM
    double ms = 0; long long cycles = 0;
    auto go = [&](auto kern) {
      ms = time_ms([&] { kern<<<blocks, 256>>>(out, 12345u, cyc); });
      cudaMemcpy(&cycles, cyc, 8, cudaMemcpyDeviceToHost);
    };
    if (mode == 0) go(k_rate<0>); else if (mode == 1) go(k_rate<1>); else if (mode == 2) go(k_rate<2>);
    else if (mode == 3) go(k_rate<3>); else if (mode == 4) go(k_rate<4>); else go(k_rate<5>);
    double ops = (double)blocks * 256 * ITERS * NACC * (mode == 4 ? 2 : 1);
    // whole-grid rate from the CUDA-event time; per-clock figures use the device's maximum SM clock (prop.clockRate), which is
    // what an unthrottled integer kernel runs at (bench.py samples the real clock next to this).  Block 0's clock64() span is
    // printed as a cross-check only: round 1 divided it by the whole-grid time and reported a meaningless 245 "MHz".
    per[mode] = ops / (ms * 1e-3 * (double)prop.clockRate * 1e3 * sms);
    (void)cycles;
    printf(", \"%s\": {\"ops_per_clk_per_sm\": %.2f, \"gops\": %.1f}", names[mode], per[mode], ops / (ms * 1e-3) / 1e9);
  }
  // Fq product throughput at several occupancies / ILP
  for (int cfg = 0; cfg < 4; cfg++) {
    int bps = (cfg == 0) ? 2 : (cfg == 1) ? 4 : (cfg == 2) ? 8 : 4;
    int blocks = sms * bps, iters = 2000;
    double ms;
    if (cfg < 3) ms = time_ms([&] { k_fqmul<1><<<blocks, 256>>>(out, 777u, iters, cyc); });
    else ms = time_ms([&] { k_fqmul<2><<<blocks, 256>>>(out, 777u, iters, cyc); });
    double muls = (double)blocks * 256 * iters * (cfg == 3 ? 2 : 1);
    printf(", \"fqmul_cfg%d\": {\"blocks_per_sm\": %d, \"chains\": %d, \"gmul_per_s\": %.2f}", cfg, bps, cfg == 3 ? 2 : 1, muls / (ms * 1e-3) / 1e9);
  }
  {
    int blocks = sms * 4, iters = 2000;
    double ms = time_ms([&] { k_fq2mul<<<blocks, 256>>>(out, 777u, iters); });
    printf(", \"fq2mul_mem\": {\"gmul_fq_equiv_per_s\": %.2f}", (double)blocks * 256 * iters * 3 / (ms * 1e-3) / 1e9);
  }
  printf("}\n");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "cuda error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
