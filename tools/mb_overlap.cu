// mb_overlap.cu -- can ALU-pipe work (IADD3 / LOP3 carry chains) hide under IMAD.WIDE.U32.X carry rows on one SM sub-partition?
// Decides how the cooperative pairing machine (csrc/coop.cuh) has to arrange its multiply-free stretches.
//   mode 0  rows only: 8 carry rows of 8 IMAD.WIDE.U32.X (the wide_mac body) per iteration
//   mode 1  rows + K ALU instructions INTERLEAVED in the same basic block (independent add chains on other registers)
//   mode 2  phases: one iteration = [rows] then [K ALU instructions], warps free-running (no barrier)
//   mode 3  phases with a named barrier over the warps of the sub-partition after each phase pair (lock step)
//   mode 4  ALU only (K instructions per iteration)
// Launch: one block per SM, `wps` warps per sub-partition (block = 4 * wps warps).  Output: one JSON line with
// SM cycles per iteration per warp for every (mode, K, wps); time from CUDA events, clock from the device attribute.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITERS 2048

__device__ __forceinline__ void rows(uint32_t (&X)[8], uint32_t (&Y)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#pragma unroll
  for (int r = 0; r < 8; r++) {
    uint32_t w = b[r];
    asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;\n\t"
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7])
        : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(w));
    asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;\n\t"
        : "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]), "+r"(Y[7])
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(w));
  }
}
// K ALU instructions as add-with-carry chains of 8 over two register sets (what fq_add / fq_csub look like)
template <int K>
__device__ __forceinline__ void alu(uint32_t (&s)[8], uint32_t (&t)[8]) {
#pragma unroll
  for (int k = 0; k < K / 16; k++) {
    asm volatile("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;\n\t"
        : "+r"(s[0]), "+r"(s[1]), "+r"(s[2]), "+r"(s[3]), "+r"(s[4]), "+r"(s[5]), "+r"(s[6]), "+r"(s[7])
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]));
    asm volatile("sub.cc.u32 %0, %0, %8;\n\t"
        "subc.cc.u32 %1, %1, %9;\n\t"
        "subc.cc.u32 %2, %2, %10;\n\t"
        "subc.cc.u32 %3, %3, %11;\n\t"
        "subc.cc.u32 %4, %4, %12;\n\t"
        "subc.cc.u32 %5, %5, %13;\n\t"
        "subc.cc.u32 %6, %6, %14;\n\t"
        "subc.u32 %7, %7, %15;\n\t"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7])
        : "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]));
  }
}

template <int MODE, int K>
__global__ void __launch_bounds__(768) k(uint32_t* out, const uint32_t* in, int wps) {
  uint32_t a[8], b[8], X[8], Y[8], s[8], t[8];
  for (int i = 0; i < 8; i++) {
    a[i] = in[(threadIdx.x + i * 31) & 1023];
    b[i] = in[(threadIdx.x * 7 + i * 13 + 5) & 1023];
    X[i] = a[i] + 1; Y[i] = b[i] + 3; s[i] = a[i] ^ 0x55; t[i] = b[i] ^ 0x33;
  }
  const int bar = 1 + ((threadIdx.x >> 5) & 3);
  const int nthr = wps * 32;
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) rows(X, Y, a, b);
    else if (MODE == 1) {
      // same basic block: ptxas is free to interleave the two independent streams
      rows(X, Y, a, b);
      alu<K>(s, t);
    } else if (MODE == 2 || MODE == 3) {
      rows(X, Y, a, b);
      if (MODE == 3) asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory");
      else __syncwarp();
      alu<K>(s, t);
      if (MODE == 3) asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory");
    } else alu<K>(s, t);
  }
  uint32_t acc = 0;
  for (int i = 0; i < 8; i++) acc ^= X[i] ^ Y[i] ^ s[i] ^ t[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE, int K>
static void run(int sms, double ghz, uint32_t* out, uint32_t* in, bool& first) {
  const int wpss[4] = {1, 2, 3, 6};
  for (int wi = 0; wi < 4; wi++) {
    int wps = wpss[wi];
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, K><<<sms, wps * 128>>>(out, in, wps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE, K><<<sms, wps * 128>>>(out, in, wps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * ghz * 1e9 / ITERS;  // SM cycles per iteration (all wps warps of a sub-partition do one each)
    printf("%s\"m%d_k%d_w%d\": {\"cyc_per_iter\": %.1f, \"cyc_per_warp_iter\": %.1f}", first ? "" : ", ", MODE, K, wps, cyc, cyc / wps);
    first = false;
  }
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz / 1e6;
  uint32_t *out, *in;
  cudaMalloc(&out, (size_t)sms * 768 * 4); cudaMalloc(&in, 4096);
  uint32_t h[1024]; for (int i = 0; i < 1024; i++) h[i] = (i * 2654435761u) >> 3;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  bool first = true;
  printf("{\"sms\": %d, \"ghz_assumed\": %.3f, ", sms, ghz);
  run<0, 0>(sms, ghz, out, in, first);
  run<4, 64>(sms, ghz, out, in, first);
  run<4, 128>(sms, ghz, out, in, first);
  run<1, 32>(sms, ghz, out, in, first);
  run<1, 64>(sms, ghz, out, in, first);
  run<1, 128>(sms, ghz, out, in, first);
  run<2, 64>(sms, ghz, out, in, first);
  run<2, 128>(sms, ghz, out, in, first);
  run<3, 64>(sms, ghz, out, in, first);
  run<3, 128>(sms, ghz, out, in, first);
  printf("}\n");
  return cudaDeviceSynchronize() != cudaSuccess;
}
