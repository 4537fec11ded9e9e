// mb_wide.cu -- issue cost of the IMAD.WIDE forms on the fmaheavy pipe (decides the Fq product layout):
//   mode 0  plain IMAD.WIDE.U32 acc64 += a*b, 9 independent 64-bit column accumulators, register operands (radix-2^29 style)
//   mode 1  carry rows as in fq.cuh (mad.lo.cc / madc.hi.cc chains -> IMAD.WIDE.U32.X), 8-limb rows
//   mode 2  plain IMAD.WIDE.U32 with one operand an immediate (reduction by the constant modulus)
//   mode 3  IADD3 / LOP3 / SHF mix of a radix-2^29 carry propagation (ALU pipe), for the issue-slot budget
// Output: one JSON line; "clk_per_inst" = SM cycles per warp-instruction per SMSP for the counted opcode.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITERS 1024

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, const uint32_t* in, long long* cycles) {
  uint32_t a[9], b[9];
  for (int i = 0; i < 9; i++) { a[i] = in[(threadIdx.x + i * 31) & 1023]; b[i] = in[(threadIdx.x * 7 + i * 13 + 5) & 1023]; }
  uint64_t c[9];
  for (int i = 0; i < 9; i++) c[i] = a[i] ^ b[i];
  uint32_t X[8], Y[8];
  for (int i = 0; i < 8; i++) { X[i] = a[i] + 1; Y[i] = b[i] + 3; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0 || MODE == 2) {  // operands change every iteration so the products cannot be hoisted
#pragma unroll
      for (int i = 0; i < 9; i++) { a[i] += (uint32_t)c[i]; b[i] ^= a[i]; }
    }
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[(i + j) % 9]) : "r"(a[i]), "r"(b[j]));
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 9; i++) {
        asm volatile("mad.wide.u32 %0, %1, 0x187cfd47, %0;" : "+l"(c[(i + 0) % 9]) : "r"(a[i]));
        asm volatile("mad.wide.u32 %0, %1, 0x1c208c16, %0;" : "+l"(c[(i + 1) % 9]) : "r"(a[i]));
        asm volatile("mad.wide.u32 %0, %1, 0x0871ca8d, %0;" : "+l"(c[(i + 2) % 9]) : "r"(a[i]));
        asm volatile("mad.wide.u32 %0, %1, 0x17816a91, %0;" : "+l"(c[(i + 3) % 9]) : "r"(a[i]));
        asm volatile("mad.wide.u32 %0, %1, 0x0181585d, %0;" : "+l"(c[(i + 4) % 9]) : "r"(a[i]));
        asm volatile("mad.wide.u32 %0, %1, 0x185045b6, %0;" : "+l"(c[(i + 5) % 9]) : "r"(a[i]));
        asm volatile("mad.wide.u32 %0, %1, 0x0131a029, %0;" : "+l"(c[(i + 6) % 9]) : "r"(a[i]));
        asm volatile("mad.wide.u32 %0, %1, 0x10644e72, %0;" : "+l"(c[(i + 7) % 9]) : "r"(a[i]));
        asm volatile("mad.wide.u32 %0, %1, 0x00000306, %0;" : "+l"(c[(i + 8) % 9]) : "r"(a[i]));
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < 8; r++) {
        uint32_t w = b[r];
        asm volatile("add.cc.u32 %0, %0, %2;\n\t"
            "madc.lo.cc.u32 %1, %9, %13, %3;\n\t"
            "madc.hi.cc.u32 %2, %9, %13, %4;\n\t"
            "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
            "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
            "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
            "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
            "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
            "madc.hi.u32 %8, %12, %13, 0;\n\t"
            : "+r"(Y[0]), "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7])
            : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(w));
        asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;\n\t"
            : "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]), "+r"(Y[7]), "+r"(X[7])
            : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(w));
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        uint64_t carry = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
          uint64_t v = c[i] + carry;
          c[i] = v & 0x1fffffffu;
          carry = v >> 29;
        }
        c[0] += carry * 3 + a[r];
      }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < 9; i++) acc ^= (uint32_t)c[i] ^ (uint32_t)(c[i] >> 32);
  for (int i = 0; i < 8; i++) acc ^= X[i] ^ Y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  uint32_t *out, *in; long long* cyc;
  cudaMalloc(&out, (size_t)sms * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMalloc(&cyc, 8);
  uint32_t h[1024]; for (int i = 0; i < 1024; i++) h[i] = (i * 2654435761u) >> 3;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  const int per_iter[4] = {81, 81, 64, 0};  // counted multiply instructions per loop iteration
  printf("{\"sms\": %d", sms);
  for (int bps = 1; bps <= 8; bps *= 2) {
    for (int mode = 0; mode < 4; mode++) {
      int blocks = sms * bps;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      auto launch = [&] {
        if (mode == 0) k<0><<<blocks, 256>>>(out, in, cyc); else if (mode == 1) k<1><<<blocks, 256>>>(out, in, cyc);
        else if (mode == 2) k<2><<<blocks, 256>>>(out, in, cyc); else k<3><<<blocks, 256>>>(out, in, cyc);
      };
      launch(); cudaDeviceSynchronize();
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long cycles; cudaMemcpy(&cycles, cyc, 8, cudaMemcpyDeviceToHost);
      // warps per SMSP = bps * 8 / 4 ; cycles per (warp-iteration) per SMSP
      double warps_per_smsp = bps * 2.0;
      double clk_per_iter = (double)cycles / ITERS / warps_per_smsp;
      int cnt = mode == 1 ? 64 : per_iter[mode];
      printf(", \"m%d_b%d\": {\"ms\": %.3f, \"clk_per_warp_iter\": %.1f, \"clk_per_inst\": %.2f}", mode, bps, ms, clk_per_iter, cnt ? clk_per_iter / cnt : 0.0);
    }
  }
  printf("}\n");
  return cudaDeviceSynchronize() != cudaSuccess;
}
