// fq.cuh -- BN254 base field Fq on 8 x 32-bit limbs, Montgomery form (R = 2^256), one thread per element.
//
// Replaces the `Fq` type of the reference's arithmetic dependency (crate zeropool-bn 0.5.11, imported as `bn`
// at /root/reference/Cargo.toml:24; call sites /root/reference/src/utils.rs:44,88-90,111-112).  Values are
// always kept fully reduced in [0, q), so two implementations that agree on the field value agree on the bits.
//
// Device path: the multiplication is a 32-bit CIOS Montgomery product written as PTX carry chains in the
// "even/odd accumulator" arrangement: every (mad.lo.cc, madc.hi.cc) pair on the same operands sits on one carry
// chain, which ptxas fuses into a single IMAD.WIDE.U32.X with predicate carry-in/out -- 128 wide multiply-adds
// + 8 IMAD for the quotient digits per product.  The same header compiles as plain C++ (g++) for the CPU-side
// simulation used by tests/hostsim: there the portable 64-bit code path below is used instead of PTX.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BN_FN __device__ __forceinline__
#define BN_NOINLINE __device__ __noinline__
#define BN_CONST __constant__ const
#else
#define BN_FN static inline
#define BN_NOINLINE static __attribute__((noinline))
#define BN_CONST static const
#endif

#include "constants.cuh"

namespace bn {

struct alignas(16) fq {
  uint32_t l[8];
};

// q limbs as literals (immediates in SASS)
#define BN_Q0 0xd87cfd47u
#define BN_Q1 0x3c208c16u
#define BN_Q2 0x6871ca8du
#define BN_Q3 0x97816a91u
#define BN_Q4 0x8181585du
#define BN_Q5 0xb85045b6u
#define BN_Q6 0xe131a029u
#define BN_Q7 0x30644e72u

BN_FN fq fq_zero() {
  fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = 0;
  return r;
}
BN_FN fq fq_from_limbs(const uint32_t* p) {
  fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = p[i];
  return r;
}
BN_FN fq fq_one() { return fq_from_limbs(K_ONE); }
BN_FN bool fq_is_zero(const fq& a) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) t |= a.l[i];
  return t == 0;
}
BN_FN bool fq_eq(const fq& a, const fq& b) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) t |= a.l[i] ^ b.l[i];
  return t == 0;
}
// plain 256-bit compare helpers on limb arrays: a >= b
BN_FN bool u256_geq(const uint32_t* a, const uint32_t* b) {
  uint32_t borrow = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t t = (uint64_t)a[i] - b[i] - borrow;
    borrow = (uint32_t)(t >> 63);
  }
  return borrow == 0;
}
// r = a - b, returns borrow
BN_FN uint32_t u256_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t borrow = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t t = (uint64_t)a[i] - b[i] - borrow;
    r[i] = (uint32_t)t;
    borrow = (uint32_t)(t >> 63);
  }
  return borrow;
}

// ------------------------------------------------------------------------------------------------ add / sub
#if defined(__CUDA_ARCH__)
// a >= q decided by the top limb; the full comparison runs only when the top limbs are equal (probability 2^-32 for field
// values), out of line, so that the common path is two compares and a branch that is never taken
__device__ __noinline__ uint32_t fq_geq_q_slow(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6) {
  uint32_t bw;
  asm("{\n\t.reg .u32 t;\n\t"
      "sub.cc.u32 t, %1, 0xd87cfd47;\n\t"
      "subc.cc.u32 t, %2, 0x3c208c16;\n\t"
      "subc.cc.u32 t, %3, 0x6871ca8d;\n\t"
      "subc.cc.u32 t, %4, 0x97816a91;\n\t"
      "subc.cc.u32 t, %5, 0x8181585d;\n\t"
      "subc.cc.u32 t, %6, 0xb85045b6;\n\t"
      "subc.cc.u32 t, %7, 0xe131a029;\n\t"
      "subc.u32 %0, 0, 0;\n\t}"
      : "=r"(bw)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6));
  return bw == 0;  // no borrow: the low seven limbs are >= those of q
}
// conditional subtraction: a in [0, 2^256) -> a - q if a >= q else a.  The subtraction is PREDICATED (eight instructions that
// do nothing when a < q) instead of computed-then-selected (eight subtractions, a borrow and eight selects): the kernels on this
// path are bound by the number of instructions they issue (profiles/r02_tuning_log.md), and every modular addition /
// subtraction / reduction ends in one of these.
#if defined(BN_FQ_PRED_CSUB)
BN_FN fq fq_csub(const fq& a) {
  uint32_t ge = a.l[7] > BN_Q7 ? 1u : 0u;
  if (a.l[7] == BN_Q7) ge = fq_geq_q_slow(a.l[0], a.l[1], a.l[2], a.l[3], a.l[4], a.l[5], a.l[6]);
  fq r = a;
  asm("{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %8, 0;\n\t"
      "@p sub.cc.u32 %0, %0, 0xd87cfd47;\n\t"
      "@p subc.cc.u32 %1, %1, 0x3c208c16;\n\t"
      "@p subc.cc.u32 %2, %2, 0x6871ca8d;\n\t"
      "@p subc.cc.u32 %3, %3, 0x97816a91;\n\t"
      "@p subc.cc.u32 %4, %4, 0x8181585d;\n\t"
      "@p subc.cc.u32 %5, %5, 0xb85045b6;\n\t"
      "@p subc.cc.u32 %6, %6, 0xe131a029;\n\t"
      "@p subc.u32 %7, %7, 0x30644e72;\n\t}"
      : "+r"(r.l[0]), "+r"(r.l[1]), "+r"(r.l[2]), "+r"(r.l[3]), "+r"(r.l[4]), "+r"(r.l[5]), "+r"(r.l[6]), "+r"(r.l[7])
      : "r"(ge));
  return r;
}
#else
BN_FN fq fq_csub(const fq& a) {
  uint32_t t0, t1, t2, t3, t4, t5, t6, t7, bw;
  asm("sub.cc.u32 %0, %9, 0xd87cfd47;\n\t"
      "subc.cc.u32 %1, %10, 0x3c208c16;\n\t"
      "subc.cc.u32 %2, %11, 0x6871ca8d;\n\t"
      "subc.cc.u32 %3, %12, 0x97816a91;\n\t"
      "subc.cc.u32 %4, %13, 0x8181585d;\n\t"
      "subc.cc.u32 %5, %14, 0xb85045b6;\n\t"
      "subc.cc.u32 %6, %15, 0xe131a029;\n\t"
      "subc.cc.u32 %7, %16, 0x30644e72;\n\t"
      "subc.u32 %8, 0, 0;\n\t"
      : "=&r"(t0), "=&r"(t1), "=&r"(t2), "=&r"(t3), "=&r"(t4), "=&r"(t5), "=&r"(t6), "=&r"(t7), "=&r"(bw)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]));
  fq r;
  const bool keep = bw != 0;  // borrow: a < q
  r.l[0] = keep ? a.l[0] : t0; r.l[1] = keep ? a.l[1] : t1; r.l[2] = keep ? a.l[2] : t2; r.l[3] = keep ? a.l[3] : t3;
  r.l[4] = keep ? a.l[4] : t4; r.l[5] = keep ? a.l[5] : t5; r.l[6] = keep ? a.l[6] : t6; r.l[7] = keep ? a.l[7] : t7;
  return r;
}
#endif
BN_FN fq fq_add(const fq& a, const fq& b) {
  fq s;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      : "=&r"(s.l[0]), "=&r"(s.l[1]), "=&r"(s.l[2]), "=&r"(s.l[3]), "=&r"(s.l[4]), "=&r"(s.l[5]), "=&r"(s.l[6]), "=&r"(s.l[7])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
        "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  return fq_csub(s);  // a + b < 2 q < 2^255: no carry out of the top limb
}
#if !defined(BN_FQ_SEL_SUB)  // predicated add-back: 19 instructions instead of 25 (-0.35 % k_coop4_run, -2.9 % k_verify_lines, A/B r02)
BN_FN fq fq_sub(const fq& a, const fq& b) {
  fq r;
  // a - b, then + q under the borrow's predicate
  asm("{\n\t.reg .pred p;\n\t.reg .u32 m;\n\t"
      "sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.cc.u32 %7, %15, %23;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "setp.ne.u32 p, m, 0;\n\t"
      "@p add.cc.u32 %0, %0, 0xd87cfd47;\n\t"
      "@p addc.cc.u32 %1, %1, 0x3c208c16;\n\t"
      "@p addc.cc.u32 %2, %2, 0x6871ca8d;\n\t"
      "@p addc.cc.u32 %3, %3, 0x97816a91;\n\t"
      "@p addc.cc.u32 %4, %4, 0x8181585d;\n\t"
      "@p addc.cc.u32 %5, %5, 0xb85045b6;\n\t"
      "@p addc.cc.u32 %6, %6, 0xe131a029;\n\t"
      "@p addc.u32 %7, %7, 0x30644e72;\n\t}"
      : "=&r"(r.l[0]), "=&r"(r.l[1]), "=&r"(r.l[2]), "=&r"(r.l[3]), "=&r"(r.l[4]), "=&r"(r.l[5]), "=&r"(r.l[6]), "=&r"(r.l[7])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
        "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  return r;
}
#else
BN_FN fq fq_sub(const fq& a, const fq& b) {
  uint32_t d0, d1, d2, d3, d4, d5, d6, d7, m;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;\n\t"
      : "=&r"(d0), "=&r"(d1), "=&r"(d2), "=&r"(d3), "=&r"(d4), "=&r"(d5), "=&r"(d6), "=&r"(d7), "=&r"(m)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
        "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  fq r;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      : "=&r"(r.l[0]), "=&r"(r.l[1]), "=&r"(r.l[2]), "=&r"(r.l[3]), "=&r"(r.l[4]), "=&r"(r.l[5]), "=&r"(r.l[6]), "=&r"(r.l[7])
      : "r"(d0), "r"(d1), "r"(d2), "r"(d3), "r"(d4), "r"(d5), "r"(d6), "r"(d7),
        "r"(m & BN_Q0), "r"(m & BN_Q1), "r"(m & BN_Q2), "r"(m & BN_Q3), "r"(m & BN_Q4), "r"(m & BN_Q5), "r"(m & BN_Q6), "r"(m & BN_Q7));
  return r;
}
#endif
#else
BN_FN fq fq_add(const fq& a, const fq& b) {
  fq s, t;
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.l[i] + b.l[i];
    s.l[i] = (uint32_t)c;
    c >>= 32;
  }
  uint32_t bw = u256_sub(t.l, s.l, K_Q);
  return bw ? s : t;
}
BN_FN fq fq_sub(const fq& a, const fq& b) {
  fq d;
  uint32_t bw = u256_sub(d.l, a.l, b.l);
  if (bw) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
      c += (uint64_t)d.l[i] + K_Q[i];
      d.l[i] = (uint32_t)c;
      c >>= 32;
    }
  }
  return d;
}
BN_FN fq fq_csub(const fq& a) {
  fq t;
  const uint32_t qq[8] = {BN_Q0, BN_Q1, BN_Q2, BN_Q3, BN_Q4, BN_Q5, BN_Q6, BN_Q7};
  uint32_t bw = u256_sub(t.l, a.l, qq);
  for (int i = 0; i < 8; i++) t.l[i] = bw ? a.l[i] : t.l[i];
  return t;
}
#endif
BN_FN fq fq_dbl(const fq& a) { return fq_add(a, a); }
BN_FN fq fq_neg(const fq& a) { return fq_sub(fq_zero(), a); }
// a / 2 (valid on Montgomery representatives as well: the map is linear): (a + (a odd ? q : 0)) >> 1, canonical, no product
BN_FN fq fq_halve(const fq& a) {
  const uint32_t qq[8] = {BN_Q0, BN_Q1, BN_Q2, BN_Q3, BN_Q4, BN_Q5, BN_Q6, BN_Q7};
  const uint32_t odd = 0u - (a.l[0] & 1u);
  uint32_t t[8];
  uint64_t c = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.l[i] + (qq[i] & odd);
    t[i] = (uint32_t)c;
    c >>= 32;
  }
  fq r;  // a + q < 2^255: nothing above t[7]
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 8; i++) r.l[i] = (t[i] >> 1) | (i < 7 ? t[i + 1] << 31 : 0u);
  return r;
}

#define BN_KQ_RECIP 0xa948e8c0u /* floor(2^59 / ((q >> 226) + 1)) */
// t (nine limbs, below 11 q) mod q: quotient estimate from the top bits, k q from the table, one conditional subtraction
BN_FN fq fq_reduce9(const uint32_t (&t)[9], const uint32_t* kq) {
  fq r;
#if defined(__CUDA_ARCH__)
  const uint32_t k = min(__umulhi(__funnelshift_r(t[7], t[8], 2), BN_KQ_RECIP) >> 27, 10u);
  const uint4 e0 = *(const uint4*)(kq + 8 * k), e1 = *(const uint4*)(kq + 8 * k + 4);
  asm("sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32 %7, %15, %23;\n\t"
      : "=&r"(r.l[0]), "=&r"(r.l[1]), "=&r"(r.l[2]), "=&r"(r.l[3]), "=&r"(r.l[4]), "=&r"(r.l[5]), "=&r"(r.l[6]), "=&r"(r.l[7])
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(e0.x), "r"(e0.y), "r"(e0.z),
        "r"(e0.w), "r"(e1.x), "r"(e1.y), "r"(e1.z), "r"(e1.w));
#else
  const uint32_t T = (t[8] << 30) | (t[7] >> 2);
  uint32_t k = (uint32_t)(((uint64_t)T * BN_KQ_RECIP) >> 32) >> 27;
  if (k > 10) k = 10;
  uint32_t bw = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)t[i] - kq[8 * k + i] - bw;
    r.l[i] = (uint32_t)d;
    bw = (uint32_t)(d >> 63);
  }
#endif
  return fq_csub(r);
}
// t = a + K - b on nine limbs, for a, b < 2^256 and a constant K (a multiple of q, eight limbs) with a + K >= b.  No reduction.
BN_FN void fq9_addk_sub(uint32_t (&t)[9], const fq& a, const uint32_t (&K)[8], const fq& b) {
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;\n\t"
      : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(t[8])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]), "r"(K[0]), "r"(K[1]), "r"(K[2]),
        "r"(K[3]), "r"(K[4]), "r"(K[5]), "r"(K[6]), "r"(K[7]));
  asm("sub.cc.u32 %0, %0, %9;\n\t"
      "subc.cc.u32 %1, %1, %10;\n\t"
      "subc.cc.u32 %2, %2, %11;\n\t"
      "subc.cc.u32 %3, %3, %12;\n\t"
      "subc.cc.u32 %4, %4, %13;\n\t"
      "subc.cc.u32 %5, %5, %14;\n\t"
      "subc.cc.u32 %6, %6, %15;\n\t"
      "subc.cc.u32 %7, %7, %16;\n\t"
      "subc.u32 %8, %8, 0;\n\t"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
      : "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
#else
  int64_t c = 0;
  for (int i = 0; i < 9; i++) {
    c += (int64_t)(i < 8 ? a.l[i] : 0) + (int64_t)(i < 8 ? K[i] : 0) - (int64_t)(i < 8 ? b.l[i] : 0);
    t[i] = (uint32_t)c;
    c >>= 32;  // arithmetic shift: borrow / carry in one
  }
#endif
}
// d = a - b mod 2^256; returns 0 or 0xffffffff (the borrow, as a signed word)
BN_FN uint32_t fq_sub_borrow(fq& d, const fq& a, const fq& b) {
  uint32_t bw;
#if defined(__CUDA_ARCH__)
  fq r;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;\n\t"
      : "=&r"(r.l[0]), "=&r"(r.l[1]), "=&r"(r.l[2]), "=&r"(r.l[3]), "=&r"(r.l[4]), "=&r"(r.l[5]), "=&r"(r.l[6]), "=&r"(r.l[7]), "=&r"(bw)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
        "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  d = r;
#else
  fq r;
  bw = 0u - u256_sub(r.l, a.l, b.l);
  d = r;
#endif
  return bw;
}
// t += a + s on nine limbs, s a SIGNED small word (0, -1, -2 as 0xffffffff, 0xfffffffe: sign-extended into the ninth limb)
BN_FN void fq9_add(uint32_t (&t)[9], const fq& a, uint32_t s) {
#if defined(__CUDA_ARCH__)
  const uint32_t sx = (uint32_t)((int32_t)s >> 31);
  asm("add.cc.u32 %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, %10;\n\t"
      "addc.cc.u32 %2, %2, %11;\n\t"
      "addc.cc.u32 %3, %3, %12;\n\t"
      "addc.cc.u32 %4, %4, %13;\n\t"
      "addc.cc.u32 %5, %5, %14;\n\t"
      "addc.cc.u32 %6, %6, %15;\n\t"
      "addc.cc.u32 %7, %7, %16;\n\t"
      "addc.u32 %8, %8, 0;\n\t"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]));
  asm("add.cc.u32 %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, %10;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %10;\n\t"
      "addc.cc.u32 %4, %4, %10;\n\t"
      "addc.cc.u32 %5, %5, %10;\n\t"
      "addc.cc.u32 %6, %6, %10;\n\t"
      "addc.cc.u32 %7, %7, %10;\n\t"
      "addc.u32 %8, %8, %10;\n\t"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
      : "r"(s), "r"(sx));
#else
  int64_t c = (int64_t)(int32_t)s;
  for (int i = 0; i < 9; i++) {
    c += (int64_t)t[i] + (int64_t)(i < 8 ? a.l[i] : 0);
    t[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
}
// r = a + b mod 2^256; returns the carry
BN_FN uint32_t fq_add_carry(fq& r, const fq& a, const fq& b) {
  uint32_t cy;
#if defined(__CUDA_ARCH__)
  fq s;
  asm("add.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;\n\t"
      : "=&r"(s.l[0]), "=&r"(s.l[1]), "=&r"(s.l[2]), "=&r"(s.l[3]), "=&r"(s.l[4]), "=&r"(s.l[5]), "=&r"(s.l[6]), "=&r"(s.l[7]), "=&r"(cy)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
        "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  r = s;
#else
  uint64_t c = 0;
  fq s;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.l[i] + b.l[i];
    s.l[i] = (uint32_t)c;
    c >>= 32;
  }
  cy = (uint32_t)c;
  r = s;
#endif
  return cy;
}
// a + b as a 256-bit number (no reduction; the caller knows it fits)
BN_FN fq fq_add_raw(const fq& a, const fq& b) {
  fq s;
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      : "=&r"(s.l[0]), "=&r"(s.l[1]), "=&r"(s.l[2]), "=&r"(s.l[3]), "=&r"(s.l[4]), "=&r"(s.l[5]), "=&r"(s.l[6]), "=&r"(s.l[7])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
        "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.l[i] + b.l[i];
    s.l[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  return s;
}
// (9 x + z) mod q in one reduction, for x < q and z <= q (canonical, or q itself): t = 9 x + z < 10 q is formed on nine
// limbs, the quotient k = floor(t / q) is estimated from the top 32 bits of t >> 226 (reciprocal multiplication; the
// estimate is k or k - 1, checked exhaustively at the multiples of q and on 3 * 10^5 random values in the tests), k q comes
// from a table (kq: entries of 8 limbs, k q mod 2^256, k = 0..10) and one conditional subtraction finishes.  Replaces the
// five modular additions of 8 x + x + z; xi x = (9 x0 - x1, 9 x1 + x0) is two of these (coop.cuh coop_put_p).
BN_FN fq fq_mul9_add(const fq& x, const fq& z, const uint32_t* kq) {
  uint32_t t[9];
#if defined(__CUDA_ARCH__)
  uint32_t s[9];
  s[0] = x.l[0] << 3;
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = __funnelshift_l(x.l[i - 1], x.l[i], 3);
  s[8] = x.l[7] >> 29;
  asm("add.cc.u32 %0, %9, %18;\n\t"
      "addc.cc.u32 %1, %10, %19;\n\t"
      "addc.cc.u32 %2, %11, %20;\n\t"
      "addc.cc.u32 %3, %12, %21;\n\t"
      "addc.cc.u32 %4, %13, %22;\n\t"
      "addc.cc.u32 %5, %14, %23;\n\t"
      "addc.cc.u32 %6, %15, %24;\n\t"
      "addc.cc.u32 %7, %16, %25;\n\t"
      "addc.u32 %8, %17, 0;\n\t"
      : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(t[8])
      : "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]), "r"(s[8]), "r"(x.l[0]), "r"(x.l[1]),
        "r"(x.l[2]), "r"(x.l[3]), "r"(x.l[4]), "r"(x.l[5]), "r"(x.l[6]), "r"(x.l[7]));
  asm("add.cc.u32 %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, %10;\n\t"
      "addc.cc.u32 %2, %2, %11;\n\t"
      "addc.cc.u32 %3, %3, %12;\n\t"
      "addc.cc.u32 %4, %4, %13;\n\t"
      "addc.cc.u32 %5, %5, %14;\n\t"
      "addc.cc.u32 %6, %6, %15;\n\t"
      "addc.cc.u32 %7, %7, %16;\n\t"
      "addc.u32 %8, %8, 0;\n\t"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
      : "r"(z.l[0]), "r"(z.l[1]), "r"(z.l[2]), "r"(z.l[3]), "r"(z.l[4]), "r"(z.l[5]), "r"(z.l[6]), "r"(z.l[7]));
  return fq_reduce9(t, kq);  // (the clamp inside keeps the table index in range for items that carry arbitrary bits)
#else
  uint64_t c = 0;
  for (int i = 0; i < 9; i++) {
    uint64_t xi = i < 8 ? x.l[i] : 0, zi = i < 8 ? z.l[i] : 0;
    c += 9 * xi + zi;
    t[i] = (uint32_t)c;
    c >>= 32;
  }
  return fq_reduce9(t, kq);
#endif
}
// (3 t + 2 z) mod q for t < q, z <= q, with one reduction: the tail of the Granger-Scott squaring, 3 t +- 2 a, is this with
// z = a or z = q - a (coop.cuh coop_commit)
BN_FN fq fq_3t_2z(const fq& t, const fq& z, const uint32_t* kq) {
  uint32_t v[9];
#if defined(__CUDA_ARCH__)
  uint32_t s[9], w[9];
  s[0] = t.l[0] << 1;
  w[0] = z.l[0] << 1;
#pragma unroll
  for (int i = 1; i < 8; i++) {
    s[i] = __funnelshift_l(t.l[i - 1], t.l[i], 1);
    w[i] = __funnelshift_l(z.l[i - 1], z.l[i], 1);
  }
  s[8] = t.l[7] >> 31;
  w[8] = z.l[7] >> 31;
  asm("add.cc.u32 %0, %9, %18;\n\t"
      "addc.cc.u32 %1, %10, %19;\n\t"
      "addc.cc.u32 %2, %11, %20;\n\t"
      "addc.cc.u32 %3, %12, %21;\n\t"
      "addc.cc.u32 %4, %13, %22;\n\t"
      "addc.cc.u32 %5, %14, %23;\n\t"
      "addc.cc.u32 %6, %15, %24;\n\t"
      "addc.cc.u32 %7, %16, %25;\n\t"
      "addc.u32 %8, %17, 0;\n\t"
      : "=&r"(v[0]), "=&r"(v[1]), "=&r"(v[2]), "=&r"(v[3]), "=&r"(v[4]), "=&r"(v[5]), "=&r"(v[6]), "=&r"(v[7]), "=&r"(v[8])
      : "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]), "r"(s[8]), "r"(t.l[0]), "r"(t.l[1]),
        "r"(t.l[2]), "r"(t.l[3]), "r"(t.l[4]), "r"(t.l[5]), "r"(t.l[6]), "r"(t.l[7]));
  asm("add.cc.u32 %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, %10;\n\t"
      "addc.cc.u32 %2, %2, %11;\n\t"
      "addc.cc.u32 %3, %3, %12;\n\t"
      "addc.cc.u32 %4, %4, %13;\n\t"
      "addc.cc.u32 %5, %5, %14;\n\t"
      "addc.cc.u32 %6, %6, %15;\n\t"
      "addc.cc.u32 %7, %7, %16;\n\t"
      "addc.u32 %8, %8, %17;\n\t"
      : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8])
      : "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 9; i++) {
    uint64_t ti = i < 8 ? t.l[i] : 0, zi = i < 8 ? z.l[i] : 0;
    c += 3 * ti + 2 * zi;
    v[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  return fq_reduce9(v, kq);
}
// q - a as a plain integer (a <= q - 1 gives 1..q; a == 0 gives q itself, which fq_mul9_add accepts as its z)
BN_FN fq fq_q_minus(const fq& a) {
  fq r;
#if defined(__CUDA_ARCH__)
  asm("sub.cc.u32 %0, 0xd87cfd47, %8;\n\t"
      "subc.cc.u32 %1, 0x3c208c16, %9;\n\t"
      "subc.cc.u32 %2, 0x6871ca8d, %10;\n\t"
      "subc.cc.u32 %3, 0x97816a91, %11;\n\t"
      "subc.cc.u32 %4, 0x8181585d, %12;\n\t"
      "subc.cc.u32 %5, 0xb85045b6, %13;\n\t"
      "subc.cc.u32 %6, 0xe131a029, %14;\n\t"
      "subc.u32 %7, 0x30644e72, %15;\n\t"
      : "=&r"(r.l[0]), "=&r"(r.l[1]), "=&r"(r.l[2]), "=&r"(r.l[3]), "=&r"(r.l[4]), "=&r"(r.l[5]), "=&r"(r.l[6]), "=&r"(r.l[7])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]));
#else
  const uint32_t qq[8] = {BN_Q0, BN_Q1, BN_Q2, BN_Q3, BN_Q4, BN_Q5, BN_Q6, BN_Q7};
  u256_sub(r.l, qq, a.l);
#endif
  return r;
}

// ------------------------------------------------------------------------------------------------ Montgomery product
// portable CIOS (host simulation; also kept on the device as the cross-check variant of the PTX product)
BN_FN fq fq_mul_portable(const fq& a, const fq& b) {
  uint32_t t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      c += (uint64_t)a.l[j] * b.l[i] + t[j];
      t[j] = (uint32_t)c;
      c >>= 32;
    }
    c += t[8];
    t[8] = (uint32_t)c;
    t[9] = (uint32_t)(c >> 32);
    uint32_t m = t[0] * K_QINV_NEG;
    const uint32_t qq[8] = {BN_Q0, BN_Q1, BN_Q2, BN_Q3, BN_Q4, BN_Q5, BN_Q6, BN_Q7};
    c = (uint64_t)m * qq[0] + t[0];
    c >>= 32;
#pragma unroll
    for (int j = 1; j < 8; j++) {
      c += (uint64_t)m * qq[j] + t[j];
      t[j - 1] = (uint32_t)c;
      c >>= 32;
    }
    c += t[8];
    t[7] = (uint32_t)c;
    t[8] = t[9] + (uint32_t)(c >> 32);
  }
  fq s, r;
#pragma unroll
  for (int i = 0; i < 8; i++) s.l[i] = t[i];
  const uint32_t qq[8] = {BN_Q0, BN_Q1, BN_Q2, BN_Q3, BN_Q4, BN_Q5, BN_Q6, BN_Q7};
  uint32_t bw = u256_sub(r.l, s.l, qq);
  bool keep = (t[8] == 0) && bw;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = keep ? s.l[i] : r.l[i];
  return r;
}

#if defined(__CUDA_ARCH__)
namespace detail {
// Y[0..7] sits at limb columns c..c+7, X[0..7] at c+1..c+8 ("even" and "odd" accumulators).
// first row: Y = a_even * w, X = a_odd * w
BN_FN void mont_row_first(uint32_t* Y, uint32_t* X, const uint32_t* a, uint32_t w) {
  asm("mul.lo.u32 %0, %8, %12;\n\t"
      "mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12;\n\t"
      "mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12;\n\t"
      "mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12;\n\t"
      "mul.hi.u32 %7, %11, %12;\n\t"
      : "=&r"(Y[0]), "=&r"(Y[1]), "=&r"(Y[2]), "=&r"(Y[3]), "=&r"(Y[4]), "=&r"(Y[5]), "=&r"(Y[6]), "=&r"(Y[7])
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(w));
  asm("mul.lo.u32 %0, %8, %12;\n\t"
      "mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12;\n\t"
      "mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12;\n\t"
      "mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12;\n\t"
      "mul.hi.u32 %7, %11, %12;\n\t"
      : "=&r"(X[0]), "=&r"(X[1]), "=&r"(X[2]), "=&r"(X[3]), "=&r"(X[4]), "=&r"(X[5]), "=&r"(X[6]), "=&r"(X[7])
      : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(w));
}
// later rows: Y is the accumulator aligned at the new base column, X is the previous base-aligned accumulator
// (its limb 0 was cleared by the reduction): fold X[1] into Y[0], shift X down two limbs while adding a_odd*w,
// then Y += a_even*w with the carry-out going to X[7].
BN_FN void mont_row_next(uint32_t* Y, uint32_t* X, const uint32_t* a, uint32_t w) {
  asm("add.cc.u32 %0, %0, %2;\n\t"
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
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
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
// Montgomery reduction of the lowest limb: m = Y[0] * (-q^-1); X += q_odd * m; Y += q_even * m (Y[0] becomes 0)
BN_FN void mont_reduce_row(uint32_t* Y, uint32_t* X) {
  uint32_t m = Y[0] * K_QINV_NEG;
  asm("mad.lo.cc.u32 %0, %8, 0x3c208c16, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, 0x3c208c16, %1;\n\t"
      "madc.lo.cc.u32 %2, %8, 0x97816a91, %2;\n\t"
      "madc.hi.cc.u32 %3, %8, 0x97816a91, %3;\n\t"
      "madc.lo.cc.u32 %4, %8, 0xb85045b6, %4;\n\t"
      "madc.hi.cc.u32 %5, %8, 0xb85045b6, %5;\n\t"
      "madc.lo.cc.u32 %6, %8, 0x30644e72, %6;\n\t"
      "madc.hi.u32 %7, %8, 0x30644e72, %7;\n\t"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7])
      : "r"(m));
  asm("mad.lo.cc.u32 %0, %9, 0xd87cfd47, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, 0xd87cfd47, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, 0x6871ca8d, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, 0x6871ca8d, %3;\n\t"
      "madc.lo.cc.u32 %4, %9, 0x8181585d, %4;\n\t"
      "madc.hi.cc.u32 %5, %9, 0x8181585d, %5;\n\t"
      "madc.lo.cc.u32 %6, %9, 0xe131a029, %6;\n\t"
      "madc.hi.cc.u32 %7, %9, 0xe131a029, %7;\n\t"
      "addc.u32 %8, %8, 0;\n\t"
      : "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]), "+r"(Y[7]), "+r"(X[7])
      : "r"(m));
}
// final merge + conditional subtraction: res[k] = X[k] + Y[k+1]; value < 2q
BN_FN fq mont_finish(const uint32_t* X, const uint32_t* Y) {
  fq s;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;\n\t"
      : "=&r"(s.l[0]), "=&r"(s.l[1]), "=&r"(s.l[2]), "=&r"(s.l[3]), "=&r"(s.l[4]), "=&r"(s.l[5]), "=&r"(s.l[6]), "=&r"(s.l[7])
      : "r"(X[0]), "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]),
        "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
  return fq_csub(s);
}
}  // namespace detail

BN_FN fq fq_mul_ptx(const fq& a, const fq& b) {
  uint32_t E[8], O[8];
  detail::mont_row_first(E, O, a.l, b.l[0]);
  detail::mont_reduce_row(E, O);
  detail::mont_row_next(O, E, a.l, b.l[1]);
  detail::mont_reduce_row(O, E);
  detail::mont_row_next(E, O, a.l, b.l[2]);
  detail::mont_reduce_row(E, O);
  detail::mont_row_next(O, E, a.l, b.l[3]);
  detail::mont_reduce_row(O, E);
  detail::mont_row_next(E, O, a.l, b.l[4]);
  detail::mont_reduce_row(E, O);
  detail::mont_row_next(O, E, a.l, b.l[5]);
  detail::mont_reduce_row(O, E);
  detail::mont_row_next(E, O, a.l, b.l[6]);
  detail::mont_reduce_row(E, O);
  detail::mont_row_next(O, E, a.l, b.l[7]);
  detail::mont_reduce_row(O, E);
  return detail::mont_finish(E, O);
}
#endif

#if defined(__CUDA_ARCH__) && !defined(BN254_PORTABLE_MUL)
#if !defined(BN254_INLINE_MUL)
// One out-of-line copy of the product for the whole kernel: operands and result travel in registers (by-value
// structs use the register ABI, no stack traffic), and the pairing kernels' instruction footprint drops from
// ~200 KB to a size the instruction cache holds -- with the product inlined at every use, two resident blocks
// per SM spent more issue slots waiting for instruction fetch than on anything else (profiles/r01_*).
#if defined(BN_FQ_MUL_WIDE)
// the product through the 512-bit lazy accumulator of the cooperative machine (coop_mac.cuh: wide_mac + wide_redc + one conditional
// subtraction), defined there -- same canonical result, different instruction mix (A/B, r02 tuning log)
BN_FN fq fq_mul_wide(const fq& a, const fq& b);
__device__ __noinline__ fq fq_mul_call(fq a, fq b) { return fq_mul_wide(a, b); }
#else
__device__ __noinline__ fq fq_mul_call(fq a, fq b) { return fq_mul_ptx(a, b); }
#endif
BN_FN fq fq_mul(const fq& a, const fq& b) { return fq_mul_call(a, b); }
#else
BN_FN fq fq_mul(const fq& a, const fq& b) { return fq_mul_ptx(a, b); }
#endif
#else
BN_FN fq fq_mul(const fq& a, const fq& b) { return fq_mul_portable(a, b); }
#endif
BN_FN fq fq_sqr(const fq& a) { return fq_mul(a, a); }
// inlined product for the few routines that issue several independent products back to back (Fq2 product /
// square / scaling): ptxas interleaves their carry chains, which hides the chains' latency
#if defined(__CUDA_ARCH__) && !defined(BN254_PORTABLE_MUL)
BN_FN fq fq_mul_inl(const fq& a, const fq& b) { return fq_mul_ptx(a, b); }
#else
BN_FN fq fq_mul_inl(const fq& a, const fq& b) { return fq_mul_portable(a, b); }
#endif

// out-of-line copies: used where code size matters more than the call
BN_NOINLINE void fq_mul_ni(fq* r, const fq* a, const fq* b) { *r = fq_mul(*a, *b); }

// ------------------------------------------------------------------------------------------------ conversions
BN_FN fq fq_to_mont(const fq& a) { return fq_mul(a, fq_from_limbs(K_R2)); }
BN_FN fq fq_from_mont(const fq& a) {
  fq one = fq_zero();
  one.l[0] = 1;
  return fq_mul(a, one);
}
BN_FN uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, 0, 0x0123);
#else
  return __builtin_bswap32(x);
#endif
}
// 32 big-endian bytes -> plain limbs (no reduction, no Montgomery)
BN_FN void u256_from_be(uint32_t* l, const uint8_t* b) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t* p = b + 4 * (7 - i);
    l[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  }
}
BN_FN void u256_to_be(uint8_t* b, const uint32_t* l) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint8_t* p = b + 4 * (7 - i);
    p[0] = (uint8_t)(l[i] >> 24); p[1] = (uint8_t)(l[i] >> 16); p[2] = (uint8_t)(l[i] >> 8); p[3] = (uint8_t)l[i];
  }
}
// Fq::from_slice: value >= q -> false (NotMember)
BN_FN bool fq_from_be(fq* r, const uint8_t* b) {
  fq t;
  u256_from_be(t.l, b);
  if (u256_geq(t.l, K_Q)) return false;
  *r = fq_to_mont(t);
  return true;
}
BN_FN void fq_to_be(uint8_t* b, const fq& a) {
  fq t = fq_from_mont(a);
  u256_to_be(b, t.l);
}

// ------------------------------------------------------------------------------------------------ fixed exponents
// a^e for a public exponent e (8 limbs), 4-bit fixed windows, MSB first.  Control flow depends on e only.
BN_NOINLINE void fq_pow_pub(fq* r, const fq* a, const uint32_t* e) {
  fq tab[16];
  tab[0] = fq_one();
  tab[1] = *a;
  for (int i = 2; i < 16; i++) fq_mul_ni(&tab[i], &tab[i - 1], a);
  fq acc = fq_one();
  bool started = false;
  for (int w = 63; w >= 0; w--) {
    uint32_t nib = (e[w >> 3] >> ((w & 7) * 4)) & 15;
    if (started) {
      for (int k = 0; k < 4; k++) fq_mul_ni(&acc, &acc, &acc);
    }
    if (nib) {
      if (started) fq_mul_ni(&acc, &acc, &tab[nib]);
      else acc = tab[nib];
      started = true;
    }
  }
  *r = acc;
}
BN_FN fq fq_inv(const fq& a) {
  fq r;
  fq_pow_pub(&r, &a, K_EXP_QM2);
  return r;
}
// Fq::sqrt of the dependency: a1 = a^((q-3)/4); root = a1*a; reject iff a1*root == -1
BN_FN bool fq_sqrt(fq* root, const fq& a) {
  fq a1;
  fq_pow_pub(&a1, &a, K_EXP_QM3D4);
  fq rt = fq_mul(a1, a);
  fq chk = fq_mul(a1, rt);
  if (fq_eq(chk, fq_neg(fq_one()))) return false;
  *root = rt;
  return true;
}
// canonical parity of y (Montgomery in)
BN_FN uint32_t fq_parity(const fq& a) { return fq_from_mont(a).l[0] & 1; }

}  // namespace bn
