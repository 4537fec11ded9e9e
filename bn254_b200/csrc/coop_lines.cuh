// coop_lines.cuh -- per-item producer of the Miller-loop line sets consumed by the cooperative machine (coop.cuh).
//
// One thread walks the G2 point of its item along the signed digits of 6u+2 (doubling_step / mixed_addition_step of
// pairing.cuh) and writes, for every step m = 0..86, the sparse line of the variable pair (H(m) or the G1 generator,
// pk) scaled by the G1 point as line set 2m, and the precomputed line of the fixed pair (sig, -G2) scaled by sig as
// line set 2m+1.  A pair holding an infinity is skipped by bn::pairing_batch (/root/reference/src/ecdsa.rs:57): its
// line sets are the constant 1, which the sparse product leaves f unchanged with.
#pragma once
#include "coop.cuh"
#include "pairing.cuh"

namespace bn {

// Fq2 product / square / scaling used by the walk, as a policy:
//   lines_mul_call  the shared by-value routines of tower.cuh, one out-of-line Fq product each: smallest code and fewest live
//                   registers -- the throughput form (big batches: every sub-partition has four warps to interleave);
//   lines_mul_ilp   out-of-line Fq2 routines whose two or three Fq products are INLINED, so that their independent carry chains
//                   interleave: 25 % slower per item when the SM is full (A/B r02), but a lone warp -- a small batch -- is bound by
//                   the latency of its dependent chains, and three chains in flight hide it: the latency form.
#if defined(__CUDACC__)
#define BN_SFN static __device__ __forceinline__
#else
#define BN_SFN static inline
#endif
struct lines_mul_call {
  BN_SFN fq2 mul(const fq2& a, const fq2& b) { return fq2_mul_v(a, b); }
  BN_SFN fq2 sqr(const fq2& a) { return fq2_sqr_v(a); }
  BN_SFN fq2 scale(const fq2& a, const fq& k) { return fq2_scale_v(a, k); }
};
#if defined(__CUDA_ARCH__)
__device__ __noinline__ fq2 fq2_mul_ilp(fq2 a, fq2 b) {
  fq aa = fq_mul_inl(a.c0, b.c0);
  fq bb = fq_mul_inl(a.c1, b.c1);
  fq s = fq_mul_inl(fq_add(a.c0, a.c1), fq_add(b.c0, b.c1));
  fq2 r;
  r.c0 = fq_sub(aa, bb);
  r.c1 = fq_sub(fq_sub(s, aa), bb);
  return r;
}
__device__ __noinline__ fq2 fq2_sqr_ilp(fq2 a) {
  fq m = fq_mul_inl(a.c0, a.c1);
  fq2 r;
  r.c0 = fq_mul_inl(fq_add(a.c0, a.c1), fq_sub(a.c0, a.c1));
  r.c1 = fq_dbl(m);
  return r;
}
__device__ __noinline__ fq2 fq2_scale_ilp(fq2 a, fq k) {
  fq2 r;
  r.c0 = fq_mul_inl(a.c0, k);
  r.c1 = fq_mul_inl(a.c1, k);
  return r;
}
// the same out-of-line Fq2 routines with their Fq products pinned ONE AFTER THE OTHER (an empty asm makes an operand of the next
// product depend on the result of the previous one): one call per Fq2 operation instead of one per Fq product -- a third of the
// calling-convention moves -- without the register pressure of interleaved carry chains
#define BN_PIN_AFTER(x, after) asm volatile("" : "+r"((x).l[0]) : "r"((after).l[7]))
__device__ __noinline__ fq2 fq2_mul_seq(fq2 a, fq2 b) {
  fq aa = fq_mul_inl(a.c0, b.c0);
  BN_PIN_AFTER(a.c1, aa);
  fq bb = fq_mul_inl(a.c1, b.c1);
  fq sa = fq_add(a.c0, a.c1), sb = fq_add(b.c0, b.c1);
  BN_PIN_AFTER(sa, bb);
  fq s = fq_mul_inl(sa, sb);
  fq2 r;
  r.c0 = fq_sub(aa, bb);
  r.c1 = fq_sub(fq_sub(s, aa), bb);
  return r;
}
__device__ __noinline__ fq2 fq2_sqr_seq(fq2 a) {
  fq m = fq_mul_inl(a.c0, a.c1);
  fq sa = fq_add(a.c0, a.c1), da = fq_sub(a.c0, a.c1);
  BN_PIN_AFTER(sa, m);
  fq2 r;
  r.c0 = fq_mul_inl(sa, da);
  r.c1 = fq_dbl(m);
  return r;
}
__device__ __noinline__ fq2 fq2_scale_seq(fq2 a, fq k) {
  fq2 r;
  r.c0 = fq_mul_inl(a.c0, k);
  BN_PIN_AFTER(a.c1, r.c0);
  r.c1 = fq_mul_inl(a.c1, k);
  return r;
}
struct lines_mul_seq {
  BN_SFN fq2 mul(const fq2& a, const fq2& b) { return fq2_mul_seq(a, b); }
  BN_SFN fq2 sqr(const fq2& a) { return fq2_sqr_seq(a); }
  BN_SFN fq2 scale(const fq2& a, const fq& k) { return fq2_scale_seq(a, k); }
};
// Fq2 product with ONE reduction per coefficient (the machine's trick, coop.cuh coop_dot_block): three unreduced products in the
// lazy accumulator, split into plain halves, c0 = redc_low(L0 - L1) + H0 - H1, c1 = redc_low(L2 - L0 - L1) + H2 - H0 - H1 (+ 2 q, 3 q),
// nine-limb results through the table-driven reduction.  320 wide multiply steps instead of 384, one call instead of three.
__device__ __noinline__ fq2 fq2_mul_lazy(fq2 a, fq2 b) {
  const uint32_t* kq = &K_KQ_TABLE[0][0];
  uint64_t E[8], O[8];
  uint32_t C[8], L[8], H[8];
  fq l0, h0, l1, h1;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    E[i] = 0;
    O[i] = 0;
    C[i] = 0;
  }
  wide_mac(E, O, C, a.c0.l, b.c0.l);
  wide_split(E, O, C, L, H);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    l0.l[i] = L[i];
    h0.l[i] = H[i];
    E[i] = 0;
    O[i] = 0;
    C[i] = 0;
  }
  wide_mac(E, O, C, a.c1.l, b.c1.l);
  wide_split(E, O, C, L, H);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    l1.l[i] = L[i];
    h1.l[i] = H[i];
    E[i] = 0;
    O[i] = 0;
    C[i] = 0;
  }
  fq2 r;
  const fq zero = fq_zero();
  {
    const uint32_t k2[8] = BN_2Q_LIMBS;
    uint32_t d[9], t[9];
    fq9_addk_sub(d, l0, zero.l, l1);
    fq dl;
#pragma unroll
    for (int i = 0; i < 8; i++) dl.l[i] = d[i];
    const fq h = redc_low(dl.l);
    fq9_addk_sub(t, h0, k2, h1);
    fq9_add(t, h, d[8]);
    r.c0 = fq_reduce9(t, kq);
  }
  const uint32_t cy = fq_add_carry(l0, l0, l1);
  h0 = fq_add_raw(h0, h1);
  const fq sa = fq_add_raw(a.c0, a.c1), sb = fq_add_raw(b.c0, b.c1);
  wide_mac(E, O, C, sa.l, sb.l);
  wide_split(E, O, C, L, H);
  {
    const uint32_t k3[8] = BN_3Q_LIMBS;
    uint32_t d[9], t[9];
    fq l2, h2;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      l2.l[i] = L[i];
      h2.l[i] = H[i];
    }
    fq9_addk_sub(d, l2, zero.l, l0);
    fq dl;
#pragma unroll
    for (int i = 0; i < 8; i++) dl.l[i] = d[i];
    const fq h = redc_low(dl.l);
    fq9_addk_sub(t, h2, k3, h0);
    fq9_add(t, h, d[8] - cy);
    r.c1 = fq_reduce9(t, kq);
  }
  return r;
}
struct lines_mul_lazy {
  BN_SFN fq2 mul(const fq2& a, const fq2& b) { return fq2_mul_lazy(a, b); }
  BN_SFN fq2 sqr(const fq2& a) { return fq2_sqr_v(a); }
  BN_SFN fq2 scale(const fq2& a, const fq& k) { return fq2_scale_v(a, k); }
};
struct lines_mul_ilp {
  BN_SFN fq2 mul(const fq2& a, const fq2& b) { return fq2_mul_ilp(a, b); }
  BN_SFN fq2 sqr(const fq2& a) { return fq2_sqr_ilp(a); }
  BN_SFN fq2 scale(const fq2& a, const fq& k) { return fq2_scale_ilp(a, k); }
};
// everything inlined: the independent Fq2 products of a curve step (five in a doubling, up to four in an addition) can be
// interleaved by the compiler as well -- 100+ KB of straight-line code per step, only sensible for a warp that is alone
struct lines_mul_flat {
  BN_SFN fq2 mul(const fq2& a, const fq2& b) {
    fq aa = fq_mul_inl(a.c0, b.c0);
    fq bb = fq_mul_inl(a.c1, b.c1);
    fq s = fq_mul_inl(fq_add(a.c0, a.c1), fq_add(b.c0, b.c1));
    fq2 r;
    r.c0 = fq_sub(aa, bb);
    r.c1 = fq_sub(fq_sub(s, aa), bb);
    return r;
  }
  BN_SFN fq2 sqr(const fq2& a) {
    fq m = fq_mul_inl(a.c0, a.c1);
    fq2 r;
    r.c0 = fq_mul_inl(fq_add(a.c0, a.c1), fq_sub(a.c0, a.c1));
    r.c1 = fq_dbl(m);
    return r;
  }
  BN_SFN fq2 scale(const fq2& a, const fq& k) {
    fq2 r;
    r.c0 = fq_mul_inl(a.c0, k);
    r.c1 = fq_mul_inl(a.c1, k);
    return r;
  }
};
#else
typedef lines_mul_call lines_mul_ilp;
typedef lines_mul_call lines_mul_flat;
typedef lines_mul_call lines_mul_seq;
typedef lines_mul_call lines_mul_lazy;
#endif

BN_FN fq2 fq2_halve(const fq2& a) {
  fq2 r;
  r.c0 = fq_halve(a.c0);
  r.c1 = fq_halve(a.c1);
  return r;
}
// the two curve steps of pairing.cuh with every Fq2 value passed in registers (same formulas, same results; x / 2 as a shift)
template <class M>
BN_FN void doubling_step_v(fq2& rx, fq2& ry, fq2& rz, fq2& ell_0, fq2& ell_vw, fq2& ell_vv) {
  const fq2 twist_b = fq2_from_limbs(K_TWIST_B);
#if defined(BN_LINES_SCALE_HALF)  // the two halvings as products by 1/2 (pairing.cuh's form; four Fq products more per doubling)
  const fq two_inv = fq_from_limbs(K_TWO_INV);
  fq2 a = M::scale(M::mul(rx, ry), two_inv);
#else
  fq2 a = fq2_halve(M::mul(rx, ry));
#endif
  fq2 b = M::sqr(ry);
  fq2 cc = M::sqr(rz);
  fq2 e = M::mul(twist_b, fq2_add(fq2_dbl(cc), cc));
  fq2 f = fq2_add(fq2_dbl(e), e);
#if defined(BN_LINES_SCALE_HALF)
  fq2 g = M::scale(fq2_add(b, f), two_inv);
#else
  fq2 g = fq2_halve(fq2_add(b, f));
#endif
  fq2 h = fq2_sub(M::sqr(fq2_add(ry, rz)), fq2_add(b, cc));
  fq2 j = M::sqr(rx);
  fq2 e2 = M::sqr(e);
  rx = M::mul(a, fq2_sub(b, f));
  ry = fq2_sub(M::sqr(g), fq2_add(fq2_dbl(e2), e2));
  rz = M::mul(b, h);
  ell_0 = fq2_mul_xi(fq2_sub(e, b));
  ell_vw = fq2_neg(h);
  ell_vv = fq2_add(fq2_dbl(j), j);
}
template <class M>
BN_FN void mixed_addition_step_v(const fq2& qx, const fq2& qy, fq2& rx, fq2& ry, fq2& rz, fq2& ell_0, fq2& ell_vw, fq2& ell_vv) {
  fq2 d = fq2_sub(rx, M::mul(qx, rz));
  fq2 e = fq2_sub(ry, M::mul(qy, rz));
  fq2 f = M::sqr(d);
  fq2 g = M::sqr(e);
  fq2 h = M::mul(d, f);
  fq2 i = M::mul(rx, f);
  fq2 j = fq2_sub(fq2_add(h, M::mul(rz, g)), fq2_dbl(i));
  fq2 t = M::mul(h, ry);
  rx = M::mul(d, j);
  ry = fq2_sub(M::mul(e, fq2_sub(i, j)), t);
  rz = M::mul(rz, h);
  ell_0 = fq2_mul_xi(fq2_sub(M::mul(e, qx), M::mul(d, qy)));
  ell_vv = fq2_neg(e);
  ell_vw = d;
}
// (non-template names: the throughput form, used by every producer but the small-batch verify)
BN_FN void doubling_step_v(fq2& rx, fq2& ry, fq2& rz, fq2& ell_0, fq2& ell_vw, fq2& ell_vv) {
  doubling_step_v<lines_mul_call>(rx, ry, rz, ell_0, ell_vw, ell_vv);
}
BN_FN void mixed_addition_step_v(const fq2& qx, const fq2& qy, fq2& rx, fq2& ry, fq2& rz, fq2& ell_0, fq2& ell_vw, fq2& ell_vv) {
  mixed_addition_step_v<lines_mul_call>(qx, qy, rx, ry, rz, ell_0, ell_vw, ell_vv);
}

// per-thread constants of the walk, kept in shared memory on the device (registers are for the running point):
// [0] qx  [1] qy  [2] (hx, hy)  [3] (sx, sy)
struct lines_consts {
  fq2 v[4];
};

BN_NOINLINE void coop_emit_scaled_v(u4* lines, size_t set, size_t n_pad, size_t item, bool use, const fq2& ell_0, const fq2& ell_vw, const fq2& ell_vv,
                              const fq2& pxy) {
  fq2 l0, l3, l4;
  if (use) {
    l0 = ell_0;
    l3 = fq2_scale_v(ell_vw, pxy.c1);  // position c1.c1 = w^3, scaled by the G1 point's y
    l4 = fq2_scale_v(ell_vv, pxy.c0);  // position c0.c2 = w^4, scaled by the G1 point's x
  } else {
    l0 = fq2_one();
    l3 = fq2_zero();
    l4 = fq2_zero();
  }
  coop_emit_line(lines, set, n_pad, item, l0, l3, l4);
}

// returns the decode status of (sig, pk); on ST_OK all 174 line sets of the item are written.  K points at this thread's
// constants block (shared memory on the device).
// publish "the line sets of the first `steps` Miller steps of this item are in memory" to a consumer that runs concurrently
BN_FN void lines_publish(unsigned* progress, unsigned steps) {
#if defined(__CUDA_ARCH__)
  if (progress) {
    __threadfence();
    *(volatile unsigned*)progress = steps;
  }
#else
  (void)progress;
  (void)steps;
#endif
}
template <class M>
BN_FN int item_verify_lines_t(u4* lines, size_t n_pad, size_t item, const g1aff* h, const uint8_t* sig, const uint8_t* pk, const line_t* table,
                              lines_consts* K, unsigned* progress = nullptr) {
  bool use_a, use_b;
  {
    g2j q;
    g1j s;
    int st = g2_from_raw(&q, pk);
    if (st) return st;
    st = g1_from_raw(&s, sig);
    if (st) return st;
    use_a = !pt_is_inf(&q);
    use_b = !pt_is_inf(&s);
    K->v[0] = q.x;
    K->v[1] = q.y;
    K->v[2].c0 = h->x;
    K->v[2].c1 = h->y;
    K->v[3].c0 = s.x;
    K->v[3].c1 = s.y;
  }
  fq2 rx = K->v[0], ry = K->v[1], rz = fq2_one();
  fq2 c0 = fq2_one(), cvw = fq2_zero(), cvv = fq2_zero();
  size_t m = 0;
#pragma unroll 1
  for (int k = 0; k < 64; k++) {
    if (use_a) doubling_step_v<M>(rx, ry, rz, c0, cvw, cvv);
    coop_emit_scaled_v(lines, 2 * m, n_pad, item, use_a, c0, cvw, cvv, K->v[2]);
    coop_emit_scaled_v(lines, 2 * m + 1, n_pad, item, use_b, table[m].ell_0, table[m].ell_vw, table[m].ell_vv, K->v[3]);
    m++;
    lines_publish(progress, (unsigned)m);
    const int d = K_ATE_DIGITS[k];
    if (d != 0) {
      if (use_a) mixed_addition_step_v<M>(K->v[0], d > 0 ? K->v[1] : fq2_neg(K->v[1]), rx, ry, rz, c0, cvw, cvv);
      coop_emit_scaled_v(lines, 2 * m, n_pad, item, use_a, c0, cvw, cvv, K->v[2]);
      coop_emit_scaled_v(lines, 2 * m + 1, n_pad, item, use_b, table[m].ell_0, table[m].ell_vw, table[m].ell_vv, K->v[3]);
      m++;
      lines_publish(progress, (unsigned)m);
    }
  }
  fq2 q1x, q1y, q2x, q2y;
  g2_frobenius_pair(&q1x, &q1y, &q2x, &q2y, K->v[0], K->v[1]);
  if (use_a) mixed_addition_step_v<M>(q1x, q1y, rx, ry, rz, c0, cvw, cvv);
  coop_emit_scaled_v(lines, 2 * m, n_pad, item, use_a, c0, cvw, cvv, K->v[2]);
  coop_emit_scaled_v(lines, 2 * m + 1, n_pad, item, use_b, table[m].ell_0, table[m].ell_vw, table[m].ell_vv, K->v[3]);
  m++;
  lines_publish(progress, (unsigned)m);
  if (use_a) mixed_addition_step_v<M>(q2x, q2y, rx, ry, rz, c0, cvw, cvv);
  coop_emit_scaled_v(lines, 2 * m, n_pad, item, use_a, c0, cvw, cvv, K->v[2]);
  coop_emit_scaled_v(lines, 2 * m + 1, n_pad, item, use_b, table[m].ell_0, table[m].ell_vw, table[m].ell_vv, K->v[3]);
  return ST_OK;
}
#ifndef BN_LINES_POLICY
#define BN_LINES_POLICY lines_mul_call
#endif
BN_FN int item_verify_lines(u4* lines, size_t n_pad, size_t item, const g1aff* h, const uint8_t* sig, const uint8_t* pk, const line_t* table,
                            lines_consts* K) {
  return item_verify_lines_t<BN_LINES_POLICY>(lines, n_pad, item, h, sig, pk, table, K);
}

// ---------------------------------------------------------------------------------------------- cooperative walk (small batches)
// The walk above is one thread per item: 3083 Fq products one after the other, ~1.9 ms however small the batch.  Here FOUR warps
// (one per sub-partition of an SM) share the walk of 32 items -- lane = item, as in the machine -- and every curve step is cut
// into LEVELS of Fq2 operations that do not depend on each other, one operation (or two short ones) per warp and level, a block
// barrier between levels; all values live in shared memory (slots of one Fq2 per lane, the machine's conflict-free [Fq][chunk][lane]
// layout).  Same formulas, same canonical values as doubling_step_v / mixed_addition_step_v -- the line sets are bit-identical
// (tests/test_hostsim.py compares the two producers) -- but the dependent chain of a doubling is 9 Fq products instead of 38 and
// that of an addition 14 instead of 45.  The scaled lines of the fixed pair (sig, -G2) fill the idle slots of the first level.
//   doubling   level 0: w0 a = x y / 2, j = x^2   w1 b = y^2, fixed l3   w2 c = z^2, e = b' 3c   w3 s = (y+z)^2, fixed l4
//              level 1: w0 x' = a (b - f), l0     w1 y' = g^2 - 3 e^2    w2 z' = b h            w3 l3 = -h py, l4 = 3j px
//   addition   level 0: w0 d = x - qx z           w1 e = y - qy z        w2 fixed l3             w3 fixed l4
//              level 1: w0 f = d^2, h = d f       w1 g = e^2, z g        w2 e qx, l4 = -e px     w3 d qy, l3 = d py
//              level 2: w0 i = x f                w1 t = h y             w2 z' = z h             w3 l0 = xi (e qx - d qy)
//              level 3: w0 x' = d j               w1 y' = e (i - j) - t                          (j = h + z g - 2 i)
enum {
  WS_X, WS_Y, WS_Z, WS_QX, WS_QY, WS_NQY, WS_Q1X, WS_Q1Y, WS_Q2X, WS_Q2Y, WS_P, WS_SG,  // running point, the key and its images, (hx, hy), (sx, sy)
  WS_A, WS_B, WS_CC, WS_E, WS_S, WS_J,                                                   // doubling
  WS_D, WS_EE, WS_F, WS_RG, WS_H, WS_I, WS_T, WS_EQX, WS_DQY,                            // addition
  WS_SLOTS
};
#define WALK_WARPS 4
#define WALK_SMEM_BYTES (WS_SLOTS * 2 * 2 * COOP_LANES * 16 + 4 * COOP_LANES * 4) /* slots + per-lane decode results */
struct walk_ctx {
  u4* sm;         // the lane's column of the slot array
  int* flags;     // [0][lane] status of the key, [1][lane] status of the signature, [2][lane] key at infinity, [3][lane] signature at infinity
  int row, lane, warp;
  size_t item, n, n_pad;
  u4* lines;
  const line_t* table;
  bool live, use_a, use_b;  // (set after walk_decode's barrier)
};
BN_FN fq2 walk_ld(const walk_ctx& c, int s) {
  fq2 r;
  r.c0 = coop_ld(c.sm, 2 * s, c.row);
  r.c1 = coop_ld(c.sm, 2 * s + 1, c.row);
  return r;
}
BN_FN void walk_st(const walk_ctx& c, int s, const fq2& v) {
  coop_st(c.sm, 2 * s, c.row, v.c0);
  coop_st(c.sm, 2 * s + 1, c.row, v.c1);
}
// coefficient `which` (0: l0, 1: l3, 2: l4) of line set `set`, as the triple (x0, x1, x0 + x1) coop_emit_line writes
BN_FN void walk_emit(const walk_ctx& c, size_t set, int which, bool use, fq2 v) {
  if (!c.live) return;
  if (!use) {
    v = fq2_zero();
    if (which == 0) v = fq2_one();
  }
  const size_t r = set * COOP_LINE_FQ + 3 * which;
  coop_gst(c.lines, r + 0, c.n_pad, c.item, v.c0);
  coop_gst(c.lines, r + 1, c.n_pad, c.item, v.c1);
  coop_gst(c.lines, r + 2, c.n_pad, c.item, fq_add(v.c0, v.c1));
}
// half of the fixed pair's line set 2m + 1: warp `half` = 0 writes l0 and l3 = ell_vw sy, 1 writes l4 = ell_vv sx
template <class M>
BN_FN void walk_fixed(const walk_ctx& c, size_t m, int half) {
  const fq2 sg = walk_ld(c, WS_SG);
  if (half == 0) {
    walk_emit(c, 2 * m + 1, 0, c.use_b, c.table[m].ell_0);
    walk_emit(c, 2 * m + 1, 1, c.use_b, M::scale(c.table[m].ell_vw, sg.c1));
  } else {
    walk_emit(c, 2 * m + 1, 2, c.use_b, M::scale(c.table[m].ell_vv, sg.c0));
  }
}
// decode: warp 0 the key (and its two Frobenius images), warp 1 the signature and the G1 point of the variable pair
BN_FN void walk_decode(const walk_ctx& c, const g1aff* h, const uint8_t* sig, const uint8_t* pk, bool skip) {
  if (c.warp == 0) {
    g2j q;
    int st = skip ? ST_OK : g2_from_raw(&q, pk);
    if (skip || st) pt_set_inf(&q);
    c.flags[0 * COOP_LANES + c.lane] = st;
    c.flags[2 * COOP_LANES + c.lane] = pt_is_inf(&q) ? 1 : 0;
    fq2 q1x, q1y, q2x, q2y;
    g2_frobenius_pair(&q1x, &q1y, &q2x, &q2y, q.x, q.y);
    walk_st(c, WS_QX, q.x);
    walk_st(c, WS_QY, q.y);
    walk_st(c, WS_NQY, fq2_neg(q.y));
    walk_st(c, WS_Q1X, q1x);
    walk_st(c, WS_Q1Y, q1y);
    walk_st(c, WS_Q2X, q2x);
    walk_st(c, WS_Q2Y, q2y);
    walk_st(c, WS_X, q.x);
    walk_st(c, WS_Y, q.y);
    walk_st(c, WS_Z, fq2_one());
  } else if (c.warp == 1) {
    g1j s;
    int st = skip ? ST_OK : g1_from_raw(&s, sig);
    if (skip || st) pt_set_inf(&s);
    c.flags[1 * COOP_LANES + c.lane] = st;
    c.flags[3 * COOP_LANES + c.lane] = pt_is_inf(&s) ? 1 : 0;
    fq2 t;
    t.c0 = s.x;
    t.c1 = s.y;
    walk_st(c, WS_SG, t);
    if (!skip) {
      t.c0 = h->x;
      t.c1 = h->y;
    }
    walk_st(c, WS_P, t);
  }
}
// status of the item as item_verify_lines returns it (key first); fills the context's flags.  Call after the barrier behind walk_decode.
BN_FN int walk_flags(walk_ctx& c, bool skip) {
  const int st_pk = c.flags[0 * COOP_LANES + c.lane], st_sig = c.flags[1 * COOP_LANES + c.lane];
  const int st = st_pk ? st_pk : st_sig;
  c.live = !skip && st == ST_OK;
  c.use_a = c.flags[2 * COOP_LANES + c.lane] == 0;
  c.use_b = c.flags[3 * COOP_LANES + c.lane] == 0;
  return st;
}
template <class M>
BN_FN void walk_dbl(const walk_ctx& c, int level, size_t m) {
  if (level == 0) {
    if (c.warp == 0) {
      const fq2 x = walk_ld(c, WS_X);
      walk_st(c, WS_A, fq2_halve(M::mul(x, walk_ld(c, WS_Y))));
      walk_st(c, WS_J, M::sqr(x));
    } else if (c.warp == 1) {
      walk_st(c, WS_B, M::sqr(walk_ld(c, WS_Y)));
      walk_fixed<M>(c, m, 0);
    } else if (c.warp == 2) {
      const fq2 cc = M::sqr(walk_ld(c, WS_Z));
      walk_st(c, WS_CC, cc);
      walk_st(c, WS_E, M::mul(fq2_from_limbs(K_TWIST_B), fq2_add(fq2_dbl(cc), cc)));
    } else {
      walk_st(c, WS_S, M::sqr(fq2_add(walk_ld(c, WS_Y), walk_ld(c, WS_Z))));
      walk_fixed<M>(c, m, 1);
    }
    return;
  }
  const fq2 b = walk_ld(c, WS_B);
  if (c.warp == 0) {
    const fq2 e = walk_ld(c, WS_E);
    const fq2 f = fq2_add(fq2_dbl(e), e);
    walk_st(c, WS_X, M::mul(walk_ld(c, WS_A), fq2_sub(b, f)));
    walk_emit(c, 2 * m, 0, c.use_a, fq2_mul_xi(fq2_sub(e, b)));
  } else if (c.warp == 1) {
    const fq2 e = walk_ld(c, WS_E);
    const fq2 f = fq2_add(fq2_dbl(e), e);
    const fq2 g = fq2_halve(fq2_add(b, f));
    const fq2 e2 = M::sqr(e);
    walk_st(c, WS_Y, fq2_sub(M::sqr(g), fq2_add(fq2_dbl(e2), e2)));
  } else {
    const fq2 h = fq2_sub(walk_ld(c, WS_S), fq2_add(b, walk_ld(c, WS_CC)));
    if (c.warp == 2) {
      walk_st(c, WS_Z, M::mul(b, h));
    } else {
      const fq2 p = walk_ld(c, WS_P);
      const fq2 j = walk_ld(c, WS_J);
      walk_emit(c, 2 * m, 1, c.use_a, M::scale(fq2_neg(h), p.c1));
      walk_emit(c, 2 * m, 2, c.use_a, M::scale(fq2_add(fq2_dbl(j), j), p.c0));
    }
  }
}
template <class M>
BN_FN void walk_add(const walk_ctx& c, int level, size_t m, int sqx, int sqy) {
  switch (level) {
    case 0:
      if (c.warp == 0)
        walk_st(c, WS_D, fq2_sub(walk_ld(c, WS_X), M::mul(walk_ld(c, sqx), walk_ld(c, WS_Z))));
      else if (c.warp == 1)
        walk_st(c, WS_EE, fq2_sub(walk_ld(c, WS_Y), M::mul(walk_ld(c, sqy), walk_ld(c, WS_Z))));
      else
        walk_fixed<M>(c, m, c.warp - 2);
      break;
    case 1:
      if (c.warp == 0) {
        const fq2 d = walk_ld(c, WS_D);
        const fq2 f = M::sqr(d);
        walk_st(c, WS_F, f);
        walk_st(c, WS_H, M::mul(d, f));
      } else if (c.warp == 1) {
        walk_st(c, WS_RG, M::mul(walk_ld(c, WS_Z), M::sqr(walk_ld(c, WS_EE))));
      } else if (c.warp == 2) {
        const fq2 e = walk_ld(c, WS_EE);
        walk_st(c, WS_EQX, M::mul(e, walk_ld(c, sqx)));
        walk_emit(c, 2 * m, 2, c.use_a, M::scale(fq2_neg(e), walk_ld(c, WS_P).c0));
      } else {
        const fq2 d = walk_ld(c, WS_D);
        walk_st(c, WS_DQY, M::mul(d, walk_ld(c, sqy)));
        walk_emit(c, 2 * m, 1, c.use_a, M::scale(d, walk_ld(c, WS_P).c1));
      }
      break;
    case 2:
      if (c.warp == 0)
        walk_st(c, WS_I, M::mul(walk_ld(c, WS_X), walk_ld(c, WS_F)));
      else if (c.warp == 1)
        walk_st(c, WS_T, M::mul(walk_ld(c, WS_H), walk_ld(c, WS_Y)));
      else if (c.warp == 2)
        walk_st(c, WS_Z, M::mul(walk_ld(c, WS_Z), walk_ld(c, WS_H)));
      else
        walk_emit(c, 2 * m, 0, c.use_a, fq2_mul_xi(fq2_sub(walk_ld(c, WS_EQX), walk_ld(c, WS_DQY))));
      break;
    default:
      if (c.warp < 2) {
        const fq2 i = walk_ld(c, WS_I);
        const fq2 j = fq2_sub(fq2_add(walk_ld(c, WS_H), walk_ld(c, WS_RG)), fq2_dbl(i));
        if (c.warp == 0)
          walk_st(c, WS_X, M::mul(walk_ld(c, WS_D), j));
        else
          walk_st(c, WS_Y, fq2_sub(M::mul(walk_ld(c, WS_EE), fq2_sub(i, j)), walk_ld(c, WS_T)));
      }
      break;
  }
}
// The schedule of the 87 steps: step(kind, level, m, sqx, sqy) for every level, bar() behind every level, done(m) once step m - 1
// is complete (all of its line-set stores are before the last bar()).
template <class Step, class Bar, class Done>
BN_FN void walk_schedule(Step step, Bar bar, Done done) {
  size_t m = 0;
  auto dbl = [&] {
    for (int l = 0; l < 2; l++) {
      step(0, l, m, 0, 0);
      bar();
    }
    done(++m);
  };
  auto add = [&](int sqx, int sqy) {
    for (int l = 0; l < 4; l++) {
      step(1, l, m, sqx, sqy);
      bar();
    }
    done(++m);
  };
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int k = 0; k < 64; k++) {
    dbl();
    const int d = K_ATE_DIGITS[k];
    if (d != 0) add(WS_QX, d > 0 ? WS_QY : WS_NQY);
  }
  add(WS_Q1X, WS_Q1Y);
  add(WS_Q2X, WS_Q2Y);
}

// Multi-pairing producer: pair (h, pk) is stream `stream` of its lane `item`; its 87 line sets go to set index
// step * mk + stream (mk = pairs per lane of the program that will consume them).  use == false (padding slot, pk at infinity, or an item that failed to decode) writes the
// constant 1 everywhere, which leaves the lane's product unchanged.
BN_FN void item_pair_lines(u4* lines, size_t n_pad, size_t item, int stream, int mk, bool use, const g1aff* h, const fq2& qx, const fq2& qy,
                           lines_consts* K) {
  K->v[0] = qx;
  K->v[1] = qy;
  if (use) {
    K->v[2].c0 = h->x;
    K->v[2].c1 = h->y;
  }
  fq2 rx = qx, ry = qy, rz = fq2_one();
  fq2 c0 = fq2_one(), cvw = fq2_zero(), cvv = fq2_zero();
  size_t m = 0;
#pragma unroll 1
  for (int k = 0; k < 64; k++) {
    if (use) doubling_step_v(rx, ry, rz, c0, cvw, cvv);
    coop_emit_scaled_v(lines, m * mk + stream, n_pad, item, use, c0, cvw, cvv, K->v[2]);
    m++;
    const int d = K_ATE_DIGITS[k];
    if (d != 0) {
      if (use) mixed_addition_step_v(K->v[0], d > 0 ? K->v[1] : fq2_neg(K->v[1]), rx, ry, rz, c0, cvw, cvv);
      coop_emit_scaled_v(lines, m * mk + stream, n_pad, item, use, c0, cvw, cvv, K->v[2]);
      m++;
    }
  }
  fq2 q1x, q1y, q2x, q2y;
  g2_frobenius_pair(&q1x, &q1y, &q2x, &q2y, K->v[0], K->v[1]);
  if (use) mixed_addition_step_v(q1x, q1y, rx, ry, rz, c0, cvw, cvv);
  coop_emit_scaled_v(lines, m * mk + stream, n_pad, item, use, c0, cvw, cvv, K->v[2]);
  m++;
  if (use) mixed_addition_step_v(q2x, q2y, rx, ry, rz, c0, cvw, cvv);
  coop_emit_scaled_v(lines, m * mk + stream, n_pad, item, use, c0, cvw, cvv, K->v[2]);
}

// ---------------------------------------------------------------------------------------------- cached key lines
// A public key that verifies many messages (a fixed validator set) needs its walk along the twist only ONCE: the 87 line
// coefficient triples (ell_0, ell_vw, ell_vv) do not depend on the message.  Layout of the cache (16-byte chunks, key-contiguous):
//   klines [87 sets][6 Fq][2 chunks][k_pad]      6 Fq = ell_0.c0, ell_0.c1, ell_vw.c0, ell_vw.c1, ell_vv.c0, ell_vv.c1
// 16 704 bytes per key.  A key at infinity is stored as the constant line 1 (its pair is skipped, as bn::pairing_batch does).
#define COOP_KLINE_FQ 6
BN_FN void coop_emit_key_line(u4* klines, size_t set, size_t k_pad, size_t key, bool use, const fq2& ell_0, const fq2& ell_vw, const fq2& ell_vv) {
  const size_t r = set * COOP_KLINE_FQ;
  fq2 a = ell_0, b = ell_vw, c = ell_vv;
  if (!use) {
    a = fq2_one();
    b = fq2_zero();
    c = fq2_zero();
  }
  coop_gst(klines, r + 0, k_pad, key, a.c0);
  coop_gst(klines, r + 1, k_pad, key, a.c1);
  coop_gst(klines, r + 2, k_pad, key, b.c0);
  coop_gst(klines, r + 3, k_pad, key, b.c1);
  coop_gst(klines, r + 4, k_pad, key, c.c0);
  coop_gst(klines, r + 5, k_pad, key, c.c1);
}
// the walk of one key (affine twist point, or `use` == false for a key at infinity / a padding slot)
BN_FN void item_key_lines(u4* klines, size_t k_pad, size_t key, bool use, const fq2& qx, const fq2& qy, lines_consts* K) {
  K->v[0] = qx;
  K->v[1] = qy;
  fq2 rx = qx, ry = qy, rz = fq2_one();
  fq2 c0 = fq2_one(), cvw = fq2_zero(), cvv = fq2_zero();
  size_t m = 0;
#pragma unroll 1
  for (int k = 0; k < 64; k++) {
    if (use) doubling_step_v(rx, ry, rz, c0, cvw, cvv);
    coop_emit_key_line(klines, m, k_pad, key, use, c0, cvw, cvv);
    m++;
    const int d = K_ATE_DIGITS[k];
    if (d != 0) {
      if (use) mixed_addition_step_v(K->v[0], d > 0 ? K->v[1] : fq2_neg(K->v[1]), rx, ry, rz, c0, cvw, cvv);
      coop_emit_key_line(klines, m, k_pad, key, use, c0, cvw, cvv);
      m++;
    }
  }
  fq2 q1x, q1y, q2x, q2y;
  g2_frobenius_pair(&q1x, &q1y, &q2x, &q2y, K->v[0], K->v[1]);
  if (use) mixed_addition_step_v(q1x, q1y, rx, ry, rz, c0, cvw, cvv);
  coop_emit_key_line(klines, m, k_pad, key, use, c0, cvw, cvv);
  m++;
  if (use) mixed_addition_step_v(q2x, q2y, rx, ry, rz, c0, cvw, cvv);
  coop_emit_key_line(klines, m, k_pad, key, use, c0, cvw, cvv);
}
// line sets 2 m and 2 m + 1 of one item from the cache: the key's line m scaled by H(msg), the -G2 table's line m scaled by sig
BN_FN void item_scale_cached_lines(u4* lines, size_t n_pad, size_t item, int m, const u4* klines, size_t k_pad, size_t key, const g1aff& h,
                                   bool use_b, const fq& sx, const fq& sy, const line_t* table) {
  const size_t r = (size_t)m * COOP_KLINE_FQ;
  fq2 e0, evw, evv;
  e0.c0 = coop_gld(klines, r + 0, k_pad, key);
  e0.c1 = coop_gld(klines, r + 1, k_pad, key);
  evw.c0 = coop_gld(klines, r + 2, k_pad, key);
  evw.c1 = coop_gld(klines, r + 3, k_pad, key);
  evv.c0 = coop_gld(klines, r + 4, k_pad, key);
  evv.c1 = coop_gld(klines, r + 5, k_pad, key);
  fq2 l3, l4;
  l3.c0 = fq_mul(evw.c0, h.y);
  l3.c1 = fq_mul(evw.c1, h.y);
  l4.c0 = fq_mul(evv.c0, h.x);
  l4.c1 = fq_mul(evv.c1, h.x);
  coop_emit_line(lines, 2 * (size_t)m, n_pad, item, e0, l3, l4);
  fq2 b0 = fq2_one(), b3 = fq2_zero(), b4 = fq2_zero();
  if (use_b) {
    b0 = table[m].ell_0;
    b3.c0 = fq_mul(table[m].ell_vw.c0, sy);
    b3.c1 = fq_mul(table[m].ell_vw.c1, sy);
    b4.c0 = fq_mul(table[m].ell_vv.c0, sx);
    b4.c1 = fq_mul(table[m].ell_vv.c1, sx);
  }
  coop_emit_line(lines, 2 * (size_t)m + 1, n_pad, item, b0, b3, b4);
}

}  // namespace bn
