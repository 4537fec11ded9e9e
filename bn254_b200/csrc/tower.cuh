// tower.cuh -- Fq2 = Fq[i]/(i^2+1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v), xi = 9 + i.
//
// Replaces the Fq2/Fq6/Fq12 tower of the reference's arithmetic dependency (crate zeropool-bn 0.5.11,
// /root/reference/Cargo.toml:24; reached through bn::pairing_batch at /root/reference/src/ecdsa.rs:57,86 and
// Fq2 at /root/reference/src/utils.rs:41-42,111-112,140-141).  One thread owns one tower element; Fq-level
// values live in registers, Fq2 and larger operands are passed by pointer (thread-local / shared memory), and
// the multiplication-heavy routines are out-of-line so the pairing kernels stay within the instruction cache.
#pragma once
#include "fq.cuh"

namespace bn {

struct alignas(16) fq2 {
  fq c0, c1;
};
struct alignas(16) fq6 {
  fq2 c0, c1, c2;
};
struct alignas(16) fq12 {
  fq6 c0, c1;
};

BN_FN fq2 fq2_from_limbs(const uint32_t* p) {
  fq2 r;
  r.c0 = fq_from_limbs(p);
  r.c1 = fq_from_limbs(p + 8);
  return r;
}
BN_FN fq2 fq2_zero() {
  fq2 r;
  r.c0 = fq_zero();
  r.c1 = fq_zero();
  return r;
}
BN_FN fq2 fq2_one() {
  fq2 r;
  r.c0 = fq_one();
  r.c1 = fq_zero();
  return r;
}
// Fq2 additions are out-of-line on the device (by-value operands travel in registers): the tower and the curve
// steps use hundreds of them, and inlining every one made the pairing kernels larger than the instruction cache.
#if defined(__CUDA_ARCH__) && !defined(BN254_INLINE_ADD)
#define BN_FQ2_LEAF __device__ __noinline__
#else
#define BN_FQ2_LEAF BN_FN
#endif
BN_FQ2_LEAF fq2 fq2_add_v(fq2 a, fq2 b) {
  fq2 r;
  r.c0 = fq_add(a.c0, b.c0);
  r.c1 = fq_add(a.c1, b.c1);
  return r;
}
BN_FQ2_LEAF fq2 fq2_sub_v(fq2 a, fq2 b) {
  fq2 r;
  r.c0 = fq_sub(a.c0, b.c0);
  r.c1 = fq_sub(a.c1, b.c1);
  return r;
}
BN_FQ2_LEAF fq2 fq2_dbl_v(fq2 a) {
  fq2 r;
  r.c0 = fq_add(a.c0, a.c0);
  r.c1 = fq_add(a.c1, a.c1);
  return r;
}
BN_FQ2_LEAF fq2 fq2_neg_v(fq2 a) {
  fq2 r;
  r.c0 = fq_neg(a.c0);
  r.c1 = fq_neg(a.c1);
  return r;
}
// multiply by xi = 9 + i : (9 a0 - a1) + (9 a1 + a0) i
BN_FQ2_LEAF fq2 fq2_mul_xi_v(fq2 a) {
  fq t0 = fq_dbl(a.c0);
  t0 = fq_dbl(t0);
  t0 = fq_dbl(t0);
  t0 = fq_add(t0, a.c0);
  fq t1 = fq_dbl(a.c1);
  t1 = fq_dbl(t1);
  t1 = fq_dbl(t1);
  t1 = fq_add(t1, a.c1);
  fq2 r;
  r.c0 = fq_sub(t0, a.c1);
  r.c1 = fq_add(t1, a.c0);
  return r;
}
BN_FN fq2 fq2_add(const fq2& a, const fq2& b) { return fq2_add_v(a, b); }
BN_FN fq2 fq2_sub(const fq2& a, const fq2& b) { return fq2_sub_v(a, b); }
BN_FN fq2 fq2_dbl(const fq2& a) { return fq2_dbl_v(a); }
BN_FN fq2 fq2_neg(const fq2& a) { return fq2_neg_v(a); }
BN_FN fq2 fq2_mul_xi(const fq2& a) { return fq2_mul_xi_v(a); }
BN_FN fq2 fq2_conj(const fq2& a) {
  fq2 r;
  r.c0 = a.c0;
  r.c1 = fq_neg(a.c1);
  return r;
}
BN_FN bool fq2_is_zero(const fq2& a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }
BN_FN bool fq2_eq(const fq2& a, const fq2& b) { return fq_eq(a.c0, b.c0) && fq_eq(a.c1, b.c1); }

// Karatsuba: 3 Fq products.  They go through the single out-of-line product: inlining them here (3 interleaved
// carry chains) was measured slower, because the hot loop then no longer fits the instruction cache
BN_NOINLINE void fq2_mul(fq2* r, const fq2* a, const fq2* b) {
  fq a0 = a->c0, a1 = a->c1, b0 = b->c0, b1 = b->c1;
  fq sa = fq_add(a0, a1), sb = fq_add(b0, b1);
  fq aa = fq_mul(a0, b0);
  fq bb = fq_mul(a1, b1);
  fq s = fq_mul(sa, sb);
  r->c0 = fq_sub(aa, bb);
  r->c1 = fq_sub(fq_sub(s, aa), bb);
}
// complex squaring: 2 Fq products
BN_NOINLINE void fq2_sqr(fq2* r, const fq2* a) {
  fq a0 = a->c0, a1 = a->c1;
  fq sa = fq_add(a0, a1), da = fq_sub(a0, a1);
  fq m = fq_mul(a0, a1);
  r->c0 = fq_mul(sa, da);
  r->c1 = fq_dbl(m);
}
BN_NOINLINE void fq2_scale(fq2* r, const fq2* a, const fq* s) {
  fq k = *s;
  fq x = fq_mul(a->c0, k);
  fq y = fq_mul(a->c1, k);
  r->c0 = x;
  r->c1 = y;
}
// by-value forms (register ABI on the device): for callers that keep their working set in registers
BN_FQ2_LEAF fq2 fq2_mul_v(fq2 a, fq2 b) {
  fq aa = fq_mul(a.c0, b.c0);
  fq bb = fq_mul(a.c1, b.c1);
  fq s = fq_mul(fq_add(a.c0, a.c1), fq_add(b.c0, b.c1));
  fq2 r;
  r.c0 = fq_sub(aa, bb);
  r.c1 = fq_sub(fq_sub(s, aa), bb);
  return r;
}
BN_FQ2_LEAF fq2 fq2_sqr_v(fq2 a) {
  fq m = fq_mul(a.c0, a.c1);
  fq2 r;
  r.c0 = fq_mul(fq_add(a.c0, a.c1), fq_sub(a.c0, a.c1));
  r.c1 = fq_dbl(m);
  return r;
}
BN_FQ2_LEAF fq2 fq2_scale_v(fq2 a, fq k) {
  fq2 r;
  r.c0 = fq_mul(a.c0, k);
  r.c1 = fq_mul(a.c1, k);
  return r;
}
// value-returning forms: operands are copied into call-local slots (by-value parameters) so that every memory
// temporary lives only around its own call
BN_FN fq2 fq2_mulv(fq2 a, fq2 b) {
  fq2 r;
  fq2_mul(&r, &a, &b);
  return r;
}
BN_FN fq2 fq2_sqrv(fq2 a) {
  fq2 r;
  fq2_sqr(&r, &a);
  return r;
}
BN_NOINLINE void fq2_inv(fq2* r, const fq2* a) {
  fq n = fq_add(fq_sqr(a->c0), fq_sqr(a->c1));
  n = fq_inv(n);
  fq t = fq_mul(a->c1, n);
  r->c0 = fq_mul(a->c0, n);
  r->c1 = fq_neg(t);
}

// ------------------------------------------------------------------------------------------------ Fq6
BN_FN void fq6_add(fq6* r, const fq6* a, const fq6* b) {
  r->c0 = fq2_add(a->c0, b->c0);
  r->c1 = fq2_add(a->c1, b->c1);
  r->c2 = fq2_add(a->c2, b->c2);
}
BN_FN void fq6_sub(fq6* r, const fq6* a, const fq6* b) {
  r->c0 = fq2_sub(a->c0, b->c0);
  r->c1 = fq2_sub(a->c1, b->c1);
  r->c2 = fq2_sub(a->c2, b->c2);
}
BN_FN void fq6_neg(fq6* r, const fq6* a) {
  r->c0 = fq2_neg(a->c0);
  r->c1 = fq2_neg(a->c1);
  r->c2 = fq2_neg(a->c2);
}
// (c0, c1, c2) * v = (xi c2, c0, c1)
BN_FN void fq6_mul_by_v(fq6* r, const fq6* a) {
  fq2 t = fq2_mul_xi(a->c2), c0 = a->c0, c1 = a->c1;
  r->c0 = t;
  r->c1 = c0;
  r->c2 = c1;
}
// Karatsuba over Fq2: 6 Fq2 products
BN_NOINLINE void fq6_mul(fq6* r, const fq6* a, const fq6* b) {
  fq2 aa, bb, cc, t1, t2, t3, s, u;
  fq2_mul(&aa, &a->c0, &b->c0);
  fq2_mul(&bb, &a->c1, &b->c1);
  fq2_mul(&cc, &a->c2, &b->c2);
  s = fq2_add(a->c1, a->c2);
  u = fq2_add(b->c1, b->c2);
  fq2_mul(&t1, &s, &u);
  s = fq2_add(a->c0, a->c1);
  u = fq2_add(b->c0, b->c1);
  fq2_mul(&t2, &s, &u);
  s = fq2_add(a->c0, a->c2);
  u = fq2_add(b->c0, b->c2);
  fq2_mul(&t3, &s, &u);
  // c0 = aa + xi((a1+a2)(b1+b2) - bb - cc); c1 = (a0+a1)(b0+b1) - aa - bb + xi cc; c2 = (a0+a2)(b0+b2) - aa - cc + bb
  t1 = fq2_sub(fq2_sub(t1, bb), cc);
  r->c0 = fq2_add(fq2_mul_xi(t1), aa);
  t2 = fq2_sub(fq2_sub(t2, aa), bb);
  r->c1 = fq2_add(t2, fq2_mul_xi(cc));
  t3 = fq2_sub(fq2_sub(t3, aa), cc);
  r->c2 = fq2_add(t3, bb);
}
BN_NOINLINE void fq6_inv(fq6* r, const fq6* a) {
  fq2 c0, c1, c2, t, t2;
  // c0 = a0^2 - xi a1 a2 ; c1 = xi a2^2 - a0 a1 ; c2 = a1^2 - a0 a2
  fq2_sqr(&c0, &a->c0);
  fq2_mul(&t, &a->c1, &a->c2);
  c0 = fq2_sub(c0, fq2_mul_xi(t));
  fq2_sqr(&c1, &a->c2);
  fq2_mul(&t, &a->c0, &a->c1);
  c1 = fq2_sub(fq2_mul_xi(c1), t);
  fq2_sqr(&c2, &a->c1);
  fq2_mul(&t, &a->c0, &a->c2);
  c2 = fq2_sub(c2, t);
  // n = a0 c0 + xi (a2 c1 + a1 c2)
  fq2_mul(&t, &a->c2, &c1);
  fq2_mul(&t2, &a->c1, &c2);
  t = fq2_mul_xi(fq2_add(t, t2));
  fq2_mul(&t2, &a->c0, &c0);
  t = fq2_add(t, t2);
  fq2_inv(&t, &t);
  fq2_mul(&r->c0, &c0, &t);
  fq2_mul(&r->c1, &c1, &t);
  fq2_mul(&r->c2, &c2, &t);
}

// ------------------------------------------------------------------------------------------------ Fq12
BN_FN void fq12_set_one(fq12* r) {
  r->c0.c0 = fq2_one();
  r->c0.c1 = fq2_zero();
  r->c0.c2 = fq2_zero();
  r->c1.c0 = fq2_zero();
  r->c1.c1 = fq2_zero();
  r->c1.c2 = fq2_zero();
}
BN_FN bool fq12_is_one(const fq12* a) {
  return fq2_eq(a->c0.c0, fq2_one()) && fq2_is_zero(a->c0.c1) && fq2_is_zero(a->c0.c2) && fq2_is_zero(a->c1.c0) &&
         fq2_is_zero(a->c1.c1) && fq2_is_zero(a->c1.c2);
}
BN_FN bool fq12_is_zero(const fq12* a) {
  return fq2_is_zero(a->c0.c0) && fq2_is_zero(a->c0.c1) && fq2_is_zero(a->c0.c2) && fq2_is_zero(a->c1.c0) &&
         fq2_is_zero(a->c1.c1) && fq2_is_zero(a->c1.c2);
}
BN_NOINLINE void fq12_mul(fq12* r, const fq12* a, const fq12* b) {
  fq6 aa, bb, s, t;
  fq6_mul(&aa, &a->c0, &b->c0);
  fq6_mul(&bb, &a->c1, &b->c1);
  fq6_add(&s, &a->c0, &a->c1);
  fq6_add(&t, &b->c0, &b->c1);
  fq6_mul(&s, &s, &t);
  fq6_sub(&s, &s, &aa);
  fq6_sub(&r->c1, &s, &bb);
  fq6_mul_by_v(&bb, &bb);
  fq6_add(&r->c0, &aa, &bb);
}
// complex squaring: c0 = (a0 + a1)(a0 + v a1) - ab - v ab ; c1 = 2 ab
BN_NOINLINE void fq12_sqr(fq12* r, const fq12* a) {
  fq6 ab, s, t;
  fq6_mul(&ab, &a->c0, &a->c1);
  fq6_add(&s, &a->c0, &a->c1);
  fq6_mul_by_v(&t, &a->c1);
  fq6_add(&t, &t, &a->c0);
  fq6_mul(&s, &s, &t);
  fq6_sub(&s, &s, &ab);
  fq6_mul_by_v(&t, &ab);
  fq6_sub(&r->c0, &s, &t);
  fq6_add(&r->c1, &ab, &ab);
}
// unitary inverse (conjugation over Fq6)
BN_FN void fq12_conj(fq12* r, const fq12* a) {
  r->c0 = a->c0;
  fq6_neg(&r->c1, &a->c1);
}
BN_NOINLINE void fq12_inv(fq12* r, const fq12* a) {
  fq6 t0, t1;
  fq6_mul(&t0, &a->c0, &a->c0);
  fq6_mul(&t1, &a->c1, &a->c1);
  fq6_mul_by_v(&t1, &t1);
  fq6_sub(&t0, &t0, &t1);
  fq6_inv(&t0, &t0);
  fq6_mul(&r->c0, &a->c0, &t0);
  fq6_mul(&t1, &a->c1, &t0);
  fq6_neg(&r->c1, &t1);
}
// Frobenius x -> x^(q^k), k = 1, 2, 3
BN_NOINLINE void fq12_frobenius(fq12* r, const fq12* a, int k) {
  const uint32_t *c61, *c62, *c12;
  if (k == 1) { c61 = K_FROB6_C1_1; c62 = K_FROB6_C2_1; c12 = K_FROB12_C1_1; }
  else if (k == 2) { c61 = K_FROB6_C1_2; c62 = K_FROB6_C2_2; c12 = K_FROB12_C1_2; }
  else { c61 = K_FROB6_C1_3; c62 = K_FROB6_C2_3; c12 = K_FROB12_C1_3; }
  fq2 k61 = fq2_from_limbs(c61), k62 = fq2_from_limbs(c62), k12 = fq2_from_limbs(c12);
  fq2 t, kk;
  bool odd = (k & 1) != 0;
  // c0 part
  t = a->c0.c0; if (odd) t = fq2_conj(t); r->c0.c0 = t;
  t = a->c0.c1; if (odd) t = fq2_conj(t); fq2_mul(&r->c0.c1, &t, &k61);
  t = a->c0.c2; if (odd) t = fq2_conj(t); fq2_mul(&r->c0.c2, &t, &k62);
  // c1 part, additionally scaled by xi^((q^k-1)/6)
  t = a->c1.c0; if (odd) t = fq2_conj(t); fq2_mul(&r->c1.c0, &t, &k12);
  fq2_mul(&kk, &k61, &k12);
  t = a->c1.c1; if (odd) t = fq2_conj(t); fq2_mul(&r->c1.c1, &t, &kk);
  fq2_mul(&kk, &k62, &k12);
  t = a->c1.c2; if (odd) t = fq2_conj(t); fq2_mul(&r->c1.c2, &t, &kk);
}

// sparse product f * (e0 + evv v^2 + evw v w): the line value of the optimal-ate loop (positions c0.c0, c0.c2, c1.c1)
BN_NOINLINE void fq12_mul_by_024(fq12* f, const fq2* e0, const fq2* evw, const fq2* evv) {
  fq2 l0 = *e0, lvw = *evw, lvv = *evv;
  fq6 aa, bb, s, full;
  fq2 p00, p22, p12, p10, t, su, uu;
  // aa = f.c0 * (l0, 0, lvv)
  fq2_mul(&p00, &f->c0.c0, &l0);
  fq2_mul(&p22, &f->c0.c2, &lvv);
  fq2_mul(&p12, &f->c0.c1, &lvv);
  fq2_mul(&p10, &f->c0.c1, &l0);
  su = fq2_add(f->c0.c0, f->c0.c2);
  uu = fq2_add(l0, lvv);
  fq2_mul(&t, &su, &uu);
  t = fq2_sub(fq2_sub(t, p00), p22);
  aa.c0 = fq2_add(p00, fq2_mul_xi(p12));
  aa.c1 = fq2_add(p10, fq2_mul_xi(p22));
  aa.c2 = t;
  // bb = f.c1 * (0, lvw, 0)
  fq2_mul(&t, &f->c1.c2, &lvw);
  bb.c0 = fq2_mul_xi(t);
  fq2_mul(&bb.c1, &f->c1.c0, &lvw);
  fq2_mul(&bb.c2, &f->c1.c1, &lvw);
  // (f.c0 + f.c1) * (l0, lvw, lvv)
  fq6_add(&s, &f->c0, &f->c1);
  full.c0 = l0;
  full.c1 = lvw;
  full.c2 = lvv;
  fq6_mul(&s, &s, &full);
  fq6_sub(&s, &s, &aa);
  fq6_sub(&f->c1, &s, &bb);
  fq6_mul_by_v(&bb, &bb);
  fq6_add(&f->c0, &aa, &bb);
}

// Granger-Scott squaring, valid in the cyclotomic subgroup (after the easy part of the final exponentiation)
BN_FN void gs_sq_pair(fq2* t_even, fq2* t_odd, const fq2& x, const fq2& y) {
  // (x + y Y)^2 with Y^2 = xi : even = x^2 + xi y^2, odd = 2xy
  fq2 tmp, s, u;
  fq2_mul(&tmp, &x, &y);
  s = fq2_add(x, y);
  u = fq2_add(fq2_mul_xi(y), x);
  fq2_mul(&s, &s, &u);
  s = fq2_sub(s, tmp);
  *t_even = fq2_sub(s, fq2_mul_xi(tmp));
  *t_odd = fq2_dbl(tmp);
}
BN_NOINLINE void fq12_cyclotomic_sqr(fq12* r, const fq12* a) {
  fq2 z0 = a->c0.c0, z4 = a->c0.c1, z3 = a->c0.c2, z2 = a->c1.c0, z1 = a->c1.c1, z5 = a->c1.c2;
  fq2 t0, t1, t2, t3, t4, t5, tmp;
  gs_sq_pair(&t0, &t1, z0, z1);
  gs_sq_pair(&t2, &t3, z2, z3);
  gs_sq_pair(&t4, &t5, z4, z5);
  z0 = fq2_add(fq2_dbl(fq2_sub(t0, z0)), t0);
  z1 = fq2_add(fq2_dbl(fq2_add(t1, z1)), t1);
  tmp = fq2_mul_xi(t5);
  z2 = fq2_add(fq2_dbl(fq2_add(tmp, z2)), tmp);
  z3 = fq2_add(fq2_dbl(fq2_sub(t4, z3)), t4);
  z4 = fq2_add(fq2_dbl(fq2_sub(t2, z4)), t2);
  z5 = fq2_add(fq2_dbl(fq2_add(t3, z5)), t3);
  r->c0.c0 = z0;
  r->c0.c1 = z4;
  r->c0.c2 = z3;
  r->c1.c0 = z2;
  r->c1.c1 = z1;
  r->c1.c2 = z5;
}
// a^u (u = BN parameter, 63 bits) with cyclotomic squarings; then conjugate: a^(-u)
BN_NOINLINE void fq12_exp_by_neg_u(fq12* r, const fq12* a) {
  fq12 acc = *a, base = *a;
  const uint64_t u = K_BN_U;
  for (int i = 61; i >= 0; i--) {  // bit 62 is the leading one
    fq12_cyclotomic_sqr(&acc, &acc);
    if ((u >> i) & 1) fq12_mul(&acc, &acc, &base);
  }
  fq12_conj(r, &acc);
}

}  // namespace bn
