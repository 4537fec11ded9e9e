// curve.cuh -- G1 (over Fq) and G2 (over Fq2, D-type twist y^2 = x^3 + 3/xi) in Jacobian coordinates.
//
// Replaces bn::G1 / bn::G2 / AffineG1 / AffineG2 of the reference's dependency as used by
// /root/reference/src/types.rs:86,130,138,146,156,200,208,216,268,276,284 (Mul<Fr>, Add, Sub, Neg),
// /root/reference/src/ecdsa.rs:31 (H(m) * sk) and /root/reference/src/utils.rs:86,133,163,184 (to affine).
// Infinity <=> z == 0.  Only affine / serialised results are observable, so the group law may use any
// complete formula set; these are dbl-2009-l, add-2007-bl and madd-2007-bl with the exceptional cases handled.
#pragma once
#include "tower.cuh"

namespace bn {

// uniform field interface over Fq and Fq2
BN_FN fq fe_add(const fq& a, const fq& b) { return fq_add(a, b); }
BN_FN fq fe_sub(const fq& a, const fq& b) { return fq_sub(a, b); }
BN_FN fq fe_dbl(const fq& a) { return fq_dbl(a); }
BN_FN fq fe_neg(const fq& a) { return fq_neg(a); }
BN_FN fq fe_mul(const fq& a, const fq& b) { return fq_mul(a, b); }
BN_FN fq fe_sqr(const fq& a) { return fq_sqr(a); }
BN_FN bool fe_is_zero(const fq& a) { return fq_is_zero(a); }
BN_FN bool fe_eq(const fq& a, const fq& b) { return fq_eq(a, b); }
BN_FN void fe_set_one(fq* a) { *a = fq_one(); }
BN_FN void fe_set_zero(fq* a) { *a = fq_zero(); }
BN_FN fq fe_inv(const fq& a) { return fq_inv(a); }

BN_FN fq2 fe_add(const fq2& a, const fq2& b) { return fq2_add(a, b); }
BN_FN fq2 fe_sub(const fq2& a, const fq2& b) { return fq2_sub(a, b); }
BN_FN fq2 fe_dbl(const fq2& a) { return fq2_dbl(a); }
BN_FN fq2 fe_neg(const fq2& a) { return fq2_neg(a); }
BN_FN fq2 fe_mul(const fq2& a, const fq2& b) { return fq2_mulv(a, b); }
BN_FN fq2 fe_sqr(const fq2& a) { return fq2_sqrv(a); }
BN_FN bool fe_is_zero(const fq2& a) { return fq2_is_zero(a); }
BN_FN bool fe_eq(const fq2& a, const fq2& b) { return fq2_eq(a, b); }
BN_FN void fe_set_one(fq2* a) { *a = fq2_one(); }
BN_FN void fe_set_zero(fq2* a) { *a = fq2_zero(); }
BN_FN fq2 fe_inv(fq2 a) {
  fq2 r;
  fq2_inv(&r, &a);
  return r;
}

template <class F>
struct alignas(16) jac {
  F x, y, z;
};
typedef jac<fq> g1j;
typedef jac<fq2> g2j;

template <class F>
BN_FN bool pt_is_inf(const jac<F>* p) { return fe_is_zero(p->z); }
template <class F>
BN_FN void pt_set_inf(jac<F>* p) {
  fe_set_zero(&p->x);
  fe_set_one(&p->y);
  fe_set_zero(&p->z);
}
template <class F>
BN_FN void pt_set_affine(jac<F>* p, const F& x, const F& y) {
  p->x = x;
  p->y = y;
  fe_set_one(&p->z);
}
template <class F>
BN_FN void pt_neg(jac<F>* r, const jac<F>* p) {
  r->x = p->x;
  r->y = fe_neg(p->y);
  r->z = p->z;
}

template <class F>
BN_NOINLINE void pt_dbl(jac<F>* r, const jac<F>* p) {
  if (pt_is_inf(p)) {
    *r = *p;
    return;
  }
  F a = fe_sqr(p->x), b = fe_sqr(p->y), c = fe_sqr(b);
  F t = fe_sqr(fe_add(p->x, b));
  F d = fe_dbl(fe_sub(fe_sub(t, a), c));
  F e = fe_add(fe_dbl(a), a);
  F f = fe_sqr(e);
  F z3 = fe_dbl(fe_mul(p->y, p->z));
  F x3 = fe_sub(f, fe_dbl(d));
  F c8 = fe_dbl(fe_dbl(fe_dbl(c)));
  F y3 = fe_sub(fe_mul(e, fe_sub(d, x3)), c8);
  r->x = x3;
  r->y = y3;
  r->z = z3;
}

template <class F>
BN_NOINLINE void pt_add(jac<F>* r, const jac<F>* p, const jac<F>* q) {
  if (pt_is_inf(p)) {
    *r = *q;
    return;
  }
  if (pt_is_inf(q)) {
    *r = *p;
    return;
  }
  F z1z1 = fe_sqr(p->z), z2z2 = fe_sqr(q->z);
  F u1 = fe_mul(p->x, z2z2), u2 = fe_mul(q->x, z1z1);
  F s1 = fe_mul(fe_mul(p->y, q->z), z2z2), s2 = fe_mul(fe_mul(q->y, p->z), z1z1);
  if (fe_eq(u1, u2)) {
    if (fe_eq(s1, s2)) pt_dbl(r, p);
    else pt_set_inf(r);
    return;
  }
  F h = fe_sub(u2, u1);
  F i = fe_sqr(fe_dbl(h));
  F j = fe_mul(h, i);
  F rr = fe_dbl(fe_sub(s2, s1));
  F v = fe_mul(u1, i);
  F x3 = fe_sub(fe_sub(fe_sqr(rr), j), fe_dbl(v));
  F y3 = fe_sub(fe_mul(rr, fe_sub(v, x3)), fe_dbl(fe_mul(s1, j)));
  F z3 = fe_mul(fe_sub(fe_sub(fe_sqr(fe_add(p->z, q->z)), z1z1), z2z2), h);
  r->x = x3;
  r->y = y3;
  r->z = z3;
}

// mixed addition with an affine, finite q = (qx, qy)
template <class F>
BN_NOINLINE void pt_madd(jac<F>* r, const jac<F>* p, const F* qx, const F* qy) {
  if (pt_is_inf(p)) {
    pt_set_affine(r, *qx, *qy);
    return;
  }
  F z1z1 = fe_sqr(p->z);
  F u2 = fe_mul(*qx, z1z1);
  F s2 = fe_mul(fe_mul(*qy, p->z), z1z1);
  if (fe_eq(p->x, u2)) {
    if (fe_eq(p->y, s2)) pt_dbl(r, p);
    else pt_set_inf(r);
    return;
  }
  F h = fe_sub(u2, p->x);
  F hh = fe_sqr(h);
  F i = fe_dbl(fe_dbl(hh));
  F j = fe_mul(h, i);
  F rr = fe_dbl(fe_sub(s2, p->y));
  F v = fe_mul(p->x, i);
  F x3 = fe_sub(fe_sub(fe_sqr(rr), j), fe_dbl(v));
  F y3 = fe_sub(fe_mul(rr, fe_sub(v, x3)), fe_dbl(fe_mul(p->y, j)));
  F z3 = fe_sub(fe_sub(fe_sqr(fe_add(p->z, h)), z1z1), hh);
  r->x = x3;
  r->y = y3;
  r->z = z3;
}

// returns false for infinity
template <class F>
BN_NOINLINE bool pt_to_affine(F* x, F* y, const jac<F>* p) {
  if (pt_is_inf(p)) return false;
  F zi = fe_inv(p->z);
  F zi2 = fe_sqr(zi);
  *x = fe_mul(p->x, zi2);
  *y = fe_mul(p->y, fe_mul(zi2, zi));
  return true;
}

// scalar multiplication by a 256-bit integer k (8 plain limbs), 4-bit fixed windows, MSB first.
// Uniform schedule: 4 doublings + one table addition per window (skipped only for a zero digit).
template <class F>
BN_NOINLINE void pt_mul(jac<F>* r, const jac<F>* p, const uint32_t* k, int windows = 64) {
  jac<F> tab[16];
  pt_set_inf(&tab[0]);
  tab[1] = *p;
  pt_dbl(&tab[2], p);
  for (int i = 3; i < 16; i++) pt_add(&tab[i], &tab[i - 1], p);
  jac<F> acc;
  pt_set_inf(&acc);
  for (int w = windows - 1; w >= 0; w--) {  // windows < 64: the caller knows that k < 16^windows
    for (int d = 0; d < 4; d++) pt_dbl(&acc, &acc);
    uint32_t nib = (k[w >> 3] >> ((w & 7) * 4)) & 15;
    if (nib) pt_add(&acc, &acc, &tab[nib]);
  }
  *r = acc;
}

// ---- GLV scalar multiplication in G1 (signing, /root/reference/src/ecdsa.rs:28-31, multiplies the per-message hash point,
// so no fixed-base table applies).  phi(x, y) = (beta x, y) equals [lambda](x, y) on G1 (beta, lambda: cube roots of unity
// in Fq, Fr), and every k < r splits as k = k1 + k2 lambda (mod r) with |k1|, |k2| < 2^128 (constants.cuh K_GLV_*, derived
// and bounded in scripts/gen_constants.py): [k]P = [k1]P + [k2]phi(P) costs 128 doublings instead of 256.  The group
// element is the same, so the serialised signature is bit-identical to the double-and-add result.
// out (n limbs) = low n limbs of a (na limbs) * b (nb limbs)
BN_FN void limbs_mul(uint32_t* out, int n, const uint32_t* a, int na, const uint32_t* b, int nb) {
  for (int i = 0; i < n; i++) out[i] = 0;
  for (int i = 0; i < na; i++) {
    uint64_t c = 0;
    for (int j = 0; j < nb && i + j < n; j++) {
      c += (uint64_t)a[i] * b[j] + out[i + j];
      out[i + j] = (uint32_t)c;
      c >>= 32;
    }
    for (int t = i + nb; t < n && c; t++) {
      c += out[t];
      out[t] = (uint32_t)c;
      c >>= 32;
    }
  }
}
// k (8 limbs, k < r) -> magnitudes (4 limbs each) and signs of k1, k2
BN_FN void glv_decompose(uint32_t* k1, bool* neg1, uint32_t* k2, bool* neg2, const uint32_t* k) {
  uint32_t t[13], c1[3], c2[5], u[8], v[8], w[8];
  limbs_mul(t, 11, k, 8, K_GLV_G1, 3);  // c1 = (k g1) >> 256
  for (int i = 0; i < 3; i++) c1[i] = t[8 + i];
  limbs_mul(t, 13, k, 8, K_GLV_G2, 5);  // c2 = (k g2) >> 256
  for (int i = 0; i < 5; i++) c2[i] = t[8 + i];
  // k1 = k - c1 a1 - c2 a2 and k2 = c1 |b1| - c2 b2, two's complement mod 2^256 (the true values are below 2^128 in size)
  limbs_mul(u, 8, c1, 3, K_GLV_A1, 2);
  limbs_mul(v, 8, c2, 5, K_GLV_A2, 4);
  u256_sub(w, k, u);
  u256_sub(w, w, v);
  *neg1 = (w[7] >> 31) != 0;
  if (*neg1) {
    for (int i = 0; i < 8; i++) u[i] = 0;
    u256_sub(w, u, w);
  }
  for (int i = 0; i < 4; i++) k1[i] = w[i];
  limbs_mul(u, 8, c1, 3, K_GLV_B1N, 4);
  limbs_mul(v, 8, c2, 5, K_GLV_B2, 2);
  u256_sub(w, u, v);
  *neg2 = (w[7] >> 31) != 0;
  if (*neg2) {
    for (int i = 0; i < 8; i++) u[i] = 0;
    u256_sub(w, u, w);
  }
  for (int i = 0; i < 4; i++) k2[i] = w[i];
}
// [k]p for k < r (8 plain limbs) is g1_mul_pair over the decomposition: 32 windows of 4 bits, one table of multiples of p
// shared by both halves
// r = [+-k1]p + [+-k2]phi(p), k1 and k2 below 16^windows
BN_NOINLINE void g1_mul_pair(jac<fq>* r, const jac<fq>* p, const uint32_t* k1, bool n1, const uint32_t* k2, bool n2, int windows) {
  jac<fq> tab[16];
  pt_set_inf(&tab[0]);
  tab[1] = *p;
  pt_dbl(&tab[2], p);
  for (int i = 3; i < 16; i++) pt_add(&tab[i], &tab[i - 1], p);
  const fq beta = fq_from_limbs(K_GLV_BETA);
  jac<fq> acc, t;
  pt_set_inf(&acc);
  for (int w = windows - 1; w >= 0; w--) {
    for (int d = 0; d < 4; d++) pt_dbl(&acc, &acc);
    uint32_t d1 = (k1[w >> 3] >> ((w & 7) * 4)) & 15, d2 = (k2[w >> 3] >> ((w & 7) * 4)) & 15;
    if (d1) {
      t = tab[d1];
      if (n1) t.y = fq_neg(t.y);
      pt_add(&acc, &acc, &t);
    }
    if (d2) {
      t = tab[d2];
      t.x = fq_mul(t.x, beta);  // phi on Jacobian coordinates: (beta X, Y, Z)
      if (n2) t.y = fq_neg(t.y);
      pt_add(&acc, &acc, &t);
    }
  }
  *r = acc;
}
BN_FN void g1_mul_glv(jac<fq>* r, const jac<fq>* p, const uint32_t* k) {
  uint32_t k1[4], k2[4];
  bool n1, n2;
  glv_decompose(k1, &n1, k2, &n2, k);
  g1_mul_pair(r, p, k1, n1, k2, n2, 32);
}

// ---- fixed-base scalar multiplication (key derivation, /root/reference/src/types.rs:85-87,155-157): a table of
// d * 16^w * G (w = 0..63, d = 1..15, affine) turns G * k into at most 64 mixed additions and no doubling.
template <class F>
struct alignas(16) aff {
  F x, y;
};
#define BN_COMB_WINDOWS 64
#define BN_COMB_ROW 15
// row w of the table: entries (d + 1) * 16^w * G for d = 0..14
template <class F>
BN_FN void comb_build_row(aff<F>* row, int w, const F& gx, const F& gy) {
  jac<F> base, acc;
  pt_set_affine(&base, gx, gy);
  for (int i = 0; i < 4 * w; i++) pt_dbl(&base, &base);
  acc = base;
  for (int d = 0; d < BN_COMB_ROW; d++) {
    if (d) pt_add(&acc, &acc, &base);
    pt_to_affine(&row[d].x, &row[d].y, &acc);  // never infinity: 16^w * (d + 1) < r
  }
}
template <class F>
BN_NOINLINE void pt_mul_fixed(jac<F>* r, const aff<F>* table, const uint32_t* k) {
  jac<F> acc;
  pt_set_inf(&acc);
  for (int w = 0; w < BN_COMB_WINDOWS; w++) {
    uint32_t nib = (k[w >> 3] >> ((w & 7) * 4)) & 15;
    if (nib) {
      aff<F> t = table[w * BN_COMB_ROW + nib - 1];
      pt_madd(&acc, &acc, &t.x, &t.y);
    }
  }
  *r = acc;
}

// curve membership (affine): y^2 == x^3 + b
BN_FN bool g1_on_curve(const fq& x, const fq& y) {
  fq l = fq_sqr(y);
  fq r = fq_add(fq_mul(fq_sqr(x), x), fq_from_limbs(K_THREE));
  return fq_eq(l, r);
}
BN_FN bool g2_on_curve(const fq2& x, const fq2& y) {
  fq2 l = fq2_sqrv(y);
  fq2 r = fq2_add(fq2_mulv(fq2_sqrv(x), x), fq2_from_limbs(K_TWIST_B));
  return fq2_eq(l, r);
}

// Fr::from_slice semantics (/root/reference/src/types.rs:37): any 256-bit value, reduced mod r
BN_FN void fr_reduce(uint32_t* k, const uint8_t* be) {
  u256_from_be(k, be);
  for (int i = 0; i < 6; i++) {  // 2^256 / r < 6
    uint32_t t[8];
    uint32_t bw = u256_sub(t, k, K_R_ORDER);
    if (!bw) {
      for (int j = 0; j < 8; j++) k[j] = t[j];
    }
  }
}

// status codes: one per Error variant of /root/reference/src/error.rs:6-29 (0 = Ok)
enum {
  ST_OK = 0,
  ST_HASH_TO_POINT = 1,
  ST_INDEX_OOB = 2,
  ST_INVALID_ENCODING = 3,
  ST_INVALID_GROUP_POINT = 4,
  ST_INVALID_LENGTH = 5,
  ST_NOT_MEMBER = 6,
  ST_TO_AFFINE = 7,
  ST_POINT_IN_JACOBIAN = 8,
  ST_VERIFICATION_FAILED = 9,
  ST_SERIALIZATION = 10,
  ST_HEX_DECODE = 11,
  ST_ENGINE_FAULT = 255  // not an Error variant: the item was NOT evaluated (pipelined small-batch verify: its line sets never arrived)
};

BN_FN bool bytes_all_zero(const uint8_t* b, int n) {
  uint32_t acc = 0;
  for (int i = 0; i < n; i++) acc |= b[i];
  return acc == 0;
}
// raw 64-byte x||y (big-endian canonical) -> Jacobian; all-zero = infinity; validated like
// from_uncompressed (/root/reference/src/utils.rs:119-127): field membership then curve equation
BN_FN int g1_from_raw(g1j* p, const uint8_t* b) {
  if (bytes_all_zero(b, 64)) {
    pt_set_inf(p);
    return ST_OK;
  }
  if (!fq_from_be(&p->x, b)) return ST_NOT_MEMBER;
  if (!fq_from_be(&p->y, b + 32)) return ST_NOT_MEMBER;
  if (!g1_on_curve(p->x, p->y)) return ST_INVALID_GROUP_POINT;
  p->z = fq_one();
  return ST_OK;
}
BN_FN void g1_to_raw(uint8_t* b, const g1j* p) {
  fq x, y;
  if (!pt_to_affine(&x, &y, p)) {
    for (int i = 0; i < 64; i++) b[i] = 0;
    return;
  }
  fq_to_be(b, x);
  fq_to_be(b + 32, y);
}
// raw 128-byte x.re||x.im||y.re||y.im (the crate's uncompressed layout, /root/reference/src/utils.rs:162-179)
BN_FN int g2_from_raw(g2j* p, const uint8_t* b) {
  if (bytes_all_zero(b, 128)) {
    pt_set_inf(p);
    return ST_OK;
  }
  if (!fq_from_be(&p->x.c0, b)) return ST_NOT_MEMBER;
  if (!fq_from_be(&p->x.c1, b + 32)) return ST_NOT_MEMBER;
  if (!fq_from_be(&p->y.c0, b + 64)) return ST_NOT_MEMBER;
  if (!fq_from_be(&p->y.c1, b + 96)) return ST_NOT_MEMBER;
  p->z = fq2_one();
  if (!g2_on_curve(p->x, p->y)) return ST_INVALID_GROUP_POINT;
  return ST_OK;
}
BN_FN void g2_to_raw(uint8_t* b, const g2j* p) {
  fq2 x, y;
  if (!pt_to_affine(&x, &y, p)) {
    for (int i = 0; i < 128; i++) b[i] = 0;
    return;
  }
  fq_to_be(b, x.c0);
  fq_to_be(b + 32, x.c1);
  fq_to_be(b + 64, y.c0);
  fq_to_be(b + 96, y.c1);
}

}  // namespace bn
