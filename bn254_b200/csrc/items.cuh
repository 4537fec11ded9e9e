// items.cuh -- per-item bodies of the batch kernels (one thread = one message / signature / key / pair group).
//
// Each function states the reference item it replaces; the CUDA kernels in bn254_b200.cu are thin index
// wrappers around these, and tests/hostsim compiles the very same bodies with g++ to check the control
// logic on a machine without a GPU.
#pragma once
#include "hash.cuh"
#include "pairing.cuh"

namespace bn {

struct alignas(16) g1aff {
  fq x, y;
};

// Fq12 <-> 12 x 32-byte big-endian canonical, tower order c0.c0.re, c0.c0.im, ..., c1.c2.im
BN_FN void fq12_to_be(uint8_t* b, const fq12* f) {
  const fq* c = &f->c0.c0.c0;
  for (int i = 0; i < 12; i++) fq_to_be(b + 32 * i, c[i]);
}
BN_FN bool fq12_from_be(fq12* f, const uint8_t* b) {
  fq* c = &f->c0.c0.c0;
  for (int i = 0; i < 12; i++)
    if (!fq_from_be(&c[i], b + 32 * i)) return false;
  return true;
}

// ---------------------------------------------------------------------------------------------- sign
// /root/reference/src/ecdsa.rs:26-35: sig = H(m) * sk, sk reduced mod r (Fr::from_slice, /root/reference/src/types.rs:37)
BN_FN void item_sign(uint8_t* sig_out, const g1aff* h, const uint8_t* sk_be) {
  uint32_t k[8];
  fr_reduce(k, sk_be);
  g1j p, s;
  pt_set_affine(&p, h->x, h->y);
  g1_mul_glv(&s, &p, k);
  g1_to_raw(sig_out, &s);
}
// generic point * 256-bit scalar (no reduction; bn256.json `mul` vectors use scalars >= r)
BN_FN int item_g1_mul(uint8_t* out, const uint8_t* pt, const uint8_t* k_be) {
  g1j p, s;
  int st = g1_from_raw(&p, pt);
  if (st) {
    for (int i = 0; i < 64; i++) out[i] = 0;
    return st;
  }
  uint32_t k[8];
  fr_reduce(k, k_be);  // every point of the curve has order r (cofactor 1), so [k]P = [k mod r]P
  g1_mul_glv(&s, &p, k);
  g1_to_raw(out, &s);
  return ST_OK;
}
BN_FN int item_g2_mul(uint8_t* out, const uint8_t* pt, const uint8_t* k_be) {
  g2j p, s;
  int st = g2_from_raw(&p, pt);
  if (st) {
    for (int i = 0; i < 128; i++) out[i] = 0;
    return st;
  }
  uint32_t k[8];
  u256_from_be(k, k_be);
  pt_mul(&s, &p, k);
  g2_to_raw(out, &s);
  return ST_OK;
}
// /root/reference/src/types.rs:85-87,155-157: G * sk with sk reduced mod r  (generic ladder; comb tables are a later row)
BN_FN void item_derive_pk_g1(uint8_t* out, const uint8_t* sk_be) {
  uint32_t k[8];
  fr_reduce(k, sk_be);
  g1j g, s;
  pt_set_affine(&g, fq_from_limbs(K_G1_GEN_X), fq_from_limbs(K_G1_GEN_Y));
  pt_mul(&s, &g, k);
  g1_to_raw(out, &s);
}
BN_FN void item_derive_pk_g2(uint8_t* out, const uint8_t* sk_be) {
  uint32_t k[8];
  fr_reduce(k, sk_be);
  g2j g, s;
  pt_set_affine(&g, fq2_from_limbs(K_G2_GEN_X), fq2_from_limbs(K_G2_GEN_Y));
  pt_mul(&s, &g, k);
  g2_to_raw(out, &s);
}

// the same through the fixed-base tables (device path; the ladder above stays as the table-free cross-check)
BN_FN void item_derive_pk_g1_comb(uint8_t* out, const uint8_t* sk_be, const aff<fq>* table) {
  uint32_t k[8];
  fr_reduce(k, sk_be);
  g1j s;
  pt_mul_fixed(&s, table, k);
  g1_to_raw(out, &s);
}
BN_FN void item_derive_pk_g2_comb(uint8_t* out, const uint8_t* sk_be, const aff<fq2>* table) {
  uint32_t k[8];
  fr_reduce(k, sk_be);
  g2j s;
  pt_mul_fixed(&s, table, k);
  g2_to_raw(out, &s);
}

// ---------------------------------------------------------------------------------------------- verify
// /root/reference/src/ecdsa.rs:49-64 after the hash: decode pk (G2) and sig (G1), skip pairs holding an infinity,
// f = miller(H, pk) * miller(sig, -G2).  Returns the decode status; on ST_OK *f is the Miller product.
BN_FN int item_verify_miller(fq12* f, const g1aff* h, const uint8_t* sig, const uint8_t* pk, const line_t* neg_g2_lines) {
  g2j q;
  g1j s;
  int st = g2_from_raw(&q, pk);
  if (st) return st;
  st = g1_from_raw(&s, sig);
  if (st) return st;
  bool use_a = !pt_is_inf(&q);  // H(m) is never infinity
  bool use_b = !pt_is_inf(&s);
  miller_loop_2(f, use_a, &h->x, &h->y, &q.x, &q.y, use_b, &s.x, &s.y, neg_g2_lines);
  return ST_OK;
}
// final exponentiation and comparison with one (/root/reference/src/ecdsa.rs:59-63)
BN_FN uint8_t item_final_exp_is_one(const fq12* f) {
  fq12 gt;
  // f == 0 cannot come out of a Miller loop over points of the curve; bn::pairing_batch would panic on it.  Fail closed, like
  // the cooperative machine does (its Fermat inversion maps 0 to 0, which is not one).
  if (!final_exponentiation(&gt, f)) return ST_VERIFICATION_FAILED;
  return fq12_is_one(&gt) ? ST_OK : ST_VERIFICATION_FAILED;
}
// Miller product of k generic pairs (bn::pairing_batch before the final exponentiation); pairs with an infinity are skipped
BN_FN int item_miller_pairs(fq12* f, const uint8_t* g1s, const uint8_t* g2s, uint64_t k) {
  fq12_set_one(f);
  for (uint64_t j = 0; j < k; j++) {
    g1j p;
    g2j q;
    int st = g1_from_raw(&p, g1s + 64 * j);
    if (st) return st;
    st = g2_from_raw(&q, g2s + 128 * j);
    if (st) return st;
    if (pt_is_inf(&p) || pt_is_inf(&q)) continue;
    fq12 t;
    miller_loop_2(&t, true, &p.x, &p.y, &q.x, &q.y, false, &p.x, &p.y, (const line_t*)0);
    fq12_mul(f, f, &t);
  }
  return ST_OK;
}

// ---------------------------------------------------------------------------------------------- codecs
// /root/reference/src/utils.rs:84-104 (G1 -> 33 bytes)
BN_FN int item_g1_compress(uint8_t* out, const uint8_t* raw) {
  g1j p;
  int st = g1_from_raw(&p, raw);
  if (st) return st;
  if (pt_is_inf(&p)) return ST_POINT_IN_JACOBIAN;
  out[0] = fq_parity(p.y) ? 3 : 2;
  fq_to_be(out + 1, p.x);
  return ST_OK;
}
// bn::G1::from_compressed (33 bytes; length is checked by the caller)
BN_FN int item_g1_decompress(uint8_t* out, const uint8_t* in) {
  for (int i = 0; i < 64; i++) out[i] = 0;
  fq x, y;
  if (!fq_from_be(&x, in + 1)) return ST_NOT_MEMBER;
  fq t = fq_add(fq_mul(fq_sqr(x), x), fq_from_limbs(K_THREE));
  if (!fq_sqrt(&y, t)) return ST_NOT_MEMBER;
  uint32_t odd = fq_parity(y);
  if (in[0] == 2) {
    if (odd) y = fq_neg(y);
  } else if (in[0] == 3) {
    if (!odd) y = fq_neg(y);
  } else {
    return ST_INVALID_ENCODING;
  }
  if (!g1_on_curve(x, y)) return ST_NOT_MEMBER;
  fq_to_be(out, x);
  fq_to_be(out + 32, y);
  return ST_OK;
}
BN_FN int item_g1_validate(const uint8_t* raw) {
  fq x, y;
  if (!fq_from_be(&x, raw)) return ST_NOT_MEMBER;
  if (!fq_from_be(&y, raw + 32)) return ST_NOT_MEMBER;
  return g1_on_curve(x, y) ? ST_OK : ST_INVALID_GROUP_POINT;
}

// psi = twist^-1 o Frobenius o twist on Jacobian coordinates: (X, Y, Z) -> (gx conj X, gy conj Y, conj Z),
// gx = xi^((q-1)/3), gy = xi^((q-1)/2)
BN_FN void g2_psi(g2j* r, const g2j* p) {
  fq2 kx = fq2_from_limbs(K_TWIST_MUL_BY_Q_X), ky = fq2_from_limbs(K_TWIST_MUL_BY_Q_Y), t;
  t = fq2_conj(p->x);
  fq2_mul(&r->x, &t, &kx);
  t = fq2_conj(p->y);
  fq2_mul(&r->y, &t, &ky);
  r->z = fq2_conj(p->z);
}
// r-torsion test of a finite affine point p of the twist (z = 1).  AffineG2::new computes [r]P == infinity
// [DEP-RECALLED]; the same predicate is decided here with one 63-bit scalar multiplication:
//     P in G2  <=>  [u+1]P + psi([u]P) + psi^2([u]P) == psi^3([2u]P)
// (Dai, Lin, Zhao, Zhou, "Fast subgroup membership testings for G1, G2 and GT on pairing-friendly curves", BN case).
// E'(Fq2) has order r * h2 with h2 = 2q - r = 10069 * 5864401 * 1875725156269 * (a 177-bit prime), all to the first
// power; psi acts on each prime-order part as a scalar, the left-minus-right polynomial in psi vanishes on the r part
// and tests/test_oracle.py::test_g2_subgroup_psi_criterion checks that it does not vanish on a point of each of the
// four prime orders dividing h2, which makes the equivalence exact for this curve (not probabilistic).
BN_FN bool g2_in_subgroup(const g2j* p) {
  g2j a, b, c, l, r2;
  a = *p;
  for (int i = 61; i >= 0; i--) {  // [u]P, u = K_BN_U (63 bits, top bit consumed by a = P)
    pt_dbl(&a, &a);
    if ((K_BN_U >> i) & 1) pt_madd(&a, &a, &p->x, &p->y);
  }
  g2_psi(&b, &a);               // psi([u]P)
  g2_psi(&c, &b);               // psi^2([u]P)
  pt_madd(&l, &a, &p->x, &p->y);  // [u+1]P
  pt_add(&l, &l, &b);
  pt_add(&l, &l, &c);
  pt_dbl(&r2, &a);              // [2u]P
  g2_psi(&r2, &r2);
  g2_psi(&r2, &r2);
  g2_psi(&r2, &r2);
  r2.y = fq2_neg(r2.y);
  pt_add(&l, &l, &r2);
  return pt_is_inf(&l);
}
BN_FN int item_g2_validate(const uint8_t* raw) {
  g2j p;
  if (!fq_from_be(&p.x.c0, raw)) return ST_NOT_MEMBER;
  if (!fq_from_be(&p.x.c1, raw + 32)) return ST_NOT_MEMBER;
  if (!fq_from_be(&p.y.c0, raw + 64)) return ST_NOT_MEMBER;
  if (!fq_from_be(&p.y.c1, raw + 96)) return ST_NOT_MEMBER;
  p.z = fq2_one();
  if (!g2_on_curve(p.x, p.y) || !g2_in_subgroup(&p)) return ST_INVALID_GROUP_POINT;
  return ST_OK;
}

// ---------------------------------------------------------------------------------------------- randomised batch verification
// SURVEY.md 8(f) row 4 -- an ADDITIONAL entry point next to ECDSA::verify (/root/reference/src/ecdsa.rs:49-64), not a
// replacement: n independent triples are accepted together iff
//     prod_i e(c_i H(m_i), pk_i) * e(sum_i c_i sig_i, -G2) == 1        (c_i: secret random 128-bit coefficients)
// which, for keys in G2, holds for all-valid batches and fails with probability 1 - 2^-128 otherwise.  This is one item's
// share: hs = c H(m) (affine, Montgomery form) and c sig (raw bytes).  Returns 0 when the item can ride in the batch
// (both points decode, pk is in the r-torsion unless the caller vouches for it); anything else sends the whole batch to
// the exact per-item path, so the statuses a caller sees are always those of verify_batch.
BN_FN int item_rlc_prepare(g1aff* hs, uint8_t* sig_c_raw, const g1aff* h, const uint8_t* sig_raw, const uint8_t* pk_raw,
                           const uint8_t* c16_be, bool check_g2) {
  g1j s, t;
  g2j q;
  int st = g1_from_raw(&s, sig_raw);
  if (st) return st;
  st = g2_from_raw(&q, pk_raw);
  if (st) return st;
  if (check_g2 && !pt_is_inf(&q) && !g2_in_subgroup(&q)) return ST_INVALID_GROUP_POINT;
  // The 16 coefficient bytes are read as two 64-bit halves and the item's coefficient is c = c_lo + c_hi * lambda (mod r):
  // still 2^128 equally likely values (the map is injective: |c_lo|, |c_hi| < 2^64 is far inside the GLV lattice's
  // fundamental cell), and [c]P = [c_lo]P + [c_hi]phi(P) needs 64 doublings instead of 128.
  uint32_t k[4];
  for (int i = 0; i < 4; i++)
    k[i] = ((uint32_t)c16_be[12 - 4 * i] << 24) | ((uint32_t)c16_be[13 - 4 * i] << 16) | ((uint32_t)c16_be[14 - 4 * i] << 8) | c16_be[15 - 4 * i];
  if ((k[0] | k[1] | k[2] | k[3]) == 0) k[0] = 1;  // a zero coefficient would drop the item from the check
  pt_set_affine(&t, h->x, h->y);
  g1_mul_pair(&t, &t, k, false, k + 2, false, 16);
  pt_to_affine(&hs->x, &hs->y, &t);  // never infinity: c != 0 mod r and H(m) has prime order r
  g1_mul_pair(&s, &s, k, false, k + 2, false, 16);
  g1_to_raw(sig_c_raw, &s);
  return ST_OK;
}

// (im, re) compared as the 512-bit integer im*q + re (to_u512, /root/reference/src/utils.rs:40-45): lexicographic
BN_FN bool fq2_u512_gt(const fq2& a, const fq2& b) {
  fq ai = fq_from_mont(a.c1), bi = fq_from_mont(b.c1);
  if (!fq_eq(ai, bi)) return !u256_geq(bi.l, ai.l);
  fq ar = fq_from_mont(a.c0), br = fq_from_mont(b.c0);
  return !u256_geq(br.l, ar.l);
}
// /root/reference/src/utils.rs:130-160 (G2 -> 65 bytes): sign byte 0x0b if y > -y as 512-bit integers else 0x0a,
// then BE64(x.im * q + x.re)
BN_FN int item_g2_compress(uint8_t* out, const uint8_t* raw) {
  g2j p;
  int st = g2_from_raw(&p, raw);
  if (st) return st;
  if (pt_is_inf(&p)) return ST_POINT_IN_JACOBIAN;
  out[0] = fq2_u512_gt(p.y, fq2_neg(p.y)) ? 0x0b : 0x0a;
  fq re = fq_from_mont(p.x.c0), im = fq_from_mont(p.x.c1);
  uint32_t w[16];
  for (int i = 0; i < 16; i++) w[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) {
      c += (uint64_t)im.l[i] * K_Q[j] + w[i + j];
      w[i + j] = (uint32_t)c;
      c >>= 32;
    }
    w[i + 8] = (uint32_t)c;
  }
  uint64_t c = 0;
  for (int i = 0; i < 16; i++) {
    c += (uint64_t)w[i] + (i < 8 ? re.l[i] : 0u);
    w[i] = (uint32_t)c;
    c >>= 32;
  }
  for (int i = 0; i < 16; i++) {
    uint8_t* p8 = out + 1 + 4 * (15 - i);
    p8[0] = (uint8_t)(w[i] >> 24); p8[1] = (uint8_t)(w[i] >> 16); p8[2] = (uint8_t)(w[i] >> 8); p8[3] = (uint8_t)w[i];
  }
  return ST_OK;
}
// a^e in Fq2 for a public exponent (plain square-and-multiply, MSB first)
BN_NOINLINE void fq2_pow_pub(fq2* r, const fq2* a, const uint32_t* e) {
  fq2 acc = fq2_one(), base = *a;
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (started) fq2_sqr(&acc, &acc);
    if ((e[i >> 5] >> (i & 31)) & 1) {
      if (started) fq2_mul(&acc, &acc, &base);
      else acc = base;
      started = true;
    }
  }
  *r = acc;
}
// Fq2::sqrt of the dependency (complex method, q = 3 mod 4)
BN_FN bool fq2_sqrt(fq2* r, const fq2& a) {
  fq2 a1, alpha, a0, x0, t, m1 = fq2_neg(fq2_one());
  fq2_pow_pub(&a1, &a, K_EXP_QM3D4);
  fq2_sqr(&alpha, &a1);
  fq2_mul(&alpha, &alpha, &a);
  t = fq2_conj(alpha);
  fq2_mul(&a0, &t, &alpha);
  if (fq2_eq(a0, m1)) return false;
  fq2_mul(&x0, &a1, &a);
  if (fq2_eq(alpha, m1)) {
    fq2 iu;
    iu.c0 = fq_zero();
    iu.c1 = fq_one();
    fq2_mul(r, &iu, &x0);
  } else {
    fq2 b;
    t = fq2_add(fq2_one(), alpha);
    fq2_pow_pub(&b, &t, K_EXP_QM1D2);
    fq2_mul(r, &b, &x0);
  }
  return true;
}
// bn::G2::from_compressed (65 bytes; length checked by the caller): x = divmod(BE64, q), sqrt, sign rule, curve + r-torsion
BN_FN int item_g2_decompress(uint8_t* out, const uint8_t* in) {
  for (int i = 0; i < 128; i++) out[i] = 0;
  uint32_t w[16];
  for (int i = 0; i < 16; i++) {
    const uint8_t* p = in + 1 + 4 * (15 - i);
    w[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  }
  // bitwise long division by q: quotient (must be < q) and remainder
  uint32_t rem[9], quot[16];
  for (int i = 0; i < 9; i++) rem[i] = 0;
  for (int i = 0; i < 16; i++) quot[i] = 0;
  for (int i = 511; i >= 0; i--) {
    for (int k = 8; k > 0; k--) rem[k] = (rem[k] << 1) | (rem[k - 1] >> 31);
    rem[0] = (rem[0] << 1) | ((w[i >> 5] >> (i & 31)) & 1);
    if (rem[8] || u256_geq(rem, K_Q)) {
      uint32_t t[8];
      uint32_t bw = u256_sub(t, rem, K_Q);
      for (int k = 0; k < 8; k++) rem[k] = t[k];
      rem[8] -= bw;
      quot[i >> 5] |= 1u << (i & 31);
    }
  }
  uint32_t hi = 0;
  for (int i = 8; i < 16; i++) hi |= quot[i];
  if (hi || u256_geq(quot, K_Q)) return ST_NOT_MEMBER;
  g2j p;
  fq t0, t1;
  for (int i = 0; i < 8; i++) {
    t0.l[i] = rem[i];
    t1.l[i] = quot[i];
  }
  p.x.c0 = fq_to_mont(t0);
  p.x.c1 = fq_to_mont(t1);
  fq2 rhs = fq2_add(fq2_mulv(fq2_sqrv(p.x), p.x), fq2_from_limbs(K_TWIST_B));
  fq2 y;
  if (!fq2_sqrt(&y, rhs)) return ST_NOT_MEMBER;
  fq2 ny = fq2_neg(y);
  bool gt = fq2_u512_gt(y, ny);
  if (in[0] == 0x0a) {
    if (gt) y = ny;
  } else if (in[0] == 0x0b) {
    if (!gt) y = ny;
  } else {
    return ST_INVALID_ENCODING;
  }
  p.y = y;
  p.z = fq2_one();
  if (!g2_on_curve(p.x, p.y) || !g2_in_subgroup(&p)) return ST_NOT_MEMBER;
  g2_to_raw(out, &p);
  return ST_OK;
}

}  // namespace bn

// ---------------------------------------------------------------------------------------------- layer hooks
// Building blocks exposed to the parity tests (GPU kernel vs g++ host simulation vs oracle), Fq arrays in / out:
//   op 0 fq2_mul (4 -> 2), 1 fq2_sqr (2 -> 2), 2 fq2_scale (3 -> 2), 3 doubling_step (6 -> 12),
//   4 mixed_addition_step (q:4, r:6 -> 12), 5 fq12_mul_by_024 (12 + 6 -> 12), 6 G1 pt_madd (jac 3 + affine 2 -> 3),
//   7 fq2_mul_xi (2 -> 2), 8 fq2_inv (2 -> 2), 11 G1 pt_madd in place, 12 G1 pt_add (6 -> 3), 13 G2 pt_madd (10 -> 6),
//   200 single-pair Miller loop (p:2, q:4 -> 12)
namespace bn {
BN_FN void debug_layer_op(int op, const fq* in, fq* out) {
  const fq2* i2 = (const fq2*)in;
  fq2* o2 = (fq2*)out;
  if (op == 0) fq2_mul(&o2[0], &i2[0], &i2[1]);
  else if (op == 1) fq2_sqr(&o2[0], &i2[0]);
  else if (op == 2) fq2_scale(&o2[0], &i2[0], &in[2]);
  else if (op == 3) {
    g2proj r;
    line_t c;
    r.x = i2[0]; r.y = i2[1]; r.z = i2[2];
    doubling_step(&r, &c);
    o2[0] = r.x; o2[1] = r.y; o2[2] = r.z; o2[3] = c.ell_0; o2[4] = c.ell_vw; o2[5] = c.ell_vv;
  } else if (op == 4) {
    g2proj r;
    line_t c;
    r.x = i2[2]; r.y = i2[3]; r.z = i2[4];
    mixed_addition_step(&i2[0], &i2[1], &r, &c);
    o2[0] = r.x; o2[1] = r.y; o2[2] = r.z; o2[3] = c.ell_0; o2[4] = c.ell_vw; o2[5] = c.ell_vv;
  } else if (op == 5) {
    fq12 f = *(const fq12*)in;
    fq12_mul_by_024(&f, &i2[6], &i2[7], &i2[8]);
    *(fq12*)out = f;
  } else if (op == 6) {
    g1j p, r;
    p.x = in[0]; p.y = in[1]; p.z = in[2];
    pt_madd(&r, &p, &in[3], &in[4]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
  } else if (op == 7) o2[0] = fq2_mul_xi(i2[0]);
  else if (op == 8) fq2_inv(&o2[0], &i2[0]);
  else if (op == 200) {
    fq12 f;
    miller_loop_2(&f, true, &in[0], &in[1], (const fq2*)&in[2], (const fq2*)&in[4], false, &in[0], &in[1], (const line_t*)0);
    *(fq12*)out = f;
  } else if (op == 11) {
    g1j p;
    p.x = in[0]; p.y = in[1]; p.z = in[2];
    pt_madd(&p, &p, &in[3], &in[4]);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
  } else if (op == 12) {
    g1j p, q, r;
    p.x = in[0]; p.y = in[1]; p.z = in[2];
    q.x = in[3]; q.y = in[4]; q.z = in[5];
    pt_add(&r, &p, &q);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
  } else if (op == 13) {
    g2j p, r;
    const fq2* i2b = (const fq2*)in;
    p.x = i2b[0]; p.y = i2b[1]; p.z = i2b[2];
    pt_madd(&r, &p, &i2b[3], &i2b[4]);
    fq2* o = (fq2*)out;
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
  }
}
}  // namespace bn
