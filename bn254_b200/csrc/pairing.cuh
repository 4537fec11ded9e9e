// pairing.cuh -- optimal-ate Miller loop over the signed digits of 6u+2 and the final exponentiation.
//
// Replaces bn::pairing_batch of the reference's dependency (called at /root/reference/src/ecdsa.rs:57 and :86).
// Semantics kept: a pair holding an infinity is skipped, an empty product is one; only "== Gt::one()" is
// ever observed by the reference (/root/reference/src/ecdsa.rs:59,88).  The G2 point walks the twist in
// homogeneous projective coordinates; each step yields a sparse line (three Fq2 coefficients) that is
// scaled by the G1 point and folded into f by fq12_mul_by_024.  A fixed second argument (-G2 for verify)
// uses a precomputed line table instead of walking the curve.
#pragma once
#include "curve.cuh"

namespace bn {

struct alignas(16) line_t {
  fq2 ell_0, ell_vw, ell_vv;
};
struct alignas(16) g2proj {
  fq2 x, y, z;
};

BN_NOINLINE void doubling_step(g2proj* r, line_t* c) {
  fq two_inv = fq_from_limbs(K_TWO_INV);
  fq2 twist_b = fq2_from_limbs(K_TWIST_B);
  fq2 a, b, cc, d, e, f, g, h, i, j, e2, t;
  fq2_mul(&a, &r->x, &r->y);
  fq2_scale(&a, &a, &two_inv);
  fq2_sqr(&b, &r->y);
  fq2_sqr(&cc, &r->z);
  d = fq2_add(fq2_dbl(cc), cc);
  fq2_mul(&e, &twist_b, &d);
  f = fq2_add(fq2_dbl(e), e);
  g = fq2_add(b, f);
  fq2_scale(&g, &g, &two_inv);
  h = fq2_add(r->y, r->z);
  fq2_sqr(&h, &h);
  h = fq2_sub(h, fq2_add(b, cc));
  i = fq2_sub(e, b);
  fq2_sqr(&j, &r->x);
  fq2_sqr(&e2, &e);
  t = fq2_sub(b, f);
  fq2_mul(&r->x, &a, &t);
  fq2_sqr(&t, &g);
  r->y = fq2_sub(t, fq2_add(fq2_dbl(e2), e2));
  fq2_mul(&r->z, &b, &h);
  c->ell_0 = fq2_mul_xi(i);
  c->ell_vw = fq2_neg(h);
  c->ell_vv = fq2_add(fq2_dbl(j), j);
}

BN_NOINLINE void mixed_addition_step(const fq2* qx, const fq2* qy, g2proj* r, line_t* c) {
  fq2 d, e, f, g, h, i, j, t, t2;
  fq2_mul(&t, qx, &r->z);
  d = fq2_sub(r->x, t);
  fq2_mul(&t, qy, &r->z);
  e = fq2_sub(r->y, t);
  fq2_sqr(&f, &d);
  fq2_sqr(&g, &e);
  fq2_mul(&h, &d, &f);
  fq2_mul(&i, &r->x, &f);
  fq2_mul(&t, &r->z, &g);
  j = fq2_sub(fq2_add(h, t), fq2_dbl(i));
  fq2_mul(&t, &h, &r->y);
  fq2_mul(&r->x, &d, &j);
  t2 = fq2_sub(i, j);
  fq2_mul(&t2, &e, &t2);
  r->y = fq2_sub(t2, t);
  fq2_mul(&r->z, &r->z, &h);
  fq2_mul(&t, &e, qx);
  fq2_mul(&t2, &d, qy);
  c->ell_0 = fq2_mul_xi(fq2_sub(t, t2));
  c->ell_vv = fq2_neg(e);
  c->ell_vw = d;
}

// Frobenius images of an affine twist point: q1 = pi(Q), q2 = -pi^2(Q)
BN_FN void g2_frobenius_pair(fq2* q1x, fq2* q1y, fq2* q2x, fq2* q2y, const fq2& qx, const fq2& qy) {
  fq2 kx = fq2_from_limbs(K_TWIST_MUL_BY_Q_X), ky = fq2_from_limbs(K_TWIST_MUL_BY_Q_Y), t;
  t = fq2_conj(qx);
  fq2_mul(q1x, &t, &kx);
  t = fq2_conj(qy);
  fq2_mul(q1y, &t, &ky);
  t = fq2_conj(*q1x);
  fq2_mul(q2x, &t, &kx);
  t = fq2_conj(*q1y);
  fq2_mul(q2y, &t, &ky);
  *q2y = fq2_neg(*q2y);
}

// all K_N_LINES line coefficients of a fixed affine Q (used once, for -G2)
BN_FN void g2_precompute_lines(line_t* out, fq2 qx_in, fq2 qy_in) {
  struct {
    g2proj r;
    fq2 qx, qy, nqy, q1x, q1y, q2x, q2y;
  } L;
  L.qx = qx_in;
  L.qy = qy_in;
  L.r.x = L.qx;
  L.r.y = L.qy;
  L.r.z = fq2_one();
  L.nqy = fq2_neg(L.qy);
  int n = 0;
  for (int k = 0; k < 64; k++) {
    doubling_step(&L.r, &out[n++]);
    int d = K_ATE_DIGITS[k];
    if (d == 1) mixed_addition_step(&L.qx, &L.qy, &L.r, &out[n++]);
    else if (d == -1) mixed_addition_step(&L.qx, &L.nqy, &L.r, &out[n++]);
  }
  g2_frobenius_pair(&L.q1x, &L.q1y, &L.q2x, &L.q2y, L.qx, L.qy);
  mixed_addition_step(&L.q1x, &L.q1y, &L.r, &out[n++]);
  mixed_addition_step(&L.q2x, &L.q2y, &L.r, &out[n++]);
}

BN_FN void ell_apply(fq12* f, const line_t* c, const fq* px, const fq* py) {
  struct {
    fq2 vw, vv;
  } L;
  fq2_scale(&L.vw, &c->ell_vw, py);
  fq2_scale(&L.vv, &c->ell_vv, px);
  fq12_mul_by_024(f, &c->ell_0, &L.vw, &L.vv);
}

// f = miller(Pa, Qa) * miller(Pb, fixed Q given by its line table), sharing the squaring chain.
//   USE_A / USE_B: whether each pair takes part (a pair holding an infinity is skipped by the caller).
//   (pax, pay), (qax, qay): affine G1 / G2 of the variable pair;  (pbx, pby): affine G1 paired with the table.
// The two flags are template parameters: each combination is its own straight-line schedule.
template <bool USE_A, bool USE_B>
BN_NOINLINE void miller_loop_t(fq12* f, const fq* pax, const fq* pay, const fq2* qax, const fq2* qay, const fq* pbx, const fq* pby,
                               const line_t* table) {
  fq12_set_one(f);
  // Every local whose address is handed to an out-of-line routine lives in ONE frame object: this toolchain's
  // stack-slot sharing has been seen to overlap separately declared locals that were still live (a slot reached
  // through a pointer select, or through a reference parameter of an inlined helper), and members of a single
  // object cannot be overlapped.
  struct {
    g2proj r;
    line_t c;
    fq2 qy_sel, q1x, q1y, q2x, q2y;
  } L;
  g2proj& r = L.r;
  line_t& c = L.c;
  if (USE_A) {
    r.x = *qax;
    r.y = *qay;
    r.z = fq2_one();
  }
  int idx = 0;
  for (int k = 0; k < 64; k++) {
    if (k > 0) fq12_sqr(f, f);
    if (USE_A) {
      doubling_step(&r, &c);
      ell_apply(f, &c, pax, pay);
    }
    if (USE_B) ell_apply(f, &table[idx], pbx, pby);
    idx++;
    int d = K_ATE_DIGITS[k];
    if (d != 0) {
      if (USE_A) {
        L.qy_sel = d > 0 ? *qay : fq2_neg(*qay);
        mixed_addition_step(qax, &L.qy_sel, &r, &c);
        ell_apply(f, &c, pax, pay);
      }
      if (USE_B) ell_apply(f, &table[idx], pbx, pby);
      idx++;
    }
  }
  if (USE_A) {
    g2_frobenius_pair(&L.q1x, &L.q1y, &L.q2x, &L.q2y, *qax, *qay);
    mixed_addition_step(&L.q1x, &L.q1y, &r, &c);
    ell_apply(f, &c, pax, pay);
    if (USE_B) ell_apply(f, &table[idx], pbx, pby);
    idx++;
    mixed_addition_step(&L.q2x, &L.q2y, &r, &c);
    ell_apply(f, &c, pax, pay);
    if (USE_B) ell_apply(f, &table[idx], pbx, pby);
  } else if (USE_B) {
    ell_apply(f, &table[idx], pbx, pby);
    idx++;
    ell_apply(f, &table[idx], pbx, pby);
  }
}
BN_FN void miller_loop_2(fq12* f, bool use_a, const fq* pax, const fq* pay, const fq2* qax, const fq2* qay, bool use_b, const fq* pbx,
                         const fq* pby, const line_t* table) {
  if (use_a && use_b) miller_loop_t<true, true>(f, pax, pay, qax, qay, pbx, pby, table);
  else if (use_a) miller_loop_t<true, false>(f, pax, pay, qax, qay, pbx, pby, table);
  else if (use_b) miller_loop_t<false, true>(f, pax, pay, qax, qay, pbx, pby, table);
  else fq12_set_one(f);
}

// f^((q^12 - 1) / r): easy part (q^6 - 1)(q^2 + 1), then the hard part with three exponentiations by u
// (cyclotomic squarings) in the addition chain used by libff / substrate-bn for alt_bn128.
// Returns false for f == 0 (impossible for valid pairing inputs).
BN_NOINLINE bool final_exponentiation(fq12* out, const fq12* in) {
  if (fq12_is_zero(in)) return false;
  fq12 e, b, d, ee, g, t;
  // easy part
  fq12_inv(&t, in);
  fq12_conj(&e, in);
  fq12_mul(&t, &e, &t);      // f^(q^6 - 1)
  fq12_frobenius(&e, &t, 2);
  fq12_mul(&e, &e, &t);      // e = f^((q^6-1)(q^2+1))
  // hard part
  fq12_exp_by_neg_u(&t, &e);          // A
  fq12_cyclotomic_sqr(&b, &t);        // B = A^2
  fq12_cyclotomic_sqr(&t, &b);        // C = B^2
  fq12_mul(&d, &t, &b);               // D = C * B
  fq12_exp_by_neg_u(&ee, &d);         // E
  fq12_cyclotomic_sqr(&t, &ee);       // F = E^2
  fq12_exp_by_neg_u(&g, &t);          // G
  fq12_conj(&d, &d);                  // H = conj(D)
  fq12_conj(&g, &g);                  // I = conj(G)
  fq12_mul(&g, &g, &ee);              // J = I * E
  fq12_mul(&g, &g, &d);               // K = J * H
  fq12_mul(&b, &g, &b);               // L = K * B
  fq12_mul(&ee, &g, &ee);             // M = K * E
  fq12_mul(&ee, &ee, &e);             // N = M * e
  fq12_frobenius(&t, &b, 1);          // O = frob(L)
  fq12_mul(&ee, &t, &ee);             // P = O * N
  fq12_frobenius(&t, &g, 2);          // Q = frob2(K)
  fq12_mul(&ee, &t, &ee);             // R = Q * P
  fq12_conj(&e, &e);                  // S = conj(e)
  fq12_mul(&e, &e, &b);               // T = S * L
  fq12_frobenius(&t, &e, 3);          // U = frob3(T)
  fq12_mul(out, &t, &ee);             // U * R
  return true;
}

}  // namespace bn
