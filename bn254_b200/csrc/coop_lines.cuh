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

BN_FN void coop_emit_scaled(u4* lines, size_t set, size_t n_pad, size_t item, bool use, const line_t* c, const fq* px, const fq* py) {
  struct {
    fq2 l0, l3, l4;
  } L;
  if (use) {
    L.l0 = c->ell_0;
    fq2_scale(&L.l3, &c->ell_vw, py);  // position c1.c1 = w^3
    fq2_scale(&L.l4, &c->ell_vv, px);  // position c0.c2 = w^4
  } else {
    L.l0 = fq2_one();
    L.l3 = fq2_zero();
    L.l4 = fq2_zero();
  }
  coop_emit_line(lines, set, n_pad, item, L.l0, L.l3, L.l4);
}

// returns the decode status of (sig, pk); on ST_OK all 174 line sets of the item are written
BN_NOINLINE int item_verify_lines(u4* lines, size_t n_pad, size_t item, const g1aff* h, const uint8_t* sig, const uint8_t* pk,
                                  const line_t* table) {
  struct {
    g2j q;
    g1j s;
    g2proj r;
    line_t c;
    fq2 qy_sel, q1x, q1y, q2x, q2y;
  } L;
  int st = g2_from_raw(&L.q, pk);
  if (st) return st;
  st = g1_from_raw(&L.s, sig);
  if (st) return st;
  const bool use_a = !pt_is_inf(&L.q), use_b = !pt_is_inf(&L.s);
  L.r.x = L.q.x;
  L.r.y = L.q.y;
  L.r.z = fq2_one();
  size_t m = 0;
  for (int k = 0; k < 64; k++) {
    if (use_a) doubling_step(&L.r, &L.c);
    coop_emit_scaled(lines, 2 * m, n_pad, item, use_a, &L.c, &h->x, &h->y);
    coop_emit_scaled(lines, 2 * m + 1, n_pad, item, use_b, &table[m], &L.s.x, &L.s.y);
    m++;
    int d = K_ATE_DIGITS[k];
    if (d != 0) {
      if (use_a) {
        L.qy_sel = d > 0 ? L.q.y : fq2_neg(L.q.y);
        mixed_addition_step(&L.q.x, &L.qy_sel, &L.r, &L.c);
      }
      coop_emit_scaled(lines, 2 * m, n_pad, item, use_a, &L.c, &h->x, &h->y);
      coop_emit_scaled(lines, 2 * m + 1, n_pad, item, use_b, &table[m], &L.s.x, &L.s.y);
      m++;
    }
  }
  if (use_a) g2_frobenius_pair(&L.q1x, &L.q1y, &L.q2x, &L.q2y, L.q.x, L.q.y);
  if (use_a) mixed_addition_step(&L.q1x, &L.q1y, &L.r, &L.c);
  coop_emit_scaled(lines, 2 * m, n_pad, item, use_a, &L.c, &h->x, &h->y);
  coop_emit_scaled(lines, 2 * m + 1, n_pad, item, use_b, &table[m], &L.s.x, &L.s.y);
  m++;
  if (use_a) mixed_addition_step(&L.q2x, &L.q2y, &L.r, &L.c);
  coop_emit_scaled(lines, 2 * m, n_pad, item, use_a, &L.c, &h->x, &h->y);
  coop_emit_scaled(lines, 2 * m + 1, n_pad, item, use_b, &table[m], &L.s.x, &L.s.y);
  return ST_OK;
}

}  // namespace bn
