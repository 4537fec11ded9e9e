// hostsim.cpp -- compiles the engine's per-item device code (bn254_b200/csrc/*.cuh) with g++ for the CPU.
// TEST INFRASTRUCTURE ONLY: lets `pytest -m "not gpu"` check the control logic of the kernels (tower formulas,
// Miller schedule, final-exponentiation chain, hash loop, codecs) against the oracle on a machine without a GPU.
// It is never loaded by the product library or by bench.py; the PTX Montgomery product is device-only and is
// checked on the GPU by the `-m gpu` tests.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include "../../bn254_b200/csrc/items.cuh"

using namespace bn;
#define API extern "C" __attribute__((visibility("default")))

static line_t g_lines[K_N_LINES];
static bool g_init = false;
static void ensure_init() {
  if (g_init) return;
  fq2 gx = fq2_from_limbs(K_G2_GEN_X), gy = fq2_neg(fq2_from_limbs(K_G2_GEN_Y));
  g2_precompute_lines(g_lines, gx, gy);
  g_init = true;
}

API int hs_hash_to_g1(const uint8_t* msg, size_t len, uint8_t* out, int* ctr) {
  g1aff h;
  int st = hash_to_g1(&h.x, &h.y, msg, len, ctr);
  if (st) { memset(out, 0, 64); return st; }
  fq_to_be(out, h.x); fq_to_be(out + 32, h.y);
  return 0;
}
// the same with the try limit of the device's test hook (bn254_set_hash_try_limit) and a first counter (the counter-parallel kernel)
API int hs_hash_to_g1_limited(const uint8_t* msg, size_t len, uint8_t* out, int* ctr, int max_tries, int ctr_first) {
  g1aff h;
  int st = hash_to_g1(&h.x, &h.y, msg, len, ctr, max_tries, ctr_first);
  if (st) { memset(out, 0, 64); return st; }
  fq_to_be(out, h.x); fq_to_be(out + 32, h.y);
  return 0;
}
API int hs_sign(const uint8_t* msg, size_t len, const uint8_t* sk, uint8_t* sig) {
  g1aff h;
  int st = hash_to_g1(&h.x, &h.y, msg, len, nullptr);
  if (st) { memset(sig, 0, 64); return st; }
  item_sign(sig, &h, sk);
  return 0;
}
API int hs_verify(const uint8_t* msg, size_t len, const uint8_t* sig, const uint8_t* pk) {
  ensure_init();
  g1aff h;
  int st = hash_to_g1(&h.x, &h.y, msg, len, nullptr);
  if (st) return st;
  fq12 f;
  st = item_verify_miller(&f, &h, sig, pk, g_lines);
  if (st) return st;
  return item_final_exp_is_one(&f);
}
API int hs_check_public_keys(const uint8_t* pk_g2, const uint8_t* pk_g1) {
  ensure_init();
  g1aff h;
  h.x = fq_from_limbs(K_G1_GEN_X); h.y = fq_from_limbs(K_G1_GEN_Y);
  fq12 f;
  int st = item_verify_miller(&f, &h, pk_g1, pk_g2, g_lines);
  if (st) return st;
  return item_final_exp_is_one(&f);
}
API int hs_verify_miller(const uint8_t* msg, size_t len, const uint8_t* sig, const uint8_t* pk, uint8_t* f_out) {
  ensure_init();
  g1aff h;
  int st = hash_to_g1(&h.x, &h.y, msg, len, nullptr);
  if (st) return st;
  fq12 f;
  st = item_verify_miller(&f, &h, sig, pk, g_lines);
  if (st) return st;
  fq12_to_be(f_out, &f);
  return 0;
}
API int hs_miller_product(const uint8_t* g1s, const uint8_t* g2s, size_t k, uint8_t* f_out) {
  fq12 f;
  int st = item_miller_pairs(&f, g1s, g2s, k);
  if (st) return st;
  fq12_to_be(f_out, &f);
  return 0;
}
API int hs_pairing_check(const uint8_t* g1s, const uint8_t* g2s, size_t k) {
  fq12 f;
  int st = item_miller_pairs(&f, g1s, g2s, k);
  if (st) return st;
  return item_final_exp_is_one(&f);
}
API int hs_final_exp(const uint8_t* f_in, uint8_t* gt_out) {
  fq12 f, gt;
  if (!fq12_from_be(&f, f_in)) return ST_NOT_MEMBER;
  if (!final_exponentiation(&gt, &f)) return ST_TO_AFFINE;
  fq12_to_be(gt_out, &gt);
  return 0;
}
API int hs_fq_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  fq x, y, r;
  if (!fq_from_be(&x, a)) return ST_NOT_MEMBER;
  if (op <= 2 && !fq_from_be(&y, b)) return ST_NOT_MEMBER;
  if (op == 0) r = fq_mul(x, y);
  else if (op == 1) r = fq_add(x, y);
  else if (op == 2) r = fq_sub(x, y);
  else if (op == 3) r = fq_inv(x);
  else if (op == 4) { if (!fq_sqrt(&r, x)) return ST_NOT_MEMBER; }
  else return ST_INVALID_ENCODING;
  fq_to_be(out, r);
  return 0;
}
API int hs_fq12_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  fq12 x, y, r;
  if (!fq12_from_be(&x, a)) return ST_NOT_MEMBER;
  if (op == 0) { if (!fq12_from_be(&y, b)) return ST_NOT_MEMBER; fq12_mul(&r, &x, &y); }
  else if (op == 1) fq12_sqr(&r, &x);
  else if (op == 2) fq12_inv(&r, &x);
  else if (op == 3) fq12_cyclotomic_sqr(&r, &x);
  else if (op >= 4 && op <= 6) fq12_frobenius(&r, &x, op - 3);
  else if (op == 7) fq12_conj(&r, &x);
  else return ST_INVALID_ENCODING;
  fq12_to_be(out, &r);
  return 0;
}
API int hs_g1_mul(const uint8_t* pt, const uint8_t* k, uint8_t* out) { return item_g1_mul(out, pt, k); }
API int hs_g2_mul(const uint8_t* pt, const uint8_t* k, uint8_t* out) { return item_g2_mul(out, pt, k); }
API int hs_g1_add(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  g1j p, q, r;
  int st;
  if ((st = g1_from_raw(&p, a)) || (st = g1_from_raw(&q, b))) return st;
  pt_add(&r, &p, &q);
  g1_to_raw(out, &r);
  return 0;
}
API int hs_g1_madd(const uint8_t* a, const uint8_t* b, uint8_t* out) {  // b finite
  g1j p, q, r;
  int st;
  if ((st = g1_from_raw(&p, a)) || (st = g1_from_raw(&q, b))) return st;
  pt_madd(&r, &p, &q.x, &q.y);
  g1_to_raw(out, &r);
  return 0;
}
API int hs_g2_add(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  g2j p, q, r;
  int st;
  if ((st = g2_from_raw(&p, a)) || (st = g2_from_raw(&q, b))) return st;
  pt_add(&r, &p, &q);
  g2_to_raw(out, &r);
  return 0;
}
API int hs_g2_madd(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  g2j p, q, r;
  int st;
  if ((st = g2_from_raw(&p, a)) || (st = g2_from_raw(&q, b))) return st;
  pt_madd(&r, &p, &q.x, &q.y);
  g2_to_raw(out, &r);
  return 0;
}
API int hs_derive_pk_g1(const uint8_t* sk, uint8_t* out) { item_derive_pk_g1(out, sk); return 0; }
API int hs_derive_pk_g2(const uint8_t* sk, uint8_t* out) { item_derive_pk_g2(out, sk); return 0; }
API int hs_g1_compress(const uint8_t* raw, uint8_t* out) { return item_g1_compress(out, raw); }
API int hs_g1_decompress(const uint8_t* in, size_t len, uint8_t* out) {
  if (len != 33) { memset(out, 0, 64); return ST_INVALID_ENCODING; }
  return item_g1_decompress(out, in);
}
API int hs_g2_compress(const uint8_t* raw, uint8_t* out) { return item_g2_compress(out, raw); }
API int hs_g2_decompress(const uint8_t* in, size_t len, uint8_t* out) {
  if (len != 65) { memset(out, 0, 128); return ST_INVALID_ENCODING; }
  return item_g2_decompress(out, in);
}
API int hs_g1_validate(const uint8_t* raw, size_t len) { return len == 64 ? item_g1_validate(raw) : ST_INVALID_LENGTH; }
API int hs_g2_validate(const uint8_t* raw, size_t len) { return len == 128 ? item_g2_validate(raw) : ST_INVALID_LENGTH; }
API int hs_layer_op(int op, const uint8_t* in, int n_in, uint8_t* out, int n_out) {
  fq a[20], r[12];
  for (int k = 0; k < 12; k++) r[k] = fq_zero();
  for (int k = 0; k < n_in && k < 20; k++) fq_from_be(&a[k], in + 32 * k);
  debug_layer_op(op, a, r);
  for (int k = 0; k < n_out && k < 12; k++) fq_to_be(out + 32 * k, r[k]);
  return 0;
}

// (9 x + z) mod q with one reduction (fq.cuh fq_mul9_add); x, z big-endian integers, z may equal q; plain (non-Montgomery) values
API void hs_mul9_add(const uint8_t* x_be, const uint8_t* z_be, uint8_t* out_be) {
  fq x, z;
  u256_from_be(x.l, x_be);
  u256_from_be(z.l, z_be);
  fq r = fq_mul9_add(x, z, &K_KQ_TABLE[0][0]);
  for (int i = 0; i < 8; i++) {
    uint32_t w = r.l[7 - i];
    out_be[4 * i] = (uint8_t)(w >> 24); out_be[4 * i + 1] = (uint8_t)(w >> 16); out_be[4 * i + 2] = (uint8_t)(w >> 8); out_be[4 * i + 3] = (uint8_t)w;
  }
}

// (3 t + 2 z) mod q with one reduction (fq.cuh fq_3t_2z); z may equal q
API void hs_3t_2z(const uint8_t* t_be, const uint8_t* z_be, uint8_t* out_be) {
  fq t, z;
  u256_from_be(t.l, t_be);
  u256_from_be(z.l, z_be);
  fq r = fq_3t_2z(t, z, &K_KQ_TABLE[0][0]);
  for (int i = 0; i < 8; i++) {
    uint32_t w = r.l[7 - i];
    out_be[4 * i] = (uint8_t)(w >> 24); out_be[4 * i + 1] = (uint8_t)(w >> 16); out_be[4 * i + 2] = (uint8_t)(w >> 8); out_be[4 * i + 3] = (uint8_t)w;
  }
}

// GLV decomposition used by signing: k (32 bytes big-endian, any value; reduced mod r first like Fr::from_slice) ->
// |k1|, |k2| as 4 little-endian u32 limbs each, signs in sg[0], sg[1]
API void hs_glv_decompose(const uint8_t* k_be, uint32_t* k1, uint32_t* k2, uint8_t* sg) {
  uint32_t k[8];
  fr_reduce(k, k_be);
  bool n1, n2;
  glv_decompose(k1, &n1, k2, &n2, k);
  sg[0] = n1; sg[1] = n2;
}

// randomised batch verification, one item's share: c * H(msg) (raw) and c * sig (raw) ; returns the ride / no-ride status
API int hs_rlc_prepare(const uint8_t* msg, size_t len, const uint8_t* sig, const uint8_t* pk, const uint8_t* c16, int check_g2,
                       uint8_t* hs_out, uint8_t* sigc_out) {
  g1aff h, hs;
  int st = hash_to_g1(&h.x, &h.y, msg, len, nullptr);
  if (st) return st;
  st = item_rlc_prepare(&hs, sigc_out, &h, sig, pk, c16, check_g2 != 0);
  if (st) return st;
  fq_to_be(hs_out, hs.x); fq_to_be(hs_out + 32, hs.y);
  return 0;
}

// ---------------------------------------------------------------------------------------------- cooperative machine
// The six warps of a block are simulated one after the other, phase by phase, for lane 0 of one block.
#include "../../bn254_b200/csrc/coop_lines.cuh"
#include <vector>

static int g_coop_wmode = 0;  // 0: block layout (32 items per row, warp k = coefficient k), 1: warp-local layout (5 items per row),
                              // 2: half-warp layout (16 items per row, two coefficients per warp)
API void hs_coop_set_layout(int wmode) { g_coop_wmode = wmode; }

struct coop_sim {
  std::vector<u4> sm, lines, gslots, fio;
  uint8_t status = 0;
  size_t n_pad = COOP_LANES;
  bool wmode = g_coop_wmode == 1;
  int row = g_coop_wmode == 1 ? COOPW_ROW : g_coop_wmode == 2 ? COOPH_ROW : COOP_LANES;
  coop_sim() : sm(COOP_SLOTS * 2 * COOP_LANES), lines((size_t)COOP_MULTI_K * K_N_LINES * COOP_LINE_FQ * 2 * COOP_LANES),
               gslots((size_t)COOP_GSLOTS * 6 * 2 * 2 * COOP_LANES), fio((size_t)6 * 2 * 2 * COOP_LANES) {}
  int lanes = 1;  // simulated items of the group (32 for the multi-pairing butterfly, block layout only)
  int line_next[COOP_WARPS][COOP_LANES] = {};
  coop_ctx ctx(int k, int lane) {
    coop_ctx c;
    c.sm = sm.data() + lane; c.row = row; c.wmode = wmode; c.plans = g_coop_wmode == 1 ? K_COOP_PLANS_W : g_coop_wmode == 2 ? K_COOP_PLANS_H : K_COOP_PLANS;
    c.kq = g_coop_wmode == 0 ? &K_KQ_TABLE[0][0] : nullptr;  // the default kernel's one-reduction xi variants; the other layouts keep the additions
    c.k = k; c.lane = lane; c.active = true; c.item = lane; c.n_pad = n_pad;
    c.lines = lines.data(); c.gslots = gslots.data(); c.fio = fio.data(); c.status = &status; c.progress = nullptr; c.sets_per_step = 1;
    return c;
  }
  void run(const uint32_t* prog) {
    std::vector<fq2> t(COOP_WARPS * COOP_LANES);
    for (int pc = 0;; pc++) {
      uint32_t ins = prog[pc];
      if ((ins & 0xff) == COP_END) break;
      if (wmode && (ins & 0xff) == COP_INVT) {  // k_coopw_run: the block's items are inverted by one warp between two block barriers
        for (int l = 0; l < lanes; l++) coop_invt(sm.data() + l, row);
        continue;
      }
      if (g_coop_wmode == 0 && (ins & 0xff) == COP_DOT) {
        // block layout: the fast path the device kernels run (coop_run_block): fetch, coop_dot_block_any, barrier, coop_commit_block
        const int plan = (ins >> 8) & 0xff;
        for (int l = 0; l < lanes; l++)
          for (int k = 0; k < COOP_WARPS; k++) {
            coop_ctx c = ctx(k, l);
            if ((ins >> 16) & 1) {
              coop_line_fetch(c, line_next[k][l], line_next[k][l] & 1);
              line_next[k][l]++;
            }
            t[k * COOP_LANES + l] = coop_dot_block_any(c.sm, plan, k, c.kq);
          }
        for (int l = 0; l < lanes; l++)
          for (int k = 0; k < COOP_WARPS; k++) {
            coop_ctx c = ctx(k, l);
            coop_commit_block(c.sm, k, K_COOP_PLANS[plan][k][0], t[k * COOP_LANES + l], (ins >> 25) & 0x3f, c.kq);
          }
        continue;
      }
      for (int l = 0; l < lanes; l++)
        for (int k = 0; k < COOP_WARPS; k++) t[k * COOP_LANES + l] = coop_phase_a(ctx(k, l), ins, line_next[k][l]);
      for (int l = 0; l < lanes; l++)
        for (int k = 0; k < COOP_WARPS; k++) coop_phase_b(ctx(k, l), ins, t[k * COOP_LANES + l], line_next[k][l]);
    }
  }
  // tower order c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 <- a0, a2, a4, a1, a3, a5
  void get_fio(fq12* f) {
    static const int pos[6] = {0, 2, 4, 1, 3, 5};
    fq2* c = &f->c0.c0;
    for (int t = 0; t < 6; t++) {
      c[t].c0 = coop_gld(fio.data(), (size_t)pos[t] * 2 + 0, n_pad, 0);
      c[t].c1 = coop_gld(fio.data(), (size_t)pos[t] * 2 + 1, n_pad, 0);
    }
  }
  void set_fio(const fq12* f) {
    static const int pos[6] = {0, 2, 4, 1, 3, 5};
    const fq2* c = &f->c0.c0;
    for (int t = 0; t < 6; t++) {
      coop_gst(fio.data(), (size_t)pos[t] * 2 + 0, n_pad, 0, c[t].c0);
      coop_gst(fio.data(), (size_t)pos[t] * 2 + 1, n_pad, 0, c[t].c1);
    }
  }
};

// full verify through the cooperative program; h_is_gen: check_public_keys form
API int hs_coop_verify(const uint8_t* msg, size_t len, const uint8_t* sig, const uint8_t* pk, int h_is_gen) {
  ensure_init();
  g1aff h;
  if (h_is_gen) {
    h.x = fq_from_limbs(K_G1_GEN_X); h.y = fq_from_limbs(K_G1_GEN_Y);
  } else {
    int st = hash_to_g1(&h.x, &h.y, msg, len, nullptr);
    if (st) return st;
  }
  coop_sim S;
  lines_consts K;
  int st = item_verify_lines(S.lines.data(), S.n_pad, 0, &h, sig, pk, g_lines, &K);
  if (st) return st;
  S.run(K_COOP_PROG_VERIFY);
  return S.status;
}
// The cooperative walk (coop_lines.cuh walk_*), simulated: per level every (warp, lane) in turn, a barrier = the end of the sweep.
// Fills the line sets of lane 0 .. lanes - 1 (each lane gets the same item) and returns the item's status.
static int walk4_sim(std::vector<u4>& lines, size_t n_pad, int lanes, const g1aff* h, const uint8_t* sig, const uint8_t* pk) {
  std::vector<u4> sm(WS_SLOTS * 2 * 2 * COOP_LANES);
  std::vector<int> flags(4 * COOP_LANES);
  auto ctx = [&](int w, int l) {
    walk_ctx c;
    c.sm = sm.data() + l; c.flags = flags.data(); c.row = COOP_LANES; c.lane = l; c.warp = w;
    c.item = (size_t)l; c.n = (size_t)lanes; c.n_pad = n_pad; c.lines = lines.data(); c.table = g_lines;
    c.live = c.use_a = c.use_b = false;
    return c;
  };
  for (int w = 0; w < WALK_WARPS; w++)
    for (int l = 0; l < lanes; l++) walk_decode(ctx(w, l), h, sig, pk, false);
  int st = 0;
  {
    walk_ctx c = ctx(0, 0);
    st = walk_flags(c, false);
  }
  int steps_done = 0;
  walk_schedule(
      [&](int kind, int level, size_t m, int sqx, int sqy) {
        for (int w = 0; w < WALK_WARPS; w++)
          for (int l = 0; l < lanes; l++) {
            walk_ctx c = ctx(w, l);
            walk_flags(c, false);
            if (kind == 0) walk_dbl<lines_mul_call>(c, level, m); else walk_add<lines_mul_call>(c, level, m, sqx, sqy);
          }
      },
      [] {}, [&](size_t m) { steps_done = (int)m; });
  if (steps_done != K_N_LINES) return -1;
  return st;
}
// both producers on one item: 0 = same status and (status OK) bit-identical line sets; out_status = the status
API int hs_walk4_matches(const uint8_t* msg, size_t len, const uint8_t* sig, const uint8_t* pk, int* out_status) {
  ensure_init();
  g1aff h;
  int st = hash_to_g1(&h.x, &h.y, msg, len, nullptr);
  if (st) return -2;
  const size_t n_pad = COOP_LANES, words = (size_t)2 * K_N_LINES * COOP_LINE_FQ * 2 * n_pad;
  u4 zero; zero.x = zero.y = zero.z = zero.w = 0;
  std::vector<u4> a(words, zero), b(words, zero);
  lines_consts K;
  const int st_a = item_verify_lines(a.data(), n_pad, 0, &h, sig, pk, g_lines, &K);
  const int st_b = walk4_sim(b, n_pad, 1, &h, sig, pk);
  *out_status = st_b;
  if (st_a != st_b) return 1;
  if (st_a) return 0;
  return memcmp(a.data(), b.data(), words * sizeof(u4)) == 0 ? 0 : 2;
}
// full verify with the cooperative walk as the producer of the machine's line sets
API int hs_coop_verify_walk4(const uint8_t* msg, size_t len, const uint8_t* sig, const uint8_t* pk) {
  ensure_init();
  g1aff h;
  int st = hash_to_g1(&h.x, &h.y, msg, len, nullptr);
  if (st) return st;
  coop_sim S;
  st = walk4_sim(S.lines, S.n_pad, 1, &h, sig, pk);
  if (st) return st;
  S.run(K_COOP_PROG_VERIFY);
  return S.status;
}
// The latency layouts of the machine (coop.cuh coop_run_block12<SPLIT>: twelve / eighteen warps per group, the Karatsuba components
// of a row on different warps, handed over through exchange slots): every simulated warp is a host THREAD running the device
// function itself, the block barrier is a std::barrier.  One item (lane 0); the line sets come from the cooperative walk.
#include <barrier>
#include <thread>
template <int SPLIT>
static int coop_split_verify(const uint8_t* msg, size_t len, const uint8_t* sig, const uint8_t* pk) {
  g1aff h;
  int st = hash_to_g1(&h.x, &h.y, msg, len, nullptr);
  if (st) return st;
  coop_sim S;
  st = walk4_sim(S.lines, S.n_pad, 1, &h, sig, pk);
  if (st) return st;
  std::vector<u4> xch(12 * 2 * COOP_LANES);
  const int nw = SPLIT * COOP_WARPS;
  std::barrier<> bar(nw);
  std::vector<std::thread> th;
  for (int w = 0; w < nw; w++)
    th.emplace_back([&, w] {
      coop_ctx c = S.ctx(w % COOP_WARPS, 0);
      c.sets_per_step = 2;
      coop_run_block12<SPLIT>(c, w / COOP_WARPS, xch.data(), K_COOP_PROG_VERIFY, [&] { bar.arrive_and_wait(); });
    });
  for (auto& t : th) t.join();
  return S.status;
}
API int hs_coop_verify_split(const uint8_t* msg, size_t len, const uint8_t* sig, const uint8_t* pk, int split) {
  ensure_init();
  if (g_coop_wmode != 0) return -1;
  return split == 3 ? coop_split_verify<3>(msg, len, sig, pk) : coop_split_verify<2>(msg, len, sig, pk);
}
// Miller product of the verify pairs (tower order, big-endian) through the cooperative program
API int hs_coop_verify_miller(const uint8_t* msg, size_t len, const uint8_t* sig, const uint8_t* pk, uint8_t* f_out) {
  ensure_init();
  g1aff h;
  int st = hash_to_g1(&h.x, &h.y, msg, len, nullptr);
  if (st) return st;
  coop_sim S;
  lines_consts K;
  st = item_verify_lines(S.lines.data(), S.n_pad, 0, &h, sig, pk, g_lines, &K);
  if (st) return st;
  S.run(K_COOP_PROG_MILLER2);
  fq12 f;
  S.get_fio(&f);
  fq12_to_be(f_out, &f);
  return 0;
}
API int hs_coop_final_exp(const uint8_t* f_in, uint8_t* gt_out) {
  fq12 f, gt;
  if (!fq12_from_be(&f, f_in)) return ST_NOT_MEMBER;
  coop_sim S;
  S.set_fio(&f);
  S.run(K_COOP_PROG_FINALEXP);
  S.get_fio(&gt);
  fq12_to_be(gt_out, &gt);
  return S.status;
}
// one plan applied to given P (and S) values: op 0 MUL (P <- S*P), 1 SQR, 3 CYCLO ; values in tower order
API int hs_coop_plan(int plan, const uint8_t* p_in, const uint8_t* s_in, uint8_t* out) {
  fq12 p, s, r;
  if (!fq12_from_be(&p, p_in)) return ST_NOT_MEMBER;
  if (s_in && !fq12_from_be(&s, s_in)) return ST_NOT_MEMBER;
  coop_sim S;
  S.set_fio(&p);
  uint32_t prog[8];
  int n = 0;
  if (s_in) {
    S.set_fio(&s);
    prog[n++] = COP_LOADF;
    prog[n++] = COP_STORE | (0 << 8);
    S.run((prog[n] = COP_END, prog));
    n = 0;
    S.set_fio(&p);
  }
  prog[n++] = COP_LOADF;
  if (s_in) prog[n++] = COP_LOADS | (0 << 8);
  prog[n++] = COP_DOT | (plan << 8);
  prog[n++] = COP_STOREF;
  prog[n++] = COP_END;
  S.run(prog);
  S.get_fio(&r);
  fq12_to_be(out, &r);
  return 0;
}

// multi-pairing program `which` (CPROG_MULTI8 / 4 / 2 / 1): n <= 32 * mk pairs (g1 64 B, g2 128 B each) -> the block's Miller
// product (tower order)
API int hs_coop_multi_miller_k(int which, const uint8_t* g1s, const uint8_t* g2s, size_t n, uint8_t* f_out) {
  const int mk = coop_multi_k(which);
  coop_sim S;
  S.lanes = COOP_LANES;
  lines_consts K;
  for (size_t slot = 0; slot < (size_t)COOP_LANES * mk; slot++) {
    size_t lane = slot % COOP_LANES;
    int stream = (int)(slot / COOP_LANES);
    bool use = false;
    g1aff h;
    g2j q;
    q.x = fq2_one();
    q.y = fq2_one();
    if (slot < n) {
      g1j p;
      int st = g1_from_raw(&p, g1s + 64 * slot);
      if (st) return st;
      st = g2_from_raw(&q, g2s + 128 * slot);
      if (st) return st;
      use = !pt_is_inf(&p) && !pt_is_inf(&q);
      h.x = p.x;
      h.y = p.y;
    }
    item_pair_lines(S.lines.data(), S.n_pad, lane, stream, mk, use, &h, q.x, q.y, &K);
  }
  S.run(coop_program(which));
  fq12 f;
  S.get_fio(&f);
  fq12_to_be(f_out, &f);
  return 0;
}
API int hs_coop_multi_miller(const uint8_t* g1s, const uint8_t* g2s, size_t n, uint8_t* f_out) {
  return hs_coop_multi_miller_k(CPROG_MULTI8, g1s, g2s, n, f_out);
}
// finish of an aggregate check (program FINISH, what bn254_finish_distinct_dev runs): f_in = product of the exchanged partials,
// sig = the G1 point paired with -G2 (all-zero = infinity: that pair is skipped).  Returns the verdict.
API int hs_coop_finish(const uint8_t* f_in, const uint8_t* sig) {
  ensure_init();
  fq12 f;
  if (!fq12_from_be(&f, f_in)) return ST_NOT_MEMBER;
  g1j s;
  int st = g1_from_raw(&s, sig);
  if (st) return st;
  coop_sim S;
  S.set_fio(&f);
  fq2 sxy;
  sxy.c0 = s.x;
  sxy.c1 = s.y;
  for (int m = 0; m < K_N_LINES; m++)
    coop_emit_scaled_v(S.lines.data(), m, S.n_pad, 0, !pt_is_inf(&s), g_lines[m].ell_0, g_lines[m].ell_vw, g_lines[m].ell_vv, sxy);
  S.run(K_COOP_PROG_FINISH);
  return S.status;
}

// fixed-base tables (curve.cuh comb_build_row / pt_mul_fixed), built on the host exactly as k_init_comb does on the device
static aff<fq> g_comb1[BN_COMB_WINDOWS * BN_COMB_ROW];
static aff<fq2> g_comb2[BN_COMB_WINDOWS * BN_COMB_ROW];
static bool g_comb_init = false;
static void ensure_comb() {
  if (g_comb_init) return;
  for (int w = 0; w < BN_COMB_WINDOWS; w++) {
    comb_build_row(g_comb1 + w * BN_COMB_ROW, w, fq_from_limbs(K_G1_GEN_X), fq_from_limbs(K_G1_GEN_Y));
    comb_build_row(g_comb2 + w * BN_COMB_ROW, w, fq2_from_limbs(K_G2_GEN_X), fq2_from_limbs(K_G2_GEN_Y));
  }
  g_comb_init = true;
}
API void hs_derive_pk_g1_comb(const uint8_t* sk, uint8_t* out) { ensure_comb(); item_derive_pk_g1_comb(out, sk, g_comb1); }
API void hs_derive_pk_g2_comb(const uint8_t* sk, uint8_t* out) { ensure_comb(); item_derive_pk_g2_comb(out, sk, g_comb2); }
