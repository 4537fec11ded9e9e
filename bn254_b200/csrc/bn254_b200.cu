// bn254_b200.cu -- CUDA kernels (sm_100a) and the C ABI of the BN254 batch engine (include/bn254_b200.h).
//
// One thread owns one item (message / signature / key / pairing product).  The path is integer-pipe bound
// (IMAD.WIDE carry chains), not HBM or tensor bound: a verify reads 224 B and issues ~3 M multiply-adds.
// Phases are separate kernels so that the divergent try-and-increment hash, the uniform Miller loop and the
// uniform final exponentiation each run with their own register budget and can be profiled on their own:
//     k_hash_to_g1 -> k_verify_miller -> k_final_exp_check            (ECDSA::verify, /root/reference/src/ecdsa.rs:49-64)
//     k_hash_to_g1 -> k_sign                                          (ECDSA::sign,   /root/reference/src/ecdsa.rs:26-35)
//     k_sum_partial<F> -> k_sum_final<F>                              (Add/Sub/Neg folds, /root/reference/src/types.rs:126-286)
//     k_hash_to_g1 -> k_distinct_partial -> k_fq12_prod_final -> k_distinct_finish   (multi-pairing, one final exponentiation)
// There is no CPU fallback anywhere in this file: without a CUDA device bn254_ctx_create fails.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/bn254_b200.h"
#include "items.cuh"
#include "coop_lines.cuh"

using namespace bn;

#define BN_BLOCK 128
#define BN_PROD_BLOCK 64
// minimum resident blocks per SM requested from ptxas for the two pairing kernels (caps registers per thread)
#ifndef BN_MINB
#define BN_MINB 1
#endif

// ------------------------------------------------------------------------------------------------ kernels
// fixed-base tables of the two generators: one thread per window row
__global__ void k_init_comb(aff<fq>* t1, aff<fq2>* t2) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= BN_COMB_WINDOWS) return;
  comb_build_row(t1 + w * BN_COMB_ROW, w, fq_from_limbs(K_G1_GEN_X), fq_from_limbs(K_G1_GEN_Y));
  comb_build_row(t2 + w * BN_COMB_ROW, w, fq2_from_limbs(K_G2_GEN_X), fq2_from_limbs(K_G2_GEN_Y));
}
__global__ void k_init_lines(line_t* out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    fq2 gx = fq2_from_limbs(K_G2_GEN_X), gy = fq2_neg(fq2_from_limbs(K_G2_GEN_Y));
    g2_precompute_lines(out, gx, gy);
  }
}

// hash_to_try_and_increment for message i = msgs[i*msg_len ..] (offsets == NULL) or msgs[offsets[i] .. offsets[i+1])
__global__ void __launch_bounds__(BN_BLOCK) k_hash_to_g1(const uint8_t* __restrict__ msgs, size_t msg_len, const uint64_t* __restrict__ offsets,
                                                         size_t n, g1aff* __restrict__ H, uint8_t* __restrict__ status,
                                                         uint8_t* __restrict__ tries, int max_tries) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* m = offsets ? msgs + offsets[i] : msgs + i * msg_len;
  uint64_t len = offsets ? offsets[i + 1] - offsets[i] : msg_len;
  g1aff h;
  int ctr = 0;
  int st = hash_to_g1(&h.x, &h.y, m, len, &ctr, max_tries);
  if (st) {
    h.x = fq_zero();
    h.y = fq_zero();
  }
  H[i] = h;
  status[i] = (uint8_t)st;
  if (tries) tries[i] = (uint8_t)ctr;
}

// ---- compacting hash: round r tries counter r for every item still without a point; items that fail are appended to the
// next round's list (warp-aggregated), so a warp never idles behind its slowest lane (the accept probability is 0.47 per
// try: the per-thread loop above spends ~3x the useful tries at warp granularity).  Messages of at most 54 bytes only.
__global__ void __launch_bounds__(BN_BLOCK) k_hash_round(const uint8_t* __restrict__ msgs, uint32_t msg_len, uint32_t n, uint32_t ctr,
                                                         const uint32_t* __restrict__ list_in, const uint32_t* __restrict__ count_in,
                                                         uint32_t* __restrict__ list_out, uint32_t* __restrict__ count_out,
                                                         g1aff* __restrict__ H, uint8_t* __restrict__ status) {
  const uint32_t total = list_in ? *count_in : n;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const uint32_t i = list_in ? list_in[t] : t;
    g1aff h;
    const bool ok = hash_try_1blk(&h.x, &h.y, msgs + (size_t)i * msg_len, msg_len, ctr);
    if (ok) {
      H[i] = h;
      status[i] = ST_OK;
    }
    const unsigned active = __activemask();
    const unsigned fail = __ballot_sync(active, !ok);
    if (!ok) {
      const unsigned lane = threadIdx.x & 31, leader = __ffs(fail) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(count_out, (uint32_t)__popc(fail));
      base = __shfl_sync(fail, base, leader);
      list_out[base + __popc(fail & ((1u << lane) - 1))] = i;
    }
  }
}
// ---- counter-parallel hash: G lanes per message (G = 32, 8, 4, 2), lane j of a group tries counter base + j, the lowest accepted
// counter wins (exactly the point the sequential loop returns), groups that found nothing go on with the next G counters.  G x the
// work of the loop per step, but one step (0.2 ms: the 254 dependent squarings of the square root) almost always decides.  Used for
// small and mid-size batches, where the loop's latency is the slowest lane's try count -- G is chosen so that a step's G n tries
// fit the ~28 k tries the GPU completes in the latency of one -- and (G = 32) for the survivors of the compacting rounds.
// list_in == NULL: items 0 .. n-1; else items list_in[0 .. *count_in).
template <int G>
__global__ void __launch_bounds__(BN_BLOCK) k_hash_wide(const uint8_t* __restrict__ msgs, size_t msg_len, const uint64_t* __restrict__ offsets,
                                                        size_t n, const uint32_t* __restrict__ list_in, const uint32_t* __restrict__ count_in,
                                                        int ctr0, int max_tries, g1aff* __restrict__ H, uint8_t* __restrict__ status,
                                                        uint8_t* __restrict__ tries) {
  const size_t total = list_in ? (size_t)*count_in : n;
  const unsigned lane = threadIdx.x & 31, sub = lane % G, grp = lane / G;
  const unsigned gmask = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u) << (grp * G);
  const size_t per_warp = 32 / G, warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t wb = (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * per_warp; wb < total; wb += warps * per_warp) {
    const size_t w = wb + grp;
    const bool valid = w < total;
    const size_t i = valid ? (list_in ? list_in[w] : w) : 0;
    const uint8_t* m = offsets ? msgs + offsets[i] : msgs + i * msg_len;
    const uint64_t len = offsets ? offsets[i + 1] - offsets[i] : msg_len;
    bool done = !valid;
    for (int base = ctr0; base < max_tries && __any_sync(0xffffffffu, !done); base += G) {
      const int ctr = base + (int)sub;
      g1aff h;
      bool ok = false;
      if (!done && ctr < max_tries) ok = hash_to_g1(&h.x, &h.y, m, len, nullptr, ctr + 1, ctr) == ST_OK;
      const unsigned hit = __ballot_sync(0xffffffffu, ok) & gmask;
      if (!done && hit) {
        if (lane == (unsigned)(__ffs(hit) - 1)) {
          H[i] = h;
          status[i] = ST_OK;
          if (tries) tries[i] = (uint8_t)ctr;
        }
        done = true;
      }
    }
    if (!done && sub == 0) {
      g1aff z;
      z.x = fq_zero();
      z.y = fq_zero();
      H[i] = z;
      status[i] = ST_HASH_TO_POINT;
      if (tries) tries[i] = 0;
    }
  }
}

__global__ void __launch_bounds__(BN_BLOCK) k_g1aff_to_raw(const g1aff* __restrict__ H, const uint8_t* __restrict__ status, size_t n,
                                                           uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (status[i]) {
    for (int k = 0; k < 64; k++) out[64 * i + k] = 0;
    return;
  }
  fq_to_be(out + 64 * i, H[i].x);
  fq_to_be(out + 64 * i + 32, H[i].y);
}

__global__ void __launch_bounds__(BN_BLOCK) k_sign(const g1aff* __restrict__ H, const uint8_t* __restrict__ sks, size_t n,
                                                   uint8_t* __restrict__ sigs, const uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (status[i]) {
    for (int k = 0; k < 64; k++) sigs[64 * i + k] = 0;
    return;
  }
  g1aff h = H[i];
  item_sign(sigs + 64 * i, &h, sks + 32 * i);
}

// H == NULL: the first G1 argument is the generator (check_public_keys, /root/reference/src/ecdsa.rs:78-93)
__global__ void __launch_bounds__(BN_BLOCK, BN_MINB) k_verify_miller(const g1aff* __restrict__ H, const uint8_t* __restrict__ sigs,
                                                            const uint8_t* __restrict__ pks, size_t n, fq12* __restrict__ F,
                                                            uint8_t* __restrict__ status, const line_t* __restrict__ lines) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (status[i]) return;  // hash error propagates (/root/reference/src/ecdsa.rs:53); so does a validation error
  g1aff h;
  if (H) {
    h = H[i];
  } else {
    h.x = fq_from_limbs(K_G1_GEN_X);
    h.y = fq_from_limbs(K_G1_GEN_Y);
  }
  fq12 f;
  int st = item_verify_miller(&f, &h, sigs + 64 * i, pks + 128 * i, lines);
  status[i] = (uint8_t)st;
  if (!st) F[i] = f;
}

__global__ void __launch_bounds__(BN_BLOCK, BN_MINB) k_final_exp_check(const fq12* __restrict__ F, size_t n, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (status[i]) return;
  fq12 f = F[i];
  status[i] = item_final_exp_is_one(&f);
}

__global__ void __launch_bounds__(BN_BLOCK) k_miller_pairs(const uint8_t* __restrict__ g1s, const uint8_t* __restrict__ g2s, size_t k, size_t n,
                                                           fq12* __restrict__ F, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fq12 f;
  int st = item_miller_pairs(&f, g1s + 64 * k * i, g2s + 128 * k * i, k);
  status[i] = (uint8_t)st;
  if (!st) F[i] = f;
}

__global__ void __launch_bounds__(BN_BLOCK) k_fq12_to_be(const fq12* __restrict__ F, const uint8_t* __restrict__ status, size_t n,
                                                         uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (status && status[i]) {
    for (int k = 0; k < 384; k++) out[384 * i + k] = 0;
    return;
  }
  fq12 f = F[i];
  fq12_to_be(out + 384 * i, &f);
}

__global__ void __launch_bounds__(BN_BLOCK) k_final_exp_bytes(const uint8_t* __restrict__ in, size_t n, uint8_t* __restrict__ out,
                                                              uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fq12 f, gt;
  int st = ST_OK;
  if (!fq12_from_be(&f, in + 384 * i)) st = ST_NOT_MEMBER;
  else if (!final_exponentiation(&gt, &f)) st = ST_TO_AFFINE;
  status[i] = (uint8_t)st;
  if (st) {
    for (int k = 0; k < 384; k++) out[384 * i + k] = 0;
  } else {
    fq12_to_be(out + 384 * i, &gt);
  }
}

__global__ void __launch_bounds__(BN_BLOCK) k_fq_op(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n,
                                                    uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fq x, y, r = fq_zero();
  int st = ST_OK;
  if (!fq_from_be(&x, a + 32 * i)) st = ST_NOT_MEMBER;
  if (!st && (op <= 2 || op >= 5) && !fq_from_be(&y, b + 32 * i)) st = ST_NOT_MEMBER;
  if (!st) {
    if (op == 0) r = fq_mul(x, y);
    else if (op == 1) r = fq_add(x, y);
    else if (op == 2) r = fq_sub(x, y);
    else if (op == 3) r = fq_inv(x);
    else if (op == 4) { if (!fq_sqrt(&r, x)) st = ST_NOT_MEMBER; }
    else if (op == 5) r = fq_mul_portable(x, y);
    else if (op == 6) r = fq_mul9_add(x, y, &K_KQ_TABLE[0][0]);              // 9 x + y
    else if (op == 7) r = fq_mul9_add(x, fq_q_minus(y), &K_KQ_TABLE[0][0]);  // 9 x - y
    else if (op == 8) r = fq_3t_2z(x, y, &K_KQ_TABLE[0][0]);                 // 3 x + 2 y
    else if (op == 9) r = fq_3t_2z(x, fq_q_minus(y), &K_KQ_TABLE[0][0]);     // 3 x - 2 y
    else st = ST_INVALID_ENCODING;
  }
  status[i] = (uint8_t)st;
  if (st) r = fq_zero();
  fq_to_be(out + 32 * i, r);
}

__global__ void __launch_bounds__(BN_BLOCK) k_fq12_op(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n,
                                                      uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fq12 x, y, r;
  int st = ST_OK;
  if (!fq12_from_be(&x, a + 384 * i)) st = ST_NOT_MEMBER;
  if (!st) {
    if (op == 0) {
      if (!fq12_from_be(&y, b + 384 * i)) st = ST_NOT_MEMBER;
      else fq12_mul(&r, &x, &y);
    } else if (op == 1) fq12_sqr(&r, &x);
    else if (op == 2) fq12_inv(&r, &x);
    else if (op == 3) fq12_cyclotomic_sqr(&r, &x);
    else if (op >= 4 && op <= 6) fq12_frobenius(&r, &x, op - 3);
    else if (op == 7) fq12_conj(&r, &x);
    else st = ST_INVALID_ENCODING;
  }
  status[i] = (uint8_t)st;
  if (st) {
    for (int k = 0; k < 384; k++) out[384 * i + k] = 0;
  } else {
    fq12_to_be(out + 384 * i, &r);
  }
}

// layer hook: n_in Fq values in, n_out Fq values out per item (see debug_layer_op)
__global__ void __launch_bounds__(BN_BLOCK) k_layer_op(int op, const uint8_t* __restrict__ in, int n_in, size_t n, uint8_t* __restrict__ out,
                                                       int n_out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fq a[20], r[12];
  for (int k = 0; k < 12; k++) r[k] = fq_zero();
  for (int k = 0; k < n_in && k < 20; k++) fq_from_be(&a[k], in + 32 * ((size_t)n_in * i + k));
  debug_layer_op(op, a, r);
  for (int k = 0; k < n_out && k < 12; k++) fq_to_be(out + 32 * ((size_t)n_out * i + k), r[k]);
}

// generic per-item map kernel for the light-weight entry points (codecs, key derivation, scalar multiplication)
enum { OP_G1_MUL, OP_G2_MUL, OP_DERIVE_G1, OP_DERIVE_G2, OP_G1_COMPRESS, OP_G1_DECOMPRESS, OP_G2_COMPRESS, OP_G2_DECOMPRESS, OP_G1_VALIDATE, OP_G2_VALIDATE };
template <int OP>
__global__ void __launch_bounds__(BN_BLOCK) k_item_op(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n,
                                                      uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int st = ST_OK;
  if (OP == OP_G1_MUL) st = item_g1_mul(out + 64 * i, a + 64 * i, b + 32 * i);
  else if (OP == OP_G2_MUL) st = item_g2_mul(out + 128 * i, a + 128 * i, b + 32 * i);
  else if (OP == OP_DERIVE_G1) item_derive_pk_g1_comb(out + 64 * i, a + 32 * i, (const aff<fq>*)b);   // b = the context's G1 table
  else if (OP == OP_DERIVE_G2) item_derive_pk_g2_comb(out + 128 * i, a + 32 * i, (const aff<fq2>*)b);  // b = the context's G2 table
  else if (OP == OP_G1_COMPRESS) {
    st = item_g1_compress(out + 33 * i, a + 64 * i);
    if (st) for (int k = 0; k < 33; k++) out[33 * i + k] = 0;
  } else if (OP == OP_G1_DECOMPRESS) st = item_g1_decompress(out + 64 * i, a + 33 * i);
  else if (OP == OP_G2_COMPRESS) {
    st = item_g2_compress(out + 65 * i, a + 128 * i);
    if (st) for (int k = 0; k < 65; k++) out[65 * i + k] = 0;
  } else if (OP == OP_G2_DECOMPRESS) st = item_g2_decompress(out + 128 * i, a + 65 * i);
  else if (OP == OP_G1_VALIDATE) st = item_g1_validate(a + 64 * i);
  else if (OP == OP_G2_VALIDATE) st = item_g2_validate(a + 128 * i);
  if (status) status[i] = (uint8_t)st;
}

// ---- cooperative pairing path (coop.cuh): per-item line sets, then six warps per 32 items run the table-driven program
// H == NULL: the first G1 argument is the generator (check_public_keys)
#ifndef BN_LINES_MINB
#define BN_LINES_MINB 4
#endif
__global__ void __launch_bounds__(BN_BLOCK, BN_LINES_MINB) k_verify_lines(const g1aff* __restrict__ H, const uint8_t* __restrict__ sigs,
                                                           const uint8_t* __restrict__ pks, size_t n, u4* __restrict__ lines, size_t n_pad,
                                                           uint8_t* __restrict__ status, const line_t* __restrict__ table) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (status[i]) return;  // a hash or validation error propagates; the item's line sets stay unwritten and its verdict is never stored
  g1aff h;
  if (H) {
    h = H[i];
  } else {
    h.x = fq_from_limbs(K_G1_GEN_X);
    h.y = fq_from_limbs(K_G1_GEN_Y);
  }
  __shared__ lines_consts consts[BN_BLOCK];
  status[i] = (uint8_t)item_verify_lines(lines, n_pad, i, &h, sigs + 64 * i, pks + 128 * i, table, &consts[threadIdx.x]);
}

#ifndef BN_LINES_LAT_POLICY
#define BN_LINES_LAT_POLICY lines_mul_ilp
#endif
// the same producer in its latency form (coop_lines.cuh lines_mul_ilp), one warp per block so that a small batch spreads over the SMs
__global__ void __launch_bounds__(32) k_verify_lines_lat(const g1aff* __restrict__ H, const uint8_t* __restrict__ sigs, const uint8_t* __restrict__ pks,
                                                         size_t n, u4* __restrict__ lines, size_t n_pad, uint8_t* __restrict__ status,
                                                         const line_t* __restrict__ table, unsigned* __restrict__ progress) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (status[i]) {
    if (progress) *((volatile unsigned*)progress + i) = 0xffffffffu;  // nothing will be written: do not keep the machine waiting
    return;
  }
  g1aff h;
  if (H) {
    h = H[i];
  } else {
    h.x = fq_from_limbs(K_G1_GEN_X);
    h.y = fq_from_limbs(K_G1_GEN_Y);
  }
  __shared__ lines_consts consts[32];
  const int st = item_verify_lines_t<BN_LINES_LAT_POLICY>(lines, n_pad, i, &h, sigs + 64 * i, pks + 128 * i, table, &consts[threadIdx.x],
                                                          progress ? progress + i : nullptr);
  // (pipelined: the machine may already be past its own status check, so the status is final before the release below;
  // a zero is never written: the machine may have reported a fault for this item in the meantime)
  if (st) status[i] = (uint8_t)st;
  if (progress) {
    __threadfence();
    *((volatile unsigned*)progress + i) = 0xffffffffu;
  }
}

// Fq2 routines of the cooperative walk: the compact by-value form.  (The form with three products in flight, which a LONE warp
// needs, buys nothing here -- four warps already interleave, and behind the machine the walk is not the critical path: 2.54 ms per
// one-item verify either way -- and as a producer of mid-size batches it is 25 % slower: 27.4 vs 21.8 ms per 2^18 items.)
#ifndef BN_WALK_MUL
#define BN_WALK_MUL lines_mul_call
#endif
// the producer of small batches: four warps walk the 32 items of a group together (coop_lines.cuh "cooperative walk"), one block per group
__global__ void __launch_bounds__(WALK_WARPS * 32, 4) k_verify_lines_walk4(const g1aff* __restrict__ H, const uint8_t* __restrict__ sigs,
                                                                        const uint8_t* __restrict__ pks, size_t n, u4* __restrict__ lines, size_t n_pad,
                                                                        uint8_t* __restrict__ status, const line_t* __restrict__ table,
                                                                        unsigned* __restrict__ progress, size_t mute_item) {
  extern __shared__ u4 walk_sm[];
  walk_ctx c;
  c.lane = threadIdx.x & 31;
  // role = warp index.  (Rotating the roles with the block index, to even out the sub-partitions when several blocks share an SM,
  // was measured: no gain as a throughput producer, and 8 % slower next to a twelve-warp machine block -- r02 tuning log.)
  c.warp = threadIdx.x >> 5;
  c.sm = walk_sm + c.lane;
  c.flags = (int*)(walk_sm + WS_SLOTS * 2 * 2 * COOP_LANES);
  c.row = COOP_LANES;
  c.item = (size_t)blockIdx.x * COOP_LANES + c.lane;
  c.n = n;
  c.n_pad = n_pad;
  c.lines = lines;
  c.table = table;
  const bool skip = c.item >= n || status[c.item] != 0;  // (an item that already failed: nothing is read, nothing is written)
  g1aff h;
  if (!skip && H) {
    h = H[c.item];
  } else {
    h.x = fq_from_limbs(K_G1_GEN_X);
    h.y = fq_from_limbs(K_G1_GEN_Y);
  }
  walk_decode(c, &h, sigs + 64 * c.item, pks + 128 * c.item, skip);
  __syncthreads();
  const int st = walk_flags(c, skip);
  if (c.item == mute_item) progress = nullptr;  // test hook (bn254_set_test_fault): this item's progress is never published
  const bool reporter = c.warp == WALK_WARPS - 1 && c.item < n;  // the warp with the shortest first level reports for the group
  if (reporter && !c.live) {  // no line set will be written: the status is final, do not keep a pipelined machine waiting
    if (!skip) status[c.item] = (uint8_t)st;
    if (progress) {
      __threadfence();
      *((volatile unsigned*)progress + c.item) = 0xffffffffu;
    }
  }
  walk_schedule(
      [&](int kind, int level, size_t m, int sqx, int sqy) {
        if (kind == 0)
          walk_dbl<BN_WALK_MUL>(c, level, m);
        else
          walk_add<BN_WALK_MUL>(c, level, m, sqx, sqy);
      },
      [] { __syncthreads(); },
      [&](size_t m) {  // every store of steps < m is before the barrier this thread has just left: fence, then publish
        if (progress && reporter && c.live) {
          __threadfence();
          *((volatile unsigned*)progress + c.item) = m < K_N_LINES ? (unsigned)m : 0xffffffffu;
        }
      });
}

// Untrusted-input policy (the default, bn254_set_input_policy): sig / pk bytes are decoded exactly as
// Signature::from_uncompressed / PublicKey::from_uncompressed would decode them (/root/reference/src/utils.rs:107-127):
// field membership, curve equation -- which (0, 0), the engine's encoding of infinity, fails -- and for G2 the r-torsion
// test of AffineG2::new.  pk is decoded first, then sig; a decode error replaces whatever the hash left in status.
__global__ void __launch_bounds__(BN_BLOCK) k_validate_inputs(const uint8_t* __restrict__ sigs, const uint8_t* __restrict__ pks, size_t n,
                                                              uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int st = item_g2_validate(pks + 128 * i);
  if (!st) st = item_g1_validate(sigs + 64 * i);
  if (st) status[i] = (uint8_t)st;
}

#if defined(COOP_ABLATE_NOBAR)  // timing experiment only: results are wrong without the barriers
#define COOP_BARRIER() __syncwarp()
#else
#define COOP_BARRIER() __syncthreads()
#endif
#ifndef BN_COOP_MINB
#define BN_COOP_MINB 4
#endif
#ifndef BN_COOP_STAGGER
#define BN_COOP_STAGGER 0
#endif
#ifndef BN_COOP_CHUNK_LOG2
#define BN_COOP_CHUNK_LOG2 19
#endif
#ifndef BN_COOP_DEFAULT_GROUPS4
#define BN_COOP_DEFAULT_GROUPS4 1
#endif
#ifndef BN_COOP_DEFAULT_H
#define BN_COOP_DEFAULT_H false
#endif
#ifndef BN_COOP_DEFAULT_W
#define BN_COOP_DEFAULT_W false
#endif
// which: 0 verify (Miller of 2 line streams + final exponentiation + verdict), 1 / 2 Miller of 1 / 2 streams -> fio,
// 3 final exponentiation of fio (+ verdict), 4 multi-pairing: COOP_MULTI_K pairs per lane, block product -> fio
__global__ void __launch_bounds__(COOP_THREADS, BN_COOP_MINB) k_coop_run(int which, size_t n, size_t n_pad, const u4* __restrict__ lines,
                                                                         u4* __restrict__ gslots, u4* __restrict__ fio,
                                                                         uint8_t* __restrict__ status, unsigned stagger, unsigned sms,
                                                                         const unsigned* progress) {
  extern __shared__ u4 coop_sm[];
  // Blocks that share an SM run the same program at the same speed: started together they stay in lockstep, so their
  // multiply-free stretches (commit, barriers) coincide and the multiplier pipe idles for all of them at once.  The k-th
  // co-resident block of the first wave therefore starts k * stagger cycles late (blocks of later waves inherit the offsets).
  if (stagger) {
    const unsigned slot = (blockIdx.x / sms) % BN_COOP_MINB;
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)slot * stagger) {
    }
  }
  uint32_t* kq = (uint32_t*)(coop_sm + COOP_SLOTS * 2 * COOP_LANES);  // the k q table of fq_mul9_add, after the group's slots
  if (threadIdx.x < 88) kq[threadIdx.x] = (&K_KQ_TABLE[0][0])[threadIdx.x];
  __syncthreads();
  coop_ctx c;
  c.k = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.sm = coop_sm + c.lane;
  c.row = COOP_LANES;
  c.wmode = false;
  c.plans = K_COOP_PLANS;
  c.kq = kq;
  c.item = (size_t)blockIdx.x * COOP_LANES + c.lane;
  c.active = c.item < n;
  c.n_pad = n_pad;
  c.lines = lines;
  c.gslots = gslots;
  c.fio = fio;
  c.status = status;
  c.progress = nullptr;
  c.sets_per_step = 1;
  c.progress = progress;
  c.sets_per_step = 2;  // (only the verify program is ever run pipelined)
  coop_run_block(c, coop_program(which), [] { COOP_BARRIER(); });
}

// ---- latency layout (coop.cuh coop_run_block12): one 32-item group per TWELVE-warp block, for launches of at most one group per SM
#define COOP12_THREADS (2 * COOP_THREADS)
// group slots + twelve exchange slots + the k q table = 67 936 bytes, REQUESTED as 68 KB (anything above 64 KB would do): a block of this kernel fills an SM's register
// file to 3/4, so the driver would pick the smallest shared-memory carve-out that holds one block (64 KB) and leave 1 KB over --
// and a pipelined producer block (10 KB) dispatched behind the machine could then never become resident next to it while the
// machine waits for its line sets.  More than 64 KB forces the 100 KB carve-out (or larger) on every configuration the SM has.
#define COOP12_SMEM_USED (COOP_SMEM_BYTES + 12 * 2 * COOP_LANES * 16 + 11 * 32) /* (twelve exchange slots: the eighteen-warp form uses all of them) */
#define COOP12_SMEM_BYTES (68 * 1024)
static_assert(COOP12_SMEM_USED <= COOP12_SMEM_BYTES, "latency layout: shared memory");
template <int SPLIT>
__device__ __forceinline__ void coop12_body(int which, size_t n, size_t n_pad, const u4* __restrict__ lines, u4* __restrict__ gslots, u4* __restrict__ fio,
                                            uint8_t* __restrict__ status, const unsigned* progress) {
  extern __shared__ u4 coop_sm[];
  u4* xch_base = coop_sm + COOP_SLOTS * 2 * COOP_LANES;
  uint32_t* kq = (uint32_t*)(xch_base + 12 * 2 * COOP_LANES);
  if (threadIdx.x < 88) kq[threadIdx.x] = (&K_KQ_TABLE[0][0])[threadIdx.x];
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  coop_ctx c;
  c.k = warp % COOP_WARPS;
  c.lane = threadIdx.x & 31;
  c.sm = coop_sm + c.lane;
  c.row = COOP_LANES;
  c.wmode = false;
  c.plans = K_COOP_PLANS;
  c.kq = kq;
  c.item = (size_t)blockIdx.x * COOP_LANES + c.lane;
  c.active = c.item < n;
  c.n_pad = n_pad;
  c.lines = lines;
  c.gslots = gslots;
  c.fio = fio;
  c.status = status;
  c.progress = progress;
  c.sets_per_step = 2;
  coop_run_block12<SPLIT>(c, warp / COOP_WARPS, xch_base + c.lane, coop_program(which), [] { __syncthreads(); });
}
__global__ void __launch_bounds__(COOP12_THREADS, 1) k_coop12_run(int which, size_t n, size_t n_pad, const u4* __restrict__ lines,
                                                                  u4* __restrict__ gslots, u4* __restrict__ fio, uint8_t* __restrict__ status,
                                                                  const unsigned* progress) {
  coop12_body<2>(which, n, n_pad, lines, gslots, fio, status, progress);
}
// eighteen warps: one per (coefficient, Karatsuba component)
__global__ void __launch_bounds__(3 * COOP_THREADS, 1) k_coop18_run(int which, size_t n, size_t n_pad, const u4* __restrict__ lines,
                                                                    u4* __restrict__ gslots, u4* __restrict__ fio, uint8_t* __restrict__ status,
                                                                    const unsigned* progress) {
  coop12_body<3>(which, n, n_pad, lines, gslots, fio, status, progress);
}

// ---- the same block-layout machine with FOUR 32-item groups in one 24-warp block (one block per SM).  Warp w of a block
// runs on sub-partition w % 4, so making group g = warps {g, g + 4, ..., g + 20} puts the six warps of a group on ONE
// sub-partition: they share one multiplier pipe fairly, reach the group barrier together, and never wait for a sibling that
// is queued behind other groups' warps on a busier sub-partition (in k_coop_run a block's warps are spread over all four
// sub-partitions and 36 % of the warp-time is spent at the block barrier).  Barriers are per group (named barrier g + 1).
#define COOP1_SMEM_BYTES (COOP_SMEM_BYTES + 11 * 32) /* + the k q table of fq_mul9_add */
#define COOP4_GROUPS 4
#define COOP4_THREADS (COOP4_GROUPS * COOP_THREADS)
#define COOP4_SMEM_BYTES (COOP4_GROUPS * COOP_SMEM_BYTES + 11 * 32) /* + the k q table of fq_mul9_add */
__global__ void __launch_bounds__(COOP4_THREADS, 1) k_coop4_run(int which, size_t n, size_t n_pad, const u4* __restrict__ lines,
                                                                u4* __restrict__ gslots, u4* __restrict__ fio,
                                                                uint8_t* __restrict__ status, unsigned stagger_ns) {
  extern __shared__ u4 coop_sm[];
  // (measured alternative: two groups sharing two sub-partitions, three warps of each on either, so that one group's commit
  // phase is covered by the other's accumulation -- no faster than k_coop_run: the cross-sub-partition waiting is back)
  const int warp = threadIdx.x >> 5, g = warp & (COOP4_GROUPS - 1);
  uint32_t* kq = (uint32_t*)(coop_sm + COOP4_GROUPS * (COOP_SLOTS * 2 * COOP_LANES));
  if (threadIdx.x < 88) kq[threadIdx.x] = (&K_KQ_TABLE[0][0])[threadIdx.x];
  __syncthreads();
  coop_ctx c;
  c.k = warp / COOP4_GROUPS;
  c.lane = threadIdx.x & 31;
  c.sm = coop_sm + g * (COOP_SLOTS * 2 * COOP_LANES) + c.lane;
  c.row = COOP_LANES;
  c.wmode = false;
  c.plans = K_COOP_PLANS;
  c.kq = kq;
  (void)stagger_ns;  // (round-1 tuning knob: start offsets between the warps of a group -- measured slower, removed from the loop)
  c.item = ((size_t)blockIdx.x * COOP4_GROUPS + g) * COOP_LANES + c.lane;
  c.active = c.item < n;
  c.n_pad = n_pad;
  c.lines = lines;
  c.gslots = gslots;
  c.fio = fio;
  c.status = status;
  c.progress = nullptr;
  c.sets_per_step = 1;
  const int bar = g + 1;
  coop_run_block(c, coop_program(which), [bar] { asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(COOP_THREADS) : "memory"); });
}

// ---- half-warp layout: a group is 16 items and THREE warps, each warp carrying two coefficients (lanes 0-15 one, lanes
// 16-31 the other), so the 24 warps of a block form EIGHT groups, two per sub-partition.  Same shared memory per item and
// the same warp count as k_coop4_run, but while one group of a sub-partition is in its multiply-free stretch (recombination,
// commit of the records, barriers) the other one can be accumulating: in k_coop4_run the six warps of a sub-partition move in
// step, so that stretch -- a quarter of the kernel's time -- leaves the multiplier pipe idle.  Coefficient pairs (0,2), (1,3),
// (4,5): squares and cyclotomic squares need one product more for even coefficients, so only the third warp pays for a mixed pair.
#define COOPH_GROUPS 8
#define COOPH_GROUP_THREADS 96
#define COOPH_GROUP_U4 (COOP_SLOTS * 2 * COOPH_ROW)
#define COOPH_SMEM_BYTES (COOPH_GROUPS * COOPH_GROUP_U4 * 16)
__global__ void __launch_bounds__(COOP4_THREADS, 1) k_cooph_run(int which, size_t n, size_t n_pad, const u4* __restrict__ lines,
                                                                u4* __restrict__ gslots, u4* __restrict__ fio,
                                                                uint8_t* __restrict__ status, unsigned offset_cycles) {
  extern __shared__ u4 coop_sm[];
  const int warp = threadIdx.x >> 5, sp = warp & 3, j = warp >> 2;  // sub-partition of this warp and its slot there
  const int g = sp + 4 * (j / 3), p = j % 3, L = threadIdx.x & 31;
  coop_ctx c;
  c.k = p == 0 ? (L < 16 ? 0 : 2) : p == 1 ? (L < 16 ? 1 : 3) : (L < 16 ? 4 : 5);
  c.lane = L & 15;
  c.sm = coop_sm + g * COOPH_GROUP_U4 + c.lane;
  c.row = COOPH_ROW;
  c.wmode = false;
  c.plans = K_COOP_PLANS_H;
  c.kq = nullptr;
  c.item = ((size_t)blockIdx.x * COOPH_GROUPS + g) * COOPH_ROW + c.lane;
  c.active = c.item < n;
  c.n_pad = n_pad;
  c.lines = lines;
  c.gslots = gslots;
  c.fio = fio;
  c.status = status;
  c.progress = nullptr;
  c.sets_per_step = 1;
  const uint32_t* prog = coop_program(which);
  // the second group of every sub-partition starts late, so that the two do not run their phases in step
  if (offset_cycles && g >= 4) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)offset_cycles) {
    }
  }
  int line_next = 0;
  const int bar = g + 1;
#pragma unroll 1
  for (int pc = 0;; pc++) {
    const uint32_t ins = prog[pc];
    if ((ins & 0xff) == COP_END) break;
    fq2 t = coop_phase_a(c, ins, line_next);
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(COOPH_GROUP_THREADS) : "memory");
    coop_phase_b(c, ins, t, line_next);
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(COOPH_GROUP_THREADS) : "memory");
  }
}

// ---- warp-local form of the same machine: the six coefficients of an item live in six lanes of ONE warp (lane = 5 k + j,
// five items per warp, lanes 30 and 31 idle), so the two synchronisation points of an instruction are __syncwarp() and
// no warp ever waits for a sibling that shares its sub-partition with other blocks.  The only block-level step is the
// single Fq inversion of the final exponentiation: the block's 30 items are gathered into one warp between two block
// barriers.  Shared memory: 54 slots x 160 B per warp, plus the plan table (lane-varying index, so not constant memory).
#define COOPW_WARPS 6
#define COOPW_ITEMS (COOPW_WARPS * COOPW_ROW)
#define COOPW_WARP_U4 (COOP_SLOTS * 2 * COOPW_ROW)
#define COOPW_PLAN_WORDS (CPLAN_COUNT * 6 * 7)
#define COOPW_SMEM_BYTES (COOPW_WARPS * COOPW_WARP_U4 * 16 + COOPW_PLAN_WORDS * 4)
__global__ void __launch_bounds__(COOPW_WARPS * 32, BN_COOP_MINB) k_coopw_run(int which, size_t n, size_t n_pad, const u4* __restrict__ lines,
                                                                              u4* __restrict__ gslots, u4* __restrict__ fio,
                                                                              uint8_t* __restrict__ status) {
  extern __shared__ u4 coop_sm[];
  uint32_t* plans = (uint32_t*)(coop_sm + COOPW_WARPS * COOPW_WARP_U4);
  for (int i = threadIdx.x; i < COOPW_PLAN_WORDS; i += blockDim.x) plans[i] = (&K_COOP_PLANS_W[0][0][0])[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, L = threadIdx.x & 31;
  const bool live = L < 6 * COOPW_ROW;
  coop_ctx c;
  c.k = live ? L / COOPW_ROW : 0;
  c.lane = live ? L % COOPW_ROW : 0;
  c.sm = coop_sm + warp * COOPW_WARP_U4 + c.lane;
  c.row = COOPW_ROW;
  c.wmode = true;
  c.plans = (const uint32_t (*)[6][7])plans;
  c.kq = nullptr;
  c.item = ((size_t)blockIdx.x * COOPW_WARPS + warp) * COOPW_ROW + c.lane;
  c.active = live && c.item < n;
  c.n_pad = n_pad;
  c.lines = lines;
  c.gslots = gslots;
  c.fio = fio;
  c.status = status;
  c.progress = nullptr;
  c.sets_per_step = 1;
  const uint32_t* prog = coop_program(which);
  int line_next = 0;
#pragma unroll 1
  for (int pc = 0;; pc++) {
    const uint32_t ins = prog[pc];
    const int op = ins & 0xff;
    if (op == COP_END) break;
    if (op == COP_INVT) {
      __syncthreads();
      if (warp == 0 && L < COOPW_ITEMS) coop_invt(coop_sm + (L / COOPW_ROW) * COOPW_WARP_U4 + (L % COOPW_ROW), COOPW_ROW);
      __syncthreads();
      continue;
    }
    fq2 t;
    if (live) t = coop_phase_a(c, ins, line_next);
    __syncwarp();
    if (live) coop_phase_b(c, ins, t, line_next);
    __syncwarp();
  }
}

// ---- multi-pairing through the cooperative machine: pair p is stream p / L of lane p % L (L lanes, COOP_MULTI_K streams each)
__device__ __forceinline__ void record_error(unsigned long long* err, size_t i, int st);
// mk = pairs per lane of the consuming program.  extra_sig != NULL: pair number n is (that G1 point, -G2) -- the rank's own
// share of the second pairing of the aggregate check, folded into its Miller product (bilinearity: the product over the
// ranks of e(S_r, -G2) is e(sum S_r, -G2)); an all-zero extra_sig (infinity) is skipped like any pair holding an infinity.
__global__ void __launch_bounds__(BN_BLOCK, BN_LINES_MINB) k_pair_lines(const g1aff* __restrict__ H, const uint8_t* __restrict__ hstatus,
                                                                        const uint8_t* __restrict__ pks, size_t n, size_t L, int mk,
                                                                        u4* __restrict__ lines, unsigned long long* __restrict__ err,
                                                                        size_t index_base, const uint8_t* __restrict__ extra_sig,
                                                                        int typed) {
  __shared__ lines_consts consts[BN_BLOCK];
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L * (size_t)mk) return;
  bool use = false;
  g1aff h;
  g2j q;
  q.x = fq2_one();
  q.y = fq2_one();
  if (extra_sig && p == n) {
    g1j s;
    int st = g1_from_raw(&s, extra_sig);
    if (st) record_error(err, index_base + p, st);
    else if (!pt_is_inf(&s)) {
      use = true;
      h.x = s.x;
      h.y = s.y;
      q.x = fq2_from_limbs(K_G2_GEN_X);
      q.y = fq2_neg(fq2_from_limbs(K_G2_GEN_Y));
    }
  } else if (p < n) {
    if (hstatus[p]) {
      record_error(err, index_base + p, hstatus[p]);
    } else {
      // untrusted keys are decoded like PublicKey::from_uncompressed (r-torsion test included, no infinity)
      int st = typed ? g2_from_raw(&q, pks + 128 * p) : item_g2_validate(pks + 128 * p);
      if (!st && !typed) st = g2_from_raw(&q, pks + 128 * p);
      if (st) record_error(err, index_base + p, st);
      else if (!pt_is_inf(&q)) {
        use = true;
        h = H[p];
      }
    }
  }
  item_pair_lines(lines, L, p % L, (int)(p / L), mk, use, &h, q.x, q.y, &consts[threadIdx.x]);
}
// ---- bn::pairing_batch(k pairs) == one through the cooperative machine (k = 1, 2): decode statuses first (first failing
// pair in order, G1 before G2, as a left-to-right decode would report), then one thread per (item, pair) writes the pair's
// line stream, then the machine runs Miller loop + final exponentiation + verdict per item
__global__ void __launch_bounds__(BN_BLOCK) k_pairs_decode(const uint8_t* __restrict__ g1s, const uint8_t* __restrict__ g2s, size_t k, size_t n,
                                                           uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int st = ST_OK;
  for (size_t j = 0; j < k && !st; j++) {
    g1j p;
    g2j q;
    st = g1_from_raw(&p, g1s + 64 * (k * i + j));
    if (!st) st = g2_from_raw(&q, g2s + 128 * (k * i + j));
  }
  status[i] = (uint8_t)st;
}
__global__ void __launch_bounds__(BN_BLOCK, BN_LINES_MINB) k_item_pair_lines(const uint8_t* __restrict__ g1s, const uint8_t* __restrict__ g2s, int k,
                                                                             size_t n, size_t n_pad, u4* __restrict__ lines,
                                                                             const uint8_t* __restrict__ status) {
  __shared__ lines_consts consts[BN_BLOCK];
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (size_t)k) return;
  const size_t i = t / k;
  const int s = (int)(t % k);
  if (status[i]) return;
  g1j p;
  g2j q;
  g1_from_raw(&p, g1s + 64 * t);
  g2_from_raw(&q, g2s + 128 * t);
  const bool use = !pt_is_inf(&p) && !pt_is_inf(&q);
  g1aff h;
  h.x = p.x;
  h.y = p.y;
  if (!use) {
    q.x = fq2_one();
    q.y = fq2_one();
  }
  item_pair_lines(lines, n_pad, i, s, k, use, &h, q.x, q.y, &consts[threadIdx.x]);
}
// ---- cached key lines (a fixed validator set): the walk of every key once, then per verify only the scaling
// kstatus[j] = decode status of key j under the context's input policy (untrusted: from_uncompressed semantics, r-torsion test)
__global__ void __launch_bounds__(BN_BLOCK, BN_LINES_MINB) k_key_lines(const uint8_t* __restrict__ pks, size_t n_keys, size_t k_pad, int typed,
                                                                       u4* __restrict__ klines, uint8_t* __restrict__ kstatus) {
  __shared__ lines_consts consts[BN_BLOCK];
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_keys) return;
  g2j q;
  int st = typed ? ST_OK : item_g2_validate(pks + 128 * j);
  if (!st) st = g2_from_raw(&q, pks + 128 * j);
  kstatus[j] = (uint8_t)st;
  const bool use = !st && !pt_is_inf(&q);
  if (!use) {
    q.x = fq2_one();
    q.y = fq2_one();
  }
  item_key_lines(klines, k_pad, j, use, q.x, q.y, &consts[threadIdx.x]);
}
// per item: key index / key status / signature decode -> status (precedence of verify_batch: under the untrusted policy a decode
// error replaces a hash error, under the typed policy the hash error stays) and the signature in Montgomery form ((0, 0) = infinity)
__global__ void __launch_bounds__(BN_BLOCK) k_cached_decode(const uint8_t* __restrict__ sigs, size_t n, const uint32_t* __restrict__ key_index,
                                                            size_t key_base, size_t n_keys, const uint8_t* __restrict__ kstatus, int typed,
                                                            g1aff* __restrict__ S, uint8_t* __restrict__ status) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t key = key_index ? key_index[i] : key_base + i;  // (identity mapping: the chunk's first item is key `key_base`)
  int st = key >= n_keys ? ST_INDEX_OOB : kstatus[key];
  g1j s;
  pt_set_inf(&s);
  if (!st && !typed) st = item_g1_validate(sigs + 64 * i);
  if (!st) st = g1_from_raw(&s, sigs + 64 * i);
  g1aff o;
  o.x = fq_zero();
  o.y = fq_zero();
  if (!st && !pt_is_inf(&s)) {
    o.x = s.x;
    o.y = s.y;
  }
  S[i] = o;
  if (st && (!typed || !status[i])) status[i] = (uint8_t)st;
}
// one thread per (line m, item): thread index = m * n_pad + item, so that a warp writes 32 neighbouring items of one line set
__global__ void __launch_bounds__(BN_BLOCK) k_scale_cached_lines(const g1aff* __restrict__ H, const g1aff* __restrict__ S, size_t n, size_t n_pad,
                                                                 const u4* __restrict__ klines, size_t k_pad, const uint32_t* __restrict__ key_index,
                                                                 size_t key_base, u4* __restrict__ lines, const uint8_t* __restrict__ status,
                                                                 const line_t* __restrict__ table) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t item = t % n_pad;
  const int m = (int)(t / n_pad);
  if (m >= K_N_LINES || item >= n || status[item]) return;
  const size_t key = key_index ? key_index[item] : key_base + item;
  const g1aff s = S[item];
  item_scale_cached_lines(lines, n_pad, item, m, klines, k_pad, key, H[item], !(fq_is_zero(s.x) && fq_is_zero(s.y)), s.x, s.y, table);
}

// lane 0 of every block holds the block's product (power-basis layout fio) -> tower-order Fq12 array
__global__ void k_coop_gather(const u4* __restrict__ fio, size_t L, size_t blocks, fq12* __restrict__ out) {
  size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= blocks) return;
  const int pos[6] = {0, 2, 4, 1, 3, 5};
  fq12 f;
  fq2* c = &f.c0.c0;
  for (int t = 0; t < 6; t++) {
    c[t].c0 = coop_gld(fio, (size_t)pos[t] * 2 + 0, L, b * COOP_LANES);
    c[t].c1 = coop_gld(fio, (size_t)pos[t] * 2 + 1, L, b * COOP_LANES);
  }
  out[b] = f;
}

// ---- randomised batch verification: per item c H(m) (in place) and c sig ; any item that cannot ride sets *bad
__global__ void __launch_bounds__(BN_BLOCK) k_rlc_prepare(g1aff* __restrict__ H, const uint8_t* __restrict__ hstatus,
                                                          const uint8_t* __restrict__ sigs, const uint8_t* __restrict__ pks,
                                                          const uint8_t* __restrict__ coeffs16, size_t n, int check_g2,
                                                          uint8_t* __restrict__ sig_c, unsigned* __restrict__ bad, size_t L, size_t W) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int st = hstatus[i];
  if (!st) {
    g1aff h = H[i], hs;
    st = item_rlc_prepare(&hs, sig_c + 64 * i, &h, sigs + 64 * i, pks + 128 * i, coeffs16 + 16 * i, check_g2 != 0);
    if (!st) H[i] = hs;
  }
  if (st) {
    for (int k = 0; k < 64; k++) sig_c[64 * i + k] = 0;  // nothing of this item enters the slice's signature sum
    atomicOr(bad + (i % L) / W, 1u);                     // its slice goes to the exact path
  }
}
// The multi-pairing machine puts pair p on lane p % L (L lanes, stream p / L): slice s = lanes [s W, (s + 1) W) owns the items
// { t L + s W + j : t < COOP_MULTI_K, j < W }.  One block per slice: the sum of its scaled signatures, affine.
__global__ void __launch_bounds__(BN_BLOCK) k_rlc_slice_sums(const uint8_t* __restrict__ sig_c, size_t n, size_t L, size_t W,
                                                             const unsigned* __restrict__ bad, uint8_t* __restrict__ agg) {
  __shared__ jac<fq> sh[BN_BLOCK];
  const size_t s = blockIdx.x;
  jac<fq> acc;
  pt_set_inf(&acc);
  if (!bad[s]) {
    for (size_t idx = threadIdx.x; idx < (size_t)COOP_MULTI_K * W; idx += BN_BLOCK) {
      size_t lane = s * W + idx % W, p = (idx / W) * L + lane;
      if (lane >= L || p >= n) continue;
      jac<fq> q;
      if (g1_from_raw(&q, sig_c + 64 * p) || pt_is_inf(&q)) continue;
      pt_madd(&acc, &acc, &q.x, &q.y);
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int t = BN_BLOCK / 2; t > 0; t >>= 1) {
    if (threadIdx.x < t) pt_add(&sh[threadIdx.x], &sh[threadIdx.x], &sh[threadIdx.x + t]);
    __syncthreads();
  }
  if (threadIdx.x == 0) g1_to_raw(agg + 64 * s, &sh[0]);
}
// verdict of slice s: prod of its groups' Miller partials * miller(sum of its scaled signatures, -G2), final exponentiation
__global__ void __launch_bounds__(32) k_rlc_finish_slices(const fq12* __restrict__ partial, size_t groups, size_t gps,
                                                          const uint8_t* __restrict__ agg, const unsigned* __restrict__ bad,
                                                          const line_t* __restrict__ lines, uint8_t* __restrict__ verdict) {
  if (threadIdx.x != 0) return;
  const size_t s = blockIdx.x;
  if (bad[s]) {
    verdict[s] = ST_VERIFICATION_FAILED;
    return;
  }
  fq12 acc, t;
  fq12_set_one(&acc);
  for (size_t g = s * gps; g < (s + 1) * gps && g < groups; g++) {
    t = partial[g];
    fq12_mul(&acc, &acc, &t);
  }
  g1j a;
  if (g1_from_raw(&a, agg + 64 * s)) {
    verdict[s] = ST_VERIFICATION_FAILED;
    return;
  }
  if (!pt_is_inf(&a)) {
    fq2 dummy = fq2_one();
    miller_loop_2(&t, false, &a.x, &a.y, &dummy, &dummy, true, &a.x, &a.y, lines);
    fq12_mul(&acc, &acc, &t);
  }
  verdict[s] = (uint8_t)item_final_exp_is_one(&acc);
}

// ---- point aggregation: strided mixed additions per thread, then a shared-memory tree per block
template <class F> struct pt_io;
template <> struct pt_io<fq> {
  static const int BYTES = 64;
  static __device__ int load(jac<fq>* p, const uint8_t* b) { return g1_from_raw(p, b); }
  static __device__ int validate(const uint8_t* b) { return item_g1_validate(b); }
  static __device__ void store(uint8_t* b, const jac<fq>* p) { g1_to_raw(b, p); }
};
template <> struct pt_io<fq2> {
  static const int BYTES = 128;
  static __device__ int load(jac<fq2>* p, const uint8_t* b) { return g2_from_raw(p, b); }
  static __device__ int validate(const uint8_t* b) { return item_g2_validate(b); }
  static __device__ void store(uint8_t* b, const jac<fq2>* p) { g2_to_raw(b, p); }
};
__device__ __forceinline__ void record_error(unsigned long long* err, size_t i, int st) {
  atomicMin(err, ((unsigned long long)i << 8) | (unsigned long long)st);
}

template <class F>
__global__ void __launch_bounds__(BN_BLOCK) k_sum_partial(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ neg, size_t n,
                                                          jac<F>* __restrict__ partial, unsigned long long* __restrict__ err, int strict) {
  __shared__ jac<F> sh[BN_BLOCK];
  const int B = pt_io<F>::BYTES;
  jac<F> acc;
  pt_set_inf(&acc);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    jac<F> p;
    // strict: bytes from outside are decoded like from_uncompressed (no infinity, G2 in the r-torsion)
    int st = strict ? pt_io<F>::validate(pts + (size_t)B * i) : ST_OK;
    if (!st) st = pt_io<F>::load(&p, pts + (size_t)B * i);
    if (st) {
      record_error(err, i, st);
      continue;
    }
    if (pt_is_inf(&p)) continue;
    if (neg && neg[i]) p.y = fe_neg(p.y);
    pt_madd(&acc, &acc, &p.x, &p.y);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = BN_BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) pt_add(&sh[threadIdx.x], &sh[threadIdx.x], &sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
template <class F>
__global__ void __launch_bounds__(BN_BLOCK) k_sum_final(const jac<F>* __restrict__ partial, int m, uint8_t* __restrict__ out,
                                                        const unsigned long long* __restrict__ err, uint8_t* __restrict__ status) {
  __shared__ jac<F> sh[BN_BLOCK];
  jac<F> acc;
  pt_set_inf(&acc);
  for (int i = threadIdx.x; i < m; i += BN_BLOCK) pt_add(&acc, &acc, &partial[i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = BN_BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) pt_add(&sh[threadIdx.x], &sh[threadIdx.x], &sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    unsigned long long e = *err;
    if (e != ~0ull) {
      *status = (uint8_t)(e & 0xff);
      for (int k = 0; k < pt_io<F>::BYTES; k++) out[k] = 0;
    } else {
      *status = 0;
      pt_io<F>::store(out, &sh[0]);
    }
  }
}

// ---- distinct-message multi-pairing: every thread folds the Miller values of its strided pairs into one Fq12
__global__ void __launch_bounds__(BN_PROD_BLOCK) k_distinct_partial(const g1aff* __restrict__ H, const uint8_t* __restrict__ hstatus,
                                                                     const uint8_t* __restrict__ pks, size_t n, fq12* __restrict__ partial,
                                                                     unsigned long long* __restrict__ err, int typed) {
  __shared__ fq12 sh[BN_PROD_BLOCK];
  fq12 acc;
  fq12_set_one(&acc);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (hstatus[i]) {
      record_error(err, i, hstatus[i]);
      continue;
    }
    g2j q;
    int st = typed ? ST_OK : item_g2_validate(pks + 128 * i);
    if (!st) st = g2_from_raw(&q, pks + 128 * i);
    if (st) {
      record_error(err, i, st);
      continue;
    }
    if (pt_is_inf(&q)) continue;
    g1aff h = H[i];
    fq12 t;
    miller_loop_2(&t, true, &h.x, &h.y, &q.x, &q.y, false, &h.x, &h.y, (const line_t*)0);
    fq12_mul(&acc, &acc, &t);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = BN_PROD_BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) fq12_mul(&sh[threadIdx.x], &sh[threadIdx.x], &sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
// one-thread-per-item mode: the Miller value of the extra pair (sig, -G2) of a rank's partial
__global__ void k_extra_pair(const uint8_t* __restrict__ sig, const line_t* __restrict__ lines, fq12* __restrict__ out,
                             unsigned long long* __restrict__ err, size_t index) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  fq12 t;
  fq12_set_one(&t);
  g1j s;
  int st = g1_from_raw(&s, sig);
  if (st) record_error(err, index, st);
  else if (!pt_is_inf(&s)) {
    fq2 dummy = fq2_one();
    miller_loop_2(&t, false, &s.x, &s.y, &dummy, &dummy, true, &s.x, &s.y, lines);
  }
  *out = t;
}
// product of m Fq12 partials -> one Fq12 (Montgomery form) and its 384-byte big-endian image
__global__ void __launch_bounds__(BN_PROD_BLOCK) k_fq12_prod_final(const fq12* __restrict__ partial, int m, fq12* __restrict__ out,
                                                                    uint8_t* __restrict__ out_be, const unsigned long long* __restrict__ err,
                                                                    uint8_t* __restrict__ status) {
  __shared__ fq12 sh[BN_PROD_BLOCK];
  fq12 acc;
  fq12_set_one(&acc);
  for (int i = threadIdx.x; i < m; i += BN_PROD_BLOCK) {
    fq12 t = partial[i];
    fq12_mul(&acc, &acc, &t);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = BN_PROD_BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) fq12_mul(&sh[threadIdx.x], &sh[threadIdx.x], &sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (out) *out = sh[0];
    if (out_be) fq12_to_be(out_be, &sh[0]);
    if (status) {
      unsigned long long e = *err;
      *status = (e == ~0ull) ? 0 : (uint8_t)(e & 0xff);  // first failing item by index
    }
  }
}
// one level of the product tree: block b folds partial[b * per .. (b + 1) * per) into out[b]
__global__ void __launch_bounds__(BN_PROD_BLOCK) k_fq12_prod_level(const fq12* __restrict__ partial, size_t m, size_t per, fq12* __restrict__ out) {
  __shared__ fq12 sh[BN_PROD_BLOCK];
  fq12 acc;
  fq12_set_one(&acc);
  const size_t lo = (size_t)blockIdx.x * per, hi = lo + per < m ? lo + per : m;
  for (size_t i = lo + threadIdx.x; i < hi; i += BN_PROD_BLOCK) {
    fq12 t = partial[i];
    fq12_mul(&acc, &acc, &t);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = BN_PROD_BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) fq12_mul(&sh[threadIdx.x], &sh[threadIdx.x], &sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// ---- finish of an aggregate check through the cooperative machine (one item): the exchanged payloads are folded here
// payload r (BN254_DISTINCT_PAYLOAD bytes): [0, 384) Miller partial (big-endian, tower order), [384] its status, [385] 1 when
// the rank's own (sum of signatures, -G2) pair is already inside its partial.  Thread t folds payload t, t + 64, ...; the
// product goes to fio in the machine's layout (item 0 of a 32-item group), the 87 scaled -G2 lines of agg_sig (or the
// constant 1 when agg_sig is NULL / infinity) to `lines`, and *status gets the first payload error in rank order (the
// machine's CHECK never overwrites an error).
#define BN_PAYLOAD 448
__global__ void __launch_bounds__(BN_BLOCK) k_finish_prepare(const uint8_t* __restrict__ payloads, int m, const uint8_t* __restrict__ agg_sig,
                                                             const line_t* __restrict__ table, u4* __restrict__ lines, u4* __restrict__ fio,
                                                             uint8_t* __restrict__ status) {
  __shared__ fq12 sh[BN_PROD_BLOCK];
  __shared__ int first_err;
  __shared__ lines_consts sconst;
  if (threadIdx.x == 0) first_err = 0x7fffffff;
  __syncthreads();
  if (threadIdx.x < BN_PROD_BLOCK) {
    fq12 acc, t;
    fq12_set_one(&acc);
    for (int i = threadIdx.x; i < m; i += BN_PROD_BLOCK) {
      const uint8_t* pl = payloads + (size_t)BN_PAYLOAD * i;
      int st = pl[384];
      if (!st && !fq12_from_be(&t, pl)) st = ST_NOT_MEMBER;
      if (st) atomicMin(&first_err, (i << 8) | st);
      else fq12_mul(&acc, &acc, &t);
    }
    sh[threadIdx.x] = acc;
  }
  __syncthreads();
  for (int s = BN_PROD_BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) fq12_mul(&sh[threadIdx.x], &sh[threadIdx.x], &sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    g1j sg;
    pt_set_inf(&sg);
    int st = agg_sig ? g1_from_raw(&sg, agg_sig) : ST_OK;
    if (st) atomicMin(&first_err, (m << 8) | st);
    sconst.v[3].c0 = sg.x;
    sconst.v[3].c1 = sg.y;
    sconst.v[2].c0 = pt_is_inf(&sg) ? fq_zero() : fq_one();  // "use" flag for the line writers
    const int pos[6] = {0, 2, 4, 1, 3, 5};  // tower order (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2) -> power-basis coefficient
    const fq2* c = &sh[0].c0.c0;
    for (int t = 0; t < 6; t++) {
      coop_gst(fio, (size_t)pos[t] * 2 + 0, COOP_LANES, 0, c[t].c0);
      coop_gst(fio, (size_t)pos[t] * 2 + 1, COOP_LANES, 0, c[t].c1);
    }
  }
  __syncthreads();
  if (threadIdx.x < K_N_LINES) {
    const int mline = threadIdx.x;
    const bool use = !fq_is_zero(sconst.v[2].c0);
    coop_emit_scaled_v(lines, mline, COOP_LANES, 0, use, table[mline].ell_0, table[mline].ell_vw, table[mline].ell_vv, sconst.v[3]);
  }
  if (threadIdx.x == 0) *status = first_err == 0x7fffffff ? 0 : (uint8_t)(first_err & 0xff);
}

// prod(partials) * miller(agg_sig, -G2), final exponentiation, verdict
// (the one-thread form: pairing mode 1, and the cross-check of the cooperative finish in the tests)
__global__ void k_distinct_finish(const uint8_t* __restrict__ payloads, int m, const uint8_t* __restrict__ agg_sig,
                                  const line_t* __restrict__ lines, uint8_t* __restrict__ status) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  fq12 acc, t;
  fq12_set_one(&acc);
  for (int i = 0; i < m; i++) {
    const uint8_t* pl = payloads + (size_t)BN_PAYLOAD * i;
    if (pl[384]) {
      *status = pl[384];
      return;
    }
    if (!fq12_from_be(&t, pl)) {
      *status = ST_NOT_MEMBER;
      return;
    }
    fq12_mul(&acc, &acc, &t);
  }
  g1j s;
  pt_set_inf(&s);
  int st = agg_sig ? g1_from_raw(&s, agg_sig) : ST_OK;
  if (st) {
    *status = (uint8_t)st;
    return;
  }
  if (!pt_is_inf(&s)) {
    fq2 dummy = fq2_one();
    miller_loop_2(&t, false, &s.x, &s.y, &dummy, &dummy, true, &s.x, &s.y, lines);
    fq12_mul(&acc, &acc, &t);
  }
  *status = item_final_exp_is_one(&acc);
}

// out = first non-zero of st[0 .. m)
__global__ void k_first_status(const uint8_t* __restrict__ st, int m, uint8_t* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint8_t r = 0;
  for (int i = 0; i < m && !r; i++) r = st[i];
  *out = r;
}
// 32 big-endian bytes -> 32 little-endian bytes
__device__ __forceinline__ void rev32(uint8_t* dst, const uint8_t* src) {
  for (int i = 0; i < 32; i++) dst[i] = src[31 - i];
}
// /root/reference/src/utils.rs:197-239; the order of the error checks is the reference's: hash, public key, signature
__global__ void __launch_bounds__(BN_BLOCK) k_format_pairing_check(const g1aff* __restrict__ H, const uint8_t* __restrict__ sigs,
                                                                   const uint8_t* __restrict__ pks, size_t n, int compressed,
                                                                   uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t* o = out + 384 * i;
  int st = status[i];  // hash status
  uint8_t sig[64], pk[128];
  if (!st) {
    if (compressed) {
      st = item_g2_decompress(pk, pks + 65 * i);
      if (!st) st = item_g1_decompress(sig, sigs + 33 * i);
    } else {
      for (int k = 0; k < 128; k++) pk[k] = pks[128 * i + k];
      for (int k = 0; k < 64; k++) sig[k] = sigs[64 * i + k];
    }
  }
  status[i] = (uint8_t)st;
  if (st) {
    for (int k = 0; k < 384; k++) o[k] = 0;
    return;
  }
  uint8_t hb[64];
  fq_to_be(hb, H[i].x);
  fq_to_be(hb + 32, H[i].y);
  rev32(o, hb);
  rev32(o + 32, hb + 32);
  for (int k = 0; k < 4; k++) rev32(o + 64 + 32 * k, pk + 32 * k);
  rev32(o + 192, sig);
  rev32(o + 224, sig + 32);
  // -G2::one(): x.re, x.im, y.re, y.im of the negated generator
  fq2 gx = fq2_from_limbs(K_G2_GEN_X), gy = fq2_neg(fq2_from_limbs(K_G2_GEN_Y));
  uint8_t gb[128];
  fq_to_be(gb, gx.c0);
  fq_to_be(gb + 32, gx.c1);
  fq_to_be(gb + 64, gy.c0);
  fq_to_be(gb + 96, gy.c1);
  for (int k = 0; k < 4; k++) rev32(o + 256 + 32 * k, gb + 32 * k);
}

// ------------------------------------------------------------------------------------------------ host side
struct bn254_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host-buffer verify: signatures and keys are uploaded here while the hash kernels run
  cudaEvent_t ev_alloc = nullptr, ev_copy = nullptr;
  cudaStream_t aux_stream = nullptr;   // small-batch verify: the line producer runs here WHILE the machine consumes its line sets
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool pipeline_small = true;          // BN254_PIPELINE=0 turns the producer / machine overlap off (measurement)
  size_t lines_lat_max = 0;            // largest launch that uses the small-batch producer (default 160 items per SM; BN254_LINES_LAT_MAX)
  bool lines_walk4 = true;             // BN254_LINES_WALK4=0: small batches use the one-thread-per-item latency producer (measurement)
  size_t piped_max_groups = 0;         // most groups of a pipelined verify (default: 3/4 of the SMs; BN254_PIPED_MAX_GROUPS)
  bool coop_tail_split = true;         // BN254_COOP_TAIL_SPLIT=0: the remainder of a launch of a few waves runs as four-group blocks too (measurement)
  size_t test_mute_item = ~(size_t)0;  // bn254_set_test_fault: the pipelined producer never publishes this item (exercises BN254_ENGINE_FAULT)
  unsigned fault_retries = 0;          // host-buffer calls that found a BN254_ENGINE_FAULT status and ran again without pipelining
  bool hash_groups = true;             // BN254_HASH_GROUPS=0: the counter-parallel hash always uses 32 lanes per message, up to 8192 messages (measurement)
  bool coop18 = true;                  // BN254_COOP18=0: twelve-warp blocks instead of eighteen (one warp per coefficient and Karatsuba component)
  bool coop12 = true;                  // BN254_COOP12=0: six-warp blocks even when a group has an SM to itself (measurement)
  line_t* d_lines = nullptr;
  aff<fq>* d_comb_g1 = nullptr;   // (d + 1) * 16^w * G1 generator
  aff<fq2>* d_comb_g2 = nullptr;  // (d + 1) * 16^w * G2 generator
  uint64_t launches = 0;
  // 0: cooperative machine (coop.cuh) in its default form (block layout, four groups per block, one per sub-partition),
  // 1: one thread per item (pairing.cuh), 2: cooperative, block layout with one 32-item group per six-warp block,
  // 3: cooperative, warp-local layout (six lanes per item)
  int pairing_mode = 0;
  unsigned coop_stagger = BN_COOP_STAGGER;  // start offset between co-resident blocks of k_coop_run, SM cycles
  int chunk_log2 = BN_COOP_CHUNK_LOG2;  // verify: items per line-set workspace chunk (50 KB of line sets per item)
  int coop_groups4 = BN_COOP_DEFAULT_GROUPS4;  // block layout: four groups per 24-warp block, one group per sub-partition (k_coop4_run)
  bool coop_h = BN_COOP_DEFAULT_H;  // verify uses the half-warp layout (k_cooph_run, pairing mode 4)
  bool coop_w = BN_COOP_DEFAULT_W;  // layout mode 0 uses for verify (BN254_COOP_W=0/1 in the environment overrides)
  cudaMemPool_t pool = nullptr;  // private stream-ordered pool: every temporary of this context comes from it and goes back on destroy
  int input_policy = BN254_INPUTS_UNTRUSTED;
  int hash_try_limit = 255;      // /root/reference/src/hash.rs:39; lowered only by the test hook bn254_set_hash_try_limit
  bool lines_throughput_only = false;  // BN254_LINES_LAT=0: never use the latency form of the line producer (measurement)
  std::string err;
  // optional per-phase timing of the verify pipeline (bn254_set_profiling): events recorded on `stream`
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;  // groups of 4: before hash, after hash, after miller, after final exp
};
static thread_local std::string g_create_err;

#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) {                                                                              \
      char buf_[512];                                                                                     \
      snprintf(buf_, sizeof buf_, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      ctx->err = buf_;                                                                                    \
      return BN254_E_CUDA;                                                                                \
    }                                                                                                     \
  } while (0)

static inline unsigned grid_for(size_t n, int block = BN_BLOCK) { return (unsigned)((n + block - 1) / block); }

// RAII device buffer on the context's stream (stream-ordered pool allocation)
struct dbuf {
  bn254_ctx* ctx;
  void* p = nullptr;
  dbuf(bn254_ctx* c) : ctx(c) {}
  cudaError_t alloc(size_t bytes) { return cudaMallocFromPoolAsync(&p, bytes ? bytes : 1, ctx->pool, ctx->stream); }
  void release() {
    if (p) cudaFreeAsync(p, ctx->stream);
    p = nullptr;
  }
  ~dbuf() {
    if (p) cudaFreeAsync(p, ctx->stream);
  }
  template <class T> T* as() { return (T*)p; }
};

extern "C" {

int bn254_ctx_create(int device, bn254_ctx** out) {
  if (!out) return BN254_E_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_err = std::string("no CUDA device available (") + cudaGetErrorString(e) + "): this engine has no CPU fallback";
    return BN254_E_CUDA;
  }
  if (device < 0 || device >= count) {
    g_create_err = "device index out of range";
    return BN254_E_ARG;
  }
  bn254_ctx* ctx = new bn254_ctx();
  ctx->device = device;
  if (const char* w = getenv("BN254_COOP_W")) ctx->coop_w = w[0] == '1';
  if (const char* w = getenv("BN254_COOP_CHUNK_LOG2")) {
    int v = atoi(w);
    if (v >= 10 && v <= 21) ctx->chunk_log2 = v;
  }
  if (const char* w = getenv("BN254_COOP_H")) ctx->coop_h = w[0] == '1';
  if (const char* w = getenv("BN254_LINES_LAT")) ctx->lines_throughput_only = w[0] == '0';
  if (const char* w = getenv("BN254_COOP_GROUPS4")) ctx->coop_groups4 = atoi(w);
  if (const char* w = getenv("BN254_COOP_STAGGER")) ctx->coop_stagger = (unsigned)atoi(w);  // tuning knob (cycles)
  auto fail = [&](const char* what, cudaError_t ee) {
    g_create_err = std::string(what) + ": " + cudaGetErrorString(ee);
    delete ctx;
    return BN254_E_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  ctx->sm_count = prop.multiProcessorCount;
  // the one-thread-per-item pairing kernels keep Fq12 temporaries on the stack; the limit is device-wide, so it is only ever raised
  size_t stack_now = 0;
  if (cudaDeviceGetLimit(&stack_now, cudaLimitStackSize) != cudaSuccess || stack_now < 16 * 1024)
    if ((e = cudaDeviceSetLimit(cudaLimitStackSize, 16 * 1024)) != cudaSuccess) return fail("cudaDeviceSetLimit(stack)", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate(copy)", e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_alloc, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate(aux)", e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  if (const char* w = getenv("BN254_PIPELINE")) ctx->pipeline_small = w[0] != '0';
  if (const char* w = getenv("BN254_COOP12")) ctx->coop12 = w[0] != '0';
  if (const char* w = getenv("BN254_COOP18")) ctx->coop18 = w[0] != '0';
  if (const char* w = getenv("BN254_HASH_GROUPS")) ctx->hash_groups = w[0] != '0';
  if (const char* w = getenv("BN254_COOP_TAIL_SPLIT")) ctx->coop_tail_split = w[0] != '0';
  ctx->piped_max_groups = (size_t)(ctx->sm_count - ctx->sm_count / 4);
  if (const char* w = getenv("BN254_PIPED_MAX_GROUPS")) ctx->piped_max_groups = (size_t)atoll(w);
  if (ctx->piped_max_groups > (size_t)ctx->sm_count) ctx->piped_max_groups = (size_t)ctx->sm_count;
  if (const char* w = getenv("BN254_LINES_WALK4")) ctx->lines_walk4 = w[0] != '0';
  ctx->lines_lat_max = (size_t)ctx->sm_count * 160;
  if (const char* w = getenv("BN254_LINES_LAT_MAX")) ctx->lines_lat_max = (size_t)atoll(w);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  // A private pool (the device's default pool is shared with the host process, e.g. torch): freed blocks stay cached here
  // between calls -- the line-set workspace of verify is allocated once, not per call -- and everything is returned to the
  // driver by bn254_ctx_destroy / bn254_trim.
  {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if ((e = cudaMemPoolCreate(&ctx->pool, &props)) != cudaSuccess) return fail("cudaMemPoolCreate", e);
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  if ((e = cudaMalloc(&ctx->d_lines, sizeof(line_t) * K_N_LINES)) != cudaSuccess) return fail("cudaMalloc(lines)", e);
  k_init_lines<<<1, 1, 0, ctx->stream>>>(ctx->d_lines);
  ctx->launches++;
  if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_init_lines launch", e);
  if ((e = cudaMalloc(&ctx->d_comb_g1, sizeof(aff<fq>) * BN_COMB_WINDOWS * BN_COMB_ROW)) != cudaSuccess) return fail("cudaMalloc(comb g1)", e);
  if ((e = cudaMalloc(&ctx->d_comb_g2, sizeof(aff<fq2>) * BN_COMB_WINDOWS * BN_COMB_ROW)) != cudaSuccess) return fail("cudaMalloc(comb g2)", e);
  k_init_comb<<<2, 32, 0, ctx->stream>>>(ctx->d_comb_g1, ctx->d_comb_g2);
  ctx->launches++;
  if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_init_comb launch", e);
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail("k_init_lines / k_init_comb", e);
  if ((e = cudaFuncSetAttribute(k_coop12_run, cudaFuncAttributeMaxDynamicSharedMemorySize, COOP12_SMEM_BYTES)) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_coop12_run)", e);
  // the pipelined pair (producer + machine) must be co-resident on every SM whichever of the two is dispatched first
  cudaFuncSetAttribute(k_coop12_run, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if ((e = cudaFuncSetAttribute(k_coop18_run, cudaFuncAttributeMaxDynamicSharedMemorySize, COOP12_SMEM_BYTES)) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_coop18_run)", e);
  cudaFuncSetAttribute(k_coop18_run, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_coop_run, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_verify_lines_lat, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if ((e = cudaFuncSetAttribute(k_verify_lines_walk4, cudaFuncAttributeMaxDynamicSharedMemorySize, WALK_SMEM_BYTES)) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_verify_lines_walk4)", e);
  cudaFuncSetAttribute(k_verify_lines_walk4, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if ((e = cudaFuncSetAttribute(k_coop_run, cudaFuncAttributeMaxDynamicSharedMemorySize, COOP1_SMEM_BYTES)) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_coop_run)", e);
  if ((e = cudaFuncSetAttribute(k_coopw_run, cudaFuncAttributeMaxDynamicSharedMemorySize, COOPW_SMEM_BYTES)) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_coopw_run)", e);
  if ((e = cudaFuncSetAttribute(k_coop4_run, cudaFuncAttributeMaxDynamicSharedMemorySize, COOP4_SMEM_BYTES)) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_coop4_run)", e);
  if ((e = cudaFuncSetAttribute(k_cooph_run, cudaFuncAttributeMaxDynamicSharedMemorySize, COOPH_SMEM_BYTES)) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_cooph_run)", e);
  *out = ctx;
  return 0;
}

void bn254_ctx_destroy(bn254_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->d_lines) cudaFree(ctx->d_lines);
  if (ctx->d_comb_g1) cudaFree(ctx->d_comb_g1);
  if (ctx->d_comb_g2) cudaFree(ctx->d_comb_g2);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->aux_stream) {
    cudaStreamSynchronize(ctx->aux_stream);
    cudaStreamDestroy(ctx->aux_stream);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_alloc) cudaEventDestroy(ctx->ev_alloc);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  delete ctx;
}
const char* bn254_last_error(bn254_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
int bn254_sync(bn254_ctx* ctx) {
  if (!ctx) return BN254_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int bn254_trim(bn254_ctx* ctx) {
  if (!ctx) return BN254_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemPoolTrimTo(ctx->pool, 0));
  return 0;
}
int bn254_set_input_policy(bn254_ctx* ctx, int policy) {
  if (!ctx || (policy != BN254_INPUTS_UNTRUSTED && policy != BN254_INPUTS_TYPED)) return BN254_E_ARG;
  ctx->input_policy = policy;
  return 0;
}
int bn254_get_input_policy(bn254_ctx* ctx) { return ctx ? ctx->input_policy : BN254_E_ARG; }
int bn254_set_hash_try_limit(bn254_ctx* ctx, int max_tries) {
  if (!ctx || max_tries < 1 || max_tries > 255) return BN254_E_ARG;
  ctx->hash_try_limit = max_tries;
  return 0;
}
void* bn254_stream(bn254_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int bn254_sm_count(bn254_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t bn254_launch_count(bn254_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

#define LAUNCH(kern, grid, block, ...)                        \
  do {                                                        \
    kern<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__);   \
    ctx->launches++;                                          \
    CK(cudaGetLastError());                                   \
  } while (0)
#define ARGCHECK(cond)                            \
  do {                                            \
    if (!(cond)) {                                \
      if (ctx) ctx->err = "bad argument: " #cond; \
      return BN254_E_ARG;                         \
    }                                             \
  } while (0)
#define ENTER()                \
  ARGCHECK(ctx != nullptr);    \
  CK(cudaSetDevice(ctx->device))
#define H2D(dst, src, bytes) CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, ctx->stream))
#define D2H(dst, src, bytes) CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, ctx->stream))
#define DALLOC(name, bytes) \
  dbuf name(ctx);           \
  CK(name.alloc(bytes))

// ---- device-pointer pipelines (asynchronous on ctx->stream)
#ifndef BN_HASH_ROUNDS
#define BN_HASH_ROUNDS 10
#endif
#define BN_HASH_WIDE_MAX (ctx->hash_groups ? 32768u : 8192u)
static int hash_dev(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint64_t* offsets, size_t n, g1aff* H, uint8_t* status,
                    uint8_t* tries) {
  if (n == 0) return 0;
  const int cap = ctx->hash_try_limit;
  // small batches: one warp per message, 32 counters at a time (latency of ONE try instead of the unluckiest lane's count)
  if (n <= BN_HASH_WIDE_MAX) {
#define HASH_WIDE(G) \
  LAUNCH(k_hash_wide<G>, grid_for(n * G), BN_BLOCK, msgs, msg_len, offsets, n, (const uint32_t*)nullptr, (const uint32_t*)nullptr, 0, cap, H, status, tries)
    if (n <= 1024 || !ctx->hash_groups) {
      HASH_WIDE(32);
    } else if (n <= 4096) {
      HASH_WIDE(8);
    } else if (n <= 16384) {
      HASH_WIDE(4);
    } else {
      HASH_WIDE(2);
    }
#undef HASH_WIDE
    return 0;
  }
  // big batches of ragged / multi-block messages, and callers that want the counters: one thread loops per item
  if (offsets || tries || msg_len > 54 || n > 0xffffffffu) {
    LAUNCH(k_hash_to_g1, grid_for(n), BN_BLOCK, msgs, msg_len, offsets, n, H, status, tries, cap);
    return 0;
  }
  // compacting rounds: expected survivors of round r = n * 0.527^r; grids are sized with a margin and stride over the list.
  // After BN_HASH_ROUNDS = 10 rounds 0.17 % of the items are left: they finish counter-parallel (one warp each) in one more step
  // (A/B r02 at 2^20 messages: 6 rounds 16.7 ms, 8 14.3, 10 14.0, 12 14.2; round 1: 12 rounds + a per-thread tail 15.8).
  DALLOC(lists, sizeof(uint32_t) * 2 * n);
  DALLOC(counts, sizeof(uint32_t) * (BN_HASH_ROUNDS + 1));
  CK(cudaMemsetAsync(counts.p, 0, sizeof(uint32_t) * (BN_HASH_ROUNDS + 1), ctx->stream));
  uint32_t* L[2] = {lists.as<uint32_t>(), lists.as<uint32_t>() + n};
  uint32_t* C = counts.as<uint32_t>();
  double expect = (double)n;
  const int rounds = cap < BN_HASH_ROUNDS ? cap : BN_HASH_ROUNDS;  // (the last step marks whatever is left after `cap` tries)
  for (int r = 0; r < rounds; r++) {
    size_t threads = r == 0 ? n : (size_t)(expect * 1.25) + 4096;
    if (threads > n) threads = n;
    LAUNCH(k_hash_round, grid_for(threads), BN_BLOCK, msgs, (uint32_t)msg_len, (uint32_t)n, (uint32_t)r, r == 0 ? (const uint32_t*)nullptr : L[(r - 1) & 1],
           r == 0 ? (const uint32_t*)nullptr : C + r, L[r & 1], C + r + 1, H, status);
    expect *= 0.5275;
  }
  size_t warps = (size_t)(expect * 1.25) + 4096;
  if (warps > n) warps = n;
  LAUNCH(k_hash_wide<32>, grid_for(warps * 32), BN_BLOCK, msgs, msg_len, (const uint64_t*)nullptr, n, L[(rounds - 1) & 1], C + rounds, rounds, cap, H, status,
         (uint8_t*)nullptr);
  return 0;
}

extern "C" {

int bn254_hash_to_g1_batch_dev(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, size_t n, uint8_t* g1_out, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  DALLOC(H, sizeof(g1aff) * n);
  int rc = hash_dev(ctx, msgs, msg_len, nullptr, n, H.as<g1aff>(), status, nullptr);
  if (rc) return rc;
  LAUNCH(k_g1aff_to_raw, grid_for(n), BN_BLOCK, H.as<g1aff>(), status, n, g1_out);
  return 0;
}
int bn254_hash_to_g1_batch(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, size_t n, uint8_t* g1_out, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(msgs || msg_len == 0);
  ARGCHECK(g1_out && status);
  DALLOC(d_msgs, msg_len * n);
  DALLOC(d_out, 64 * n);
  DALLOC(d_st, n);
  if (msg_len) H2D(d_msgs.p, msgs, msg_len * n);
  int rc = bn254_hash_to_g1_batch_dev(ctx, d_msgs.as<uint8_t>(), msg_len, n, d_out.as<uint8_t>(), d_st.as<uint8_t>());
  if (rc) return rc;
  D2H(g1_out, d_out.p, 64 * n);
  D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int bn254_hash_to_g1_var(bn254_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t n, uint8_t* g1_out, uint8_t* status,
                         uint8_t* tries_out) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(offsets && g1_out && status);
  size_t total = offsets[n];
  DALLOC(d_msgs, total);
  DALLOC(d_off, 8 * (n + 1));
  DALLOC(d_out, 64 * n);
  DALLOC(d_st, n);
  DALLOC(d_tr, n);
  DALLOC(H, sizeof(g1aff) * n);
  if (total) H2D(d_msgs.p, msgs, total);
  H2D(d_off.p, offsets, 8 * (n + 1));
  int rc = hash_dev(ctx, d_msgs.as<uint8_t>(), 0, d_off.as<uint64_t>(), n, H.as<g1aff>(), d_st.as<uint8_t>(), d_tr.as<uint8_t>());
  if (rc) return rc;
  LAUNCH(k_g1aff_to_raw, grid_for(n), BN_BLOCK, H.as<g1aff>(), d_st.as<uint8_t>(), n, d_out.as<uint8_t>());
  D2H(g1_out, d_out.p, 64 * n);
  D2H(status, d_st.p, n);
  if (tries_out) D2H(tries_out, d_tr.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int bn254_sign_batch_dev(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sks, size_t n, uint8_t* sigs, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  DALLOC(H, sizeof(g1aff) * n);
  int rc = hash_dev(ctx, msgs, msg_len, nullptr, n, H.as<g1aff>(), status, nullptr);
  if (rc) return rc;
  LAUNCH(k_sign, grid_for(n), BN_BLOCK, H.as<g1aff>(), sks, n, sigs, status);
  return 0;
}
int bn254_sign_batch(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sks, size_t n, uint8_t* sigs, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK((msgs || msg_len == 0) && sks && sigs && status);
  DALLOC(d_msgs, msg_len * n);
  DALLOC(d_sks, 32 * n);
  DALLOC(d_out, 64 * n);
  DALLOC(d_st, n);
  if (msg_len) H2D(d_msgs.p, msgs, msg_len * n);
  H2D(d_sks.p, sks, 32 * n);
  int rc = bn254_sign_batch_dev(ctx, d_msgs.as<uint8_t>(), msg_len, d_sks.as<uint8_t>(), n, d_out.as<uint8_t>(), d_st.as<uint8_t>());
  if (rc) return rc;
  D2H(sigs, d_out.p, 64 * n);
  D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// one launch of the block-layout cooperative machine over `groups` 32-item groups: four groups per 24-warp block (one per
// sub-partition, k_coop4_run) unless the context asks for the one-group-per-block kernel (pairing mode 2)
static int launch_coop_groups(bn254_ctx* ctx, int which, size_t n, size_t n_pad, const u4* lines, u4* gslots, u4* fio, uint8_t* status,
                              size_t groups, const unsigned* progress = nullptr) {
  const size_t sms = (size_t)ctx->sm_count;
  auto one_group_blocks = [&](size_t g0, size_t cnt) {  // groups g0 .. g0 + cnt - 1, one block each
    const size_t i0 = g0 * COOP_LANES;
    // a group has an SM to itself: twelve warps per group (latency layout)
    if (cnt <= sms && ctx->coop12 && ctx->pairing_mode == 0) {
      (ctx->coop18 ? k_coop18_run : k_coop12_run)<<<(unsigned)cnt, ctx->coop18 ? 3 * COOP_THREADS : COOP12_THREADS, COOP12_SMEM_BYTES, ctx->stream>>>(which, n > i0 ? n - i0 : 0, n_pad, lines + i0,
                                                                                    gslots ? gslots + i0 : gslots, fio ? fio + i0 : fio,
                                                                                    status ? status + i0 : status, progress ? progress + i0 : progress);
      ctx->launches++;
      return;
    }
    k_coop_run<<<(unsigned)cnt, COOP_THREADS, COOP1_SMEM_BYTES, ctx->stream>>>(which, n > i0 ? n - i0 : 0, n_pad, lines + i0, gslots ? gslots + i0 : gslots,
                                                                            fio ? fio + i0 : fio, status ? status + i0 : status, ctx->coop_stagger,
                                                                            (unsigned)sms, progress ? progress + i0 : progress);
    ctx->launches++;
  };
  if (ctx->pairing_mode == 2 || !ctx->coop_groups4) {
    one_group_blocks(0, groups);
    CK(cudaGetLastError());
    return 0;
  }
  // Default: four groups per 24-warp block (one group per sub-partition).  Launches of at most two groups per SM use one-group
  // blocks instead: the groups spread over all SMs and a group's six warps over the four sub-partitions of its SM (2 + 2 + 1 + 1),
  // and the machine's share of a one-item verify drops from 7.6 ms to 3.1 ms.  (From three one-group blocks per SM on, two
  // sub-partitions carry six warps again and the four-group block is the faster layout.  Cutting big launches into whole waves
  // plus a one-group tail was measured too: +5 %, the tail blocks run two deep.)
  if (groups <= 2 * sms) {
    one_group_blocks(0, groups);
    CK(cudaGetLastError());
    return 0;
  }
  // Launches of a few waves: a remainder of at most two groups per SM behind the last whole wave of four-group blocks runs as
  // one-group blocks (eighteen warps when it is at most one group per SM) instead of a mostly empty wave of its own: a wave costs
  // 6.6 ms, the one-group blocks 2.3 ms (one per SM) or 4.7 ms (two per SM).  Only up to eight whole waves -- for big launches
  // the split was measured slower (r02 tuning log, v6) and the remainder is under a few per cent of the time anyway.
  const size_t wave = COOP4_GROUPS * sms, whole = groups / wave, rest = groups % wave;
  if (ctx->coop_tail_split && whole >= 1 && whole <= 8 && rest > 0 && rest <= 2 * sms) {
    const size_t main_groups = whole * wave, main_items = main_groups * COOP_LANES;
    k_coop4_run<<<(unsigned)(main_groups / COOP4_GROUPS), COOP4_THREADS, COOP4_SMEM_BYTES, ctx->stream>>>(which, n < main_items ? n : main_items, n_pad, lines,
                                                                                                        gslots, fio, status, ctx->coop_stagger);
    ctx->launches++;
    one_group_blocks(main_groups, rest);
    CK(cudaGetLastError());
    return 0;
  }
  k_coop4_run<<<(unsigned)((groups + COOP4_GROUPS - 1) / COOP4_GROUPS), COOP4_THREADS, COOP4_SMEM_BYTES, ctx->stream>>>(which, n, n_pad, lines, gslots,
                                                                                                                      fio, status, ctx->coop_stagger);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

// msgs == NULL: check_public_keys form (first G1 argument = generator, no hashing; the caller has zeroed `status`)
static int verify_dev_impl(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                           uint8_t* status, cudaEvent_t inputs_ready = nullptr) {
  const bool coop = ctx->pairing_mode != 1;
  const bool wl = ctx->pairing_mode == 3 || (ctx->pairing_mode == 0 && ctx->coop_w);
  const bool hl = !wl && (ctx->pairing_mode == 4 || (ctx->pairing_mode == 0 && ctx->coop_h));
  // chunking bounds the workspace: the cooperative path stores 174 line sets (50 KB) per item, 27.5 GB for the default chunk of
  // 2^19 items (BN254_COOP_CHUNK_LOG2 changes it; 2^20-item chunks are 0.4 % faster, 2^17-item chunks 1.1 % slower).  The
  // workspace comes from the context's pool and stays cached there between calls; if the device cannot hold it the chunk is halved.
  // (warp-local layout: a whole number of waves of 30-item blocks, so that the last wave of a chunk is not mostly empty)
  size_t CHUNK = !coop ? ((size_t)1 << 20) : wl ? (size_t)ctx->sm_count * BN_COOP_MINB * COOPW_ITEMS * 7 : ((size_t)1 << ctx->chunk_log2);
  // the hash runs once over the whole batch: its compacting rounds are latency-bound when a round gets small, so one
  // pass over n items costs far less than n / CHUNK passes over CHUNK items
  DALLOC(H, sizeof(g1aff) * (msgs ? n : 1));
  dbuf F(ctx), LN(ctx), GS(ctx);
  for (;;) {
    size_t cap = n < CHUNK ? n : CHUNK;
    size_t cap_pad = (cap + COOP_LANES - 1) / COOP_LANES * COOP_LANES;
    cudaError_t e1 = F.alloc(coop ? 16 : sizeof(fq12) * cap);
    cudaError_t e2 = e1 == cudaSuccess ? LN.alloc(coop ? sizeof(u4) * 2 * COOP_LINE_FQ * 2 * K_N_LINES * cap_pad : 16) : e1;
    cudaError_t e3 = e2 == cudaSuccess ? GS.alloc(coop ? sizeof(u4) * COOP_GSLOTS * 6 * 2 * 2 * cap_pad : 16) : e2;
    if (e3 == cudaSuccess) break;
    cudaGetLastError();  // clear the allocation failure and retry with half the chunk
    F.release();
    LN.release();
    GS.release();
    if (e3 != cudaErrorMemoryAllocation || CHUNK <= 1024) {
      ctx->err = std::string("verify workspace allocation failed: ") + cudaGetErrorString(e3);
      return e3 == cudaErrorMemoryAllocation ? BN254_E_NOMEM : BN254_E_CUDA;
    }
    CHUNK >>= 1;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemPoolTrimTo(ctx->pool, 0));  // give cached blocks of other shapes back before the retry
  }
  for (size_t off = 0; off < n; off += CHUNK) {
    size_t m = n - off < CHUNK ? n - off : CHUNK;
    g1aff* h = nullptr;
    auto mark = [&]() -> cudaError_t {
      if (!ctx->prof) return cudaSuccess;
      cudaEvent_t ev;
      cudaError_t e = cudaEventCreate(&ev);
      if (e != cudaSuccess) return e;
      ctx->prof_ev.push_back(ev);
      return cudaEventRecord(ev, ctx->stream);
    };
    CK(mark());
    if (msgs) {
      if (off == 0) {
        int rc = hash_dev(ctx, msgs, msg_len, nullptr, n, H.as<g1aff>(), status, nullptr);
        if (rc) return rc;
      }
      h = H.as<g1aff>() + off;
    }
    if (inputs_ready && off == 0) CK(cudaStreamWaitEvent(ctx->stream, inputs_ready, 0));  // sigs / pks arrive on the copy stream
    CK(mark());
    if (ctx->input_policy == BN254_INPUTS_UNTRUSTED)
      LAUNCH(k_validate_inputs, grid_for(m), BN_BLOCK, sigs + 64 * off, pks + 128 * off, m, status + off);
    if (coop) {
      size_t m_pad = (m + COOP_LANES - 1) / COOP_LANES * COOP_LANES;
      // small batches: every warp is alone on its sub-partition and bound by the latency of its dependent carry chains -> the
      // form with three products in flight (profiles/r02_tuning_log.md section 4); big batches: the compact form
      const bool lat = m <= (size_t)ctx->lines_lat_max && !ctx->lines_throughput_only;
      // Small launches, default layout: producer and machine run CONCURRENTLY.  The producer (aux stream) publishes, per item, how
      // many Miller steps' line sets are in memory; the machine waits for a step's count before it fetches the step's sets, so the
      // walk hides under the Miller loop.  A machine block and a producer block fit one SM together (registers per sub-partition:
      // 3 x 4096 + 4096; shared memory 66 + 55 KB), but nothing here DEPENDS on that: the launch is pipelined only while it leaves
      // a quarter of the SMs free of machine blocks, so producer blocks (four fit an empty SM) always have somewhere to run, whichever
      // kernel the hardware dispatches first.  Launches of more groups run the two kernels one after the other (measured: from
      // ~120 groups on that is also the faster order, the two kernels then compete for the same multipliers).
      const bool piped = lat && !wl && !hl && ctx->pairing_mode == 0 && ctx->coop_groups4 && ctx->pipeline_small && !ctx->prof && n <= CHUNK &&
                         m_pad / COOP_LANES <= ctx->piped_max_groups;
      if (piped) {
        DALLOC(PR, sizeof(unsigned) * m_pad);
        CK(cudaMemsetAsync(PR.p, 0, sizeof(unsigned) * m_pad, ctx->stream));
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
        if (ctx->lines_walk4)
          k_verify_lines_walk4<<<(unsigned)(m_pad / COOP_LANES), WALK_WARPS * 32, WALK_SMEM_BYTES, ctx->aux_stream>>>(
              h, sigs + 64 * off, pks + 128 * off, m, LN.as<u4>(), m_pad, status + off, ctx->d_lines, PR.as<unsigned>(), ctx->test_mute_item);
        else
          k_verify_lines_lat<<<grid_for(m, 32), 32, 0, ctx->aux_stream>>>(h, sigs + 64 * off, pks + 128 * off, m, LN.as<u4>(), m_pad, status + off,
                                                                          ctx->d_lines, PR.as<unsigned>());
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
        int rc = launch_coop_groups(ctx, 0, m, m_pad, LN.as<u4>(), GS.as<u4>(), (u4*)nullptr, status + off, m_pad / COOP_LANES, PR.as<unsigned>());
        if (rc) {
          cudaStreamSynchronize(ctx->aux_stream);
          return rc;
        }
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));  // everything after this call (frees included) is ordered after the producer
        continue;
      }
      if (lat && ctx->lines_walk4) {
        k_verify_lines_walk4<<<(unsigned)(m_pad / COOP_LANES), WALK_WARPS * 32, WALK_SMEM_BYTES, ctx->stream>>>(
            h, sigs + 64 * off, pks + 128 * off, m, LN.as<u4>(), m_pad, status + off, ctx->d_lines, (unsigned*)nullptr, ~(size_t)0);
        ctx->launches++;
        CK(cudaGetLastError());
      } else if (lat)
        LAUNCH(k_verify_lines_lat, grid_for(m, 32), 32, h, sigs + 64 * off, pks + 128 * off, m, LN.as<u4>(), m_pad, status + off, ctx->d_lines,
               (unsigned*)nullptr);
      else
        LAUNCH(k_verify_lines, grid_for(m), BN_BLOCK, h, sigs + 64 * off, pks + 128 * off, m, LN.as<u4>(), m_pad, status + off, ctx->d_lines);
      CK(mark());
      if (wl)
        k_coopw_run<<<(unsigned)((m + COOPW_ITEMS - 1) / COOPW_ITEMS), COOPW_WARPS * 32, COOPW_SMEM_BYTES, ctx->stream>>>(
            0, m, m_pad, LN.as<u4>(), GS.as<u4>(), (u4*)nullptr, status + off);
      else if (hl)
        k_cooph_run<<<(unsigned)((m_pad / COOPH_ROW + COOPH_GROUPS - 1) / COOPH_GROUPS), COOP4_THREADS, COOPH_SMEM_BYTES, ctx->stream>>>(
            0, m, m_pad, LN.as<u4>(), GS.as<u4>(), (u4*)nullptr, status + off, ctx->coop_stagger);
      else {
        int rc = launch_coop_groups(ctx, 0, m, m_pad, LN.as<u4>(), GS.as<u4>(), (u4*)nullptr, status + off, m_pad / COOP_LANES);
        if (rc) return rc;
      }
      if (wl || hl) {
        ctx->launches++;
        CK(cudaGetLastError());
      }
      CK(mark());
    } else {
      LAUNCH(k_verify_miller, grid_for(m), BN_BLOCK, h, sigs + 64 * off, pks + 128 * off, m, F.as<fq12>(), status + off, ctx->d_lines);
      CK(mark());
      LAUNCH(k_final_exp_check, grid_for(m), BN_BLOCK, F.as<fq12>(), m, status + off);
      CK(mark());
    }
  }
  return 0;
}
// ---- verify against cached key lines (bn254_key_lines_prepare_dev): hash, scaling of the cached / fixed lines, the machine
static int verify_cached_dev_impl(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const u4* klines,
                                  const uint8_t* kstatus, size_t n_keys, const uint32_t* key_index, size_t n, uint8_t* status) {
  const size_t k_pad = (n_keys + COOP_LANES - 1) / COOP_LANES * COOP_LANES;
  size_t CHUNK = (size_t)1 << ctx->chunk_log2;
  const int typed = ctx->input_policy == BN254_INPUTS_TYPED ? 1 : 0;
  DALLOC(H, sizeof(g1aff) * n);
  DALLOC(SG, sizeof(g1aff) * (n < CHUNK ? n : CHUNK));
  dbuf LN(ctx), GS(ctx);
  for (;;) {
    size_t cap = n < CHUNK ? n : CHUNK;
    size_t cap_pad = (cap + COOP_LANES - 1) / COOP_LANES * COOP_LANES;
    cudaError_t e2 = LN.alloc(sizeof(u4) * 2 * COOP_LINE_FQ * 2 * K_N_LINES * cap_pad);
    cudaError_t e3 = e2 == cudaSuccess ? GS.alloc(sizeof(u4) * COOP_GSLOTS * 6 * 2 * 2 * cap_pad) : e2;
    if (e3 == cudaSuccess) break;
    cudaGetLastError();
    LN.release();
    GS.release();
    if (e3 != cudaErrorMemoryAllocation || CHUNK <= 1024) {
      ctx->err = std::string("verify workspace allocation failed: ") + cudaGetErrorString(e3);
      return e3 == cudaErrorMemoryAllocation ? BN254_E_NOMEM : BN254_E_CUDA;
    }
    CHUNK >>= 1;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemPoolTrimTo(ctx->pool, 0));
  }
  auto mark = [&]() -> cudaError_t {
    if (!ctx->prof) return cudaSuccess;
    cudaEvent_t ev;
    cudaError_t e = cudaEventCreate(&ev);
    if (e != cudaSuccess) return e;
    ctx->prof_ev.push_back(ev);
    return cudaEventRecord(ev, ctx->stream);
  };
  for (size_t off = 0; off < n; off += CHUNK) {
    const size_t m = n - off < CHUNK ? n - off : CHUNK, m_pad = (m + COOP_LANES - 1) / COOP_LANES * COOP_LANES;
    CK(mark());
    if (off == 0) {
      int rc = hash_dev(ctx, msgs, msg_len, nullptr, n, H.as<g1aff>(), status, nullptr);
      if (rc) return rc;
    }
    CK(mark());
    LAUNCH(k_cached_decode, grid_for(m), BN_BLOCK, sigs + 64 * off, m, key_index ? key_index + off : key_index, off, n_keys, kstatus, typed,
           SG.as<g1aff>(), status + off);
    LAUNCH(k_scale_cached_lines, grid_for(m_pad * K_N_LINES), BN_BLOCK, H.as<g1aff>() + off, SG.as<g1aff>(), m, m_pad, klines, k_pad,
           key_index ? key_index + off : key_index, off, LN.as<u4>(), status + off, ctx->d_lines);
    CK(mark());
    int rc = launch_coop_groups(ctx, CPROG_VERIFY, m, m_pad, LN.as<u4>(), GS.as<u4>(), (u4*)nullptr, status + off, m_pad / COOP_LANES);
    if (rc) return rc;
    CK(mark());
  }
  return 0;
}
size_t bn254_key_lines_bytes(size_t n_keys) {
  const size_t k_pad = (n_keys + COOP_LANES - 1) / COOP_LANES * COOP_LANES;
  return sizeof(u4) * 2 * COOP_KLINE_FQ * K_N_LINES * (k_pad ? k_pad : COOP_LANES);
}
int bn254_key_lines_prepare_dev(bn254_ctx* ctx, const uint8_t* pks, size_t n_keys, uint8_t* key_lines, uint8_t* key_status) {
  ENTER();
  if (n_keys == 0) return 0;
  ARGCHECK(pks && key_lines && key_status && ((uintptr_t)key_lines & 15) == 0);
  const size_t k_pad = (n_keys + COOP_LANES - 1) / COOP_LANES * COOP_LANES;
  LAUNCH(k_key_lines, grid_for(n_keys), BN_BLOCK, pks, n_keys, k_pad, ctx->input_policy == BN254_INPUTS_TYPED ? 1 : 0, (u4*)key_lines, key_status);
  return 0;
}
int bn254_verify_batch_cached_dev(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* key_lines,
                                  const uint8_t* key_status, size_t n_keys, const uint32_t* key_index, size_t n, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK((msgs || msg_len == 0) && sigs && key_lines && key_status && status && n_keys > 0);
  ARGCHECK(key_index != nullptr || n <= n_keys);
  return verify_cached_dev_impl(ctx, msgs, msg_len, sigs, (const u4*)key_lines, key_status, n_keys, key_index, n, status);
}
int bn254_set_pairing_mode(bn254_ctx* ctx, int mode) {
  ENTER();
  ARGCHECK(mode >= 0 && mode <= 4);
  ctx->pairing_mode = mode;
  return 0;
}
int bn254_set_profiling(bn254_ctx* ctx, int on) {
  ENTER();
  ctx->prof = on != 0;
  return 0;
}
// accumulated device time (ms) of the hash / Miller / final-exponentiation kernels of the verify calls made since the
// last query; synchronises the stream
int bn254_phase_ms(bn254_ctx* ctx, float* out3) {
  ENTER();
  ARGCHECK(out3 != nullptr);
  CK(cudaStreamSynchronize(ctx->stream));
  out3[0] = out3[1] = out3[2] = 0.f;
  for (size_t g = 0; g + 3 < ctx->prof_ev.size(); g += 4)
    for (int k = 0; k < 3; k++) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, ctx->prof_ev[g + k], ctx->prof_ev[g + k + 1]));
      out3[k] += ms;
    }
  for (cudaEvent_t ev : ctx->prof_ev) cudaEventDestroy(ev);
  ctx->prof_ev.clear();
  return 0;
}
int bn254_verify_batch_dev(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                           uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(msgs != nullptr);
  return verify_dev_impl(ctx, msgs, msg_len, sigs, pks, n, status);
}
// a host-side status array of a batch small enough to have been pipelined holds BN254_ENGINE_FAULT
static bool faulted(const bn254_ctx* ctx, const uint8_t* status, size_t n) {
  return ctx->pipeline_small && n <= ctx->piped_max_groups * COOP_LANES && memchr(status, ST_ENGINE_FAULT, n) != nullptr;
}
// test hook: the pipelined line producer never publishes item `item` (~0 = off); *retries = host-buffer calls that ran again
int bn254_set_test_fault(bn254_ctx* ctx, size_t item, uint32_t* retries) {
  ENTER();
  ctx->test_mute_item = item;
  if (retries) *retries = ctx->fault_retries;
  return 0;
}
int bn254_verify_batch(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                       uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK((msgs || msg_len == 0) && sigs && pks && status);
  DALLOC(d_msgs, msg_len * n + 1);
  DALLOC(d_sigs, 64 * n);
  DALLOC(d_pks, 128 * n);
  DALLOC(d_st, n);
  // the messages go first on the compute stream (the hash needs nothing else); signatures and keys (6/7 of the bytes) are
  // uploaded on the copy stream while the hash kernels run, and the line-set kernel waits for them
  CK(cudaEventRecord(ctx->ev_alloc, ctx->stream));  // the buffers exist (stream-ordered allocation) from here on
  CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_alloc, 0));
  // from here on copies may be in flight on the copy stream into buffers that the destructors free on the compute stream:
  // every exit path first drains both streams
  auto body = [&]() -> int {
    CK(cudaMemcpyAsync(d_sigs.p, sigs, 64 * n, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaMemcpyAsync(d_pks.p, pks, 128 * n, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    if (msg_len) H2D(d_msgs.p, msgs, msg_len * n);
    int rc = verify_dev_impl(ctx, d_msgs.as<uint8_t>(), msg_len, d_sigs.as<uint8_t>(), d_pks.as<uint8_t>(), n, d_st.as<uint8_t>(), ctx->ev_copy);
    if (rc) return rc;
    D2H(status, d_st.p, n);
    CK(cudaStreamSynchronize(ctx->stream));
    if (faulted(ctx, status, n)) {  // never observed outside the test hook: run again, producer and machine one after the other
      const bool keep = ctx->pipeline_small;
      ctx->pipeline_small = false;
      ctx->fault_retries++;
      rc = verify_dev_impl(ctx, d_msgs.as<uint8_t>(), msg_len, d_sigs.as<uint8_t>(), d_pks.as<uint8_t>(), n, d_st.as<uint8_t>());
      ctx->pipeline_small = keep;
      if (rc) return rc;
      D2H(status, d_st.p, n);
      CK(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
  };
  int rc = body();
  if (rc) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
  }
  return rc;
}
int bn254_check_public_keys_batch(bn254_ctx* ctx, const uint8_t* pk_g2, const uint8_t* pk_g1, size_t n, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(pk_g2 && pk_g1 && status);
  DALLOC(d_g1, 64 * n);
  DALLOC(d_g2, 128 * n);
  DALLOC(d_st, n);
  H2D(d_g1.p, pk_g1, 64 * n);
  H2D(d_g2.p, pk_g2, 128 * n);
  CK(cudaMemsetAsync(d_st.p, 0, n, ctx->stream));
  int rc = verify_dev_impl(ctx, nullptr, 0, d_g1.as<uint8_t>(), d_g2.as<uint8_t>(), n, d_st.as<uint8_t>());
  if (rc) return rc;
  D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  if (faulted(ctx, status, n)) {
    const bool keep = ctx->pipeline_small;
    ctx->pipeline_small = false;
    ctx->fault_retries++;
    cudaError_t e = cudaMemsetAsync(d_st.p, 0, n, ctx->stream);
    rc = e == cudaSuccess ? verify_dev_impl(ctx, nullptr, 0, d_g1.as<uint8_t>(), d_g2.as<uint8_t>(), n, d_st.as<uint8_t>()) : BN254_E_CUDA;
    ctx->pipeline_small = keep;
    if (rc) return rc;
    D2H(status, d_st.p, n);
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

int bn254_pairing_check_batch_dev(bn254_ctx* ctx, const uint8_t* g1s, const uint8_t* g2s, size_t k, size_t n, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(status && (k == 0 || (g1s && g2s)));
  if (ctx->pairing_mode == 1 || ctx->pairing_mode >= 3 || (k != 1 && k != 2)) {  // one thread per item: any k
    DALLOC(F, sizeof(fq12) * n);
    LAUNCH(k_miller_pairs, grid_for(n), BN_BLOCK, g1s, g2s, k, n, F.as<fq12>(), status);
    LAUNCH(k_final_exp_check, grid_for(n), BN_BLOCK, F.as<fq12>(), n, status);
    return 0;
  }
  // cooperative machine: k line streams per item (25 KB each), chunked like verify
  size_t CHUNK = (size_t)1 << ctx->chunk_log2;
  const size_t cap = n < CHUNK ? n : CHUNK, cap_pad = (cap + COOP_LANES - 1) / COOP_LANES * COOP_LANES;
  DALLOC(LN, sizeof(u4) * k * COOP_LINE_FQ * 2 * K_N_LINES * cap_pad);
  DALLOC(GS, sizeof(u4) * COOP_GSLOTS * 6 * 2 * 2 * cap_pad);
  LAUNCH(k_pairs_decode, grid_for(n), BN_BLOCK, g1s, g2s, k, n, status);
  for (size_t off = 0; off < n; off += CHUNK) {
    const size_t m = n - off < CHUNK ? n - off : CHUNK, m_pad = (m + COOP_LANES - 1) / COOP_LANES * COOP_LANES;
    LAUNCH(k_item_pair_lines, grid_for(m * k), BN_BLOCK, g1s + 64 * k * off, g2s + 128 * k * off, (int)k, m, m_pad, LN.as<u4>(), status + off);
    int rc = launch_coop_groups(ctx, k == 1 ? CPROG_PAIRING1 : CPROG_VERIFY, m, m_pad, LN.as<u4>(), GS.as<u4>(), (u4*)nullptr, status + off,
                                m_pad / COOP_LANES);
    if (rc) return rc;
  }
  return 0;
}
int bn254_pairing_check_batch(bn254_ctx* ctx, const uint8_t* g1s, const uint8_t* g2s, size_t k, size_t n, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(status && (k == 0 || (g1s && g2s)));
  DALLOC(d_g1, 64 * k * n);
  DALLOC(d_g2, 128 * k * n);
  DALLOC(d_st, n);
  if (k) {
    H2D(d_g1.p, g1s, 64 * k * n);
    H2D(d_g2.p, g2s, 128 * k * n);
  }
  int rc = bn254_pairing_check_batch_dev(ctx, d_g1.as<uint8_t>(), d_g2.as<uint8_t>(), k, n, d_st.as<uint8_t>());
  if (rc) return rc;
  D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int bn254_miller_loop_batch(bn254_ctx* ctx, const uint8_t* g1s, const uint8_t* g2s, size_t k, size_t n, uint8_t* f_out384, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(status && f_out384 && (k == 0 || (g1s && g2s)));
  DALLOC(d_g1, 64 * k * n);
  DALLOC(d_g2, 128 * k * n);
  DALLOC(d_st, n);
  DALLOC(F, sizeof(fq12) * n);
  DALLOC(d_out, 384 * n);
  if (k) {
    H2D(d_g1.p, g1s, 64 * k * n);
    H2D(d_g2.p, g2s, 128 * k * n);
  }
  LAUNCH(k_miller_pairs, grid_for(n), BN_BLOCK, d_g1.as<uint8_t>(), d_g2.as<uint8_t>(), k, n, F.as<fq12>(), d_st.as<uint8_t>());
  LAUNCH(k_fq12_to_be, grid_for(n), BN_BLOCK, F.as<fq12>(), d_st.as<uint8_t>(), n, d_out.as<uint8_t>());
  D2H(f_out384, d_out.p, 384 * n);
  D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int bn254_final_exp_batch(bn254_ctx* ctx, const uint8_t* f_in384, size_t n, uint8_t* gt_out384, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(f_in384 && gt_out384 && status);
  DALLOC(d_in, 384 * n);
  DALLOC(d_out, 384 * n);
  DALLOC(d_st, n);
  H2D(d_in.p, f_in384, 384 * n);
  LAUNCH(k_final_exp_bytes, grid_for(n), BN_BLOCK, d_in.as<uint8_t>(), n, d_out.as<uint8_t>(), d_st.as<uint8_t>());
  D2H(gt_out384, d_out.p, 384 * n);
  D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int bn254_fq_op_batch(bn254_ctx* ctx, int op, const uint8_t* a32, const uint8_t* b32, size_t n, uint8_t* out32, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(a32 && out32 && status);
  DALLOC(d_a, 32 * n);
  DALLOC(d_b, 32 * n);
  DALLOC(d_out, 32 * n);
  DALLOC(d_st, n);
  H2D(d_a.p, a32, 32 * n);
  if (b32) H2D(d_b.p, b32, 32 * n);
  else CK(cudaMemsetAsync(d_b.p, 0, 32 * n, ctx->stream));
  LAUNCH(k_fq_op, grid_for(n), BN_BLOCK, op, d_a.as<uint8_t>(), d_b.as<uint8_t>(), n, d_out.as<uint8_t>(), d_st.as<uint8_t>());
  D2H(out32, d_out.p, 32 * n);
  D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int bn254_layer_op_batch(bn254_ctx* ctx, int op, const uint8_t* in, size_t n_in, size_t n, uint8_t* out, size_t n_out) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(in && out && n_in <= 20 && n_out <= 12);
  DALLOC(d_in, 32 * n_in * n);
  DALLOC(d_out, 32 * n_out * n);
  H2D(d_in.p, in, 32 * n_in * n);
  LAUNCH(k_layer_op, grid_for(n), BN_BLOCK, op, d_in.as<uint8_t>(), (int)n_in, n, d_out.as<uint8_t>(), (int)n_out);
  D2H(out, d_out.p, 32 * n_out * n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int bn254_fq12_op_batch(bn254_ctx* ctx, int op, const uint8_t* a384, const uint8_t* b384, size_t n, uint8_t* out384, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(a384 && out384 && status);
  DALLOC(d_a, 384 * n);
  DALLOC(d_b, 384 * n);
  DALLOC(d_out, 384 * n);
  DALLOC(d_st, n);
  H2D(d_a.p, a384, 384 * n);
  if (b384) H2D(d_b.p, b384, 384 * n);
  else CK(cudaMemsetAsync(d_b.p, 0, 384 * n, ctx->stream));
  LAUNCH(k_fq12_op, grid_for(n), BN_BLOCK, op, d_a.as<uint8_t>(), d_b.as<uint8_t>(), n, d_out.as<uint8_t>(), d_st.as<uint8_t>());
  D2H(out384, d_out.p, 384 * n);
  D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"

// generic host wrapper of k_item_op
template <int OP>
static int item_op_host(bn254_ctx* ctx, const uint8_t* a, size_t a_bytes, const uint8_t* b, size_t b_bytes, size_t n, uint8_t* out,
                        size_t out_bytes, uint8_t* status, const void* dev_b = nullptr) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK(a != nullptr && (b_bytes == 0 || b != nullptr));
  DALLOC(d_a, a_bytes * n);
  DALLOC(d_b, b_bytes * n);
  DALLOC(d_out, out_bytes * n);
  DALLOC(d_st, n);
  H2D(d_a.p, a, a_bytes * n);
  if (b_bytes) H2D(d_b.p, b, b_bytes * n);
  // dev_b: a device-resident second operand (the context's fixed-base table) instead of per-item host bytes
  LAUNCH(k_item_op<OP>, grid_for(n), BN_BLOCK, d_a.as<uint8_t>(), dev_b ? (const uint8_t*)dev_b : d_b.as<uint8_t>(), n, d_out.as<uint8_t>(),
         d_st.as<uint8_t>());
  if (out && out_bytes) D2H(out, d_out.p, out_bytes * n);
  if (status) D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

template <class F>
static int sum_dev_impl(bn254_ctx* ctx, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out, uint8_t* status, int strict = 0) {
  size_t want = (n + 7) / 8;  // >= 8 points per thread before the tree
  size_t max_blocks = (size_t)ctx->sm_count * 4;
  size_t blocks = (want + BN_BLOCK - 1) / BN_BLOCK;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks == 0) blocks = 1;
  DALLOC(partial, sizeof(jac<F>) * blocks);
  DALLOC(err, 8);
  CK(cudaMemsetAsync(err.p, 0xff, 8, ctx->stream));
  LAUNCH(k_sum_partial<F>, (unsigned)blocks, BN_BLOCK, pts, neg, n, partial.as<jac<F>>(), err.as<unsigned long long>(), strict);
  LAUNCH(k_sum_final<F>, 1, BN_BLOCK, partial.as<jac<F>>(), (int)blocks, out, err.as<unsigned long long>(), status);
  return 0;
}
template <class F>
static int sum_host_impl(bn254_ctx* ctx, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out, uint8_t* status) {
  const size_t B = pt_io<F>::BYTES;
  ARGCHECK(out && status && (n == 0 || pts));
  DALLOC(d_pts, B * n);
  DALLOC(d_neg, n);
  DALLOC(d_out, B);
  DALLOC(d_st, 1);
  if (n) H2D(d_pts.p, pts, B * n);
  if (n && neg) H2D(d_neg.p, neg, n);
  int rc = sum_dev_impl<F>(ctx, d_pts.as<uint8_t>(), neg ? d_neg.as<uint8_t>() : nullptr, n, d_out.as<uint8_t>(), d_st.as<uint8_t>());
  if (rc) return rc;
  D2H(out, d_out.p, B);
  D2H(status, d_st.p, 1);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

struct dptr {  // a borrowed device pointer with the accessor of dbuf
  void* p;
  template <class T> T* as() { return (T*)p; }
};
// product of `count` Fq12 values on the device -> *f_be (384 bytes) and *status (first error by index, or 0); folds in
// levels of 64 so that no single block multiplies more than a few thousand values in sequence
static int fq12_product_dev(bn254_ctx* ctx, fq12* vals, size_t count, dbuf& scratch, uint8_t* f_be, unsigned long long* err, uint8_t* status) {
  fq12* cur = vals;
  size_t m = count;
  fq12* sc = scratch.as<fq12>();
  while (m > 4096) {
    size_t per = 64, blocks = (m + per - 1) / per;
    LAUNCH(k_fq12_prod_level, (unsigned)blocks, BN_PROD_BLOCK, cur, m, per, sc);
    cur = sc;
    sc += blocks;
    m = blocks;
  }
  LAUNCH(k_fq12_prod_final, 1, BN_PROD_BLOCK, cur, (int)m, (fq12*)nullptr, f_be, err, status);
  return 0;
}
// Miller-product partial of (P_i, pk_i), i < n, for G1 points already on the device (affine, Montgomery form; hst_p = their
// per-item status)  ->  f_be (384 bytes, device) ; status = first error or 0.  extra_sig (device, 64 bytes, may be NULL):
// one more pair (extra_sig, -G2) is folded into the product.
static int distinct_partial_points_dev(bn254_ctx* ctx, g1aff* H_p, uint8_t* hst_p, const uint8_t* pks, size_t n, uint8_t* f_be,
                                       uint8_t* status, const uint8_t* extra_sig = nullptr) {
  dptr H{H_p}, hst{hst_p};
  DALLOC(err, 8);
  CK(cudaMemsetAsync(err.p, 0xff, 8, ctx->stream));
  const int typed = ctx->input_policy == BN254_INPUTS_TYPED ? 1 : 0;
  const size_t N = n + (extra_sig ? 1 : 0);  // pairs, the optional extra one last
  if (ctx->pairing_mode != 1 && N > 0) {
    // Cooperative multi-pairing.  A block of the default layout is four 32-lane groups, every lane folds mk pairs (shared
    // squarings) and every group leaves one partial product.  Work is cut into launches of whole waves: full waves of
    // mk = 8 blocks (1024 pairs each, at most 6 waves = 23 GB of line sets per launch), then ONE more wave for the
    // remainder with the smallest mk in {8, 4, 2, 1} whose block count still fits a wave -- a partial last wave of
    // mk = 8 blocks would cost a full wave's time (15 % of the step when 2^22 pairs are spread over 8 GPUs).
    const size_t sms = (size_t)ctx->sm_count, gpb = (ctx->pairing_mode == 2 || !ctx->coop_groups4) ? 1 : COOP4_GROUPS;
    struct seg { size_t off, cnt; int mk, prog; };
    std::vector<seg> segs;
    const size_t pairs_wave8 = sms * gpb * COOP_LANES * 8;
    size_t full = N / pairs_wave8 * pairs_wave8;
    for (size_t off = 0; off < full;) {
      size_t c = full - off < 6 * pairs_wave8 ? full - off : 6 * pairs_wave8;
      segs.push_back({off, c, 8, CPROG_MULTI8});
      off += c;
    }
    if (N > full) {
      const size_t R = N - full;
      int mk = 8, prog = CPROG_MULTI8;
      const int mks[3] = {1, 2, 4}, progs[3] = {CPROG_MULTI1, CPROG_MULTI2, CPROG_MULTI4};
      for (int t = 0; t < 3; t++)
        if ((R + gpb * COOP_LANES * mks[t] - 1) / (gpb * COOP_LANES * mks[t]) <= sms) {
          mk = mks[t];
          prog = progs[t];
          break;
        }
      segs.push_back({full, R, mk, prog});
    }
    size_t total_groups = 0, max_pairs_lanes = 0, maxL = 0;
    for (auto& sg : segs) {
      size_t groups = (sg.cnt + (size_t)COOP_LANES * sg.mk - 1) / ((size_t)COOP_LANES * sg.mk);
      total_groups += groups;
      size_t L = groups * COOP_LANES;
      if (L * sg.mk > max_pairs_lanes) max_pairs_lanes = L * sg.mk;
      if (L > maxL) maxL = L;
    }
    DALLOC(LN, sizeof(u4) * COOP_LINE_FQ * 2 * K_N_LINES * max_pairs_lanes);
    DALLOC(FIO, sizeof(u4) * 6 * 2 * 2 * maxL);
    DALLOC(partial, sizeof(fq12) * total_groups);
    DALLOC(scratch, sizeof(fq12) * (total_groups / 32 + 64));
    size_t done_groups = 0;
    for (auto& sg : segs) {
      size_t groups = (sg.cnt + (size_t)COOP_LANES * sg.mk - 1) / ((size_t)COOP_LANES * sg.mk), L = groups * COOP_LANES;
      // pairs of this launch that come from the caller's arrays (the extra pair, if any, is the very last pair of all)
      size_t own = sg.off + sg.cnt > n ? (n > sg.off ? n - sg.off : 0) : sg.cnt;
      const uint8_t* ex = (extra_sig && sg.off + sg.cnt == N) ? extra_sig : nullptr;
      LAUNCH(k_pair_lines, grid_for(L * sg.mk), BN_BLOCK, H.as<g1aff>() + sg.off, hst.as<uint8_t>() + sg.off, pks + 128 * sg.off, own, L, sg.mk,
             LN.as<u4>(), err.as<unsigned long long>(), sg.off, ex, typed);
      int rc2 = launch_coop_groups(ctx, sg.prog, L, L, LN.as<u4>(), (u4*)nullptr, FIO.as<u4>(), (uint8_t*)nullptr, groups);
      if (rc2) return rc2;
      LAUNCH(k_coop_gather, grid_for(groups), BN_BLOCK, FIO.as<u4>(), L, groups, partial.as<fq12>() + done_groups);
      done_groups += groups;
    }
    return fq12_product_dev(ctx, partial.as<fq12>(), total_groups, scratch, f_be, err.as<unsigned long long>(), status);
  }
  size_t max_blocks = (size_t)ctx->sm_count * 4;
  size_t blocks = (n + BN_PROD_BLOCK - 1) / BN_PROD_BLOCK;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks == 0) blocks = 1;
  DALLOC(partial, sizeof(fq12) * (blocks + 1));
  LAUNCH(k_distinct_partial, (unsigned)blocks, BN_PROD_BLOCK, H.as<g1aff>(), hst.as<uint8_t>(), pks, n, partial.as<fq12>(),
         err.as<unsigned long long>(), typed);
  if (extra_sig) {
    LAUNCH(k_extra_pair, 1, 32, extra_sig, ctx->d_lines, partial.as<fq12>() + blocks, err.as<unsigned long long>(), n);
    blocks++;
  }
  LAUNCH(k_fq12_prod_final, 1, BN_PROD_BLOCK, partial.as<fq12>(), (int)blocks, (fq12*)nullptr, f_be, err.as<unsigned long long>(), status);
  return 0;
}
// Miller-product partial of (H(msg_i), pk_i), i < n  ->  f_be (384 bytes, device) ; status = first error or 0
static int distinct_partial_dev(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* pks, size_t n, uint8_t* f_be,
                                uint8_t* status, const uint8_t* extra_sig = nullptr) {
  size_t cap = n ? n : 1;
  DALLOC(H, sizeof(g1aff) * cap);
  DALLOC(hst, cap);
  int rc = hash_dev(ctx, msgs, msg_len, nullptr, n, H.as<g1aff>(), hst.as<uint8_t>(), nullptr);
  if (rc) return rc;
  return distinct_partial_points_dev(ctx, H.as<g1aff>(), hst.as<uint8_t>(), pks, n, f_be, status, extra_sig);
}

// ---- randomised batch verification (items.cuh item_rlc_prepare): fast path when every item rides, exact path otherwise
static int verify_rlc_dev_impl(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                               const uint8_t* coeffs16, int flags, uint8_t* status, int* took_fast_path) {
  if (took_fast_path) *took_fast_path = 0;
  if (n == 0) return 0;
  bool all_fast = true;
  // chunks of 2^20 triples (the line sets of a chunk take 26 GB); a chunk is cut into up to 64 slices of whole multi-pairing
  // groups, every slice gets its own verdict from the one pass, and only failing slices are redone by the exact path
  const size_t CH = (size_t)1 << 20, per_group = (size_t)COOP_LANES * COOP_MULTI_K;
  const size_t cap = n < CH ? n : CH, cap_groups = (cap + per_group - 1) / per_group, capL = cap_groups * COOP_LANES;
  DALLOC(H, sizeof(g1aff) * cap);
  DALLOC(hst, cap);
  DALLOC(sigc, 64 * cap);
  DALLOC(bad, 4 * 64);
  DALLOC(agg, 64 * 64);
  DALLOC(verdict, 64);
  DALLOC(err, 8);
  DALLOC(LN, sizeof(u4) * COOP_MULTI_K * COOP_LINE_FQ * 2 * K_N_LINES * capL);
  DALLOC(FIO, sizeof(u4) * 6 * 2 * 2 * capL);
  DALLOC(partial, sizeof(fq12) * cap_groups);
  for (size_t off = 0; off < n; off += CH) {
    const size_t m = n - off < CH ? n - off : CH;
    const size_t groups = (m + per_group - 1) / per_group, L = groups * COOP_LANES;
    const size_t S = groups < 64 ? groups : 64, gps = (groups + S - 1) / S, W = gps * COOP_LANES;
    const size_t slices = (groups + gps - 1) / gps;
    const uint8_t *cm = msgs + msg_len * off, *cs = sigs + 64 * off, *cp = pks + 128 * off;
    CK(cudaMemsetAsync(bad.p, 0, 4 * 64, ctx->stream));
    CK(cudaMemsetAsync(err.p, 0xff, 8, ctx->stream));
    int rc = hash_dev(ctx, cm, msg_len, nullptr, m, H.as<g1aff>(), hst.as<uint8_t>(), nullptr);
    if (rc) return rc;
    LAUNCH(k_rlc_prepare, grid_for(m), BN_BLOCK, H.as<g1aff>(), hst.as<uint8_t>(), cs, cp, coeffs16 + 16 * off, m, (flags & 1) ? 0 : 1,
           sigc.as<uint8_t>(), bad.as<unsigned>(), L, W);
    // pairs of items that cannot ride are harmless here (skipped or multiplied into a slice that is already marked)
    LAUNCH(k_pair_lines, grid_for(L * COOP_MULTI_K), BN_BLOCK, H.as<g1aff>(), hst.as<uint8_t>(), cp, m, L, COOP_MULTI_K, LN.as<u4>(),
           err.as<unsigned long long>(), (size_t)0, (const uint8_t*)nullptr, 1);
    rc = launch_coop_groups(ctx, CPROG_MULTI8, L, L, LN.as<u4>(), (u4*)nullptr, FIO.as<u4>(), (uint8_t*)nullptr, groups);
    if (rc) return rc;
    LAUNCH(k_coop_gather, grid_for(groups), BN_BLOCK, FIO.as<u4>(), L, groups, partial.as<fq12>());
    LAUNCH(k_rlc_slice_sums, (unsigned)slices, BN_BLOCK, sigc.as<uint8_t>(), m, L, W, bad.as<unsigned>(), agg.as<uint8_t>());
    LAUNCH(k_rlc_finish_slices, (unsigned)slices, 32, partial.as<fq12>(), groups, gps, agg.as<uint8_t>(), bad.as<unsigned>(), ctx->d_lines,
           verdict.as<uint8_t>());
    uint8_t h_v[64];
    D2H(h_v, verdict.p, slices);
    CK(cudaStreamSynchronize(ctx->stream));
    // the items of a slice are COOP_MULTI_K runs of W consecutive triples; passing runs get status 0, failing runs are packed
    // into one contiguous batch for a single exact pass (a failing slice alone would fill a fraction of the GPU)
    std::vector<std::pair<size_t, size_t>> redo;
    size_t redo_items = 0;
    for (size_t sl = 0; sl < slices; sl++) {
      for (size_t t = 0; t < COOP_MULTI_K; t++) {
        size_t lo = t * L + sl * W, hi = lo + W;
        if (lo >= m) break;
        if (hi > m) hi = m;
        if (hi > (t + 1) * L) hi = (t + 1) * L;
        if (hi <= lo) continue;
        if (h_v[sl] == 0) {
          CK(cudaMemsetAsync(status + off + lo, 0, hi - lo, ctx->stream));
        } else {
          redo.push_back({lo, hi});
          redo_items += hi - lo;
        }
      }
    }
    if (redo_items) {
      all_fast = false;
      DALLOC(pm, msg_len * redo_items);
      DALLOC(ps, 64 * redo_items);
      DALLOC(pp, 128 * redo_items);
      DALLOC(pst, redo_items);
      size_t at = 0;
      for (auto& r : redo) {
        size_t c = r.second - r.first;
        if (msg_len) CK(cudaMemcpyAsync(pm.as<uint8_t>() + msg_len * at, cm + msg_len * r.first, msg_len * c, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ps.as<uint8_t>() + 64 * at, cs + 64 * r.first, 64 * c, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(pp.as<uint8_t>() + 128 * at, cp + 128 * r.first, 128 * c, cudaMemcpyDeviceToDevice, ctx->stream));
        at += c;
      }
      rc = verify_dev_impl(ctx, pm.as<uint8_t>(), msg_len, ps.as<uint8_t>(), pp.as<uint8_t>(), redo_items, pst.as<uint8_t>());
      if (rc) return rc;
      at = 0;
      for (auto& r : redo) {
        size_t c = r.second - r.first;
        CK(cudaMemcpyAsync(status + off + r.first, pst.as<uint8_t>() + at, c, cudaMemcpyDeviceToDevice, ctx->stream));
        at += c;
      }
    }
  }
  if (took_fast_path) *took_fast_path = all_fast ? 1 : 0;
  return 0;
}

extern "C" {

int bn254_verify_batch_rlc_dev(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                               const uint8_t* coeffs16, int flags, uint8_t* status, int* took_fast_path) {
  ENTER();
  ARGCHECK(n == 0 || (msgs && sigs && pks && coeffs16 && status));
  return verify_rlc_dev_impl(ctx, msgs, msg_len, sigs, pks, n, coeffs16, flags, status, took_fast_path);
}
int bn254_verify_batch_rlc(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                           const uint8_t* coeffs16, int flags, uint8_t* status, int* took_fast_path) {
  ENTER();
  ARGCHECK(n == 0 || (msgs && sigs && pks && status));
  if (took_fast_path) *took_fast_path = 0;
  if (n == 0) return 0;
  std::vector<uint8_t> drawn;
  if (!coeffs16) {  // the engine draws the coefficients itself (OS entropy): they must stay unknown to whoever made the signatures
    drawn.resize(16 * n);
    FILE* f = fopen("/dev/urandom", "rb");
    size_t got = f ? fread(drawn.data(), 1, drawn.size(), f) : 0;
    if (f) fclose(f);
    if (got != drawn.size()) {
      ctx->err = "bn254_verify_batch_rlc: cannot read /dev/urandom for the batch coefficients";
      return BN254_E_ARG;
    }
    coeffs16 = drawn.data();
  }
  DALLOC(d_msgs, msg_len * n);
  DALLOC(d_sigs, 64 * n);
  DALLOC(d_pks, 128 * n);
  DALLOC(d_c, 16 * n);
  DALLOC(d_st, n);
  if (msg_len) H2D(d_msgs.p, msgs, msg_len * n);
  H2D(d_sigs.p, sigs, 64 * n);
  H2D(d_pks.p, pks, 128 * n);
  H2D(d_c.p, coeffs16, 16 * n);
  int rc = verify_rlc_dev_impl(ctx, d_msgs.as<uint8_t>(), msg_len, d_sigs.as<uint8_t>(), d_pks.as<uint8_t>(), n, d_c.as<uint8_t>(), flags,
                               d_st.as<uint8_t>(), took_fast_path);
  if (rc) return rc;
  D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int bn254_g1_mul_batch(bn254_ctx* ctx, const uint8_t* pts, const uint8_t* scalars, size_t n, uint8_t* out64, uint8_t* status) {
  return item_op_host<OP_G1_MUL>(ctx, pts, 64, scalars, 32, n, out64, 64, status);
}
int bn254_g2_mul_batch(bn254_ctx* ctx, const uint8_t* pts, const uint8_t* scalars, size_t n, uint8_t* out128, uint8_t* status) {
  return item_op_host<OP_G2_MUL>(ctx, pts, 128, scalars, 32, n, out128, 128, status);
}
int bn254_derive_pk_g1_batch(bn254_ctx* ctx, const uint8_t* sks, size_t n, uint8_t* out64) {
  return item_op_host<OP_DERIVE_G1>(ctx, sks, 32, nullptr, 0, n, out64, 64, nullptr, ctx ? ctx->d_comb_g1 : nullptr);
}
int bn254_derive_pk_g2_batch(bn254_ctx* ctx, const uint8_t* sks, size_t n, uint8_t* out128) {
  return item_op_host<OP_DERIVE_G2>(ctx, sks, 32, nullptr, 0, n, out128, 128, nullptr, ctx ? ctx->d_comb_g2 : nullptr);
}
int bn254_g1_compress_batch(bn254_ctx* ctx, const uint8_t* raw64, size_t n, uint8_t* out33, uint8_t* status) {
  return item_op_host<OP_G1_COMPRESS>(ctx, raw64, 64, nullptr, 0, n, out33, 33, status);
}
int bn254_g1_decompress_batch(bn254_ctx* ctx, const uint8_t* in33, size_t n, uint8_t* out64, uint8_t* status) {
  return item_op_host<OP_G1_DECOMPRESS>(ctx, in33, 33, nullptr, 0, n, out64, 64, status);
}
int bn254_g2_compress_batch(bn254_ctx* ctx, const uint8_t* raw128, size_t n, uint8_t* out65, uint8_t* status) {
  return item_op_host<OP_G2_COMPRESS>(ctx, raw128, 128, nullptr, 0, n, out65, 65, status);
}
int bn254_g2_decompress_batch(bn254_ctx* ctx, const uint8_t* in65, size_t n, uint8_t* out128, uint8_t* status) {
  return item_op_host<OP_G2_DECOMPRESS>(ctx, in65, 65, nullptr, 0, n, out128, 128, status);
}
int bn254_g1_validate_batch(bn254_ctx* ctx, const uint8_t* raw64, size_t n, uint8_t* status) {
  return item_op_host<OP_G1_VALIDATE>(ctx, raw64, 64, nullptr, 0, n, nullptr, 0, status);
}
int bn254_g2_validate_batch(bn254_ctx* ctx, const uint8_t* raw128, size_t n, uint8_t* status) {
  return item_op_host<OP_G2_VALIDATE>(ctx, raw128, 128, nullptr, 0, n, nullptr, 0, status);
}

int bn254_g1_sum_dev(bn254_ctx* ctx, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out64, uint8_t* status) {
  ENTER();
  return sum_dev_impl<fq>(ctx, pts, neg, n, out64, status);
}
int bn254_g2_sum_dev(bn254_ctx* ctx, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out128, uint8_t* status) {
  ENTER();
  return sum_dev_impl<fq2>(ctx, pts, neg, n, out128, status);
}
int bn254_g1_sum(bn254_ctx* ctx, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out64, uint8_t* status) {
  ENTER();
  return sum_host_impl<fq>(ctx, pts, neg, n, out64, status);
}
int bn254_g2_sum(bn254_ctx* ctx, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out128, uint8_t* status) {
  ENTER();
  return sum_host_impl<fq2>(ctx, pts, neg, n, out128, status);
}

int bn254_aggregate_verify_same_msg(bn254_ctx* ctx, const uint8_t* msg, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                                    uint8_t* status) {
  ENTER();
  ARGCHECK(status && (msg || msg_len == 0) && (n == 0 || (sigs && pks)));
  DALLOC(d_msg, msg_len + 1);
  DALLOC(d_sigs, 64 * n);
  DALLOC(d_pks, 128 * n);
  DALLOC(d_st, 4);
  if (msg_len) H2D(d_msg.p, msg, msg_len);
  if (n) {
    H2D(d_sigs.p, sigs, 64 * n);
    H2D(d_pks.p, pks, 128 * n);
  }
  int rc = bn254_aggregate_verify_same_msg_dev(ctx, d_msg.as<uint8_t>(), msg_len, d_sigs.as<uint8_t>(), d_pks.as<uint8_t>(), n, d_st.as<uint8_t>());
  if (rc) return rc;
  D2H(status, d_st.p, 1);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
// device form: sums + one verify, no host round trip; *status (device) = first failing stage (signature sum, key sum, verify)
int bn254_aggregate_verify_same_msg_dev(bn254_ctx* ctx, const uint8_t* msg, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                                        uint8_t* status) {
  ENTER();
  ARGCHECK(status != nullptr);
  DALLOC(d_asig, 64);
  DALLOC(d_apk, 128);
  DALLOC(d_st, 4);
  uint8_t* st = d_st.as<uint8_t>();
  const int strict = ctx->input_policy == BN254_INPUTS_UNTRUSTED;
  int rc = sum_dev_impl<fq>(ctx, sigs, nullptr, n, d_asig.as<uint8_t>(), st + 0, strict);
  if (rc) return rc;
  rc = sum_dev_impl<fq2>(ctx, pks, nullptr, n, d_apk.as<uint8_t>(), st + 1, strict);
  if (rc) return rc;
  // the sums are values of the crate's types (they may be infinity, and the key sum is in G2 because every key is)
  const int saved = ctx->input_policy;
  ctx->input_policy = BN254_INPUTS_TYPED;
  rc = verify_dev_impl(ctx, msg, msg_len, d_asig.as<uint8_t>(), d_apk.as<uint8_t>(), 1, st + 2);
  ctx->input_policy = saved;
  if (rc) return rc;
  LAUNCH(k_first_status, 1, 32, st, 3, status);
  return 0;
}

int bn254_miller_partial_distinct_dev(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* pks, size_t n, uint8_t* f_out384,
                                      uint8_t* status) {
  ENTER();
  ARGCHECK(f_out384 && status && (n == 0 || (pks && (msgs || msg_len == 0))));
  return distinct_partial_dev(ctx, msgs, msg_len, pks, n, f_out384, status);
}
int bn254_miller_partial_distinct(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* pks, size_t n, uint8_t* f_out384,
                                  uint8_t* status) {
  ENTER();
  ARGCHECK(f_out384 && status && (n == 0 || (pks && (msgs || msg_len == 0))));
  DALLOC(d_msgs, msg_len * n + 1);
  DALLOC(d_pks, 128 * n);
  DALLOC(d_f, 384);
  DALLOC(d_st, 1);
  if (n && msg_len) H2D(d_msgs.p, msgs, msg_len * n);
  if (n) H2D(d_pks.p, pks, 128 * n);
  int rc = distinct_partial_dev(ctx, d_msgs.as<uint8_t>(), msg_len, d_pks.as<uint8_t>(), n, d_f.as<uint8_t>(), d_st.as<uint8_t>());
  if (rc) return rc;
  D2H(f_out384, d_f.p, 384);
  D2H(status, d_st.p, 1);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
// One rank's share of a distinct-message aggregate check as ONE fixed-size device record (BN254_DISTINCT_PAYLOAD_BYTES): the
// Miller product of its (H(msg_i), pk_i) pairs and -- when sigs != NULL -- of the pair (sum of its signatures, -G2), plus the
// status byte.  The records of all ranks are exchanged (all-gather straight into a device buffer) and bn254_finish_distinct_dev
// folds them; with sigs given on every rank no signature material has to travel at all.
int bn254_distinct_payload_dev(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* pks, const uint8_t* sigs, size_t n,
                               uint8_t* payload) {
  ENTER();
  ARGCHECK(payload && (n == 0 || (pks && (msgs || msg_len == 0))));
  CK(cudaMemsetAsync(payload, 0, BN_PAYLOAD, ctx->stream));
  if (!sigs) return distinct_partial_dev(ctx, msgs, msg_len, pks, n, payload, payload + 384);
  DALLOC(d_sum, 64);
  int rc = sum_dev_impl<fq>(ctx, sigs, nullptr, n, d_sum.as<uint8_t>(), payload + 386, ctx->input_policy == BN254_INPUTS_UNTRUSTED);
  if (rc) return rc;
  rc = distinct_partial_dev(ctx, msgs, msg_len, pks, n, payload, payload + 385, d_sum.as<uint8_t>());
  if (rc) return rc;
  LAUNCH(k_first_status, 1, 32, payload + 385, 2, payload + 384);
  return 0;
}
// payloads: m records of BN254_DISTINCT_PAYLOAD_BYTES (device); agg_sig: 64 bytes (device) or NULL when every record already
// holds its rank's signature pair.  *status (device) = first failing record's status, else the verdict.
static int finish_payloads_dev(bn254_ctx* ctx, const uint8_t* payloads, size_t m, const uint8_t* agg_sig, uint8_t* status) {
  if (ctx->pairing_mode == 1) {
    LAUNCH(k_distinct_finish, 1, 32, payloads, (int)m, agg_sig, ctx->d_lines, status);
    return 0;
  }
  // one item through the cooperative machine (one block, latency layout): Miller loop over the 87 scaled -G2 lines of agg_sig,
  // times the product of the records, final exponentiation, verdict
  DALLOC(LNf, sizeof(u4) * COOP_LINE_FQ * 2 * K_N_LINES * COOP_LANES);
  DALLOC(FIOf, sizeof(u4) * 6 * 2 * 2 * COOP_LANES);
  DALLOC(GSf, sizeof(u4) * COOP_GSLOTS * 6 * 2 * 2 * COOP_LANES);
  LAUNCH(k_finish_prepare, 1, BN_BLOCK, payloads, (int)m, agg_sig, ctx->d_lines, LNf.as<u4>(), FIOf.as<u4>(), status);
  if (ctx->coop12)
    (ctx->coop18 ? k_coop18_run : k_coop12_run)<<<1, ctx->coop18 ? 3 * COOP_THREADS : COOP12_THREADS, COOP12_SMEM_BYTES, ctx->stream>>>(CPROG_FINISH, (size_t)1, (size_t)COOP_LANES, LNf.as<u4>(), GSf.as<u4>(), FIOf.as<u4>(),
                                                                        status, (const unsigned*)nullptr);
  else
    k_coop_run<<<1, COOP_THREADS, COOP1_SMEM_BYTES, ctx->stream>>>(CPROG_FINISH, (size_t)1, (size_t)COOP_LANES, LNf.as<u4>(), GSf.as<u4>(), FIOf.as<u4>(), status,
                                                                  0u, (unsigned)ctx->sm_count, (const unsigned*)nullptr);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}
int bn254_finish_distinct_dev(bn254_ctx* ctx, const uint8_t* payloads, size_t n_payloads, const uint8_t* agg_sig, uint8_t* status) {
  ENTER();
  ARGCHECK(status && (n_payloads == 0 || payloads) && n_payloads <= 0x7fffff);
  return finish_payloads_dev(ctx, payloads, n_payloads, agg_sig, status);
}
int bn254_finish_distinct(bn254_ctx* ctx, const uint8_t* partials384, size_t n_partials, const uint8_t* agg_sig, uint8_t* status) {
  ENTER();
  ARGCHECK(status && agg_sig && (n_partials == 0 || partials384) && n_partials <= 0x7fffff);
  std::vector<uint8_t> host(BN_PAYLOAD * (n_partials ? n_partials : 1), 0);
  for (size_t i = 0; i < n_partials; i++) memcpy(&host[BN_PAYLOAD * i], partials384 + 384 * i, 384);
  DALLOC(d_p, host.size());
  DALLOC(d_sig, 64);
  DALLOC(d_st, 1);
  H2D(d_p.p, host.data(), host.size());
  H2D(d_sig.p, agg_sig, 64);
  int rc = finish_payloads_dev(ctx, d_p.as<uint8_t>(), n_partials, d_sig.as<uint8_t>(), d_st.as<uint8_t>());
  if (rc) {
    cudaStreamSynchronize(ctx->stream);  // `host` is the source of an enqueued copy
    return rc;
  }
  D2H(status, d_st.p, 1);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int bn254_aggregate_verify_distinct(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* pks, size_t n, const uint8_t* agg_sig,
                                    uint8_t* status) {
  ENTER();
  ARGCHECK(status && agg_sig && (n == 0 || (pks && (msgs || msg_len == 0))));
  DALLOC(d_msgs, msg_len * n + 1);
  DALLOC(d_pks, 128 * n + 1);
  DALLOC(d_sig, 64);
  DALLOC(d_pl, BN_PAYLOAD);
  DALLOC(d_st, 1);
  if (n && msg_len) H2D(d_msgs.p, msgs, msg_len * n);
  if (n) H2D(d_pks.p, pks, 128 * n);
  H2D(d_sig.p, agg_sig, 64);
  int rc = bn254_distinct_payload_dev(ctx, d_msgs.as<uint8_t>(), msg_len, d_pks.as<uint8_t>(), nullptr, n, d_pl.as<uint8_t>());
  if (rc) return rc;
  rc = finish_payloads_dev(ctx, d_pl.as<uint8_t>(), 1, d_sig.as<uint8_t>(), d_st.as<uint8_t>());
  if (rc) return rc;
  D2H(status, d_st.p, 1);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// format_pairing_check_values / format_pairing_check_uncompressed_values (/root/reference/src/utils.rs:197-239): per item
// [(H(msg) 64 B, pk 128 B), (sig 64 B, -G2 128 B)] = 384 bytes, every 32-byte coordinate LITTLE-endian (the dependency's Borsh
// form).  compressed != 0: sig is 33 bytes, pk 65 bytes, both decoded (and validated) like Signature / PublicKey::from_compressed;
// compressed == 0: 64 / 128 bytes, re-ordered without validation exactly like the reference (:218-239).
int bn254_format_pairing_check_batch(bn254_ctx* ctx, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                                     int compressed, uint8_t* out384, uint8_t* status) {
  ENTER();
  if (n == 0) return 0;
  ARGCHECK((msgs || msg_len == 0) && sigs && pks && out384 && status);
  const size_t sb = compressed ? 33 : 64, pb = compressed ? 65 : 128;
  DALLOC(d_msgs, msg_len * n + 1);
  DALLOC(d_sigs, sb * n);
  DALLOC(d_pks, pb * n);
  DALLOC(H, sizeof(g1aff) * n);
  DALLOC(d_out, 384 * n);
  DALLOC(d_st, n);
  if (msg_len) H2D(d_msgs.p, msgs, msg_len * n);
  H2D(d_sigs.p, sigs, sb * n);
  H2D(d_pks.p, pks, pb * n);
  int rc = hash_dev(ctx, d_msgs.as<uint8_t>(), msg_len, nullptr, n, H.as<g1aff>(), d_st.as<uint8_t>(), nullptr);
  if (rc) return rc;
  LAUNCH(k_format_pairing_check, grid_for(n), BN_BLOCK, H.as<g1aff>(), d_sigs.as<uint8_t>(), d_pks.as<uint8_t>(), n, compressed, d_out.as<uint8_t>(),
         d_st.as<uint8_t>());
  D2H(out384, d_out.p, 384 * n);
  D2H(status, d_st.p, n);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
