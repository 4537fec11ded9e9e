// hash.cuh -- SHA-256 (FIPS 180-4) and the try-and-increment hash to G1.
//
// Replaces /root/reference/src/hash.rs:29-63 (hash_to_try_and_increment), /root/reference/src/utils.rs:27-37
// (mod_u256, strict '>') and :56-63 (arbitrary_string_to_g1 = G1::from_compressed(0x02 || x)), plus the sha2
// crate (/root/reference/Cargo.toml:30).  For ctr = 0..254: h = SHA-256(msg || ctr) as a big-endian integer;
// skip if h >= 5q (/root/reference/src/hash.rs:11-14,49-51); x = h reduced by repeated subtraction while x > q;
// x == q fails Fq::from_slice; accept when x^3 + 3 is a square and take the even root.
#pragma once
#include "curve.cuh"

namespace bn {

BN_CONST uint32_t K_SHA256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

BN_FN uint32_t rotr32(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(x, x, n);
#else
  return (x >> n) | (x << (32 - n));
#endif
}

// one compression: state += F(state, block); the 16-word schedule window is updated in place
BN_NOINLINE void sha256_compress(uint32_t* state, const uint32_t* block) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = block[i];
  uint32_t a = state[0], b = state[1], c = state[2], d = state[3], e = state[4], f = state[5], g = state[6], h = state[7];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    uint32_t wi;
    if (i < 16) {
      wi = w[i];
    } else {
      uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
      uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
      w[i & 15] = wi;
    }
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = h + S1 + ch + K_SHA256[i] + wi;
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  state[0] += a; state[1] += b; state[2] += c; state[3] += d; state[4] += e; state[5] += f; state[6] += g; state[7] += h;
}

BN_FN void sha256_init(uint32_t* s) {
  s[0] = 0x6a09e667; s[1] = 0xbb67ae85; s[2] = 0x3c6ef372; s[3] = 0xa54ff53a;
  s[4] = 0x510e527f; s[5] = 0x9b05688c; s[6] = 0x1f83d9ab; s[7] = 0x5be0cd19;
}

// Hash `msg` to a G1 point (affine, Montgomery form).  Returns ST_OK or ST_HASH_TO_POINT; *tries_out (optional)
// receives the accepted counter.  max_tries is 255 (/root/reference/src/hash.rs:39: `for ctr in 0..255`); a smaller
// value exists only so that tests can reach the HashToPointError exit (/root/reference/src/hash.rs:62), which no
// real message does (2^-235).
// Counters ctr_first .. max_tries - 1 are tried (ctr_first > 0: the warp-parallel batch kernel, where lane l of a warp owns one counter).
BN_NOINLINE int hash_to_g1(fq* hx, fq* hy, const uint8_t* msg, uint64_t len, int* ctr_out, int max_tries = 255, int ctr_first = 0) {
  // midstate over the full 64-byte blocks of msg
  uint32_t mid[8], blk[16];
  sha256_init(mid);
  uint64_t off = 0;
  while (len - off >= 64) {
    for (int i = 0; i < 16; i++) {
      const uint8_t* p = msg + off + 4 * i;
      blk[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
    sha256_compress(mid, blk);
    off += 64;
  }
  // tail: msg[off..len) || ctr || 0x80 || 0* || BE64(bit length), in one or two blocks (ctr byte left zero here)
  uint32_t tail[32];
  for (int i = 0; i < 32; i++) tail[i] = 0;
  int rem = (int)(len - off);
  for (int i = 0; i < rem; i++) tail[i >> 2] |= (uint32_t)msg[off + i] << (24 - 8 * (i & 3));
  int pos80 = rem + 1;
  tail[pos80 >> 2] |= 0x80u << (24 - 8 * (pos80 & 3));
  int nblk = (rem + 2 + 8 <= 64) ? 1 : 2;
  uint64_t bits = (len + 1) * 8;
  tail[nblk * 16 - 2] = (uint32_t)(bits >> 32);
  tail[nblk * 16 - 1] = (uint32_t)bits;
  const int cw = rem >> 2, cs = 24 - 8 * (rem & 3);
  const uint32_t base_word = tail[cw];

  for (int ctr = ctr_first; ctr < max_tries; ctr++) {
    uint32_t st[8];
    for (int i = 0; i < 8; i++) st[i] = mid[i];
    tail[cw] = base_word | ((uint32_t)ctr << cs);
    sha256_compress(st, tail);
    if (nblk == 2) sha256_compress(st, tail + 16);
    // digest as a big-endian integer -> little-endian limbs
    fq x;
    for (int i = 0; i < 8; i++) x.l[i] = st[7 - i];
    if (u256_geq(x.l, K_FIVE_Q)) continue;
    // mod_u256: while x > q { x -= q }   (strict: x == q stays and then fails the membership test)
    for (int k = 0; k < 4; k++) {
      uint32_t t[8];
      bool gt = !u256_geq(K_Q, x.l);  // x > q
      if (gt) {
        u256_sub(t, x.l, K_Q);
        for (int i = 0; i < 8; i++) x.l[i] = t[i];
      }
    }
    if (u256_geq(x.l, K_Q)) continue;  // Fq::from_slice rejects x >= q (only x == q can reach here)
    fq xm = fq_to_mont(x);
    fq t = fq_add(fq_mul(fq_sqr(xm), xm), fq_from_limbs(K_THREE));
    fq y;
    if (!fq_sqrt(&y, t)) continue;
    if (fq_parity(y)) y = fq_neg(y);  // sign byte 0x02: even y
    *hx = xm;
    *hy = y;
    if (ctr_out) *ctr_out = ctr;
    return ST_OK;
  }
  return ST_HASH_TO_POINT;
}

// One try of the loop above for a message that shares a single SHA-256 block with its counter byte and the padding
// (len <= 54): used by the compacting batch kernels, where every surviving item of round r tries counter r.
BN_NOINLINE bool hash_try_1blk(fq* hx, fq* hy, const uint8_t* msg, uint32_t len, uint32_t ctr) {
  uint32_t blk[16], st[8];
#pragma unroll
  for (int i = 0; i < 16; i++) blk[i] = 0;
  for (uint32_t i = 0; i < len; i++) blk[i >> 2] |= (uint32_t)msg[i] << (24 - 8 * (i & 3));
  blk[len >> 2] |= ctr << (24 - 8 * (len & 3));
  blk[(len + 1) >> 2] |= 0x80u << (24 - 8 * ((len + 1) & 3));
  blk[15] = (len + 1) * 8;
  sha256_init(st);
  sha256_compress(st, blk);
  fq x;
  for (int i = 0; i < 8; i++) x.l[i] = st[7 - i];
  if (u256_geq(x.l, K_FIVE_Q)) return false;
  for (int k = 0; k < 4; k++) {  // mod_u256: while x > q { x -= q }
    uint32_t t[8];
    bool gt = !u256_geq(K_Q, x.l);
    if (gt) {
      u256_sub(t, x.l, K_Q);
      for (int i = 0; i < 8; i++) x.l[i] = t[i];
    }
  }
  if (u256_geq(x.l, K_Q)) return false;
  fq xm = fq_to_mont(x);
  fq t = fq_add(fq_mul(fq_sqr(xm), xm), fq_from_limbs(K_THREE));
  fq y;
  if (!fq_sqrt(&y, t)) return false;
  if (fq_parity(y)) y = fq_neg(y);
  *hx = xm;
  *hy = y;
  return true;
}

}  // namespace bn
