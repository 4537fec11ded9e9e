/*
 * bn254_b200.h -- C ABI of the B200-native BN254 batch engine (libbn254_b200.so).
 *
 * This is the drop-in boundary for the sign / aggregate / pairing-verify hot path of sedaprotocol/bn254.
 * The reference has no FFI of its own: its boundary is the Rust public API re-exported at
 * /root/reference/src/lib.rs:60-63.  Each entry point below names the reference item it replaces; a thin Rust
 * wrapper keeps the crate's types (PrivateKey / PublicKey / PublicKeyG1 / Signature, ECDSA::sign / verify,
 * + - aggregation) and converts through the crate's own big-endian byte formats.
 *
 * Conventions
 *   - Every integer is 32-byte big-endian canonical.  A G1 point is x||y (64 B, the crate's uncompressed form,
 *     /root/reference/src/utils.rs:182-194); a G2 point is x.re||x.im||y.re||y.im (128 B, :162-179).  The point
 *     at infinity, which the crate cannot serialise (/root/reference/src/utils.rs:86), is all-zero bytes.
 *   - Per-item results are status bytes: one code per variant of the crate's Error enum
 *     (/root/reference/src/error.rs:6-29), 0 = Ok(()).  A failed verification is BN254_VERIFICATION_FAILED,
 *     exactly as Err(Error::VerificationFailed) at /root/reference/src/ecdsa.rs:62.
 *   - Functions return 0 on success or a negative BN254_E_* engine error (CUDA failure, bad argument); the
 *     message is available from bn254_last_error().  There is no CPU fallback: without a CUDA device
 *     bn254_ctx_create fails.
 *   - verify / check_public_keys / aggregate entry points validate their point inputs by default (BN254_INPUTS_UNTRUSTED below);
 *     callers that hold values of the crate's types switch the context to BN254_INPUTS_TYPED.
 *   - Entry points without a suffix take HOST buffers and include the host<->device copies; the *_dev variants
 *     take DEVICE pointers (4-byte aligned), enqueue on the context's stream and return without synchronising
 *     (call bn254_sync).  A context is bound to one GPU; calls on one context must not overlap.
 */
#ifndef BN254_B200_H
#define BN254_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-item status = Error variant (/root/reference/src/error.rs:6-29) */
enum {
  BN254_OK = 0,
  BN254_HASH_TO_POINT_ERROR = 1,
  BN254_INDEX_OUT_OF_BOUNDS = 2,
  BN254_INVALID_ENCODING = 3,
  BN254_INVALID_GROUP_POINT = 4,
  BN254_INVALID_LENGTH = 5,
  BN254_NOT_MEMBER_ERROR = 6,
  BN254_TO_AFFINE_CONVERSION = 7,
  BN254_POINT_IN_JACOBIAN = 8,
  BN254_VERIFICATION_FAILED = 9,
  BN254_SERIALIZATION_ERROR = 10,
  BN254_HEX_DECODE_FAILED = 11,
  /* not an Error variant: the engine did NOT evaluate the item (the producer / machine handshake of the pipelined small-batch
   * verify timed out -- never observed; it exists so that such an item is reported instead of guessed).  Retry the call. */
  BN254_ENGINE_FAULT = 255
};

/* engine-level errors (function return values) */
enum { BN254_E_CUDA = -1, BN254_E_ARG = -2, BN254_E_NOMEM = -3 };

/* How the G1 / G2 byte inputs of verify, check_public_keys and the aggregate checks are read (bn254_set_input_policy).
 *   BN254_INPUTS_UNTRUSTED (default): bytes from outside.  Every signature / key is decoded exactly as
 *     Signature::from_uncompressed / PublicKey::from_uncompressed decode it (/root/reference/src/utils.rs:107-127, which end in
 *     AffineG1::new / AffineG2::new): field membership -> NotMemberError, curve equation -> InvalidGroupPoint -- all-zero bytes,
 *     the engine's encoding of infinity, are (0, 0) and fail here -- and, for G2, the r-torsion test -> InvalidGroupPoint.
 *     pk is decoded first, then sig; a decode error takes precedence over a hash error.
 *   BN254_INPUTS_TYPED: the bytes are values of the crate's types (made by its constructors and + - operators): all-zero bytes are
 *     the point at infinity, which bn::pairing_batch skips (so sig = infinity with pk = infinity VERIFIES, as in the crate), and G2
 *     membership is a type invariant that is not re-checked.  The Rust wrapper uses this; never feed network bytes this way.
 * The sum / pairing / codec building blocks always take typed values. */
enum { BN254_INPUTS_UNTRUSTED = 0, BN254_INPUTS_TYPED = 1 };

/* size of one rank's record of a distinct-message aggregate check (bn254_distinct_payload_dev): bytes [0, 384) the Miller
 * product (12 x 32 B big-endian, tower order), byte 384 the status, the rest reserved */
#define BN254_DISTINCT_PAYLOAD_BYTES 448

typedef struct bn254_ctx bn254_ctx;

/* lifecycle: one context per GPU; owns a stream, the -G2 line table, comb tables and the workspace */
int bn254_ctx_create(int device, bn254_ctx** out);
void bn254_ctx_destroy(bn254_ctx* ctx);
const char* bn254_last_error(bn254_ctx* ctx); /* ctx may be NULL: error of the last failed create */
int bn254_sync(bn254_ctx* ctx);
void* bn254_stream(bn254_ctx* ctx);            /* the cudaStream_t used by every launch of this context */
int bn254_sm_count(bn254_ctx* ctx);
uint64_t bn254_launch_count(bn254_ctx* ctx);   /* kernels launched by this context so far */
/* Temporaries (the 27 GB line-set workspace of a 2^19-item verify chunk included) come from a pool private to the context and stay
 * cached there between calls; bn254_trim returns the cached memory to the driver (bn254_ctx_destroy does too).  If the device
 * cannot hold the workspace, verify halves its chunk until it fits. */
int bn254_trim(bn254_ctx* ctx);
int bn254_set_input_policy(bn254_ctx* ctx, int policy);
int bn254_get_input_policy(bn254_ctx* ctx);
/* test hook: number of counters hash_to_try_and_increment tries (255 in the reference, /root/reference/src/hash.rs:39); lowering
 * it is the only way to reach HashToPointError (/root/reference/src/hash.rs:62), whose natural probability is 2^-235 */
int bn254_set_hash_try_limit(bn254_ctx* ctx, int max_tries);
/* test hook for BN254_ENGINE_FAULT: the line producer of a pipelined small-batch verify never publishes item `item` ((size_t)-1 =
 * off), so the machine's bounded wait for it times out (~1 s) and the item gets status 255.  The host-buffer entry points
 * (bn254_verify_batch, bn254_check_public_keys_batch) then run the batch again without pipelining and return real statuses;
 * *retries (may be NULL) = how many calls on this context did that.  The _dev entry points leave the 255 for the caller. */
int bn254_set_test_fault(bn254_ctx* ctx, size_t item, uint32_t* retries);

/* measurement support: when on, the verify pipeline brackets its three kernels (hash, Miller loop, final
 * exponentiation) with CUDA events on the context's stream; bn254_phase_ms returns and clears the accumulated times */
/* the verify pipeline's three phases: hash / Miller / final exponentiation with mode 1, hash / line sets / cooperative
 * Miller + final exponentiation with mode 0 */
int bn254_set_profiling(bn254_ctx* ctx, int on);
int bn254_phase_ms(bn254_ctx* ctx, float* out3);
/* pairing kernels used by verify / check_public_keys: 0 (default) = cooperative machine (csrc/coop.cuh), six warps share the
 * Fq12 values of a 32-item group, four groups per 24-warp block so that each group owns one SM sub-partition; 1 = one thread
 * per item (csrc/pairing.cuh); 2 = cooperative, one group per six-warp block; 3 = cooperative, warp-local layout (six lanes
 * share the Fq12 value of one item, five items per warp); 4 = cooperative, half-warp layout (three warps per 16-item group,
 * two coefficients per warp, two groups per sub-partition).  All give identical verdicts (tests compare them); 2 - 4 are
 * the measured design alternatives of DESIGN.md 4.3, slower than the default. */
int bn254_set_pairing_mode(bn254_ctx* ctx, int mode);

/* hash_to_try_and_increment (/root/reference/src/hash.rs:29-63): n messages of msg_len bytes each -> G1 */
int bn254_hash_to_g1_batch(bn254_ctx*, const uint8_t* msgs, size_t msg_len, size_t n, uint8_t* g1_out, uint8_t* status);
int bn254_hash_to_g1_batch_dev(bn254_ctx*, const uint8_t* msgs, size_t msg_len, size_t n, uint8_t* g1_out, uint8_t* status);
/* ragged messages: message i = msgs[offsets[i] .. offsets[i+1])  (offsets has n+1 entries) */
int bn254_hash_to_g1_var(bn254_ctx*, const uint8_t* msgs, const uint64_t* offsets, size_t n, uint8_t* g1_out, uint8_t* status,
                         uint8_t* tries_out /* optional: accepted counter per message */);

/* ECDSA::sign (/root/reference/src/ecdsa.rs:26-35): sig_i = H(msg_i) * sk_i ; sk is any 32 bytes, reduced mod r
 * like Fr::from_slice (/root/reference/src/types.rs:37) */
int bn254_sign_batch(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* sks, size_t n, uint8_t* sigs, uint8_t* status);
int bn254_sign_batch_dev(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* sks, size_t n, uint8_t* sigs, uint8_t* status);

/* ECDSA::verify (/root/reference/src/ecdsa.rs:49-64): status_i = verdict of e(H(msg_i), pk_i) * e(sig_i, -G2) == 1 */
int bn254_verify_batch(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n, uint8_t* status);
int bn254_verify_batch_dev(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n, uint8_t* status);

/* ECDSA::verify for keys that verify many messages (a fixed validator set) -- ADDITIONAL entry points with verify_batch's statuses.
 * The walk of a key along the twist (the 87 line coefficient triples of the Miller loop) does not depend on the message:
 * bn254_key_lines_prepare_dev does it once per key into a caller-owned device buffer of bn254_key_lines_bytes(n_keys) bytes
 * (16 704 bytes per key, 16-byte aligned) and records each key's decode status under the context's input policy
 * (key_status, n_keys bytes); bn254_verify_batch_cached_dev then verifies triple i against key key_index[i] (uint32, device;
 * NULL = key i) and per item only scales the cached lines by H(msg_i) and the -G2 lines by sig_i.  status[i] = the key's decode
 * status, the signature's, the hash's or the verdict, with the precedence of verify_batch; an index >= n_keys gives
 * BN254_INDEX_OUT_OF_BOUNDS. */
size_t bn254_key_lines_bytes(size_t n_keys);
int bn254_key_lines_prepare_dev(bn254_ctx*, const uint8_t* pks, size_t n_keys, uint8_t* key_lines, uint8_t* key_status);
int bn254_verify_batch_cached_dev(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* key_lines,
                                  const uint8_t* key_status, size_t n_keys, const uint32_t* key_index, size_t n, uint8_t* status);

/* Randomised batch form of ECDSA::verify -- an ADDITIONAL entry point, never used by verify_batch: the n
 * triples are accepted together iff  prod_i e(c_i H(msg_i), pk_i) * e(sum_i c_i sig_i, -G2) == 1  for 128-bit coefficients
 * c_i (coeffs16: n x 16 bytes, secret from whoever produced the signatures; NULL in the host-buffer form = drawn from
 * /dev/urandom; the bytes are read as two 64-bit halves and c_i = lo + hi * lambda mod r, lambda the GLV eigenvalue, which
 * keeps 2^128 distinct coefficients and halves the scalar multiplications).  One pass gives a verdict per SLICE (a 2^20-triple
 * chunk is cut into up to 64 slices of whole multi-pairing groups, each with its own product, signature sum and final
 * exponentiation): triples of a passing slice get status 0 (what verify_batch returns, up to a 2^-128 false-accept
 * probability); slices that fail, or hold an item that does not decode or a pk outside G2, are packed together and redone by
 * the exact per-item path, so status[] is always verify_batch's.  *took_fast_path = 1 iff no slice had to be redone.
 * flags bit 0: the caller vouches that every pk is in the r-torsion (e.g. it came from from_compressed) -- skips that test. */
int bn254_verify_batch_rlc(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                           const uint8_t* coeffs16, int flags, uint8_t* status, int* took_fast_path);
int bn254_verify_batch_rlc_dev(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                               const uint8_t* coeffs16, int flags, uint8_t* status, int* took_fast_path);

/* check_public_keys (/root/reference/src/ecdsa.rs:78-93): e(G1, pk_g2_i) * e(pk_g1_i, -G2) == 1 */
int bn254_check_public_keys_batch(bn254_ctx*, const uint8_t* pk_g2, const uint8_t* pk_g1, size_t n, uint8_t* status);

/* bn::pairing_batch(...) == Gt::one() for n independent products of k pairs each (pairs with an infinity are skipped) */
int bn254_pairing_check_batch(bn254_ctx*, const uint8_t* g1s, const uint8_t* g2s, size_t k, size_t n, uint8_t* status);
int bn254_pairing_check_batch_dev(bn254_ctx*, const uint8_t* g1s, const uint8_t* g2s, size_t k, size_t n, uint8_t* status);

/* Add / Sub / Neg folds (/root/reference/src/types.rs:126-148,196-218,264-286): out = sum of n points
 * (optionally sum of (-1)^neg[i] * P_i when neg != NULL).  *status = first decode error or 0; an infinite sum is all-zero. */
int bn254_g1_sum(bn254_ctx*, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out64, uint8_t* status);
int bn254_g2_sum(bn254_ctx*, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out128, uint8_t* status);
int bn254_g1_sum_dev(bn254_ctx*, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out64, uint8_t* status);
int bn254_g2_sum_dev(bn254_ctx*, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out128, uint8_t* status);

/* PublicKey::from_private_key / PublicKeyG1::from_private_key (/root/reference/src/types.rs:85-87,155-157) */
int bn254_derive_pk_g2_batch(bn254_ctx*, const uint8_t* sks, size_t n, uint8_t* out128);
int bn254_derive_pk_g1_batch(bn254_ctx*, const uint8_t* sks, size_t n, uint8_t* out64);

/* generic G1 / G2 scalar multiplication by a 256-bit big-endian integer (bn256.json `mul` semantics) */
int bn254_g1_mul_batch(bn254_ctx*, const uint8_t* pts, const uint8_t* scalars, size_t n, uint8_t* out64, uint8_t* status);
int bn254_g2_mul_batch(bn254_ctx*, const uint8_t* pts, const uint8_t* scalars, size_t n, uint8_t* out128, uint8_t* status);

/* codecs (/root/reference/src/utils.rs:84-194, bn::G1/G2::from_compressed): 33-byte G1 / 65-byte G2 compressed forms */
int bn254_g1_compress_batch(bn254_ctx*, const uint8_t* raw64, size_t n, uint8_t* out33, uint8_t* status);
int bn254_g1_decompress_batch(bn254_ctx*, const uint8_t* in33, size_t n, uint8_t* out64, uint8_t* status);
int bn254_g2_compress_batch(bn254_ctx*, const uint8_t* raw128, size_t n, uint8_t* out65, uint8_t* status);
int bn254_g2_decompress_batch(bn254_ctx*, const uint8_t* in65, size_t n, uint8_t* out128, uint8_t* status);
/* from_uncompressed validation (/root/reference/src/utils.rs:107-127): membership, curve, and for G2 the r-torsion check */
int bn254_g1_validate_batch(bn254_ctx*, const uint8_t* raw64, size_t n, uint8_t* status);
int bn254_g2_validate_batch(bn254_ctx*, const uint8_t* raw128, size_t n, uint8_t* status);

/* same-message aggregate verify (the flow of /root/reference/examples/bn254.rs:25-32): sum n sigs in G1 and n pks in G2,
 * then one verify of (msg, sum_sig, sum_pk) */
int bn254_aggregate_verify_same_msg(bn254_ctx*, const uint8_t* msg, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                                    uint8_t* status);
int bn254_aggregate_verify_same_msg_dev(bn254_ctx*, const uint8_t* msg, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                                        uint8_t* status);
/* distinct-message aggregate verify: prod_i e(H(msg_i), pk_i) * e(agg_sig, -G2) == 1, one shared final exponentiation */
int bn254_aggregate_verify_distinct(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* pks, size_t n, const uint8_t* agg_sig,
                                    uint8_t* status);
/* multi-GPU building blocks of the above: each GPU reduces its slice to one Fq12 Miller product (12 x 32 B big-endian,
 * tower order c0.c0.re .. c1.c2.im), the partials are exchanged (all-gather) and one rank finishes */
int bn254_miller_partial_distinct(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* pks, size_t n, uint8_t* f_out384,
                                  uint8_t* status);
int bn254_miller_partial_distinct_dev(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* pks, size_t n, uint8_t* f_out384,
                                      uint8_t* status);
int bn254_finish_distinct(bn254_ctx*, const uint8_t* partials384, size_t n_partials, const uint8_t* agg_sig, uint8_t* status);
/* the same exchange without leaving the device: every rank reduces its slice to ONE record of BN254_DISTINCT_PAYLOAD_BYTES
 * (with sigs != NULL the pair (sum of the rank's signatures, -G2) is folded into the rank's Miller product, so no signature
 * has to travel: prod_r e(S_r, -G2) = e(sum_r S_r, -G2)); the records are all-gathered into one device buffer and
 * bn254_finish_distinct_dev multiplies them, adds the pair (agg_sig, -G2) when agg_sig != NULL, and runs the single final
 * exponentiation through the cooperative machine.  *status (device, 1 byte) = first failing record's status, else the verdict. */
int bn254_distinct_payload_dev(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* pks, const uint8_t* sigs, size_t n,
                               uint8_t* payload);
int bn254_finish_distinct_dev(bn254_ctx*, const uint8_t* payloads, size_t n_payloads, const uint8_t* agg_sig, uint8_t* status);

/* format_pairing_check_values (compressed != 0: 33-byte sig, 65-byte pk, decoded like from_compressed) and
 * format_pairing_check_uncompressed_values (compressed == 0: 64 / 128 bytes, re-ordered without validation, as the reference does)
 * (/root/reference/src/utils.rs:197-239): per item [(H(msg), pk), (sig, -G2)] = 64 + 128 + 64 + 128 bytes, every 32-byte
 * coordinate little-endian (the dependency's Borsh form).  Errors in the reference's order: hash, public key, signature. */
int bn254_format_pairing_check_batch(bn254_ctx*, const uint8_t* msgs, size_t msg_len, const uint8_t* sigs, const uint8_t* pks, size_t n,
                                     int compressed, uint8_t* out384, uint8_t* status);

/* building blocks exposed for parity tests and profiling of the two pairing phases */
int bn254_miller_loop_batch(bn254_ctx*, const uint8_t* g1s, const uint8_t* g2s, size_t k, size_t n, uint8_t* f_out384, uint8_t* status);
int bn254_final_exp_batch(bn254_ctx*, const uint8_t* f_in384, size_t n, uint8_t* gt_out384, uint8_t* status);
/* Fq self-test hook: op 0 mul, 1 add, 2 sub, 3 inv, 4 sqrt (status 6 for a non-residue), 5 mul (portable cross-check variant),
 * 6 / 7: 9 a + b / 9 a - b, 8 / 9: 3 a + 2 b / 3 a - 2 b through the one-reduction routines the cooperative machine uses for
 * xi * x and for the tail of the cyclotomic squaring */
int bn254_fq_op_batch(bn254_ctx*, int op, const uint8_t* a32, const uint8_t* b32, size_t n, uint8_t* out32, uint8_t* status);
/* Fq12 self-test hook: op 0 mul, 1 sqr, 2 inv, 3 cyclotomic sqr, 4..6 frobenius 1..3, 7 conj */
int bn254_fq12_op_batch(bn254_ctx*, int op, const uint8_t* a384, const uint8_t* b384, size_t n, uint8_t* out384, uint8_t* status);

/* layer hook for the parity tests: n_in Fq values in, n_out Fq values out per item; ops listed at debug_layer_op in
 * bn254_b200/csrc/items.cuh (Fq2 product / square / scale, Miller doubling and addition steps, sparse line product, G1 mixed add) */
int bn254_layer_op_batch(bn254_ctx*, int op, const uint8_t* in, size_t n_in, size_t n, uint8_t* out, size_t n_out);

#ifdef __cplusplus
}
#endif
#endif /* BN254_B200_H */
