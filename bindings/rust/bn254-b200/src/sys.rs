//! Raw declarations: one per line of `include/bn254_b200.h` that the safe layer uses (same shapes, plain pointers and sizes).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct bn254_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn bn254_ctx_create(device: c_int, out: *mut *mut bn254_ctx) -> c_int;
    pub fn bn254_ctx_destroy(ctx: *mut bn254_ctx);
    pub fn bn254_last_error(ctx: *mut bn254_ctx) -> *const c_char;
    pub fn bn254_hash_to_g1_var(ctx: *mut bn254_ctx, msgs: *const u8, offsets: *const u64, n: usize, g1_out: *mut u8,
                                status: *mut u8, tries_out: *mut u8) -> c_int;
    pub fn bn254_sign_batch(ctx: *mut bn254_ctx, msgs: *const u8, msg_len: usize, sks: *const u8, n: usize, sigs: *mut u8,
                            status: *mut u8) -> c_int;
    pub fn bn254_verify_batch(ctx: *mut bn254_ctx, msgs: *const u8, msg_len: usize, sigs: *const u8, pks: *const u8, n: usize,
                              status: *mut u8) -> c_int;
    pub fn bn254_verify_batch_rlc(ctx: *mut bn254_ctx, msgs: *const u8, msg_len: usize, sigs: *const u8, pks: *const u8,
                                  n: usize, coeffs16: *const u8, flags: c_int, status: *mut u8, took_fast_path: *mut c_int) -> c_int;
    pub fn bn254_check_public_keys_batch(ctx: *mut bn254_ctx, pk_g2: *const u8, pk_g1: *const u8, n: usize, status: *mut u8) -> c_int;
    pub fn bn254_g1_sum(ctx: *mut bn254_ctx, pts: *const u8, neg: *const u8, n: usize, out64: *mut u8, status: *mut u8) -> c_int;
    pub fn bn254_g2_sum(ctx: *mut bn254_ctx, pts: *const u8, neg: *const u8, n: usize, out128: *mut u8, status: *mut u8) -> c_int;
    pub fn bn254_derive_pk_g2_batch(ctx: *mut bn254_ctx, sks: *const u8, n: usize, out128: *mut u8) -> c_int;
    pub fn bn254_derive_pk_g1_batch(ctx: *mut bn254_ctx, sks: *const u8, n: usize, out64: *mut u8) -> c_int;
    pub fn bn254_g1_compress_batch(ctx: *mut bn254_ctx, raw64: *const u8, n: usize, out33: *mut u8, status: *mut u8) -> c_int;
    pub fn bn254_g1_decompress_batch(ctx: *mut bn254_ctx, in33: *const u8, n: usize, out64: *mut u8, status: *mut u8) -> c_int;
    pub fn bn254_g2_compress_batch(ctx: *mut bn254_ctx, raw128: *const u8, n: usize, out65: *mut u8, status: *mut u8) -> c_int;
    pub fn bn254_g2_decompress_batch(ctx: *mut bn254_ctx, in65: *const u8, n: usize, out128: *mut u8, status: *mut u8) -> c_int;
    pub fn bn254_g1_validate_batch(ctx: *mut bn254_ctx, raw64: *const u8, n: usize, status: *mut u8) -> c_int;
    pub fn bn254_g2_validate_batch(ctx: *mut bn254_ctx, raw128: *const u8, n: usize, status: *mut u8) -> c_int;
    pub fn bn254_aggregate_verify_same_msg(ctx: *mut bn254_ctx, msg: *const u8, msg_len: usize, sigs: *const u8, pks: *const u8,
                                           n: usize, status: *mut u8) -> c_int;
    pub fn bn254_aggregate_verify_distinct(ctx: *mut bn254_ctx, msgs: *const u8, msg_len: usize, pks: *const u8, n: usize,
                                           agg_sig: *const u8, status: *mut u8) -> c_int;
}
