//! `bn254-b200`: the public API of the `bn254` crate (sedaprotocol/bn254, `src/lib.rs:60-63`) over the B200 batch engine.
//!
//! Same names and meaning as the reference: [`PrivateKey`], [`PublicKey`] (G2), [`PublicKeyG1`], [`Signature`] (G1),
//! [`ECDSA::sign`] / [`ECDSA::verify`], [`check_public_keys`], `+ - neg` aggregation, the 33 / 65-byte compressed and
//! 64 / 128-byte uncompressed encodings, the 11-variant [`Error`].  Values are held as the crate's own big-endian byte
//! encodings, so nothing here depends on `bn` (zeropool-bn): every arithmetic step is a call into `libbn254_b200.so`.
//! On top of the reference API there are the batch entry points the engine exists for ([`ECDSA::verify_batch`],
//! [`ECDSA::sign_batch`], [`ECDSA::verify_batch_randomized`], [`aggregate_verify_distinct`]).
//!
//! The raw declarations in `sys.rs` are generated from `include/bn254_b200.h` (`scripts/gen_rust_sys.py`) and a test of the
//! engine's repository fails when the two drift apart.
//! This crate is NOT built in the engine's own repository (its image has no Rust toolchain); the Python mirror
//! `bn254_b200/api.py` is what the parity tests drive.  Differences from the reference, all forced by the byte-level
//! representation: the inner field of the newtypes is a byte array instead of a `bn` struct, and the point at infinity
//! (which the reference cannot serialise, `src/utils.rs:86`) is the all-zero encoding.
mod sys;

use std::ops::{Add, Neg, Sub};
use std::os::raw::c_int;
use std::sync::Mutex;

/// `src/error.rs:5-29`, plus `Engine` for failures of the GPU library itself (never a per-item verdict).
#[derive(thiserror::Error, Debug)]
pub enum Error {
    #[error("errored to find a valid point while converting hash to point")]
    HashToPointError,
    #[error("errored to get data from an index out of bounds")]
    IndexOutOfBounds,
    #[error("errored to create group or field due to invalid input encoding")]
    InvalidEncoding,
    #[error("errored to map point to a curve")]
    InvalidGroupPoint,
    #[error("errored to create group or field due to invalid input length")]
    InvalidLength,
    #[error("errored to create a field element")]
    NotMemberError,
    #[error("errored to convert to affine coordinates")]
    ToAffineConversion,
    #[error("Point was already in affine coordinates (division-by-zero)")]
    PointInJacobian,
    #[error("Bn254 verification failed")]
    VerificationFailed,
    #[error("Serialization failed")]
    SerializationError,
    #[error(transparent)]
    HexDecodeFailed(#[from] hex::FromHexError),
    #[error("bn254_b200 engine error: {0}")]
    Engine(String),
}
pub type Result<T, E = Error> = core::result::Result<T, E>;

/// status byte of the C ABI -> `Result` (the codes follow the order of the enum above, 0 = Ok)
fn status_to_result(code: u8) -> Result<()> {
    match code {
        0 => Ok(()),
        1 => Err(Error::HashToPointError),
        2 => Err(Error::IndexOutOfBounds),
        3 => Err(Error::InvalidEncoding),
        4 => Err(Error::InvalidGroupPoint),
        5 => Err(Error::InvalidLength),
        6 => Err(Error::NotMemberError),
        7 => Err(Error::ToAffineConversion),
        8 => Err(Error::PointInJacobian),
        9 => Err(Error::VerificationFailed),
        10 => Err(Error::SerializationError),
        255 => Err(Error::Engine("item not evaluated (BN254_ENGINE_FAULT): retry the call".into())),
        other => Err(Error::Engine(format!("unknown status byte {other}"))),
    }
}

// ------------------------------------------------------------------------------------------------ engine context
struct Ctx(*mut sys::bn254_ctx);
unsafe impl Send for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::bn254_ctx_destroy(self.0) }
    }
}
/// One context (GPU `BN254_B200_DEVICE`, default 0) behind a mutex: calls on one context must not overlap.
static ENGINE: Mutex<Option<Ctx>> = Mutex::new(None);

fn engine_call<F: FnOnce(*mut sys::bn254_ctx) -> c_int>(f: F) -> Result<()> {
    let mut guard = ENGINE.lock().map_err(|_| Error::Engine("engine mutex poisoned".into()))?;
    if guard.is_none() {
        let device: c_int = std::env::var("BN254_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut raw: *mut sys::bn254_ctx = std::ptr::null_mut();
        let rc = unsafe { sys::bn254_ctx_create(device, &mut raw) };
        if rc != 0 {
            return Err(Error::Engine(last_error(std::ptr::null_mut())));
        }
        // Every point this crate hands to the engine is a value of its own types: made by a validating constructor
        // (from_compressed / from_uncompressed / from_private_key) or by + - on such values, so infinity is a legal value
        // and G2 membership is an invariant -- exactly the engine's BN254_INPUTS_TYPED policy (include/bn254_b200.h).
        unsafe { sys::bn254_set_input_policy(raw, sys::BN254_INPUTS_TYPED) };
        *guard = Some(Ctx(raw));
    }
    let ctx = guard.as_ref().unwrap().0;
    let rc = f(ctx);
    if rc != 0 {
        return Err(Error::Engine(last_error(ctx)));
    }
    Ok(())
}
fn last_error(ctx: *mut sys::bn254_ctx) -> String {
    unsafe {
        let p = sys::bn254_last_error(ctx);
        if p.is_null() {
            "unknown".into()
        } else {
            std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

// ------------------------------------------------------------------------------------------------ PrivateKey
/// `src/types.rs:13-77`.  Any 32 bytes are accepted and reduced mod r, as `Fr::from_slice` does; the field always holds
/// the CANONICAL big-endian scalar, so equal keys compare equal (the reference derives `PartialEq` on the `Fr` inside).
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub struct PrivateKey([u8; 32]);

const R_ORDER: [u8; 32] = [
    0x30, 0x64, 0x4e, 0x72, 0xe1, 0x31, 0xa0, 0x29, 0xb8, 0x50, 0x45, 0xb6, 0x81, 0x81, 0x58, 0x5d, 0x28, 0x33, 0xe8, 0x48, 0x79, 0xb9,
    0x70, 0x91, 0x43, 0xe1, 0xf5, 0x93, 0xf0, 0x00, 0x00, 0x01,
];
/// any 256-bit big-endian integer mod r (2^256 / r < 6: at most five subtractions)
fn reduce_mod_r(mut k: [u8; 32]) -> [u8; 32] {
    while k >= R_ORDER {
        let mut borrow = 0i16;
        for i in (0..32).rev() {
            let d = k[i] as i16 - R_ORDER[i] as i16 - borrow;
            borrow = (d < 0) as i16;
            k[i] = (d + 256 * borrow) as u8;
        }
    }
    k
}

impl PrivateKey {
    /// `src/types.rs:17-25`: uniform in [0, r) like `Fr::random` (rejection sampling on 254-bit draws: 2^254 / r < 1.33)
    pub fn random<R: rand::Rng>(rng: &mut R) -> Self {
        loop {
            let mut b = [0u8; 32];
            rng.fill_bytes(&mut b);
            b[0] &= 0x3f;
            if b < R_ORDER {
                return PrivateKey(b);
            }
        }
    }
    /// `src/types.rs:27-29`: canonical big-endian scalar
    pub fn to_bytes(&self) -> Result<Vec<u8>> {
        Ok(self.0.to_vec())
    }
    /// the canonical 32 bytes (what the engine is given)
    pub fn as_bytes(&self) -> &[u8; 32] {
        &self.0
    }
}
impl TryFrom<&[u8]> for PrivateKey {
    type Error = Error;
    fn try_from(b: &[u8]) -> Result<Self> {
        let a: [u8; 32] = b.try_into().map_err(|_| Error::InvalidLength)?; // src/types_test.rs:29-46
        Ok(PrivateKey(reduce_mod_r(a)))
    }
}
impl TryFrom<&str> for PrivateKey {
    type Error = Error;
    fn try_from(s: &str) -> Result<Self> {
        PrivateKey::try_from(hex::decode(s)?.as_slice())
    }
}
impl TryFrom<String> for PrivateKey {
    type Error = Error;
    fn try_from(s: String) -> Result<Self> {
        PrivateKey::try_from(s.as_str())
    }
}
impl TryFrom<PrivateKey> for String {
    type Error = Error;
    fn try_from(k: PrivateKey) -> Result<String> {
        Ok(hex::encode(k.to_bytes()?))
    }
}

// ------------------------------------------------------------------------------------------------ group elements
macro_rules! point_type {
    ($name:ident, $raw:expr, $comp:expr, $sum:path, $compress:path, $decompress:path, $validate:path, $doc:expr $(, $derive:ident)*) => {
        #[doc = $doc]
        #[derive(Copy, Clone, Debug $(, $derive)*)]
        pub struct $name(pub [u8; $raw]);

        impl $name {
            pub fn from_compressed<T: AsRef<[u8]>>(bytes: T) -> Result<Self> {
                let b = bytes.as_ref();
                if b.len() != $comp {
                    return Err(Error::InvalidEncoding); // bn::G1 / G2::from_compressed: CurveError::InvalidEncoding
                }
                let (mut out, mut st) = ([0u8; $raw], 0u8);
                engine_call(|c| unsafe { $decompress(c, b.as_ptr(), 1, out.as_mut_ptr(), &mut st) })?;
                status_to_result(st)?;
                Ok($name(out))
            }
            pub fn from_uncompressed<T: AsRef<[u8]>>(bytes: T) -> Result<Self> {
                let b = bytes.as_ref();
                let a: [u8; $raw] = b.try_into().map_err(|_| Error::InvalidLength)?;
                let mut st = 0u8;
                engine_call(|c| unsafe { $validate(c, a.as_ptr(), 1, &mut st) })?;
                status_to_result(st)?;
                Ok($name(a))
            }
            pub fn to_compressed(&self) -> Result<Vec<u8>> {
                let (mut out, mut st) = (vec![0u8; $comp], 0u8);
                engine_call(|c| unsafe { $compress(c, self.0.as_ptr(), 1, out.as_mut_ptr(), &mut st) })?;
                status_to_result(st)?; // infinity -> PointInJacobian, as src/utils.rs:86
                Ok(out)
            }
            pub fn to_uncompressed(&self) -> Result<Vec<u8>> {
                if self.0.iter().all(|&b| b == 0) {
                    return Err(Error::PointInJacobian);
                }
                Ok(self.0.to_vec())
            }
            /// sum of `(-1)^neg[i] * points[i]` in one engine call (the fold the `+` / `-` operators are made of)
            pub fn sum(points: &[$name], neg: Option<&[u8]>) -> Result<Self> {
                let flat: Vec<u8> = points.iter().flat_map(|p| p.0).collect();
                let (mut out, mut st) = ([0u8; $raw], 0u8);
                let negp = neg.map_or(std::ptr::null(), |n| n.as_ptr());
                engine_call(|c| unsafe { $sum(c, flat.as_ptr(), negp, points.len(), out.as_mut_ptr(), &mut st) })?;
                status_to_result(st)?;
                Ok($name(out))
            }
        }
        impl Add for $name {
            type Output = $name;
            fn add(self, other: $name) -> $name {
                $name::sum(&[self, other], None).expect("bn254_b200 engine")
            }
        }
        impl Sub for $name {
            type Output = $name;
            fn sub(self, other: $name) -> $name {
                $name::sum(&[self, other], Some(&[0, 1])).expect("bn254_b200 engine")
            }
        }
        impl Neg for $name {
            type Output = $name;
            fn neg(self) -> $name {
                $name::sum(&[self], Some(&[1])).expect("bn254_b200 engine")
            }
        }
    };
}
point_type!(PublicKey, 128, 65, sys::bn254_g2_sum, sys::bn254_g2_compress_batch, sys::bn254_g2_decompress_batch,
            sys::bn254_g2_validate_batch, "`src/types.rs:81-148`: a G2 point, `x.re || x.im || y.re || y.im` big-endian.", PartialEq, Eq);
point_type!(PublicKeyG1, 64, 33, sys::bn254_g1_sum, sys::bn254_g1_compress_batch, sys::bn254_g1_decompress_batch,
            sys::bn254_g1_validate_batch, "`src/types.rs:151-218`: a G1 point, `x || y` big-endian.", PartialEq, Eq);
// (the reference derives no PartialEq on Signature, src/types.rs:221)
point_type!(Signature, 64, 33, sys::bn254_g1_sum, sys::bn254_g1_compress_batch, sys::bn254_g1_decompress_batch,
            sys::bn254_g1_validate_batch, "`src/types.rs:221-286`: a G1 point, `x || y` big-endian.");

impl PublicKey {
    /// `src/types.rs:85-87`: `G2::one() * sk`
    pub fn from_private_key(private_key: &PrivateKey) -> Self {
        let mut out = [0u8; 128];
        engine_call(|c| unsafe { sys::bn254_derive_pk_g2_batch(c, private_key.as_bytes().as_ptr(), 1, out.as_mut_ptr()) }).expect("bn254_b200 engine");
        PublicKey(out)
    }
}
impl PublicKeyG1 {
    /// `src/types.rs:155-157`: `G1::one() * sk`
    pub fn from_private_key(private_key: &PrivateKey) -> Self {
        let mut out = [0u8; 64];
        engine_call(|c| unsafe { sys::bn254_derive_pk_g1_batch(c, private_key.as_bytes().as_ptr(), 1, out.as_mut_ptr()) }).expect("bn254_b200 engine");
        PublicKeyG1(out)
    }
}

// ------------------------------------------------------------------------------------------------ ECDSA
/// `src/ecdsa.rs:13`
pub struct ECDSA;

impl ECDSA {
    /// `src/ecdsa.rs:26-35`: `H(message) * sk`
    pub fn sign<T: AsRef<[u8]>>(message: T, private_key: &PrivateKey) -> Result<Signature> {
        let m = message.as_ref();
        let (mut sig, mut st) = ([0u8; 64], 0u8);
        engine_call(|c| unsafe { sys::bn254_sign_batch(c, m.as_ptr(), m.len(), private_key.as_bytes().as_ptr(), 1, sig.as_mut_ptr(), &mut st) })?;
        status_to_result(st)?;
        Ok(Signature(sig))
    }
    /// `src/ecdsa.rs:49-64`: `e(H(m), pk) * e(sig, -G2) == 1`, `Err(VerificationFailed)` otherwise
    pub fn verify<T: AsRef<[u8]>>(message: T, signature: &Signature, public_key: &PublicKey) -> Result<()> {
        let m = message.as_ref();
        let mut st = 0u8;
        engine_call(|c| unsafe { sys::bn254_verify_batch(c, m.as_ptr(), m.len(), signature.0.as_ptr(), public_key.0.as_ptr(), 1, &mut st) })?;
        status_to_result(st)
    }
    /// n messages of `msg_len` bytes each, one key per message -> n signatures (one GPU pass)
    pub fn sign_batch(msgs: &[u8], msg_len: usize, keys: &[PrivateKey]) -> Result<Vec<Result<Signature>>> {
        let n = keys.len();
        if msgs.len() != n * msg_len {
            return Err(Error::InvalidLength);
        }
        let sks: Vec<u8> = keys.iter().flat_map(|k| *k.as_bytes()).collect();
        let (mut sigs, mut st) = (vec![0u8; 64 * n], vec![0u8; n]);
        engine_call(|c| unsafe { sys::bn254_sign_batch(c, msgs.as_ptr(), msg_len, sks.as_ptr(), n, sigs.as_mut_ptr(), st.as_mut_ptr()) })?;
        Ok((0..n).map(|i| status_to_result(st[i]).map(|_| Signature(sigs[64 * i..64 * i + 64].try_into().unwrap()))).collect())
    }
    /// n independent (message, signature, key) triples -> n exact verdicts, each what [`ECDSA::verify`] returns
    pub fn verify_batch(msgs: &[u8], msg_len: usize, sigs: &[Signature], pks: &[PublicKey]) -> Result<Vec<Result<()>>> {
        let n = sigs.len();
        if pks.len() != n || msgs.len() != n * msg_len {
            return Err(Error::InvalidLength);
        }
        let s: Vec<u8> = sigs.iter().flat_map(|x| x.0).collect();
        let p: Vec<u8> = pks.iter().flat_map(|x| x.0).collect();
        let mut st = vec![0u8; n];
        engine_call(|c| unsafe { sys::bn254_verify_batch(c, msgs.as_ptr(), msg_len, s.as_ptr(), p.as_ptr(), n, st.as_mut_ptr()) })?;
        Ok(st.into_iter().map(status_to_result).collect())
    }
    /// Same verdicts as [`ECDSA::verify_batch`] (up to a 2^-128 false-accept probability), computed with one shared final
    /// exponentiation when the whole batch is valid; the engine draws the secret coefficients.  `keys_in_g2`: every key came
    /// from `from_compressed` / `from_uncompressed` (which check the r-torsion), so the engine need not test it again.
    pub fn verify_batch_randomized(msgs: &[u8], msg_len: usize, sigs: &[Signature], pks: &[PublicKey], keys_in_g2: bool) -> Result<Vec<Result<()>>> {
        let n = sigs.len();
        if pks.len() != n || msgs.len() != n * msg_len {
            return Err(Error::InvalidLength);
        }
        let s: Vec<u8> = sigs.iter().flat_map(|x| x.0).collect();
        let p: Vec<u8> = pks.iter().flat_map(|x| x.0).collect();
        let (mut st, mut fast) = (vec![0u8; n], 0 as c_int);
        engine_call(|c| unsafe {
            sys::bn254_verify_batch_rlc(c, msgs.as_ptr(), msg_len, s.as_ptr(), p.as_ptr(), n, std::ptr::null(), keys_in_g2 as c_int, st.as_mut_ptr(), &mut fast)
        })?;
        Ok(st.into_iter().map(status_to_result).collect())
    }
}

/// `src/ecdsa.rs:78-93`: `e(G1, pk_g2) * e(pk_g1, -G2) == 1`
pub fn check_public_keys(public_key_g2: &PublicKey, public_key_g1: &PublicKeyG1) -> Result<()> {
    let mut st = 0u8;
    engine_call(|c| unsafe { sys::bn254_check_public_keys_batch(c, public_key_g2.0.as_ptr(), public_key_g1.0.as_ptr(), 1, &mut st) })?;
    status_to_result(st)
}

/// The flow of `examples/bn254.rs:25-32` in one call: sum the signatures and the keys, verify the sums against `message`.
pub fn aggregate_verify_same_message<T: AsRef<[u8]>>(message: T, sigs: &[Signature], pks: &[PublicKey]) -> Result<()> {
    let m = message.as_ref();
    if sigs.len() != pks.len() {
        return Err(Error::InvalidLength);
    }
    let s: Vec<u8> = sigs.iter().flat_map(|x| x.0).collect();
    let p: Vec<u8> = pks.iter().flat_map(|x| x.0).collect();
    let mut st = 0u8;
    engine_call(|c| unsafe { sys::bn254_aggregate_verify_same_msg(c, m.as_ptr(), m.len(), s.as_ptr(), p.as_ptr(), sigs.len(), &mut st) })?;
    status_to_result(st)
}

/// Distinct-message aggregate verification: `prod_i e(H(m_i), pk_i) * e(agg_sig, -G2) == 1`, one final exponentiation.
pub fn aggregate_verify_distinct(msgs: &[u8], msg_len: usize, pks: &[PublicKey], agg_sig: &Signature) -> Result<()> {
    let n = pks.len();
    if msgs.len() != n * msg_len {
        return Err(Error::InvalidLength);
    }
    let p: Vec<u8> = pks.iter().flat_map(|x| x.0).collect();
    let mut st = 0u8;
    engine_call(|c| unsafe { sys::bn254_aggregate_verify_distinct(c, msgs.as_ptr(), msg_len, p.as_ptr(), n, agg_sig.0.as_ptr(), &mut st) })?;
    status_to_result(st)
}

/// `[(H(m), pk), (sig, -G2)]` with little-endian coordinates, the value `format_pairing_check_*` return (`src/utils.rs:197-239`)
pub type PairingCheckValues = [([u8; 64], [u8; 128]); 2];

fn format_values(message: &[u8], sig: &[u8], pk: &[u8], compressed: bool) -> Result<PairingCheckValues> {
    let (mut out, mut st) = ([0u8; 384], 0u8);
    engine_call(|c| unsafe {
        sys::bn254_format_pairing_check_batch(c, message.as_ptr(), message.len(), sig.as_ptr(), pk.as_ptr(), 1, compressed as c_int, out.as_mut_ptr(), &mut st)
    })?;
    status_to_result(st)?;
    Ok([(out[0..64].try_into().unwrap(), out[64..192].try_into().unwrap()), (out[192..256].try_into().unwrap(), out[256..384].try_into().unwrap())])
}
/// `src/utils.rs:197-216`: 33-byte compressed signature, 65-byte compressed public key
pub fn format_pairing_check_values(message: Vec<u8>, signature: Vec<u8>, public_key: Vec<u8>) -> Result<PairingCheckValues> {
    if public_key.len() != 65 || signature.len() != 33 {
        return Err(Error::InvalidEncoding); // from_compressed of either point
    }
    format_values(&message, &signature, &public_key, true)
}
/// `src/utils.rs:218-239`: 64 / 128-byte uncompressed inputs, re-ordered without validation like the reference; a wrong
/// length is the reference's `Vec<u8> -> [u8; N]` failure, `SerializationError` (`src/error.rs:64-68`)
pub fn format_pairing_check_uncompressed_values(message: Vec<u8>, signature: Vec<u8>, public_key: Vec<u8>) -> Result<PairingCheckValues> {
    if signature.len() != 64 || public_key.len() != 128 {
        return Err(Error::SerializationError);
    }
    format_values(&message, &signature, &public_key, false)
}

/// `hash_to_try_and_increment` (`src/hash.rs:29-63`) for n messages of `msg_len` bytes: the uncompressed G1 points
pub fn hash_to_g1_batch(msgs: &[u8], msg_len: usize) -> Result<Vec<Result<[u8; 64]>>> {
    if msg_len == 0 || msgs.len() % msg_len != 0 {
        return Err(Error::InvalidLength);
    }
    let n = msgs.len() / msg_len;
    let (mut out, mut st) = (vec![0u8; 64 * n], vec![0u8; n]);
    engine_call(|c| unsafe { sys::bn254_hash_to_g1_batch(c, msgs.as_ptr(), msg_len, n, out.as_mut_ptr(), st.as_mut_ptr()) })?;
    Ok((0..n).map(|i| status_to_result(st[i]).map(|_| out[64 * i..64 * i + 64].try_into().unwrap())).collect())
}

// ------------------------------------------------------------------------------------------------ serde (src/serde.rs:10-56)
#[cfg(feature = "serde")]
mod serde_impl {
    use super::{PrivateKey, PublicKey};
    use serde::de::Error as _;
    use serde::ser::Error as _;
    use serde::{Deserialize, Deserializer, Serialize, Serializer};

    impl Serialize for PrivateKey {
        fn serialize<S: Serializer>(&self, s: S) -> Result<S::Ok, S::Error> {
            let b = self.to_bytes().map_err(S::Error::custom)?; // sequence of 32 u8
            b.serialize(s)
        }
    }
    impl<'de> Deserialize<'de> for PrivateKey {
        fn deserialize<D: Deserializer<'de>>(d: D) -> Result<Self, D::Error> {
            let b = <[u8; 32]>::deserialize(d)?; // src/serde.rs:29
            PrivateKey::try_from(&b[..]).map_err(D::Error::custom)
        }
    }
    impl Serialize for PublicKey {
        fn serialize<S: Serializer>(&self, s: S) -> Result<S::Ok, S::Error> {
            let b = self.to_compressed().map_err(S::Error::custom)?; // sequence of 65 u8, compressed (src/serde.rs:39-44)
            b.serialize(s)
        }
    }
    impl<'de> Deserialize<'de> for PublicKey {
        fn deserialize<D: Deserializer<'de>>(d: D) -> Result<Self, D::Error> {
            let b = Vec::<u8>::deserialize(d)?;
            PublicKey::from_compressed(b).map_err(D::Error::custom) // src/serde.rs:53-54
        }
    }
}
