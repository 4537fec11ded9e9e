"""ctypes binding of oracle/libbn254_oracle.so (the CPU checker).  Test infrastructure only."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "libbn254_oracle.so")

OK, HASH_TO_POINT, INDEX_OOB, INVALID_ENCODING, INVALID_GROUP_POINT, INVALID_LENGTH = 0, 1, 2, 3, 4, 5
NOT_MEMBER, TO_AFFINE, POINT_IN_JACOBIAN, VERIFICATION_FAILED, SERIALIZATION, HEX_DECODE = 6, 7, 8, 9, 10, 11

_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "bn254_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        _lib = ctypes.CDLL(_SO)
    return _lib


def _buf(n):
    return ctypes.create_string_buffer(n)


def hash_to_g1(msg):
    out, ctr = _buf(64), ctypes.c_int(-1)
    st = lib().bn254o_hash_to_g1(bytes(msg), ctypes.c_size_t(len(msg)), out, ctypes.byref(ctr))
    return st, out.raw, ctr.value


def sign(msg, sk):
    out = _buf(64)
    st = lib().bn254o_sign(bytes(msg), ctypes.c_size_t(len(msg)), bytes(sk), out)
    return st, out.raw


def verify(msg, sig, pk):
    return lib().bn254o_verify(bytes(msg), ctypes.c_size_t(len(msg)), bytes(sig), bytes(pk))


def check_public_keys(pk_g2, pk_g1):
    return lib().bn254o_check_public_keys(bytes(pk_g2), bytes(pk_g1))


def pairing_check(g1s, g2s, k):
    gt = _buf(384)
    st = lib().bn254o_pairing_check(bytes(g1s), bytes(g2s), ctypes.c_size_t(k), gt)
    return st, gt.raw


def miller_product(g1s, g2s, k):
    f = _buf(384)
    st = lib().bn254o_miller_product(bytes(g1s), bytes(g2s), ctypes.c_size_t(k), f)
    return st, f.raw


def final_exp(f):
    gt = _buf(384)
    st = lib().bn254o_final_exp(bytes(f), gt)
    return st, gt.raw


def fq12_op(op, a, b=None):
    out = _buf(384)
    st = lib().bn254o_fq12_op(op, bytes(a), bytes(b) if b is not None else None, out)
    return st, out.raw


def fq_op(op, a, b=None):
    out = _buf(32)
    st = lib().bn254o_fq_op(op, bytes(a), bytes(b) if b is not None else bytes(32), out)
    return st, out.raw


def _pt(fn, n, *args):
    out = _buf(n)
    st = fn(*args, out)
    return st, out.raw


def g1_add(a, b):
    return _pt(lib().bn254o_g1_add, 64, bytes(a), bytes(b))


def g1_neg(a):
    return _pt(lib().bn254o_g1_neg, 64, bytes(a))


def g1_mul(a, k):
    return _pt(lib().bn254o_g1_mul, 64, bytes(a), bytes(k))


def g2_add(a, b):
    return _pt(lib().bn254o_g2_add, 128, bytes(a), bytes(b))


def g2_neg(a):
    return _pt(lib().bn254o_g2_neg, 128, bytes(a))


def g2_mul(a, k):
    return _pt(lib().bn254o_g2_mul, 128, bytes(a), bytes(k))


def g1_sum(pts, n):
    return _pt(lib().bn254o_g1_sum, 64, bytes(pts), ctypes.c_size_t(n))


def g2_sum(pts, n):
    return _pt(lib().bn254o_g2_sum, 128, bytes(pts), ctypes.c_size_t(n))


def derive_pk_g2(sk):
    return _pt(lib().bn254o_derive_pk_g2, 128, bytes(sk))


def derive_pk_g1(sk):
    return _pt(lib().bn254o_derive_pk_g1, 64, bytes(sk))


def sk_canonical(b):
    out = _buf(32)
    st = lib().bn254o_sk_canonical(bytes(b), ctypes.c_size_t(len(b)), out)
    return st, out.raw


def g1_compress(raw):
    return _pt(lib().bn254o_g1_compress, 33, bytes(raw))


def g1_decompress(b):
    out = _buf(64)
    st = lib().bn254o_g1_decompress(bytes(b), ctypes.c_size_t(len(b)), out)
    return st, out.raw


def g2_compress(raw):
    return _pt(lib().bn254o_g2_compress, 65, bytes(raw))


def g2_decompress(b):
    out = _buf(128)
    st = lib().bn254o_g2_decompress(bytes(b), ctypes.c_size_t(len(b)), out)
    return st, out.raw


def g1_validate_uncompressed(b):
    return lib().bn254o_g1_validate_uncompressed(bytes(b), ctypes.c_size_t(len(b)))


def g2_validate_uncompressed(b):
    return lib().bn254o_g2_validate_uncompressed(bytes(b), ctypes.c_size_t(len(b)))


# ---- threaded batch drivers (CPU baseline legs of bench.py and bulk expected values in tests)
def verify_batch(msgs, msg_len, sigs, pks, n, nthreads=1):
    st = _buf(n)
    lib().bn254o_verify_batch(bytes(msgs), ctypes.c_size_t(msg_len), bytes(sigs), bytes(pks), ctypes.c_size_t(n), st, nthreads)
    return st.raw


def sign_batch(msgs, msg_len, sks, n, nthreads=1):
    out, st = _buf(64 * n), _buf(n)
    lib().bn254o_sign_batch(bytes(msgs), ctypes.c_size_t(msg_len), bytes(sks), ctypes.c_size_t(n), out, st, nthreads)
    return out.raw, st.raw


def hash_to_g1_batch(msgs, msg_len, n, nthreads=1):
    out, st = _buf(64 * n), _buf(n)
    lib().bn254o_hash_to_g1_batch(bytes(msgs), ctypes.c_size_t(msg_len), ctypes.c_size_t(n), out, st, nthreads)
    return out.raw, st.raw


def derive_pk_g2_batch(sks, n, nthreads=1):
    out = _buf(128 * n)
    lib().bn254o_derive_pk_g2_batch(bytes(sks), ctypes.c_size_t(n), out, nthreads)
    return out.raw


def derive_pk_g1_batch(sks, n, nthreads=1):
    out = _buf(64 * n)
    lib().bn254o_derive_pk_g1_batch(bytes(sks), ctypes.c_size_t(n), out, nthreads)
    return out.raw
