"""Pins the CPU oracle (oracle/bn254_oracle.c) against every golden vector the reference's own tests hold
(tests/golden/reference_vectors.json, extracted by tests/golden/make_golden.py) and cross-checks it against
the independent big-int oracle (oracle/pyoracle.py).  CPU only."""
import json
import os
import random
import sys

import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as P  # noqa: E402

G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
H = bytes.fromhex
G1_GEN = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")


def g1_raw(p):
    return bytes(64) if p is None else p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")


def g2_raw(p):
    return bytes(128) if p is None else b"".join(c.to_bytes(32, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))


def G2_GEN():
    return g2_raw(P.G2_GEN)


def test_hash_to_g1_kats():  # src/hash_test.rs:9-30
    for v in G["hash_to_g1"]:
        st, raw, _ = O.hash_to_g1(H(v["msg"]))
        assert st == 0
        assert O.g1_compress(raw) == (0, H(v["compressed"]))


def test_sign_kat_and_verify():  # src/ecdsa_test.rs:5-38
    for v in G["sign"]:
        st, sig = O.sign(H(v["msg"]), H(v["sk"]))
        assert st == 0 and O.g1_compress(sig) == (0, H(v["sig_compressed"]))
    for v in G["verify_ok"]:
        st, sig = O.g1_decompress(H(v["sig_compressed"]))
        assert st == 0
        _, pk = O.derive_pk_g2(H(v["sk"]))
        assert O.verify(H(v["msg"]), sig, pk) == 0


def test_aggregate_verify():  # src/ecdsa_test.rs:41-79, examples/bn254.rs
    for v in G["aggregate_verify_ok"] + [G["example"]]:
        msg = H(v["msg"])
        sigs = [O.sign(msg, H(sk))[1] for sk in v["sks"]]
        pks = [O.derive_pk_g2(H(sk))[1] for sk in v["sks"]]
        _, asig = O.g1_add(sigs[0], sigs[1])
        _, apk = O.g2_add(pks[0], pks[1])
        assert O.verify(msg, asig, apk) == 0
        assert O.verify(msg, sigs[0], apk) == O.VERIFICATION_FAILED
    # derived vectors of SURVEY.md 8c for the example (keys > r, reduced mod r)
    v = G["example"]
    sigs = [O.sign(H(v["msg"]), H(sk))[1] for sk in v["sks"]]
    pks = [O.derive_pk_g2(H(sk))[1] for sk in v["sks"]]
    assert O.g1_compress(O.g1_add(*sigs)[1])[1].hex() == "020aea1d3390792923eb34734b1669efad8a388997061c42fb9f7d310eec339ada"
    assert O.g2_compress(O.g2_add(*pks)[1])[1].hex() == (
        "0a08685b8899122e2d7da466f39d7698b3918cf3e7ae2a2d8a3cbb6713ee4238c7e3fbfc289492c9c6b70f01a86182add29428b2ce4a6ec6ba67603ee22d7670ec")


def test_check_public_keys():  # src/ecdsa_test.rs:82-112
    for v in G["check_public_keys_ok"]:
        assert O.check_public_keys(O.derive_pk_g2(H(v["sk"]))[1], O.derive_pk_g1(H(v["sk"]))[1]) == 0
    for v in G["check_public_keys_fail"]:
        assert O.check_public_keys(O.derive_pk_g2(H(v["sk_g2"]))[1], O.derive_pk_g1(H(v["sk_g1"]))[1]) == O.VERIFICATION_FAILED


def test_uncompressed_roundtrips():  # src/ecdsa_test.rs:115-154
    for v in G["pk_g1_uncompressed_roundtrip"]:
        _, pk1 = O.derive_pk_g1(H(v["sk"]))
        assert O.g1_validate_uncompressed(pk1) == 0
        assert O.check_public_keys(O.derive_pk_g2(H(v["sk"]))[1], pk1) == 0
    for v in G["sig_uncompressed_roundtrip"]:
        st, sig = O.g1_decompress(H(v["sig_compressed"]))
        assert st == 0 and O.g1_validate_uncompressed(sig) == 0
        assert O.verify(H(v["msg"]), sig, O.derive_pk_g2(H(v["sk"]))[1]) == 0


def test_private_key():  # src/types_test.rs:14-46
    for v in G["private_key_roundtrip"]:
        assert O.sk_canonical(H(v["sk"])) == (0, H(v["sk"]))
    for v in G["private_key_invalid_length"]:
        assert O.sk_canonical(H(v["bytes"]))[0] == O.INVALID_LENGTH
    big = (P.R + 5).to_bytes(32, "big")  # Fr::from_slice reduces, never rejects (examples/bn254.rs:8,12)
    assert O.sk_canonical(big) == (0, (5).to_bytes(32, "big"))


def test_g2_codecs():  # src/types_test.rs:48-69
    for v in G["g2_compressed_roundtrip"]:
        st, raw = O.g2_decompress(H(v["compressed"]))
        assert st == 0 and O.g2_compress(raw) == (0, H(v["compressed"]))
    for v in G["g2_uncompressed_roundtrip"]:
        assert O.g2_validate_uncompressed(H(v["uncompressed"])) == 0
        c = O.g2_compress(H(v["uncompressed"]))
        assert c[0] == 0 and O.g2_decompress(c[1]) == (0, H(v["uncompressed"]))


def test_sk_to_pk_g2():  # src/types_test.rs:71-129
    for v in G["sk_to_pk_g2"]:
        assert O.derive_pk_g2(H(v["sk"])) == (0, H(v["pk_uncompressed"]))


def test_gen_plus_gen():  # src/types_test.rs:132-159
    assert O.g2_compress(O.g2_add(G2_GEN(), G2_GEN())[1]) == (0, H(G["g2_gen_plus_gen_compressed"]))
    assert O.g1_compress(O.g1_add(G1_GEN, G1_GEN)[1]) == (0, H(G["g1_gen_plus_gen_compressed"]))


def test_bn256_json():  # src/bn256.json (EVM precompile vectors; zero bytes = infinity)
    for v in G["bn256_json"]["add"]:
        st, r = O.g1_add(H(v["x1"]) + H(v["y1"]), H(v["x2"]) + H(v["y2"]))
        assert st == 0 and r == H(v["result"])
    for v in G["bn256_json"]["mul"]:
        st, r = O.g1_mul(H(v["x"]) + H(v["y"]), H(v["scalar"]))
        assert st == 0 and r == H(v["result"])


def test_cross_check_python_oracle_groups():
    rng = random.Random(7)
    for _ in range(20):
        k1, k2 = rng.randrange(1, P.R), rng.randrange(1, P.R)
        a, b = P.g1_mul(P.G1_GEN, k1), P.g1_mul(P.G1_GEN, k2)
        assert O.g1_mul(G1_GEN, k1.to_bytes(32, "big"))[1] == g1_raw(a)
        assert O.g1_add(g1_raw(a), g1_raw(b))[1] == g1_raw(P.g1_add(a, b))
        assert O.g1_compress(g1_raw(a))[1] == P.g1_to_compressed(a)
    for _ in range(4):
        k1, k2 = rng.randrange(1, P.R), rng.randrange(1, P.R)
        a, b = P.g2_mul(P.G2_GEN, k1), P.g2_mul(P.G2_GEN, k2)
        assert O.g2_mul(G2_GEN(), k1.to_bytes(32, "big"))[1] == g2_raw(a)
        assert O.g2_add(g2_raw(a), g2_raw(b))[1] == g2_raw(P.g2_add(a, b))
        c = P.g2_to_compressed(a)
        assert O.g2_compress(g2_raw(a))[1] == c and O.g2_decompress(c) == (0, g2_raw(a))


def test_cross_check_python_oracle_hash():
    rng = random.Random(11)
    for _ in range(200):
        msg = rng.randbytes(rng.choice([0, 1, 31, 32, 33, 55, 56, 63, 64, 65, 119, 120, 200]))
        p, ctr = P.hash_to_try_and_increment(msg, want_counter=True)
        st, raw, c = O.hash_to_g1(msg)
        assert st == 0 and raw == g1_raw(p) and c == ctr


def test_cross_check_python_oracle_pairing_verdicts():
    rng = random.Random(13)
    sk = rng.randrange(1, P.R)
    msg = b"cross-check"
    sig, pk = P.sign(msg, sk), P.pk_g2_from_sk(sk)
    assert P.verify(msg, sig, pk) == 0 == O.verify(msg, g1_raw(sig), g2_raw(pk))
    bad = P.g1_add(sig, P.G1_GEN)
    assert P.verify(msg, bad, pk) == O.VERIFICATION_FAILED == O.verify(msg, g1_raw(bad), g2_raw(pk))


def test_infinity_semantics():  # pairs holding an infinity are skipped
    msg = b"inf"
    _, pk = O.derive_pk_g2((5).to_bytes(32, "big"))
    assert O.verify(msg, bytes(64), bytes(128)) == 0  # both pairs skipped -> Gt::one
    assert O.verify(msg, bytes(64), pk) == O.VERIFICATION_FAILED
    assert O.g1_compress(bytes(64))[0] == O.POINT_IN_JACOBIAN


def test_decode_rejections():
    q = P.Q
    assert O.g1_decompress(b"\x02" + q.to_bytes(32, "big"))[0] == O.NOT_MEMBER  # x >= q
    assert O.g1_decompress(b"\x04" + (1).to_bytes(32, "big"))[0] == O.INVALID_ENCODING
    assert O.g1_decompress(b"\x02" + (1).to_bytes(31, "big"))[0] == O.INVALID_ENCODING
    assert O.g1_validate_uncompressed((1).to_bytes(32, "big") + (3).to_bytes(32, "big")) == O.INVALID_GROUP_POINT
    assert O.g1_validate_uncompressed(bytes(63)) == O.INVALID_LENGTH
    good = H(G["g2_compressed_roundtrip"][0]["compressed"])
    assert O.g2_decompress(b"\x0c" + good[1:])[0] == O.INVALID_ENCODING  # bad sign byte (after the sqrt, as upstream)
    assert O.g2_decompress(good[:-1])[0] == O.INVALID_ENCODING
    # a point on the twist outside the r-torsion is rejected by the subgroup check
    x = 1
    while True:
        y = P.f2_sqrt(P.f2_add(P.f2_mul(P.f2_mul((x, 0), (x, 0)), (x, 0)), P.B2))
        if y is not None and not P.g2_in_subgroup(((x, 0), y)):
            break
        x += 1
    assert O.g2_validate_uncompressed(g2_raw(((x, 0), y))) == O.INVALID_GROUP_POINT


def _rand_twist_point(rng):
    while True:
        x = (rng.randrange(P.Q), rng.randrange(P.Q))
        y2 = P.f2_add(P.f2_mul(P.f2_mul(x, x), x), P.B2)
        y = P.f2_sqrt(y2)
        if y is not None and P.f2_mul(y, y) == y2:
            return (x, y)


def _is_prime(n):  # deterministic enough: 24 Miller-Rabin bases
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d, s = d // 2, s + 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def test_g2_subgroup_psi_criterion():
    """The engine's r-torsion test (one 63-bit scalar multiplication and the psi endomorphism, csrc/items.cuh) decides
    exactly what [r]P == infinity decides (AffineG2::new).  E'(Fq2) has order r * h2 with h2 squarefree,
    so it is the direct sum of its prime-order parts, psi acts on each part as a scalar, and the criterion's polynomial in
    psi is zero on the r part (the generator passes) -- it is exact iff it is non-zero on each of the four other parts,
    which one point of each prime order shows."""
    rng = random.Random(5)
    h2 = 2 * P.Q - P.R
    fs = P.TWIST_COFACTOR_FACTORS
    prod = 1
    for f in fs:
        assert _is_prime(f)
        prod *= f
    assert prod == h2 and len(set(fs)) == len(fs) and h2 % P.R != 0
    n_twist = P.R * h2
    assert P.g2_mul(_rand_twist_point(rng), n_twist) is None          # the group order of the twist
    assert P.g2_psi(P.G2_GEN) == P.g2_mul(P.G2_GEN, P.Q % P.R)        # psi has eigenvalue q on G2
    assert P.g2_in_subgroup_psi(P.G2_GEN) and P.g2_in_subgroup_psi(P.g2_mul(P.G2_GEN, rng.randrange(1, P.R)))
    for f in fs:
        s = None
        while s is None:
            s = P.g2_mul(_rand_twist_point(rng), n_twist // f)
        assert P.g2_mul(s, f) is None                                 # order exactly f (prime)
        assert not P.g2_in_subgroup_psi(s) and not P.g2_in_subgroup(s)
        mixed = P.g2_add(s, P.g2_mul(P.G2_GEN, rng.randrange(1, P.R)))
        assert not P.g2_in_subgroup_psi(mixed) and not P.g2_in_subgroup(mixed)
    for _ in range(3):                                                # and a few random points, both ways
        t = _rand_twist_point(rng)
        assert P.g2_in_subgroup_psi(t) == P.g2_in_subgroup(t) is False
        c = P.g2_mul(t, h2)
        assert P.g2_in_subgroup_psi(c) == P.g2_in_subgroup(c) is True
