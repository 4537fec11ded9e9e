"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle on the same seeded inputs,
against the reference's golden vectors, and -- at full BASELINE sizes -- through size-independent properties.
Bit-exact everywhere (integer arithmetic).  Run on a B200: `python -m pytest tests -m gpu`."""
import json
import os
import random

import numpy as np
import pytest

import oracle_lib as O
import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
H = bytes.fromhex
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
G1_GEN = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
NTHREADS = os.cpu_count() or 1


def be(x, n=32):
    return x.to_bytes(n, "big")


@pytest.fixture(scope="module")
def E():
    import bn254_b200
    from bn254_b200 import build, engine
    build.build()
    engine.context(0)  # raises without a GPU / the CUDA library: no fallback
    # The oracle restates the crate's functions over values of its TYPES (infinity is a value, pairing_batch skips it), so the
    # module's default context runs under the typed input policy; the untrusted (default) policy has its own tests below.
    engine.set_input_policy(engine.INPUTS_TYPED)
    return engine


# ---------------------------------------------------------------------------------------------- field layers
def test_fq_ptx_product_matches_portable_and_oracle(E):
    n = 1 << 18
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    b = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x1F  # < 2^253 < q
    b[:, 0] &= 0x1F
    edge = [0, 1, 2, Q - 1, Q - 2, (1 << 253) - 1, (Q - 1) // 2]
    for i, x in enumerate(edge):
        for j, y in enumerate(edge):
            a[i * len(edge) + j] = np.frombuffer(be(x), dtype=np.uint8)
            b[i * len(edge) + j] = np.frombuffer(be(y), dtype=np.uint8)
    ab, bb = a.tobytes(), b.tobytes()
    ptx, st = E.fq_op_batch(0, ab, bb)
    port, st2 = E.fq_op_batch(5, ab, bb)
    assert not any(st) and not any(st2)
    assert ptx == port  # PTX even/odd IMAD.WIDE chains == portable CIOS, 2^18 pairs
    for i in list(range(len(edge) ** 2)) + list(range(1000, 1000 + 4096)):
        x, y = ab[32 * i:32 * i + 32], bb[32 * i:32 * i + 32]
        assert ptx[32 * i:32 * i + 32] == O.fq_op(0, x, y)[1]
    for op in (1, 2):
        r, st = E.fq_op_batch(op, ab[:32 * 8192], bb[:32 * 8192])
        for i in range(0, 8192, 3):
            assert r[32 * i:32 * i + 32] == O.fq_op(op, ab[32 * i:32 * i + 32], bb[32 * i:32 * i + 32])[1]
    for op in (3, 4):
        r, st = E.fq_op_batch(op, ab[:32 * 512])
        for i in range(512):
            est, e = O.fq_op(op, ab[32 * i:32 * i + 32])
            assert st[i] == est and (est != 0 or r[32 * i:32 * i + 32] == e)
    # x >= q is rejected like Fq::from_slice
    r, st = E.fq_op_batch(0, be(Q), be(1))
    assert st[0] == O.NOT_MEMBER
    # 9 a +- b with one reduction (quotient estimate + table, fq.cuh fq_mul9_add): the routine sees Montgomery-form integers, so
    # the operands are chosen there: sums that land on and around every multiple of q, the extremes, and random pairs
    RINV = pow(1 << 256, -1, Q)
    prng = random.Random(41)
    pairs = [(0, 0), (Q - 1, Q - 1), (Q - 1, 0), (0, Q - 1), (1, Q - 9), (1, Q - 10)]
    for k in range(1, 10):
        for d in (-2, -1, 0, 1, 2):
            t = k * Q + d
            x = min(Q - 1, t // 9)
            if 0 <= t - 9 * x < Q:
                pairs.append((x, t - 9 * x))
    pairs += [(prng.randrange(Q), prng.randrange(Q)) for _ in range(4096)]
    xa = b"".join(be(x * RINV % Q) for x, _ in pairs)
    za = b"".join(be(z * RINV % Q) for _, z in pairs)
    plus, st = E.fq_op_batch(6, xa, za)
    minus, st2 = E.fq_op_batch(7, xa, za)
    assert not any(st) and not any(st2)
    for i, (x, z) in enumerate(pairs):
        assert plus[32 * i:32 * i + 32] == be((9 * x + z) % Q * RINV % Q), i
        assert minus[32 * i:32 * i + 32] == be((9 * x - z) % Q * RINV % Q), i
    for k in range(1, 5):                       # 3 t + 2 z = k q + d for the tail of the cyclotomic squaring
        for d in (-2, -1, 0, 1, 2):
            v = k * Q + d
            t = min(Q - 1, v // 3)
            if (v - 3 * t) % 2 == 0 and 0 <= (v - 3 * t) // 2 < Q:
                pairs.append((t, (v - 3 * t) // 2))
    xa = b"".join(be(x * RINV % Q) for x, _ in pairs)
    za = b"".join(be(z * RINV % Q) for _, z in pairs)
    plus, st = E.fq_op_batch(8, xa, za)
    minus, st2 = E.fq_op_batch(9, xa, za)
    assert not any(st) and not any(st2)
    for i, (x, z) in enumerate(pairs):
        assert plus[32 * i:32 * i + 32] == be((3 * x + 2 * z) % Q * RINV % Q), i
        assert minus[32 * i:32 * i + 32] == be((3 * x - 2 * z) % Q * RINV % Q), i


def test_fq12_ops(E):
    rng = random.Random(2)
    n = 64
    a = b"".join(be(rng.randrange(Q)) for _ in range(12 * n))
    b = b"".join(be(rng.randrange(Q)) for _ in range(12 * n))
    for op in range(8):
        r, st = E.fq12_op_batch(op, a, b)
        assert not any(st)
        for i in range(n):
            assert r[384 * i:384 * (i + 1)] == O.fq12_op(op, a[384 * i:384 * (i + 1)], b[384 * i:384 * (i + 1)])[1], (op, i)


# ---------------------------------------------------------------------------------------------- hash to G1
def test_hash_to_g1_kats(E):  # /root/reference/src/hash_test.rs:9-30
    for v in G["hash_to_g1"]:
        m = H(v["msg"])
        out, st = E.hash_to_g1_batch(m, len(m), 1)
        assert st[0] == 0
        c, cst = E.g1_compress_batch(out)
        assert cst[0] == 0 and c == H(v["compressed"])


def test_hash_to_g1_random_fixed_len(E):
    n = 8192
    msgs = synth.messages(n, 32, seed=1)
    out, st = E.hash_to_g1_batch(msgs, 32, n)
    eout, est = O.hash_to_g1_batch(msgs, 32, n, NTHREADS)
    assert st == est and out == eout


def test_hash_to_g1_compacting_path_lengths(E):
    """Batches of >= 4096 single-block messages go through the compacting rounds (k_hash_round / k_hash_tail); smaller or
    longer ones through the per-thread loop.  Both must give the oracle's points, at the block-boundary lengths too."""
    for msg_len, n in ((0, 4096), (1, 4100), (31, 5000), (54, 4097), (55, 4096), (64, 4096), (32, 4095)):
        msgs = synth.messages(n, msg_len, seed=100 + msg_len) if msg_len else b""
        out, st = E.hash_to_g1_batch(msgs, msg_len, n)
        eout, est = O.hash_to_g1_batch(msgs, msg_len, n, NTHREADS)
        assert st == est and out == eout, (msg_len, n)


def test_hash_to_g1_ragged(E):
    rng = random.Random(3)
    lens = [0, 1, 3, 31, 32, 33, 54, 55, 56, 62, 63, 64, 65, 118, 119, 120, 127, 128, 129, 300, 1000]
    msgs = [rng.randbytes(rng.choice(lens)) for _ in range(600)] + [b""]
    out, st, tries = E.hash_to_g1_var(msgs)
    for i, m in enumerate(msgs):
        est, e, ectr = O.hash_to_g1(m)
        assert (st[i], out[64 * i:64 * i + 64], tries[i]) == (est, e, ectr), i
    assert max(tries) >= 3  # the retry loop really ran


# ---------------------------------------------------------------------------------------------- sign / keys
def test_sign_kat(E):  # /root/reference/src/ecdsa_test.rs:5-17
    for v in G["sign"]:
        m = H(v["msg"])
        sig, st = E.sign_batch(m, len(m), H(v["sk"]))
        assert st[0] == 0
        assert E.g1_compress_batch(sig)[0] == H(v["sig_compressed"])


def test_sign_random(E):
    n = 2048
    msgs, sks = synth.messages(n, 32, seed=11), synth.rand_bytes(12, 32 * n)  # arbitrary 32-byte keys, many >= r
    sig, st = E.sign_batch(msgs, 32, sks)
    esig, est = O.sign_batch(msgs, 32, sks, n, NTHREADS)
    assert st == est and sig == esig


def test_sk_to_pk_kats(E):  # /root/reference/src/types_test.rs:71-129 + the example's keys (> r)
    for v in G["sk_to_pk_g2"]:
        assert E.derive_pk_g2_batch(H(v["sk"])) == H(v["pk_uncompressed"])
    sks = b"".join(H(s) for s in G["example"]["sks"]) + synth.rand_bytes(21, 32 * 254)
    n = len(sks) // 32
    assert E.derive_pk_g2_batch(sks) == O.derive_pk_g2_batch(sks, n, NTHREADS)
    assert E.derive_pk_g1_batch(sks) == O.derive_pk_g1_batch(sks, n, NTHREADS)


def test_bn256_json(E):  # /root/reference/src/bn256.json
    for v in G["bn256_json"]["add"]:
        r, st = E.g1_sum(H(v["x1"]) + H(v["y1"]) + H(v["x2"]) + H(v["y2"]))
        assert st == 0 and r == H(v["result"])
    pts = b"".join(H(v["x"]) + H(v["y"]) for v in G["bn256_json"]["mul"])
    ks = b"".join(H(v["scalar"]) for v in G["bn256_json"]["mul"])
    r, st = E.g1_mul_batch(pts, ks)
    assert not any(st)
    for i, v in enumerate(G["bn256_json"]["mul"]):
        assert r[64 * i:64 * i + 64] == H(v["result"])


# ---------------------------------------------------------------------------------------------- verify
def _signed_set(E, n, seed):
    msgs, sks = synth.messages(n, 32, seed=seed), synth.secret_keys(n, seed=seed + 1)
    sigs, st = E.sign_batch(msgs, 32, sks)
    assert not any(st)
    pks = E.derive_pk_g2_batch(sks)
    return msgs, sks, sigs, pks


def test_verify_batch_with_adversarial_items(E):
    n = 1024
    msgs, sks, sigs, pks = _signed_set(E, n, seed=31)
    msgs, sigs, pks = bytearray(msgs), bytearray(sigs), bytearray(pks)
    rng = random.Random(7)
    kinds = {}
    for i in rng.sample(range(n), 200):
        kind = rng.randrange(8)
        kinds[i] = kind
        if kind == 0:  # wrong message
            msgs[32 * i] ^= 1
        elif kind == 1:  # negated signature
            sigs[64 * i:64 * i + 64] = O.g1_neg(bytes(sigs[64 * i:64 * i + 64]))[1]
        elif kind == 2:  # key of another signer
            j = (i + 1) % n
            pks[128 * i:128 * i + 128] = pks[128 * j:128 * j + 128]
        elif kind == 3:  # sig + G1
            sigs[64 * i:64 * i + 64] = O.g1_add(bytes(sigs[64 * i:64 * i + 64]), G1_GEN)[1]
        elif kind == 4:  # signature at infinity, real key -> reject
            sigs[64 * i:64 * i + 64] = bytes(64)
        elif kind == 5:  # both at infinity -> both pairs skipped -> accept (dependency semantics)
            sigs[64 * i:64 * i + 64] = bytes(64)
            pks[128 * i:128 * i + 128] = bytes(128)
        elif kind == 6:  # off-curve signature
            sigs[64 * i + 63] ^= 1
        elif kind == 7:  # coordinate >= q
            sigs[64 * i:64 * i + 32] = be(Q + 5)
    msgs, sigs, pks = bytes(msgs), bytes(sigs), bytes(pks)
    st = E.verify_batch(msgs, 32, sigs, pks)
    est = O.verify_batch(msgs, 32, sigs, pks, n, NTHREADS)
    assert st == est
    assert sum(1 for s in st if s == 0) >= n - 200 and O.VERIFICATION_FAILED in st and O.INVALID_GROUP_POINT in st and O.NOT_MEMBER in st
    for i, k in kinds.items():
        assert (st[i] == 0) == (k == 5), (i, k)


def test_miller_and_final_exp_layers(E):
    n = 64
    msgs, sks, sigs, pks = _signed_set(E, n, seed=41)
    hs, _ = E.hash_to_g1_batch(msgs, 32, n)
    neg_g2 = O.g2_neg(O.derive_pk_g2(be(1))[1])[1]
    g1s = b"".join(hs[64 * i:64 * i + 64] + sigs[64 * i:64 * i + 64] for i in range(n))
    g2s = b"".join(pks[128 * i:128 * i + 128] + neg_g2 for i in range(n))
    f, st = E.miller_loop_batch(g1s, g2s, 2, n)
    assert not any(st)
    for i in range(n):
        assert f[384 * i:384 * i + 384] == O.miller_product(g1s[128 * i:128 * i + 128], g2s[256 * i:256 * i + 256], 2)[1]
    gt, st = E.final_exp_batch(f)
    one = be(1) + bytes(352)
    for i in range(n):
        assert st[i] == 0 and gt[384 * i:384 * i + 384] == one
    rnd = b"".join(be(random.Random(i).randrange(Q)) for i in range(12 * 16))
    gt, st = E.final_exp_batch(rnd)
    for i in range(16):
        assert gt[384 * i:384 * i + 384] == O.final_exp(rnd[384 * i:384 * i + 384])[1]
    assert E.pairing_check_batch(g1s, g2s, 2, n) == bytes(n)
    # a product with a skipped (infinite) pair and k = 0
    assert E.pairing_check_batch(g1s[:128] + bytes(64), g2s[:256] + neg_g2, 3, 1) == b"\x00"
    assert E.pairing_check_batch(b"", b"", 0, 3) == bytes(3)


def test_reference_api_vectors(E):
    """The reference's own tests, restated against the API mirror (src/ecdsa_test.rs, src/types_test.rs, examples/bn254.rs)."""
    from bn254_b200 import ECDSA, Error, PrivateKey, PublicKey, PublicKeyG1, Signature, check_public_keys
    v = G["verify_ok"][0]
    sk = PrivateKey(H(v["sk"]))
    pk = PublicKey.from_private_key(sk)
    sig = Signature.from_compressed(H(v["sig_compressed"]))
    assert ECDSA.verify(H(v["msg"]), sig, pk) is None
    assert ECDSA.sign(H(v["msg"]), sk).to_compressed() == H(v["sig_compressed"])
    # aggregate of two signers (src/ecdsa_test.rs:41-79) and the example (keys > r)
    for v in G["aggregate_verify_ok"] + [G["example"]]:
        msg = H(v["msg"])
        sks = [PrivateKey(H(s)) for s in v["sks"]]
        pks = [PublicKey.from_private_key(s) for s in sks]
        sigs = [ECDSA.sign(msg, s) for s in sks]
        assert ECDSA.verify(msg, sigs[0] + sigs[1], pks[0] + pks[1]) is None
        with pytest.raises(Error) as e:
            ECDSA.verify(msg, sigs[0], pks[0] + pks[1])
        assert e.value.variant == "VerificationFailed"
        # Sub / Neg: (s0 + s1) - s1 == s0 ; -(-s0) == s0
        assert ((sigs[0] + sigs[1]) - sigs[1]).raw == sigs[0].raw and (-(-sigs[0])).raw == sigs[0].raw
        assert ((pks[0] + pks[1]) - pks[1]).raw == pks[0].raw
    ex = G["example"]
    sks = [PrivateKey(H(s)) for s in ex["sks"]]
    agg_sig = ECDSA.sign(H(ex["msg"]), sks[0]) + ECDSA.sign(H(ex["msg"]), sks[1])
    agg_pk = PublicKey.from_private_key(sks[0]) + PublicKey.from_private_key(sks[1])
    assert agg_sig.to_compressed().hex() == "020aea1d3390792923eb34734b1669efad8a388997061c42fb9f7d310eec339ada"
    assert agg_pk.to_compressed().hex() == (
        "0a08685b8899122e2d7da466f39d7698b3918cf3e7ae2a2d8a3cbb6713ee4238c7e3fbfc289492c9c6b70f01a86182add29428b2ce4a6ec6ba67603ee22d7670ec")
    # check_public_keys accept / reject (src/ecdsa_test.rs:82-112)
    v = G["check_public_keys_ok"][0]
    s = PrivateKey(H(v["sk"]))
    assert check_public_keys(PublicKey.from_private_key(s), PublicKeyG1.from_private_key(s)) is None
    v = G["check_public_keys_fail"][0]
    with pytest.raises(Error) as e:
        check_public_keys(PublicKey.from_private_key(PrivateKey(H(v["sk_g2"]))), PublicKeyG1.from_private_key(PrivateKey(H(v["sk_g1"]))))
    assert e.value.variant == "VerificationFailed"
    # uncompressed round trips through verify (src/ecdsa_test.rs:115-154)
    v = G["sig_uncompressed_roundtrip"][0]
    sig = Signature.from_compressed(H(v["sig_compressed"]))
    sig2 = Signature.from_uncompressed(sig.to_uncompressed())
    assert ECDSA.verify(H(v["msg"]), sig2, PublicKey.from_private_key(PrivateKey(H(v["sk"])))) is None
    # G2 codecs (src/types_test.rs:48-69) and generator sums (:132-159)
    c = H(G["g2_compressed_roundtrip"][0]["compressed"])
    assert PublicKey.from_compressed(c).to_compressed() == c
    u = H(G["g2_uncompressed_roundtrip"][0]["uncompressed"])
    assert PublicKey.from_uncompressed(u).to_uncompressed() == u
    one = PrivateKey(be(1))
    g2, g1 = PublicKey.from_private_key(one), PublicKeyG1.from_private_key(one)
    assert (g2 + g2).to_compressed() == H(G["g2_gen_plus_gen_compressed"])
    assert (g1 + g1).to_compressed() == H(G["g1_gen_plus_gen_compressed"])
    # serde forms
    assert PublicKey.deserialize(g2.serialize()).raw == g2.raw and len(g2.serialize()) == 65
    # infinity cannot be serialised (PointInJacobian)
    with pytest.raises(Error) as e:
        (g1 - g1).to_compressed()
    assert e.value.variant == "PointInJacobian"


# ---------------------------------------------------------------------------------------------- codecs
def test_codecs(E):
    n = 256
    sks = synth.secret_keys(n, seed=51)
    p1, p2 = E.derive_pk_g1_batch(sks), E.derive_pk_g2_batch(sks)
    c1, st = E.g1_compress_batch(p1)
    assert not any(st)
    c2, st = E.g2_compress_batch(p2[:128 * 32])
    assert not any(st)
    for i in range(n):
        assert c1[33 * i:33 * i + 33] == O.g1_compress(p1[64 * i:64 * i + 64])[1]
    for i in range(32):
        assert c2[65 * i:65 * i + 65] == O.g2_compress(p2[128 * i:128 * i + 128])[1]
    d1, st = E.g1_decompress_batch(c1)
    assert not any(st) and d1 == p1
    d2, st = E.g2_decompress_batch(c2)
    assert not any(st) and d2 == p2[:128 * 32]
    assert E.g1_validate_batch(p1) == bytes(n) and E.g2_validate_batch(p2[:128 * 16]) == bytes(16)
    # rejection paths
    rng = random.Random(9)
    bad = b"".join(bytes([rng.choice([2, 3, 4])]) + be(rng.randrange(Q + 1000)) for _ in range(64))
    out, st = E.g1_decompress_batch(bad)
    for i in range(64):
        est, e = O.g1_decompress(bad[33 * i:33 * i + 33])
        assert st[i] == est and out[64 * i:64 * i + 64] == e
    flip = bytes([c2[0] ^ 1]) + c2[1:65]
    assert E.g2_decompress_batch(flip)[0] == O.g2_neg(p2[:128])[1]
    assert E.g2_decompress_batch(b"\x0c" + c2[1:65])[1][0] == O.INVALID_ENCODING
    assert E.g1_validate_batch(be(1) + be(3))[0] == O.INVALID_GROUP_POINT
    assert E.g1_compress_batch(bytes(64))[1][0] == O.POINT_IN_JACOBIAN
    # twist point outside the r-torsion is rejected by the subgroup check
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as P
    x = 1
    while True:
        y = P.f2_sqrt(P.f2_add(P.f2_mul(P.f2_mul((x, 0), (x, 0)), (x, 0)), P.B2))
        if y is not None and not P.g2_in_subgroup(((x, 0), y)):
            break
        x += 1
    raw = be(x) + be(0) + be(y[0]) + be(y[1])
    assert E.g2_validate_batch(raw)[0] == O.INVALID_GROUP_POINT
    import edge_points
    edge = edge_points.subgroup_edge_points()   # every prime-order part of the twist cofactor, mixed points, cofactor-cleared points
    got = E.g2_validate_batch(b"".join(pt for pt, _ in edge))
    assert list(got) == [0 if inside else O.INVALID_GROUP_POINT for _, inside in edge]
    assert list(got) == [O.g2_validate_uncompressed(pt) for pt, _ in edge]
    comp = E.g2_compress_batch(b"".join(pt for pt, _ in edge))[0]
    st = E.g2_decompress_batch(comp)[1]
    assert list(st) == [0 if inside else O.NOT_MEMBER for _, inside in edge] == [O.g2_decompress(comp[65 * i:65 * i + 65])[0] for i in range(len(edge))]


# ---------------------------------------------------------------------------------------------- aggregation
def test_sums_vs_oracle(E):
    n = 5000  # ragged: not a multiple of the block size
    sks = synth.secret_keys(n, seed=61)
    p1, p2 = E.derive_pk_g1_batch(sks), E.derive_pk_g2_batch(sks[:32 * 700])
    assert E.g1_sum(p1) == (O.g1_sum(p1, n)[1], 0)
    assert E.g2_sum(p2) == (O.g2_sum(p2, 700)[1], 0)
    assert E.g1_sum(b"") == (bytes(64), 0) and E.g2_sum(b"") == (bytes(128), 0)
    # duplicates (doubling inside the tree), infinities, P + (-P)
    d = p1[:64] * 9 + bytes(64) + p1[64:128]
    assert E.g1_sum(d) == (O.g1_sum(d, 11)[1], 0)
    assert E.g1_sum(p1[:64] + O.g1_neg(p1[:64])[1]) == (bytes(64), 0)
    assert E.g1_sum(p1[:128], bytes([0, 1])) == (O.g1_add(p1[:64], O.g1_neg(p1[64:128])[1])[1], 0)
    # an invalid point is reported with the status of the first failing item
    bad = p1[:640] + be(1) + be(3) + p1[640:1280]
    assert E.g1_sum(bad)[1] == O.INVALID_GROUP_POINT


def test_aggregate_verify_same_message(E):
    n = 3000
    msg = b"same message for everyone"
    sks = synth.secret_keys(n, seed=71)
    sigs, st = E.sign_batch(msg * n, len(msg), sks)
    pks = E.derive_pk_g2_batch(sks)
    assert E.aggregate_verify_same_msg(msg, sigs, pks) == 0
    assert O.verify(msg, E.g1_sum(sigs)[0], E.g2_sum(pks)[0]) == 0  # the sums themselves verify under the oracle
    assert E.aggregate_verify_same_msg(msg, sigs[64:], pks) == O.VERIFICATION_FAILED
    assert E.aggregate_verify_same_msg(msg + b"!", sigs, pks) == O.VERIFICATION_FAILED


def test_aggregate_verify_distinct_messages(E):
    n = 1500
    msgs, sks, sigs, pks = _signed_set(E, n, seed=81)
    agg = E.g1_sum(sigs)[0]
    assert E.aggregate_verify_distinct(msgs, 32, pks, agg) == 0
    assert E.aggregate_verify_distinct(msgs, 32, pks, E.g1_sum(sigs[64:])[0]) == O.VERIFICATION_FAILED
    # sharded the way bench.py shards across GPUs: partial Miller products, then one shared final exponentiation
    cuts = [0, 400, 401, 1100, n]
    parts = b"".join(E.miller_partial_distinct(msgs[32 * a:32 * b], 32, pks[128 * a:128 * b])[0] for a, b in zip(cuts, cuts[1:]))
    assert E.finish_distinct(parts, agg) == 0
    # the partial of a small slice equals the oracle's Miller product (field elements are canonical)
    hs, _ = E.hash_to_g1_batch(msgs, 32, 8)
    assert E.miller_partial_distinct(msgs[:256], 32, pks[:1024])[0] == O.miller_product(hs, pks[:1024], 8)[1]
    assert E.miller_partial_distinct(b"", 32, b"") == (be(1) + bytes(352), 0)


# ---------------------------------------------------------------------------------------------- full BASELINE sizes
def test_full_size_verify_2_20(E):
    """configs[1]: 2^20 independent triples; 1 % flipped invalid (seeded).  The verdict vector must equal the
    construction (valid -> 0, flipped -> 9) and, on a 2^12 sample, the oracle's."""
    n = 1 << 20
    msgs, sks, sigs, pks = _signed_set(E, n, seed=1)
    rng = np.random.default_rng(99)
    bad = np.sort(rng.choice(n, size=n // 100, replace=False))
    m = np.frombuffer(msgs, dtype=np.uint8).reshape(n, 32).copy()
    m[bad, 5] ^= 0x40
    msgs = m.tobytes()
    st = np.frombuffer(E.verify_batch(msgs, 32, sigs, pks), dtype=np.uint8)
    expect = np.zeros(n, dtype=np.uint8)
    expect[bad] = O.VERIFICATION_FAILED
    assert np.array_equal(st, expect)
    idx = np.concatenate([bad[:2048], np.arange(2048)])
    sm = b"".join(msgs[32 * i:32 * i + 32] for i in idx)
    ss = b"".join(sigs[64 * i:64 * i + 64] for i in idx)
    sp = b"".join(pks[128 * i:128 * i + 128] for i in idx)
    assert O.verify_batch(sm, 32, ss, sp, len(idx), NTHREADS) == bytes(st[idx])


def test_full_size_sums_linearity(E):
    """configs[3] at 2^20: sum(k_i * G) == (sum k_i) * G for G1 and G2, and the aggregate verifies."""
    n = 1 << 20
    sks = synth.secret_keys(n, seed=2)
    total = sum(int.from_bytes(sks[32 * i:32 * i + 32], "big") for i in range(n)) % R
    p1 = E.derive_pk_g1_batch(sks)
    assert E.g1_sum(p1) == (O.derive_pk_g1(be(total))[1], 0)
    p2 = E.derive_pk_g2_batch(sks)
    assert E.g2_sum(p2) == (O.derive_pk_g2(be(total))[1], 0)
    msg = synth.messages(1, 32, seed=3)
    sigs, st = E.sign_batch(msg * n, 32, sks)
    assert E.aggregate_verify_same_msg(msg, sigs, p2) == 0
    assert E.g1_sum(sigs)[0] == O.sign(msg, be(total))[1]


# ---------------------------------------------------------------------------------------------- building blocks
def test_layer_hooks_match_host_simulation(E):
    """Every out-of-line building block of the pairing / group code, GPU kernel vs the g++ build of the same source
    (tests/hostsim, itself pinned to the oracle by tests/test_hostsim.py).  Guards against device-compiler
    stack-slot sharing errors seen during bring-up ('toolchain hazards')."""
    import ctypes
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "hostsim")], stdout=subprocess.DEVNULL)
    hs = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
    rng = random.Random(1)
    ops = {0: (4, 2), 1: (2, 2), 2: (3, 2), 3: (6, 12), 4: (10, 12), 5: (18, 12), 6: (5, 3), 7: (2, 2), 8: (2, 2), 11: (5, 3), 12: (6, 3),
           13: (10, 6)}
    for op, (ni, no) in ops.items():
        n = 64
        data = b"".join(be(rng.randrange(Q)) for _ in range(ni * n))
        got = E.layer_op_batch(op, data, ni, no)
        for i in range(n):
            o = ctypes.create_string_buffer(32 * no)
            hs.hs_layer_op(op, data[32 * ni * i:32 * ni * (i + 1)], ni, o, no)
            assert o.raw == got[32 * no * i:32 * no * (i + 1)], (op, i)
    g1, g2 = O.derive_pk_g1(be(12345))[1], O.derive_pk_g2(be(6789))[1]
    assert E.layer_op_batch(200, g1 + g2, 6, 12) == O.miller_product(g1, g2, 1)[1]


def test_pairing_modes_agree_on_ragged_batches(E):
    """The cooperative machine (mode 0, csrc/coop.cuh) and the one-thread-per-item kernels (mode 1, csrc/pairing.cuh) are two
    independent implementations of bn::pairing_batch: identical verdicts on batches whose size is not a multiple of the
    32-item group, including a single item, with invalid items in every position class."""
    from bn254_b200._native import I
    ctx = E.context(0)
    n = 1000
    msgs, sks, sigs, pks = _signed_set(E, n, seed=77)
    sigs, pks = bytearray(sigs), bytearray(pks)
    for i in (0, 31, 32, 33, 63, 500, 991, 999):  # group boundaries and the ragged tail
        sigs[64 * i:64 * i + 64] = O.g1_neg(bytes(sigs[64 * i:64 * i + 64]))[1]
    sigs[64 * 7:64 * 8] = bytes(64)
    pks[128 * 7:128 * 8] = bytes(128)      # both infinite: accepted
    sigs[64 * 9 + 63] ^= 1                 # off curve
    sigs, pks = bytes(sigs), bytes(pks)
    want = O.verify_batch(msgs, 32, sigs, pks, n, NTHREADS)
    # modes 2 / 3: the cooperative machine in its block layout (32-item groups) and its warp-local layout (5 items per
    # warp, 30 per block): sizes around both group sizes
    for size in (1000, 1, 4, 5, 6, 15, 16, 17, 29, 30, 31, 33, 95, 127, 129):
        got = {}
        for mode in (0, 1, 2, 3, 4):
            ctx.call("bn254_set_pairing_mode", I(mode))
            got[mode] = E.verify_batch(msgs[:32 * size], 32, sigs[:64 * size], pks[:128 * size], ctx=ctx)
        ctx.call("bn254_set_pairing_mode", I(0))
        assert got[0] == got[1] == got[2] == got[3] == got[4] == want[:size], size
    # check_public_keys goes through the same two paths (generator instead of H(m))
    pk1 = E.derive_pk_g1_batch(sks[:32 * 40], ctx=ctx)
    pk2 = bytearray(pks[:128 * 40])
    pk2[128 * 3:128 * 4] = pks[128 * 4:128 * 5]
    for mode in (0, 1, 2, 3, 4):
        ctx.call("bn254_set_pairing_mode", I(mode))
        st = E.check_public_keys_batch(bytes(pk2), pk1, ctx=ctx)
        assert st[3] == O.VERIFICATION_FAILED and st[7] != 0 and sum(1 for s in st if s == 0) == 38, (mode, st)
    ctx.call("bn254_set_pairing_mode", I(0))


def test_format_pairing_check_values(E):
    """/root/reference/src/utils.rs:197-239 has no reference test (parity unpinned): check the two formatters against each
    other, the little-endian layout against the oracle's points, and that the formatted pairs satisfy the pairing check."""
    from bn254_b200 import (ECDSA, PrivateKey, PublicKey, format_pairing_check_uncompressed_values, format_pairing_check_values)
    v = G["verify_ok"][0]
    sk = PrivateKey(H(v["sk"]))
    pk = PublicKey.from_private_key(sk)
    msg = H(v["msg"])
    sig = ECDSA.sign(msg, sk)
    a = format_pairing_check_values(msg, sig.to_compressed(), pk.to_compressed())
    b = format_pairing_check_uncompressed_values(msg, sig.to_uncompressed(), pk.to_uncompressed())
    assert a == b and [len(x) for p in a for x in p] == [64, 128, 64, 128]
    rev = lambda x: b"".join(x[i:i + 32][::-1] for i in range(0, len(x), 32))
    assert rev(a[0][0]) == O.hash_to_g1(msg)[1] and rev(a[0][1]) == pk.to_uncompressed()
    assert rev(a[1][0]) == sig.to_uncompressed() and rev(a[1][1]) == O.g2_neg(O.derive_pk_g2(be(1))[1])[1]
    assert E.pairing_check_batch(rev(a[0][0]) + rev(a[1][0]), rev(a[0][1]) + rev(a[1][1]), 2, 1) == b"\x00"


def test_distinct_partial_modes_bit_identical(E):
    """The Miller product of a slice of (H(m_i), pk_i) pairs is a canonical field value: the cooperative multi-pairing program
    (8 pairs per lane, block butterfly) and the one-thread-per-pair kernels must return the same 384 bytes, and the oracle's."""
    from bn254_b200._native import I
    ctx = E.context(0)
    n = 777  # not a multiple of 8 * 32: padding slots carry constant-1 lines
    msgs, sks, sigs, pks = _signed_set(E, n, seed=91)
    pks = bytearray(pks)
    pks[128 * 13:128 * 14] = bytes(128)  # a pair with an infinite G2 point is skipped
    pks = bytes(pks)
    got = {}
    for mode in (0, 1):
        ctx.call("bn254_set_pairing_mode", I(mode))
        got[mode] = E.miller_partial_distinct(msgs, 32, pks, ctx=ctx)
    ctx.call("bn254_set_pairing_mode", I(0))
    assert got[0] == got[1] and got[0][1] == 0
    hs = E.hash_to_g1_batch(msgs, 32, n, ctx=ctx)[0]
    assert got[0][0] == O.miller_product(hs, pks, n)[1]
    # a bad key in the slice: first error by index, in both modes
    bad = bytearray(pks)
    bad[128 * 500 + 127] ^= 1
    for mode in (0, 1):
        ctx.call("bn254_set_pairing_mode", I(mode))
        assert E.miller_partial_distinct(msgs, 32, bytes(bad), ctx=ctx)[1] == O.INVALID_GROUP_POINT
    ctx.call("bn254_set_pairing_mode", I(0))
    assert E.miller_partial_distinct(b"", 32, b"", ctx=ctx)[0] == be(1) + bytes(352)


def test_verify_batch_rlc(E):
    """Randomised batch verification: the fast path is taken exactly when every
    item is valid, and the statuses are always those of verify_batch / the oracle -- including a pair of forged signatures
    whose errors cancel in an unrandomised aggregate check."""
    import edge_points
    n = 600
    msgs, sks, sigs, pks = _signed_set(E, n, seed=57)
    coeffs = synth.rand_bytes(1234, 16 * n)
    st, fast = E.verify_batch_rlc(msgs, 32, sigs, pks, coeffs)
    assert fast and st == bytes(n)
    st, fast = E.verify_batch_rlc(msgs, 32, sigs, pks)                      # coefficients drawn by the engine
    assert fast and st == bytes(n)
    st, fast = E.verify_batch_rlc(msgs, 32, sigs, pks, coeffs, pks_in_g2=True)
    assert fast and st == bytes(n)
    st, fast = E.verify_batch_rlc(msgs, 32, sigs, pks, bytes(16 * n))       # zero coefficients count as one
    assert fast and st == bytes(n)
    for size in (1, 7, 255, 257):                                           # around the 256-pair group of the multi-pairing machine
        st, fast = E.verify_batch_rlc(msgs[:32 * size], 32, sigs[:64 * size], pks[:128 * size], coeffs[:16 * size])
        assert fast and st == bytes(size), size
    # infinite signature + infinite key: both pairs are skipped, the item is accepted (bn::pairing_batch semantics)
    s2, p2 = bytearray(sigs), bytearray(pks)
    s2[64 * 5:64 * 6] = bytes(64)
    p2[128 * 5:128 * 6] = bytes(128)
    st, fast = E.verify_batch_rlc(msgs, 32, bytes(s2), bytes(p2), coeffs)
    assert fast and st == bytes(n) == E.verify_batch(msgs, 32, bytes(s2), bytes(p2))
    # one wrong signature: the combined check fails, the exact path decides
    s3 = bytearray(sigs)
    s3[64 * 77:64 * 78] = O.g1_neg(sigs[64 * 77:64 * 78])[1]
    st, fast = E.verify_batch_rlc(msgs, 32, bytes(s3), pks, coeffs)
    want = O.verify_batch(msgs, 32, bytes(s3), pks, n, NTHREADS)
    assert not fast and st == want and st[77] == O.VERIFICATION_FAILED and sum(st) == O.VERIFICATION_FAILED
    # cancelling forgeries: sig_a + D and sig_b - D leave the plain sum of signatures unchanged
    d = O.derive_pk_g1(be(4242))[1]
    s4 = bytearray(sigs)
    s4[64 * 10:64 * 11] = O.g1_add(sigs[64 * 10:64 * 11], d)[1]
    s4[64 * 20:64 * 21] = O.g1_add(sigs[64 * 20:64 * 21], O.g1_neg(d)[1])[1]
    st, fast = E.verify_batch_rlc(msgs, 32, bytes(s4), pks, coeffs)
    want = O.verify_batch(msgs, 32, bytes(s4), pks, n, NTHREADS)
    assert not fast and st == want and st[10] == st[20] == O.VERIFICATION_FAILED
    ones = (bytes(15) + b"\x01") * n   # a caller that does not randomise is fooled: this is why the coefficients must be secret and random
    st, fast = E.verify_batch_rlc(msgs, 32, bytes(s4), pks, ones)
    assert fast and st == bytes(n)
    # undecodable items and a key outside G2 never ride
    s5 = bytearray(sigs)
    s5[64 * 3 + 63] ^= 1
    st, fast = E.verify_batch_rlc(msgs, 32, bytes(s5), pks, coeffs)
    assert not fast and st == O.verify_batch(msgs, 32, bytes(s5), pks, n, NTHREADS) and st[3] != 0
    outside = [pt for pt, inside in edge_points.subgroup_edge_points() if not inside][1]
    p6 = bytearray(pks)
    p6[128 * 9:128 * 10] = outside
    st, fast = E.verify_batch_rlc(msgs, 32, sigs, bytes(p6), coeffs)
    assert not fast and st == O.verify_batch(msgs, 32, sigs, bytes(p6), n, NTHREADS) == E.verify_batch(msgs, 32, sigs, bytes(p6))
    assert E.verify_batch_rlc(b"", 32, b"", b"") == (b"", False)


def test_verify_batch_rlc_slices_isolate_failures(E):
    """The randomised pass gives one verdict per slice of whole multi-pairing groups (up to 64 per 2^20-triple chunk); only
    failing slices are redone by the exact path.  20 000 triples with forged, undecodable and out-of-subgroup items spread over
    several slices: the statuses are exactly verify_batch's."""
    import edge_points
    n = 20000
    msgs, sks, sigs, pks = _signed_set(E, n, seed=63)
    coeffs = synth.rand_bytes(4321, 16 * n)
    st, fast = E.verify_batch_rlc(msgs, 32, sigs, pks, coeffs)
    assert fast and st == bytes(n)
    s2, p2 = bytearray(sigs), bytearray(pks)
    for i in (3, 777, 9999, 19999):
        s2[64 * i:64 * i + 64] = O.g1_neg(sigs[64 * i:64 * i + 64])[1]
    s2[64 * 5000 + 63] ^= 1                                                       # undecodable signature
    p2[128 * 15000:128 * 15001] = [pt for pt, inside in edge_points.subgroup_edge_points() if not inside][0]   # key outside G2
    s2[64 * 12345:64 * 12346] = bytes(64)                                         # infinite signature, real key: rejected
    s2, p2 = bytes(s2), bytes(p2)
    st, fast = E.verify_batch_rlc(msgs, 32, s2, p2, coeffs)
    want = E.verify_batch(msgs, 32, s2, p2)
    assert not fast and st == want
    bad = {i for i in range(n) if want[i]}
    assert bad == {3, 777, 5000, 9999, 12345, 15000, 19999}
    assert want == O.verify_batch(msgs, 32, s2, p2, n, NTHREADS)


def test_verify_workspace_chunking(E):
    """verify_batch walks the batch in workspace chunks (2^20 items by default, BN254_COOP_CHUNK_LOG2): with 2^12-item chunks a
    10 000-triple batch takes three passes (the hash runs once over the whole batch, line sets and the machine per chunk) and
    must give the same verdicts as the single-chunk context, invalid items on both sides of the chunk borders included."""
    from bn254_b200._native import Context
    n = 10000
    msgs, sks, sigs, pks = _signed_set(E, n, seed=71)
    sigs = bytearray(sigs)
    for i in (0, 4095, 4096, 4097, 8191, 8192, 9999):
        sigs[64 * i:64 * i + 64] = O.g1_neg(bytes(sigs[64 * i:64 * i + 64]))[1]
    sigs = bytes(sigs)
    want = E.verify_batch(msgs, 32, sigs, pks)
    assert {i for i in range(n) if want[i]} == {0, 4095, 4096, 4097, 8191, 8192, 9999}
    old = os.environ.get("BN254_COOP_CHUNK_LOG2")
    os.environ["BN254_COOP_CHUNK_LOG2"] = "12"
    try:
        small = Context(0)
    finally:
        if old is None:
            del os.environ["BN254_COOP_CHUNK_LOG2"]
        else:
            os.environ["BN254_COOP_CHUNK_LOG2"] = old
    try:
        assert E.verify_batch(msgs, 32, sigs, pks, ctx=small) == want
        st, fast = E.verify_batch_rlc(msgs, 32, sigs, pks, synth.rand_bytes(5, 16 * n), ctx=small)
        assert not fast and st == want
    finally:
        small.close()


# ---------------------------------------------------------------------------------------------- input policy, hooks, long messages
def _untrusted_expect(msg, sig, pk):
    """What a caller of the crate gets from bytes: PublicKey::from_uncompressed, Signature::from_uncompressed, ECDSA::verify."""
    return O.g2_validate_uncompressed(pk) or O.g1_validate_uncompressed(sig) or O.verify(msg, sig, pk)


def test_untrusted_input_policy(E):
    """Default policy of a fresh context: bytes are decoded like from_uncompressed.  Raw (0, 0) inputs -- a universal
    forgery under the typed policy, where both pairs would be skipped -- fail with InvalidGroupPoint, and so does a key on the
    twist but outside G2; statuses equal the oracle's composition of the crate's three calls."""
    import edge_points
    from bn254_b200._native import Context
    ctx = Context(0)
    try:
        assert ctx.input_policy == E.INPUTS_UNTRUSTED
        n = 64
        msgs, sks, sigs, pks = _signed_set(E, n, seed=301)
        msgs, sigs, pks = bytearray(msgs), bytearray(sigs), bytearray(pks)
        sigs[0:64] = bytes(64); pks[0:128] = bytes(128)                    # both "infinity"
        sigs[64:128] = bytes(64)                                          # signature (0, 0), real key
        pks[128 * 2:128 * 3] = bytes(128)                                 # key (0, 0), real signature
        pks[128 * 3:128 * 4] = edge_points.twist_point_outside_g2()       # on the twist, not in the r-torsion
        sigs[64 * 4 + 63] ^= 1                                            # off-curve signature
        pks[128 * 5:128 * 5 + 32] = be(Q + 1)                             # coordinate >= q
        msgs[32 * 6] ^= 1                                                 # plain wrong message
        msgs, sigs, pks = bytes(msgs), bytes(sigs), bytes(pks)
        got = E.verify_batch(msgs, 32, sigs, pks, ctx=ctx)
        want = bytes(_untrusted_expect(msgs[32 * i:32 * i + 32], sigs[64 * i:64 * i + 64], pks[128 * i:128 * i + 128]) for i in range(n))
        assert got == want
        assert list(got[:7]) == [O.INVALID_GROUP_POINT] * 5 + [O.NOT_MEMBER, O.VERIFICATION_FAILED] and not any(got[7:])
        # the same inputs under the typed policy: item 0 is accepted (both pairs skipped) -- the reason the default is untrusted
        assert E.verify_batch(msgs, 32, sigs, pks)[0] == 0
        # check_public_keys and the aggregate checks follow the policy too
        g1 = E.derive_pk_g1_batch(sks)
        assert E.check_public_keys_batch(bytes(128) + pks[128:256], bytes(64) + g1[64:128], ctx=ctx) == bytes([O.INVALID_GROUP_POINT, 0])
        assert E.aggregate_verify_same_msg(b"m", bytes(64), bytes(128), ctx=ctx) == O.INVALID_GROUP_POINT
        assert E.aggregate_verify_same_msg(b"m", bytes(64), bytes(128)) == 0          # typed: infinity sums, both pairs skipped
        vm, vs, vp = msgs[32 * 7:32 * 40], sigs[64 * 7:64 * 40], pks[128 * 7:128 * 40]
        agg = E.g1_sum(vs)[0]
        assert E.aggregate_verify_distinct(vm, 32, vp, agg, ctx=ctx) == 0
        assert E.aggregate_verify_distinct(vm, 32, vp[:128 * 5] + edge_points.twist_point_outside_g2() + vp[128 * 6:], agg, ctx=ctx) == O.INVALID_GROUP_POINT
        # randomised batch verification keeps verify_batch's statuses under the same policy
        st, fast = E.verify_batch_rlc(msgs, 32, sigs, pks, synth.rand_bytes(9, 16 * n), ctx=ctx)
        assert st == want and not fast
    finally:
        ctx.close()


def test_hash_to_point_error_hook(E):
    """/root/reference/src/hash.rs:62: HashToPointError after the last counter.  With the try limit lowered to k (test hook)
    every message whose accepted counter is >= k must come back with status 1 from hash, sign and verify -- through the
    per-thread loop (small batch) and through the compacting rounds (large batch)."""
    from bn254_b200._native import Context
    ctx = Context(0)
    try:
        E.set_input_policy(E.INPUTS_TYPED, ctx=ctx)
        for n, k in ((300, 1), (300, 3), (6000, 2), (6000, 14)):
            msgs, sks = synth.messages(n, 32, seed=400 + k), synth.secret_keys(n, seed=500 + k)
            ctrs = [O.hash_to_g1(msgs[32 * i:32 * i + 32])[2] for i in range(n)]
            want = bytes(O.HASH_TO_POINT if c >= k else 0 for c in ctrs)
            assert any(want) and not all(want)
            E.set_hash_try_limit(255, ctx=ctx)
            sigs, st = E.sign_batch(msgs, 32, sks, ctx=ctx)
            pks = E.derive_pk_g2_batch(sks, ctx=ctx)
            assert not any(st)
            E.set_hash_try_limit(k, ctx=ctx)
            pts, st = E.hash_to_g1_batch(msgs, 32, n, ctx=ctx)
            assert st == want
            assert all(pts[64 * i:64 * i + 64] == (bytes(64) if want[i] else O.hash_to_g1(msgs[32 * i:32 * i + 32])[1]) for i in range(0, n, 37))
            s2, st = E.sign_batch(msgs, 32, sks, ctx=ctx)
            assert st == want and all(s2[64 * i:64 * i + 64] == (bytes(64) if want[i] else sigs[64 * i:64 * i + 64]) for i in range(n))
            assert E.verify_batch(msgs, 32, sigs, pks, ctx=ctx) == want           # the hash error propagates (src/ecdsa.rs:53)
            E.set_input_policy(E.INPUTS_UNTRUSTED, ctx=ctx)
            assert E.verify_batch(msgs, 32, sigs, pks, ctx=ctx) == want
            E.set_input_policy(E.INPUTS_TYPED, ctx=ctx)
    finally:
        ctx.close()


@pytest.mark.parametrize("msg_len", [55, 64, 65, 200])
def test_verify_long_messages(E, msg_len):
    """Messages that do not share one SHA-256 block with the counter (>= 55 bytes) take the per-thread hash loop inside verify:
    4096 triples per length, a seeded 3 % invalid, verdicts equal to the oracle's."""
    n = 4096
    msgs, sks = synth.messages(n, msg_len, seed=600 + msg_len), synth.secret_keys(n, seed=700 + msg_len)
    sigs, st = E.sign_batch(msgs, msg_len, sks)
    assert not any(st)
    assert sigs[:64 * 64] == O.sign_batch(msgs[:64 * msg_len], msg_len, sks[:64 * 32], 64, NTHREADS)[0]
    pks = E.derive_pk_g2_batch(sks)
    m = bytearray(msgs)
    bad = sorted(random.Random(msg_len).sample(range(n), n // 32))
    for i in bad:
        m[msg_len * i + msg_len - 1] ^= 0x80
    got = E.verify_batch(bytes(m), msg_len, sigs, pks)
    assert {i for i in range(n) if got[i]} == set(bad) and all(got[i] == O.VERIFICATION_FAILED for i in bad)
    assert got == O.verify_batch(bytes(m), msg_len, sigs, pks, n, NTHREADS)


def test_full_size_sign_2_20(E):
    """configs[2]: 2^20 distinct messages and keys signed in one call; a seeded sample of 1024 signatures equals the oracle's
    bit for bit, and every signature verifies (so none is wrong in a way the sample missed)."""
    n = 1 << 20
    msgs, sks, sigs, pks = _signed_set(E, n, seed=11)
    idx = np.sort(np.random.default_rng(12).choice(n, size=1024, replace=False))
    sm = b"".join(msgs[32 * i:32 * i + 32] for i in idx)
    sk = b"".join(sks[32 * i:32 * i + 32] for i in idx)
    want, st = O.sign_batch(sm, 32, sk, len(idx), NTHREADS)
    assert not any(st) and want == b"".join(sigs[64 * i:64 * i + 64] for i in idx)
    assert E.verify_batch(msgs, 32, sigs, pks) == bytes(n)


def test_full_size_rlc_2_20(E):
    """Randomised batch verification at 2^20 triples: all valid -> fast path and all-zero statuses; with 100 seeded invalid
    items -> exactly verify_batch's statuses (themselves checked against the oracle on the invalid items and a sample)."""
    n = 1 << 20
    msgs, sks, sigs, pks = _signed_set(E, n, seed=21)
    coeffs = synth.rand_bytes(22, 16 * n)
    st, fast = E.verify_batch_rlc(msgs, 32, sigs, pks, coeffs, pks_in_g2=True)
    assert fast and st == bytes(n)
    bad = np.sort(np.random.default_rng(23).choice(n, size=100, replace=False))
    s = np.frombuffer(sigs, dtype=np.uint8).reshape(n, 64).copy()
    s[bad] = s[(bad + 1) % n]
    sigs2 = s.tobytes()
    st, fast = E.verify_batch_rlc(msgs, 32, sigs2, pks, coeffs)
    want = E.verify_batch(msgs, 32, sigs2, pks)
    assert not fast and st == want
    assert {int(i) for i in np.nonzero(np.frombuffer(want, dtype=np.uint8))[0]} == {int(i) for i in bad}
    idx = np.concatenate([bad, np.arange(0, n, n // 400)])
    pick = lambda b, w: b"".join(b[w * i:w * i + w] for i in idx)
    assert O.verify_batch(pick(msgs, 32), 32, pick(sigs2, 64), pick(pks, 128), len(idx), NTHREADS) == bytes(want[i] for i in idx)


# ---------------------------------------------------------------------------------------------- aggregate checks on the device
def _dev(b):
    import torch
    return torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()


def test_distinct_aggregate_device_path_vs_oracle(E):
    """bn254_distinct_payload_dev + bn254_finish_distinct_dev (what the multi-GPU step runs, here with the records of several
    slices on one GPU): verdicts equal the oracle's fold over the same pairs -- accept, a forged signature, a key that is not on
    the curve (first failing item's status), a slice holding infinity keys -- and the cooperative finish agrees with the
    one-thread finish (pairing mode 1)."""
    import torch
    from bn254_b200 import dist as D
    from bn254_b200._native import I, S
    ctx = E.context(0)
    n = 2500
    msgs, sks, sigs, pks = _signed_set(E, n, seed=801)
    neg_g2 = O.g2_neg(O.derive_pk_g2(be(1))[1])[1]
    hs = E.hash_to_g1_batch(msgs, 32, n)[0]

    def oracle_fold(pk_bytes, agg):
        for i in range(n):
            stp = O.miller_product(hs[64 * i:64 * i + 64], pk_bytes[128 * i:128 * i + 128], 1)[0]   # decode status of pair i
            if stp:
                return stp
        return O.pairing_check(hs + agg, pk_bytes + neg_g2, n + 1)[0]

    def run(pk_bytes, sig_bytes, cuts, local_sigs, mode=0):
        ctx.call("bn254_set_pairing_mode", I(mode))
        try:
            recs = torch.zeros(448 * (len(cuts) - 1), dtype=torch.uint8, device="cuda")
            d_m, d_p, d_s = _dev(msgs), _dev(pk_bytes), _dev(sig_bytes)
            for r, (a, b) in enumerate(zip(cuts, cuts[1:])):
                ctx.call("bn254_distinct_payload_dev", d_m[32 * a:], S(32), d_p[128 * a:], d_s[64 * a:] if local_sigs else None, S(b - a),
                         recs[448 * r:448 * r + 448])
            verdict = torch.zeros(1, dtype=torch.uint8, device="cuda")
            agg = None if local_sigs else _dev(E.g1_sum(sig_bytes)[0])
            ctx.call("bn254_finish_distinct_dev", recs, S(len(cuts) - 1), agg, verdict)
            ctx.sync()
            return int(verdict.cpu()[0])
        finally:
            ctx.call("bn254_set_pairing_mode", I(0))

    cuts = [0, 700, 701, 1900, n]
    agg = E.g1_sum(sigs)[0]
    assert oracle_fold(pks, agg) == 0
    for local in (True, False):
        for mode in (0, 1):
            assert run(pks, sigs, cuts, local, mode) == 0
    forged = sigs[:64 * 1234] + sigs[64 * 1235:64 * 1236] + sigs[64 * 1235:]       # item 1234 carries item 1235's signature
    assert oracle_fold(pks, E.g1_sum(forged)[0]) == O.VERIFICATION_FAILED
    for local in (True, False):
        assert run(pks, forged, cuts, local) == O.VERIFICATION_FAILED
    assert run(pks, forged, cuts, True, 1) == O.VERIFICATION_FAILED
    badkey = bytearray(pks)
    badkey[128 * 2000 + 127] ^= 1                                                 # third slice: not on the curve
    badkey[128 * 2300:128 * 2301] = be(Q) + bytes(96)                             # fourth slice: a later, different error
    assert oracle_fold(bytes(badkey), agg) == O.INVALID_GROUP_POINT
    assert run(bytes(badkey), sigs, cuts, True) == O.INVALID_GROUP_POINT
    assert run(bytes(badkey), sigs, cuts, False, 1) == O.INVALID_GROUP_POINT
    # the class bench.py drives, world = 1
    agg_dev = D.DistinctAggregate(ctx, world=1)
    agg_dev.step(_dev(msgs), 32, _dev(pks), _dev(sigs), n)
    assert agg_dev.status() == 0
    agg_dev.step(_dev(msgs), 32, _dev(pks), _dev(forged), n)
    assert agg_dev.status() == O.VERIFICATION_FAILED


def test_distinct_aggregate_2_20_pairs(E):
    """configs[4] at 2^20 pairs on one GPU through the device path: several launches of whole waves plus the re-cut last wave.
    Accept for the honest aggregate; one forged signature, one key off the curve and one key outside G2 (untrusted policy) are
    each reported -- the verdicts the oracle's fold gives by construction (its full fold is checked at 2 500 pairs above)."""
    import edge_points
    from bn254_b200 import dist as D
    from bn254_b200._native import Context
    n = 1 << 20
    msgs, sks, sigs, pks = _signed_set(E, n, seed=811)
    ctx = E.context(0)
    agg = D.DistinctAggregate(ctx, world=1)
    d_m, d_p, d_s = _dev(msgs), _dev(pks), _dev(sigs)
    agg.step(d_m, 32, d_p, d_s, n)
    assert agg.status() == 0
    agg.step(d_m, 32, d_p, d_s[64:], n - 1)                # pairs 0 .. n-2 against signatures 1 .. n-1: rejected
    assert agg.status() == O.VERIFICATION_FAILED
    forged = bytearray(sigs)
    forged[64 * 777777:64 * 777778] = sigs[64 * 5:64 * 6]
    agg.step(d_m, 32, d_p, _dev(bytes(forged)), n)
    assert agg.status() == O.VERIFICATION_FAILED
    bad = bytearray(pks)
    bad[128 * 999999 + 64] ^= 2
    agg.step(d_m, 32, _dev(bytes(bad)), d_s, n)
    assert agg.status() == O.INVALID_GROUP_POINT
    strict = Context(0)
    try:
        out = bytearray(pks)
        out[128 * 424242:128 * 424243] = edge_points.twist_point_outside_g2()
        a2 = D.DistinctAggregate(strict, world=1)
        a2.step(d_m, 32, _dev(bytes(out)), d_s, n)
        assert a2.status() == O.INVALID_GROUP_POINT
    finally:
        strict.close()


def test_same_message_aggregate_device_path(E):
    from bn254_b200 import dist as D
    ctx = E.context(0)
    n = 5000
    msg = b"one message, many signers"
    sks = synth.secret_keys(n, seed=821)
    sigs, st = E.sign_batch(msg * n, len(msg), sks)
    pks = E.derive_pk_g2_batch(sks)
    agg = D.SameMessageAggregate(ctx, world=1)
    agg.step(_dev(msg), len(msg), _dev(sigs), _dev(pks), n)
    assert agg.status() == 0 and O.verify(msg, E.g1_sum(sigs)[0], E.g2_sum(pks)[0]) == 0
    agg.step(_dev(msg), len(msg), _dev(sigs[64:]), _dev(pks), n - 1)
    assert agg.status() == O.VERIFICATION_FAILED


def test_two_rank_nccl_aggregates():
    """world_size 2 over NCCL on two GPUs of the box (skipped on a one-GPU box): tests/dist_gpu_worker.py drives the device
    paths of bn254_b200/dist.py with the CUDA engine -- a forged signature and an undecodable key on rank 1 included."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_format_pairing_check_c_abi(E):
    """bn254_format_pairing_check_batch (/root/reference/src/utils.rs:197-239) in batch form: both input forms agree, the layout
    is the little-endian image of the oracle's points, the formatted pairs pass the pairing check, and errors come in the
    reference's order (hash, public key, signature)."""
    n = 200
    msgs, sks, sigs, pks = _signed_set(E, n, seed=831)
    cs, st1 = E.g1_compress_batch(sigs)
    cp, st2 = E.g2_compress_batch(pks)
    assert not any(st1) and not any(st2)
    a, sta = E.format_pairing_check_batch(msgs, 32, cs, cp, True)
    b, stb = E.format_pairing_check_batch(msgs, 32, sigs, pks, False)
    assert a == b and not any(sta) and not any(stb)
    rev = lambda x: b"".join(x[i:i + 32][::-1] for i in range(0, len(x), 32))
    neg_g2 = O.g2_neg(O.derive_pk_g2(be(1))[1])[1]
    for i in range(0, n, 17):
        blob = rev(a[384 * i:384 * i + 384])
        assert blob == O.hash_to_g1(msgs[32 * i:32 * i + 32])[1] + pks[128 * i:128 * i + 128] + sigs[64 * i:64 * i + 64] + neg_g2
    g1s = b"".join(rev(a[384 * i:384 * i + 64]) + rev(a[384 * i + 192:384 * i + 256]) for i in range(n))
    g2s = b"".join(rev(a[384 * i + 64:384 * i + 192]) + rev(a[384 * i + 256:384 * i + 384]) for i in range(n))
    assert E.pairing_check_batch(g1s, g2s, 2, n) == bytes(n)
    bad_pk = bytes([0x0c]) + cp[1:65]            # bad sign byte -> InvalidEncoding
    bad_sig = bytes([0x04]) + cs[1:33]
    _, st = E.format_pairing_check_batch(msgs[:64], 32, bad_sig + cs[33:66], bad_pk + cp[65:130], True)
    assert st == bytes([O.INVALID_ENCODING, 0])
    _, st = E.format_pairing_check_batch(msgs[:32], 32, bad_sig, cp[:65], True)
    assert st == bytes([O.INVALID_ENCODING])
    from bn254_b200 import Error, format_pairing_check_uncompressed_values
    with pytest.raises(Error) as e:
        format_pairing_check_uncompressed_values(b"m", sigs[:64] + b"x", pks[:128])
    assert e.value.variant == "SerializationError"


def test_verify_with_cached_key_lines(E):
    """A fixed key set (validators) verifying many messages: bn254_key_lines_prepare_dev walks every key once,
    bn254_verify_batch_cached_dev only scales the cached lines per triple.  Statuses must be verify_batch's / the oracle's for
    the same triples -- identity and permuted key indices, forged items, a bad key in the set, an index out of range, a
    signature at infinity -- under both input policies."""
    import edge_points
    from bn254_b200._native import Context
    nk = 300
    sks = synth.secret_keys(nk, seed=901)
    pks = bytearray(E.derive_pk_g2_batch(sks))
    pks[128 * 7 + 127] ^= 1                                               # key 7 is not on the curve
    pks[128 * 9:128 * 10] = bytes(128)                                    # key 9 is the point at infinity
    pks = bytes(pks)
    n = 5000
    rng = random.Random(5)
    idx = [rng.randrange(nk) for _ in range(n)]
    idx[17] = nk + 3                                                      # out of range
    msgs = synth.messages(n, 32, seed=902)
    key_sk = lambda j: sks[32 * j:32 * j + 32]
    sk_per_item = b"".join(key_sk(j if j < nk else 0) for j in idx)
    sigs = bytearray(E.sign_batch(msgs, 32, sk_per_item)[0])
    for i in (3, 1000, 4999):                                             # forged: signature of the neighbour
        sigs[64 * i:64 * i + 64] = sigs[64 * (i - 1):64 * i]
    sigs[64 * 40:64 * 41] = bytes(64)                                     # signature at infinity
    sigs[64 * 41 + 10] ^= 4                                               # off-curve signature
    sigs = bytes(sigs)
    pk_per_item = b"".join(pks[128 * j:128 * j + 128] if j < nk else bytes(128) for j in idx)

    cache = E.KeyLineCache(pks)                                           # module context: typed policy
    ks = cache.key_status()
    assert ks[7] == O.INVALID_GROUP_POINT and ks[9] == 0 and sum(1 for b in ks if b) == 1
    got = cache.verify(msgs, 32, sigs, idx)
    want = bytearray(E.verify_batch(msgs, 32, sigs, pk_per_item))
    want[17] = O.INDEX_OOB
    assert got == bytes(want)
    bad = {i for i in range(n) if got[i]}
    assert {3, 17, 40, 41, 1000, 4999} <= bad and all(idx[i] in (7, 9) or i in (3, 17, 40, 41, 1000, 4999) for i in bad)
    sample = [i for i in range(0, n, 97)] + [3, 40, 41, 1000]
    for i in sample:
        if i != 17:
            assert got[i] == O.verify(msgs[32 * i:32 * i + 32], sigs[64 * i:64 * i + 64], pk_per_item[128 * i:128 * i + 128]), i
    # identity mapping (key_index == NULL): triple i against key i
    m = nk
    m_msgs = synth.messages(m, 32, seed=903)
    m_sigs = E.sign_batch(m_msgs, 32, sks)[0]
    assert cache.verify(m_msgs, 32, m_sigs) == E.verify_batch(m_msgs, 32, m_sigs, pks)
    # more items than one workspace chunk, identity mapping (the second chunk's first item is key 4096, not key 0)
    from bn254_b200._native import Context as _C
    old = os.environ.get("BN254_COOP_CHUNK_LOG2")
    os.environ["BN254_COOP_CHUNK_LOG2"] = "12"
    try:
        small = _C(0)
    finally:
        if old is None:
            del os.environ["BN254_COOP_CHUNK_LOG2"]
        else:
            os.environ["BN254_COOP_CHUNK_LOG2"] = old
    try:
        E.set_input_policy(E.INPUTS_TYPED, ctx=small)
        big = 9000
        bm, bk = synth.messages(big, 32, seed=904), synth.secret_keys(big, seed=905)
        bs = bytearray(E.sign_batch(bm, 32, bk)[0])
        bs[64 * 8000:64 * 8001] = bs[64 * 1:64 * 2]
        bp = E.derive_pk_g2_batch(bk)
        c3 = E.KeyLineCache(bp, ctx=small)
        st3 = c3.verify(bm, 32, bytes(bs))
        assert st3 == E.verify_batch(bm, 32, bytes(bs), bp) and {i for i in range(big) if st3[i]} == {8000}
    finally:
        small.close()
    # untrusted policy: the cache records from_uncompressed statuses (infinity and off-subgroup keys rejected), signatures validated
    strict = Context(0)
    try:
        p2 = bytearray(pks)
        p2[128 * 11:128 * 12] = edge_points.twist_point_outside_g2()
        c2 = E.KeyLineCache(bytes(p2), ctx=strict)
        ks2 = c2.key_status()
        assert ks2[7] == ks2[9] == ks2[11] == O.INVALID_GROUP_POINT and sum(1 for b in ks2 if b) == 3
        pk2_item = b"".join(bytes(p2)[128 * j:128 * j + 128] if j < nk else bytes(128) for j in idx)
        want2 = bytearray(E.verify_batch(msgs, 32, sigs, pk2_item, ctx=strict))
        want2[17] = O.INDEX_OOB
        assert c2.verify(msgs, 32, sigs, idx) == bytes(want2)
        assert want2[40] == O.INVALID_GROUP_POINT
    finally:
        strict.close()


def test_engine_fault_is_reported_and_retried(E):
    """The bounded producer / machine handshake of the pipelined small-batch verify: with the test hook the producer never
    publishes item 5, the machine's wait for it times out, and (a) the device-resident entry point reports status 255
    (BN254_ENGINE_FAULT) for that item and real verdicts for all others, (b) the host-buffer entry point notices the 255, runs the
    batch again without pipelining and returns the oracle's statuses."""
    import ctypes
    import torch
    from bn254_b200._native import Context, S
    n = 40
    msgs, sks, sigs, pks = _signed_set(E, n, seed=77)
    bad = bytearray(sigs)
    bad[64 * 9:64 * 10] = O.g1_neg(bytes(sigs[64 * 9:64 * 10]))[1]
    bad = bytes(bad)
    want = O.verify_batch(msgs, 32, bad, pks, n, NTHREADS)
    assert want[9] == 9 and sum(1 for b in want if b) == 1
    ctx = Context(0)
    try:
        E.set_input_policy(E.INPUTS_TYPED, ctx=ctx)
        retries = ctypes.c_uint32(0)
        ctx.call("bn254_set_test_fault", S(5), retries)
        assert retries.value == 0
        dev = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
        d_st = torch.zeros(n, dtype=torch.uint8, device="cuda")
        ctx.call("bn254_verify_batch_dev", dev(msgs), S(32), dev(bad), dev(pks), S(n), d_st)
        ctx.sync()
        got = bytes(d_st.cpu().numpy().tobytes())
        assert got[5] == 255 and got[:5] + got[6:] == want[:5] + want[6:]
        assert E.verify_batch(msgs, 32, bad, pks, ctx=ctx) == want
        ctx.call("bn254_set_test_fault", S(2 ** 64 - 1), retries)
        assert retries.value == 1
        assert E.verify_batch(msgs, 32, bad, pks, ctx=ctx) == want
        ctx.call("bn254_set_test_fault", S(2 ** 64 - 1), retries)
        assert retries.value == 1
    finally:
        ctx.close()


@pytest.mark.parametrize("groups", [2 * 148 + 1, 4 * 148 + 1, 4 * 148 + 148, 4 * 148 + 149, 4 * 148 + 2 * 148 + 1])
def test_mid_size_launch_shapes(E, groups):
    """Batches of one to two waves of the machine: the cooperative walk as the producer (up to 160 items per SM), four-group
    blocks for the whole waves and one-group blocks (eighteen warps / six warps) for a short remainder.  Verdicts = the oracle's,
    with forged, undecodable and infinite items placed in the first block, at the wave boundary and in the remainder; and the same
    bytes with the remainder split off (BN254_COOP_TAIL_SPLIT=0)."""
    from bn254_b200._native import Context
    n = groups * 32 - 5
    msgs, sks, sigs, pks = _signed_set(E, n, seed=4000 + groups)
    msgs, sigs, pks = bytearray(msgs), bytearray(sigs), bytearray(pks)
    spots = [0, 31, 32, 4 * 148 * 32 - 1, 4 * 148 * 32, 4 * 148 * 32 + 33, n - 40, n - 1]
    for t, i in enumerate(i for i in spots if 0 <= i < n):
        kind = t % 4
        if kind == 0:
            msgs[32 * i + 1] ^= 0x40
        elif kind == 1:
            sigs[64 * i:64 * i + 64] = O.g1_neg(bytes(sigs[64 * i:64 * i + 64]))[1]
        elif kind == 2:
            sigs[64 * i + 63] ^= 1
        else:
            pks[128 * i:128 * i + 128] = bytes(128)
    msgs, sigs, pks = bytes(msgs), bytes(sigs), bytes(pks)
    want = O.verify_batch(msgs, 32, sigs, pks, n, NTHREADS)
    assert sum(1 for b in want if b) >= 4
    assert E.verify_batch(msgs, 32, sigs, pks) == want
    os.environ["BN254_COOP_TAIL_SPLIT"] = "0"
    try:
        whole = Context(0)
    finally:
        del os.environ["BN254_COOP_TAIL_SPLIT"]
    try:
        E.set_input_policy(E.INPUTS_TYPED, ctx=whole)
        assert E.verify_batch(msgs, 32, sigs, pks, ctx=whole) == want
    finally:
        whole.close()


@pytest.mark.parametrize("n", [1, 2, 31, 33, 111 * 32, 111 * 32 + 1, 148 * 32, 148 * 32 + 1])
def test_small_batch_latency_path(E, n):
    """Small batches take the low-latency route (counter-parallel hash, the four-warp cooperative walk as the line producer on its
    own stream, eighteen-warp one-group blocks of the machine consuming the line sets WHILE they are produced): verdicts must be the
    oracle's, with forged, undecodable and infinite items in the batch, under both input policies, and equal to the unpipelined
    route (BN254_PIPELINE=0).  On a 148-SM device 111 * 32 is the last batch that is pipelined, 148 * 32 the last whose groups get
    eighteen-warp blocks, 148 * 32 + 1 the first on six-warp blocks."""
    from bn254_b200._native import Context
    msgs, sks, sigs, pks = _signed_set(E, n, seed=1000 + n)
    msgs, sigs, pks = bytearray(msgs), bytearray(sigs), bytearray(pks)
    rng = random.Random(n)
    marks = {}
    for i in rng.sample(range(n), min(n, 7)):
        kind = rng.randrange(5) if n > 1 else 0
        marks[i] = kind
        if kind == 0:
            msgs[32 * i + 3] ^= 0x10                                  # wrong message
        elif kind == 1:
            j = (i + 1) % n
            sigs[64 * i:64 * i + 64] = sigs[64 * j:64 * j + 64] if j != i else O.g1_neg(bytes(sigs[64 * i:64 * i + 64]))[1]
        elif kind == 2:
            sigs[64 * i + 63] ^= 1                                    # off-curve signature: decode error, no line sets written
        elif kind == 3:
            sigs[64 * i:64 * i + 64] = bytes(64)                      # infinity (typed: skipped pair -> reject; untrusted: decode error)
        else:
            pks[128 * i:128 * i + 32] = be(Q + 2)                     # coordinate >= q
    msgs, sigs, pks = bytes(msgs), bytes(sigs), bytes(pks)
    want = O.verify_batch(msgs, 32, sigs, pks, n, NTHREADS)
    assert E.verify_batch(msgs, 32, sigs, pks) == want
    strict = Context(0)
    os.environ["BN254_PIPELINE"] = "0"
    try:
        plain = Context(0)
    finally:
        del os.environ["BN254_PIPELINE"]
    os.environ["BN254_LINES_WALK4"] = os.environ["BN254_COOP12"] = "0"  # the one-thread latency producer and six-warp blocks
    try:
        legacy = Context(0)
    finally:
        del os.environ["BN254_LINES_WALK4"], os.environ["BN254_COOP12"]
    os.environ["BN254_COOP18"] = "0"                                    # twelve-warp blocks instead of eighteen
    try:
        twelve = Context(0)
    finally:
        del os.environ["BN254_COOP18"]
    try:
        E.set_input_policy(E.INPUTS_TYPED, ctx=plain)
        assert E.verify_batch(msgs, 32, sigs, pks, ctx=plain) == want
        E.set_input_policy(E.INPUTS_TYPED, ctx=legacy)
        assert E.verify_batch(msgs, 32, sigs, pks, ctx=legacy) == want
        E.set_input_policy(E.INPUTS_TYPED, ctx=twelve)
        assert E.verify_batch(msgs, 32, sigs, pks, ctx=twelve) == want
        want_strict = bytes(_untrusted_expect(msgs[32 * i:32 * i + 32], sigs[64 * i:64 * i + 64], pks[128 * i:128 * i + 128]) for i in range(min(n, 64)))
        assert E.verify_batch(msgs, 32, sigs, pks, ctx=strict)[:len(want_strict)] == want_strict
    finally:
        strict.close()
        plain.close()
        legacy.close()
        twelve.close()
