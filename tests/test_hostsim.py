"""CPU check of the engine's per-item device code (compiled by g++ in tests/hostsim) against the oracle:
tower formulas, Miller schedule with the -G2 line table, final-exponentiation chain, hash loop, group law
and codecs.  The GPU tests repeat these through the real kernels; this catches logic errors without a GPU."""
import ctypes
import json
import os
import random
import subprocess

import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS_DIR = os.path.join(ROOT, "tests", "hostsim")
G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
H = bytes.fromhex
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
G1_GEN = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")


@pytest.fixture(scope="module")
def hs():
    subprocess.check_call(["make", "-C", HS_DIR], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(os.path.join(HS_DIR, "libhostsim.so"))


def buf(n):
    return ctypes.create_string_buffer(n)


def be(x, n=32):
    return x.to_bytes(n, "big")


def test_fq_ops(hs):
    rng = random.Random(1)
    cases = [(0, 0), (1, 1), (Q - 1, Q - 1), (Q - 1, 1), (0, Q - 1)] + [(rng.randrange(Q), rng.randrange(Q)) for _ in range(300)]
    for a, b in cases:
        for op in (0, 1, 2):
            out = buf(32)
            assert hs.hs_fq_op(op, be(a), be(b), out) == 0
            assert out.raw == O.fq_op(op, be(a), be(b))[1], (op, a, b)
    for a, _ in cases[:40]:
        for op in (3, 4):
            out = buf(32)
            st = hs.hs_fq_op(op, be(a), be(0), out)
            est, e = O.fq_op(op, be(a))
            assert st == est and (st != 0 or out.raw == e)


def rand_fq12(rng):
    return b"".join(be(rng.randrange(Q)) for _ in range(12))


def test_fq12_ops(hs):
    rng = random.Random(2)
    for _ in range(10):
        a, b = rand_fq12(rng), rand_fq12(rng)
        for op in range(8):
            out = buf(384)
            assert hs.hs_fq12_op(op, a, b, out) == 0
            assert out.raw == O.fq12_op(op, a, b)[1], op


def test_hash_kats_and_random(hs):
    for v in G["hash_to_g1"]:
        out, ctr = buf(64), ctypes.c_int(-1)
        assert hs.hs_hash_to_g1(H(v["msg"]), len(H(v["msg"])), out, ctypes.byref(ctr)) == 0
        assert O.g1_compress(out.raw) == (0, H(v["compressed"]))
    rng = random.Random(3)
    for _ in range(300):
        msg = rng.randbytes(rng.choice([0, 1, 3, 31, 32, 33, 54, 55, 56, 62, 63, 64, 65, 118, 119, 120, 127, 128, 300]))
        out, ctr = buf(64), ctypes.c_int(-1)
        st = hs.hs_hash_to_g1(msg, len(msg), out, ctypes.byref(ctr))
        est, e, ectr = O.hash_to_g1(msg)
        assert (st, out.raw, ctr.value) == (est, e, ectr)


def test_hash_to_point_error_exit(hs):
    """/root/reference/src/hash.rs:62: all counters fail -> HashToPointError.  Unreachable with 255 counters (2^-235), so the
    device code takes the counter limit as a parameter (bn254_set_hash_try_limit): a message whose accepted counter is c must
    fail with every limit <= c and succeed with c + 1, and a search that starts at counter c' > c must find the NEXT accepted
    counter (what a lane of the counter-parallel kernel does)."""
    rng = random.Random(31)
    seen_fail = 0
    for _ in range(60):
        msg = rng.randbytes(rng.choice([5, 32, 60, 64, 100]))
        est, e, c = O.hash_to_g1(msg)
        assert est == 0
        out, ctr = buf(64), ctypes.c_int(-1)
        if c > 0:
            assert hs.hs_hash_to_g1_limited(msg, len(msg), out, ctypes.byref(ctr), c, 0) == O.HASH_TO_POINT and out.raw == bytes(64)
            seen_fail += 1
        assert hs.hs_hash_to_g1_limited(msg, len(msg), out, ctypes.byref(ctr), c + 1, 0) == 0 and (out.raw, ctr.value) == (e, c)
        assert hs.hs_hash_to_g1_limited(msg, len(msg), out, ctypes.byref(ctr), c + 1, c) == 0 and (out.raw, ctr.value) == (e, c)
        st = hs.hs_hash_to_g1_limited(msg, len(msg), out, ctypes.byref(ctr), 255, c + 1)
        assert st == 0 and ctr.value > c and out.raw != e
    assert seen_fail > 20


def test_sign_kat_and_group_vectors(hs):
    for v in G["sign"]:
        sig = buf(64)
        assert hs.hs_sign(H(v["msg"]), len(H(v["msg"])), H(v["sk"]), sig) == 0
        assert O.g1_compress(sig.raw) == (0, H(v["sig_compressed"]))
    for v in G["sk_to_pk_g2"]:
        out = buf(128)
        hs.hs_derive_pk_g2(H(v["sk"]), out)
        assert out.raw == H(v["pk_uncompressed"])
    for v in G["bn256_json"]["add"]:
        out = buf(64)
        assert hs.hs_g1_add(H(v["x1"]) + H(v["y1"]), H(v["x2"]) + H(v["y2"]), out) == 0
        assert out.raw == H(v["result"])
        if any(H(v["x2"]) + H(v["y2"])):
            assert hs.hs_g1_madd(H(v["x1"]) + H(v["y1"]), H(v["x2"]) + H(v["y2"]), out) == 0
            assert out.raw == H(v["result"])
    for v in G["bn256_json"]["mul"]:
        out = buf(64)
        assert hs.hs_g1_mul(H(v["x"]) + H(v["y"]), H(v["scalar"]), out) == 0
        assert out.raw == H(v["result"])
    for sk in G["example"]["sks"]:  # keys > r
        o1, o2 = buf(64), buf(128)
        hs.hs_derive_pk_g1(H(sk), o1)
        hs.hs_derive_pk_g2(H(sk), o2)
        assert o1.raw == O.derive_pk_g1(H(sk))[1] and o2.raw == O.derive_pk_g2(H(sk))[1]


def test_g2_group(hs):
    rng = random.Random(4)
    g2 = O.derive_pk_g2(be(1))[1]
    pts = [O.derive_pk_g2(be(rng.randrange(1, R)))[1] for _ in range(4)] + [bytes(128), g2]
    for a in pts:
        for b in pts:
            out = buf(128)
            assert hs.hs_g2_add(a, b, out) == 0 and out.raw == O.g2_add(a, b)[1]
            if any(b):
                assert hs.hs_g2_madd(a, b, out) == 0 and out.raw == O.g2_add(a, b)[1]
    neg = O.g2_neg(pts[0])[1]
    out = buf(128)
    assert hs.hs_g2_add(pts[0], neg, out) == 0 and out.raw == bytes(128)
    assert hs.hs_g2_madd(pts[0], neg, out) == 0 and out.raw == bytes(128)
    k = be(rng.randrange(1 << 256))
    assert hs.hs_g2_mul(pts[1], k, out) == 0 and out.raw == O.g2_mul(pts[1], k)[1]


def test_verify_and_miller(hs):
    rng = random.Random(5)
    for t in range(6):
        sk = be(rng.randrange(1, R))
        msg = rng.randbytes(32)
        sig = O.sign(msg, sk)[1]
        pk = O.derive_pk_g2(sk)[1]
        assert hs.hs_verify(msg, len(msg), sig, pk) == 0
        # Miller product is bit-identical to the oracle's (same field values, canonical form)
        h = O.hash_to_g1(msg)[1]
        neg_g2 = O.g2_neg(O.derive_pk_g2(be(1))[1])[1]
        f = buf(384)
        assert hs.hs_verify_miller(msg, len(msg), sig, pk, f) == 0
        assert f.raw == O.miller_product(h + sig, pk + neg_g2, 2)[1]
        gt = buf(384)
        assert hs.hs_final_exp(f.raw, gt) == 0 and gt.raw == O.final_exp(f.raw)[1]
        # adversarial
        bad = O.g1_add(sig, G1_GEN)[1]
        assert hs.hs_verify(msg, len(msg), bad, pk) == O.VERIFICATION_FAILED == O.verify(msg, bad, pk)
        assert hs.hs_verify(msg + b"x", len(msg) + 1, sig, pk) == O.VERIFICATION_FAILED
        assert hs.hs_verify(msg, len(msg), O.g1_neg(sig)[1], pk) == O.VERIFICATION_FAILED
    assert hs.hs_verify(b"m", 1, bytes(64), bytes(128)) == 0 == O.verify(b"m", bytes(64), bytes(128))
    assert hs.hs_verify(b"m", 1, bytes(64), pk) == O.VERIFICATION_FAILED
    assert hs.hs_verify(b"m", 1, sig, bytes(128)) == O.VERIFICATION_FAILED == O.verify(b"m", sig, bytes(128))
    assert hs.hs_verify(b"m", 1, be(1) + be(3), pk) == O.INVALID_GROUP_POINT == O.verify(b"m", be(1) + be(3), pk)
    assert hs.hs_verify(b"m", 1, be(Q) + be(3), pk) == O.NOT_MEMBER == O.verify(b"m", be(Q) + be(3), pk)
    for v in G["check_public_keys_ok"]:
        assert hs.hs_check_public_keys(O.derive_pk_g2(H(v["sk"]))[1], O.derive_pk_g1(H(v["sk"]))[1]) == 0
    for v in G["check_public_keys_fail"]:
        assert hs.hs_check_public_keys(O.derive_pk_g2(H(v["sk_g2"]))[1], O.derive_pk_g1(H(v["sk_g1"]))[1]) == O.VERIFICATION_FAILED


def test_pairing_pairs(hs):
    rng = random.Random(6)
    a, b = rng.randrange(1, R), rng.randrange(1, R)
    pa = O.derive_pk_g1(be(a))[1]
    qb = O.derive_pk_g2(be(b))[1]
    pab = O.derive_pk_g1(be(a * b % R))[1]
    neg_g2 = O.g2_neg(O.derive_pk_g2(be(1))[1])[1]
    g1s, g2s = pa + pab + bytes(64), qb + neg_g2 + qb
    assert hs.hs_pairing_check(g1s, g2s, 3) == 0 == O.pairing_check(g1s, g2s, 3)[0]
    f = buf(384)
    assert hs.hs_miller_product(g1s, g2s, 3, f) == 0 and f.raw == O.miller_product(g1s, g2s, 3)[1]
    assert hs.hs_pairing_check(pa + pa, qb + neg_g2, 2) == O.VERIFICATION_FAILED
    assert hs.hs_pairing_check(b"", b"", 0) == 0


def test_codecs(hs):
    rng = random.Random(8)
    for v in G["g2_compressed_roundtrip"]:
        out = buf(128)
        assert hs.hs_g2_decompress(H(v["compressed"]), 65, out) == 0
        assert out.raw == O.g2_decompress(H(v["compressed"]))[1]
        c = buf(65)
        assert hs.hs_g2_compress(out.raw, c) == 0 and c.raw == H(v["compressed"])
    for _ in range(3):
        pk = O.derive_pk_g2(be(rng.randrange(1, R)))[1]
        c = buf(65)
        assert hs.hs_g2_compress(pk, c) == 0 and c.raw == O.g2_compress(pk)[1]
        out = buf(128)
        assert hs.hs_g2_decompress(c.raw, 65, out) == 0 and out.raw == pk
        assert hs.hs_g2_validate(pk, 128) == 0
        bad = bytes([c.raw[0] ^ 1]) + c.raw[1:]
        assert hs.hs_g2_decompress(bad, 65, out) == 0 and out.raw == O.g2_neg(pk)[1]
        assert hs.hs_g2_decompress(b"\x0c" + c.raw[1:], 65, out) == O.INVALID_ENCODING
    for _ in range(20):
        p = O.derive_pk_g1(be(rng.randrange(1, R)))[1]
        c = buf(33)
        assert hs.hs_g1_compress(p, c) == 0 and c.raw == O.g1_compress(p)[1]
        out = buf(64)
        assert hs.hs_g1_decompress(c.raw, 33, out) == 0 and out.raw == p
        x = be(rng.randrange(Q))
        st = hs.hs_g1_decompress(b"\x03" + x, 33, out)
        est, e = O.g1_decompress(b"\x03" + x)
        assert st == est and out.raw == e
    assert hs.hs_g1_decompress(b"\x02" + be(Q), 33, buf(64)) == O.NOT_MEMBER
    assert hs.hs_g1_decompress(b"\x04" + be(1), 33, buf(64)) == O.INVALID_ENCODING
    assert hs.hs_g1_validate(be(1) + be(3), 64) == O.INVALID_GROUP_POINT
    assert hs.hs_g1_compress(bytes(64), buf(33)) == O.POINT_IN_JACOBIAN
    # twist point outside the r-torsion
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
    assert hs.hs_g2_validate(raw, 128) == O.INVALID_GROUP_POINT == O.g2_validate_uncompressed(raw)
    import edge_points
    for pt, inside in edge_points.subgroup_edge_points():   # small-order components of the twist cofactor, mixed points, cofactor-cleared points
        want = 0 if inside else O.INVALID_GROUP_POINT
        assert hs.hs_g2_validate(pt, 128) == want == O.g2_validate_uncompressed(pt)


# ---------------------------------------------------------------------------------------------- cooperative machine
@pytest.fixture(params=[0, 1, 2], ids=["block-layout", "warp-local-layout", "half-warp-layout"])
def layout(hs, request):
    """All shared-memory layouts of the machine run the same programs: six warps per 32 items (k_coop_run / k_coop4_run), six
    lanes per item with five items per warp (k_coopw_run), three warps per 16 items (k_cooph_run)."""
    hs.hs_coop_set_layout(request.param)
    yield request.param
    hs.hs_coop_set_layout(0)


def test_coop_plans_match_tower(hs, layout):
    """One plan of the six-warp machine == the tower formula of the oracle (product, square, cyclotomic square)."""
    rng = random.Random(11)
    for _ in range(6):
        a, b = rand_fq12(rng), rand_fq12(rng)
        out = buf(384)
        assert hs.hs_coop_plan(0, a, b, out) == 0 and out.raw == O.fq12_op(0, a, b)[1]   # P <- S * P
        assert hs.hs_coop_plan(1, a, None, out) == 0 and out.raw == O.fq12_op(1, a)[1]   # P <- P^2
    # cyclotomic squaring is only defined on the cyclotomic subgroup: use final-exponentiation outputs
    for _ in range(3):
        g = O.final_exp(rand_fq12(rng))[1]
        out = buf(384)
        assert hs.hs_coop_plan(4, g, None, out) == 0 and out.raw == O.fq12_op(3, g)[1] == O.fq12_op(1, g)[1]
    edge = be(Q - 1) * 12
    out = buf(384)
    assert hs.hs_coop_plan(0, edge, edge, out) == 0 and out.raw == O.fq12_op(0, edge, edge)[1]
    assert hs.hs_coop_plan(1, edge, None, out) == 0 and out.raw == O.fq12_op(1, edge)[1]


def test_coop_final_exp(hs, layout):
    rng = random.Random(12)
    for _ in range(3):
        f = rand_fq12(rng)
        gt = buf(384)
        st = hs.hs_coop_final_exp(f, gt)
        assert gt.raw == O.final_exp(f)[1]
        assert st == O.VERIFICATION_FAILED  # a random value does not map to one


def test_coop_verify_matches_oracle(hs, layout):
    rng = random.Random(13)
    neg_g2 = O.g2_neg(O.derive_pk_g2(be(1))[1])[1]
    for t in range(4):
        sk = be(rng.randrange(1, R))
        msg = rng.randbytes(32)
        sig = O.sign(msg, sk)[1]
        pk = O.derive_pk_g2(sk)[1]
        f = buf(384)
        assert hs.hs_coop_verify_miller(msg, len(msg), sig, pk, f) == 0
        assert f.raw == O.miller_product(O.hash_to_g1(msg)[1] + sig, pk + neg_g2, 2)[1]
        assert hs.hs_coop_verify(msg, len(msg), sig, pk, 0) == 0
        bad = O.g1_add(sig, G1_GEN)[1]
        assert hs.hs_coop_verify(msg, len(msg), bad, pk, 0) == O.VERIFICATION_FAILED
        assert hs.hs_coop_verify(msg + b"x", len(msg) + 1, sig, pk, 0) == O.VERIFICATION_FAILED
    # infinity semantics of bn::pairing_batch (pairs holding an infinity are skipped) and decode errors
    assert hs.hs_coop_verify(b"m", 1, bytes(64), bytes(128), 0) == 0 == O.verify(b"m", bytes(64), bytes(128))
    assert hs.hs_coop_verify(b"m", 1, bytes(64), pk, 0) == O.VERIFICATION_FAILED
    assert hs.hs_coop_verify(b"m", 1, sig, bytes(128), 0) == O.VERIFICATION_FAILED
    assert hs.hs_coop_verify(b"m", 1, be(1) + be(3), pk, 0) == O.INVALID_GROUP_POINT
    for v in G["check_public_keys_ok"]:
        assert hs.hs_coop_verify(b"", 0, O.derive_pk_g1(H(v["sk"]))[1], O.derive_pk_g2(H(v["sk"]))[1], 1) == 0
    for v in G["check_public_keys_fail"]:
        assert hs.hs_coop_verify(b"", 0, O.derive_pk_g1(H(v["sk_g1"]))[1], O.derive_pk_g2(H(v["sk_g2"]))[1], 1) == O.VERIFICATION_FAILED


def test_cooperative_walk_matches_one_thread_walk(hs):
    """The four-warp producer of small batches (coop_lines.cuh walk_*) writes the SAME 174 line sets, bit for bit, as the
    one-thread-per-item producer, returns the same decode statuses, and the machine run on its sets agrees with the oracle."""
    import ctypes
    rng = random.Random(29)
    st = ctypes.c_int(0)
    sk = be(rng.randrange(1, R))
    msg = rng.randbytes(32)
    sig, pk = O.sign(msg, sk)[1], O.derive_pk_g2(sk)[1]
    for t in range(3):
        sk = be(rng.randrange(1, R))
        msg = rng.randbytes(1 + 40 * t)
        sig, pk = O.sign(msg, sk)[1], O.derive_pk_g2(sk)[1]
        assert hs.hs_walk4_matches(msg, len(msg), sig, pk, ctypes.byref(st)) == 0 and st.value == 0
    assert hs.hs_coop_verify_walk4(msg, len(msg), sig, pk) == 0
    assert hs.hs_coop_verify_walk4(msg, len(msg), O.g1_add(sig, G1_GEN)[1], pk) == O.VERIFICATION_FAILED
    # infinities (constant line sets) and decode errors
    for s_, p_ in ((bytes(64), bytes(128)), (bytes(64), pk), (sig, bytes(128))):
        assert hs.hs_walk4_matches(msg, len(msg), s_, p_, ctypes.byref(st)) == 0 and st.value == 0
        assert hs.hs_coop_verify_walk4(msg, len(msg), s_, p_) == O.verify(msg, s_, p_)
    assert hs.hs_walk4_matches(msg, len(msg), be(1) + be(3), pk, ctypes.byref(st)) == 0 and st.value == O.INVALID_GROUP_POINT
    bad_pk = pk[:127] + bytes([pk[127] ^ 1])
    assert hs.hs_walk4_matches(msg, len(msg), sig, bad_pk, ctypes.byref(st)) == 0 and st.value == O.verify(msg, sig, bad_pk) != 0


def test_latency_layouts_of_the_machine(hs):
    """coop_run_block12<2> / <3> (twelve / eighteen warps per group: the Karatsuba components of a row on different warps, exchanged
    through shared memory) run as host threads with a std::barrier: verdicts = the oracle's."""
    rng = random.Random(41)
    sk = be(rng.randrange(1, R))
    msg = rng.randbytes(32)
    sig, pk = O.sign(msg, sk)[1], O.derive_pk_g2(sk)[1]
    bad = O.g1_add(sig, G1_GEN)[1]
    for split in (2, 3):
        assert hs.hs_coop_verify_split(msg, len(msg), sig, pk, split) == 0
        assert hs.hs_coop_verify_split(msg, len(msg), bad, pk, split) == O.VERIFICATION_FAILED
        assert hs.hs_coop_verify_split(msg + b"!", len(msg) + 1, sig, pk, split) == O.VERIFICATION_FAILED
        assert hs.hs_coop_verify_split(msg, len(msg), bytes(64), bytes(128), split) == 0   # both pairs skipped
        assert hs.hs_coop_verify_split(msg, len(msg), sig, bytes(128), split) == O.VERIFICATION_FAILED


def test_coop_multi_pairing_program(hs):
    """COOP_MULTI_K pairs per lane share one squaring chain, then the 32 lanes are multiplied by a butterfly: the block's
    product equals the oracle's Miller product over all pairs (canonical field values, any multiplication order)."""
    rng = random.Random(17)
    n = 41  # lanes 0..31 hold stream 0, lanes 0..8 also stream 1; the other slots are padding (constant-1 lines)
    g1s = b"".join(O.derive_pk_g1(be(rng.randrange(1, R)))[1] for _ in range(n))
    g2s = b"".join(O.derive_pk_g2(be(rng.randrange(1, R)))[1] for _ in range(n))
    g1s = g1s[:64 * 5] + bytes(64) + g1s[64 * 6:]      # one pair with an infinite G1 point: skipped
    f = buf(384)
    assert hs.hs_coop_multi_miller(g1s, g2s, n, f) == 0
    assert f.raw == O.miller_product(g1s, g2s, n)[1]


def test_coop_multi_pairing_programs_all_k(hs):
    """The shorter multi-pairing programs (4, 2, 1 pairs per lane: the re-cut last wave of a big multi-pairing) give the
    oracle's Miller product too."""
    rng = random.Random(18)
    for which, mk in ((6, 4), (7, 2), (8, 1)):
        n = 32 * mk - 5
        g1s = b"".join(O.derive_pk_g1(be(rng.randrange(1, R)))[1] for _ in range(n))
        g2s = b"".join(O.derive_pk_g2(be(rng.randrange(1, R)))[1] for _ in range(n))
        f = buf(384)
        assert hs.hs_coop_multi_miller_k(which, g1s, g2s, n, f) == 0
        assert f.raw == O.miller_product(g1s, g2s, n)[1], mk


def test_coop_finish_program(hs):
    """Program FINISH (bn254_finish_distinct_dev): Miller value of (sig, -G2) times the exchanged product, final
    exponentiation, verdict -- against the oracle's verdict for the same aggregate, accept and reject."""
    rng = random.Random(19)
    n = 3
    msgs = [bytes([i]) * 32 for i in range(n)]
    sks = [be(rng.randrange(1, R)) for _ in range(n)]
    hs_pts = b"".join(O.hash_to_g1(m)[1] for m in msgs)
    pks = b"".join(O.derive_pk_g2(k)[1] for k in sks)
    sigs = [O.sign(m, k)[1] for m, k in zip(msgs, sks)]
    agg = bytes(64)
    for s in sigs:
        agg = O.g1_add(agg, s)[1]
    part = O.miller_product(hs_pts, pks, n)[1]
    assert hs.hs_coop_finish(part, agg) == 0
    assert hs.hs_coop_finish(part, O.g1_add(agg, sigs[0])[1]) == O.VERIFICATION_FAILED
    assert hs.hs_coop_finish(part, bytes(64)) == O.VERIFICATION_FAILED        # signature pair skipped: product != 1
    one = be(1) + bytes(352)
    assert hs.hs_coop_finish(one, bytes(64)) == 0                             # empty product
    # the rank-local form: the (sum sig, -G2) pair already inside the partial, nothing to add at the finish
    neg_g2 = O.g2_neg(O.derive_pk_g2(be(1))[1])[1]
    full = O.miller_product(hs_pts + agg, pks + neg_g2, n + 1)[1]
    assert hs.hs_coop_finish(full, bytes(64)) == 0


def test_fixed_base_tables(hs):
    """G * sk through the 64 x 15 fixed-base tables == the ladder == the oracle, incl. the reference's key KATs, keys > r
    (reduced mod r like Fr::from_slice), 0, 1, r - 1 and scalars with zero nibbles."""
    for v in G["sk_to_pk_g2"]:
        out = buf(128)
        hs.hs_derive_pk_g2_comb(H(v["sk"]), out)
        assert out.raw == H(v["pk_uncompressed"])
    rng = random.Random(23)
    sks = [be(0), be(1), be(R - 1), be(R), be(R + 5), be((1 << 256) - 1), be(0x10000000000000000000000000000f00), be(15 << 252)]
    sks += [H(s) for s in G["example"]["sks"]] + [be(rng.randrange(1 << 256)) for _ in range(12)]
    for sk in sks:
        o1, o2 = buf(64), buf(128)
        hs.hs_derive_pk_g1_comb(sk, o1)
        hs.hs_derive_pk_g2_comb(sk, o2)
        assert o1.raw == O.derive_pk_g1(sk)[1] and o2.raw == O.derive_pk_g2(sk)[1], sk.hex()


def test_rlc_prepare_item(hs):
    """One item's share of the randomised batch check: c * H(m) and c * sig for a 128-bit coefficient (zero is replaced by one),
    and the conditions under which an item may not ride in the batch (bad encodings, a key outside G2)."""
    import edge_points
    rng = random.Random(23)
    sk = be(rng.randrange(1, R))
    msg = rng.randbytes(32)
    sig, pk, h = O.sign(msg, sk)[1], O.derive_pk_g2(sk)[1], O.hash_to_g1(msg)[1]
    for c in (rng.randrange(1, 1 << 128), 1, (1 << 128) - 1, 0, 1 << 127, 0xf):
        hs_out, sc_out = buf(64), buf(64)
        assert hs.hs_rlc_prepare(msg, len(msg), sig, pk, be(c, 16), 1, hs_out, sc_out) == 0
        c = c or 1
        lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd   # the coefficient is lo + hi * lambda (GLV-shaped, one 64-bit ladder)
        k = be(((c & ((1 << 64) - 1)) + (c >> 64) * lam) % R)
        assert hs_out.raw == O.g1_mul(h, k)[1] and sc_out.raw == O.g1_mul(sig, k)[1]
    hs_out, sc_out = buf(64), buf(64)
    assert hs.hs_rlc_prepare(msg, len(msg), bytes(64), bytes(128), be(5, 16), 1, hs_out, sc_out) == 0 and sc_out.raw == bytes(64)
    assert hs.hs_rlc_prepare(msg, len(msg), be(1) + be(3), pk, be(5, 16), 1, hs_out, sc_out) == O.INVALID_GROUP_POINT
    assert hs.hs_rlc_prepare(msg, len(msg), sig, pk[:127] + bytes([pk[127] ^ 1]), be(5, 16), 1, hs_out, sc_out) == O.INVALID_GROUP_POINT
    outside = [pt for pt, inside in edge_points.subgroup_edge_points() if not inside][0]
    assert hs.hs_rlc_prepare(msg, len(msg), sig, outside, be(5, 16), 1, hs_out, sc_out) == O.INVALID_GROUP_POINT
    assert hs.hs_rlc_prepare(msg, len(msg), sig, outside, be(5, 16), 0, hs_out, sc_out) == 0   # the caller vouched for the key


def test_sign_glv_edge_scalars(hs):
    """Signing multiplies the hash point by k = k1 + k2 * lambda (GLV, csrc/curve.cuh g1_mul_glv): same group element, hence the
    same bytes as the oracle's double-and-add, for the scalars where the decomposition changes sign or hits its bounds."""
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    a1, a2 = 0x89d3256894d213e3, 0x6f4d8248eeb859fd0be4e1541221250b
    rng = random.Random(29)
    ks = [0, 1, 2, 15, 16, R - 1, R - 2, R, R + 1, (1 << 256) - 1, lam, lam - 1, lam + 1, R - lam, a1, a2, a1 * a2 % R, (1 << 128) - 1, 1 << 128,
          (1 << 253), (1 << 254) - 1, R // 2, R // 2 + 1, R // 3]
    ks += [rng.randrange(1 << 256) for _ in range(40)]
    for k in ks:
        msg = rng.randbytes(rng.randrange(0, 70))
        out = buf(64)
        assert hs.hs_sign(msg, len(msg), be(k), out) == 0
        assert out.raw == O.sign(msg, be(k))[1], hex(k)


def test_glv_decomposition_identity_and_bounds(hs):
    """k = k1 + k2 * lambda (mod r) with |k1|, |k2| < 2^128 (32 four-bit windows) for every scalar the signing path can see:
    random 256-bit inputs (reduced mod r first) and the values around the lattice's corners."""
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    a1, b1n, a2, b2 = 0x89d3256894d213e3, 0x6f4d8248eeb859fc8211bbeb7d4f1128, 0x6f4d8248eeb859fd0be4e1541221250b, 0x89d3256894d213e3
    assert (lam * lam + lam + 1) % R == 0 and (a1 - b1n * lam) % R == 0 and (a2 + b2 * lam) % R == 0
    rng = random.Random(31)
    ks = [0, 1, R - 1, R, R + 1, (1 << 256) - 1, lam, R - lam, a1, a2, b1n, a1 * a2 % R, R // 2, R // 2 + 1]
    ks += [(i * a1 + j * a2) % R for i in range(-2, 3) for j in range(-2, 3)]
    ks += [rng.randrange(1 << 256) for _ in range(20000)]
    k1, k2, sg = (ctypes.c_uint32 * 4)(), (ctypes.c_uint32 * 4)(), (ctypes.c_uint8 * 2)()
    worst = 0
    for k in ks:
        hs.hs_glv_decompose(be(k), k1, k2, sg)
        v1 = sum(int(k1[i]) << (32 * i) for i in range(4)) * (-1 if sg[0] else 1)
        v2 = sum(int(k2[i]) << (32 * i) for i in range(4)) * (-1 if sg[1] else 1)
        assert (v1 + v2 * lam - k) % R == 0, hex(k)
        worst = max(worst, abs(v1).bit_length(), abs(v2).bit_length())
    assert worst <= 128   # the magnitudes are returned in four limbs, so anything larger would already have failed the identity


def test_mul9_add_one_reduction(hs):
    """(9 x + z) mod q through the quotient estimate (fq.cuh fq_mul9_add, used for xi * x in the cooperative machine): exact
    at every multiple of q the sum can cross, at the operand extremes, and on random operands."""
    rng = random.Random(37)
    cases = [(0, 0), (0, Q), (Q - 1, Q), (Q - 1, Q - 1), (Q - 1, 0), (1, Q - 9), (1, Q - 10), (0, Q - 1)]
    for k in range(1, 10):                      # 9 x + z = k q + d for small d on both sides
        for d in (-2, -1, 0, 1, 2):
            t = k * Q + d
            x = min(Q - 1, t // 9)
            z = t - 9 * x
            if 0 <= z <= Q:
                cases.append((x, z))
    cases += [(rng.randrange(Q), rng.randrange(Q + 1)) for _ in range(20000)]
    out = buf(32)
    for x, z in cases:
        hs.hs_mul9_add(be(x), be(z), out)
        assert int.from_bytes(out.raw, "big") == (9 * x + z) % Q, (hex(x), hex(z))
    cases2 = [(0, 0), (0, Q), (Q - 1, Q), (Q - 1, Q - 1), (Q - 1, 0)]
    for k in range(1, 5):                       # 3 t + 2 z = k q + d
        for d in (-2, -1, 0, 1, 2):
            v = k * Q + d
            for t in (min(Q - 1, v // 3), max(0, (v - 2 * Q + 2) // 3)):
                if (v - 3 * t) % 2 == 0 and 0 <= (v - 3 * t) // 2 <= Q and 0 <= t < Q:
                    cases2.append((t, (v - 3 * t) // 2))
    cases2 += [(rng.randrange(Q), rng.randrange(Q + 1)) for _ in range(20000)]
    for t, z in cases2:
        hs.hs_3t_2z(be(t), be(z), out)
        assert int.from_bytes(out.raw, "big") == (3 * t + 2 * z) % Q, (hex(t), hex(z))
