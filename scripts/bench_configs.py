#!/usr/bin/env python3
"""Secondary measurements for the other BASELINE.json configs (3: hash + sign, 4: same-message aggregate, 5: distinct-message
aggregate) and for the ingest path (compressed-key decode with the r-torsion check).  One JSON line per config; device-resident
inputs where a *_dev entry point exists, otherwise the host-buffer call (copies included).  Not the bench line: bench.py is."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import synth
from bn254_b200 import engine as E
from bn254_b200._native import I, S


def timed(ctx, fn, reps=3):
    fn()
    ctx.sync()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    ctx = E.context(0)
    E.set_input_policy(E.INPUTS_TYPED, ctx=ctx)  # values of the crate's types, like the bench line
    dev = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    msgs, sks = synth.messages(n, 32, seed=1), synth.secret_keys(n, seed=2)
    d_msgs, d_sks = dev(msgs), dev(sks)
    d_sigs = torch.empty(64 * n, dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    dt = timed(ctx, lambda: ctx.call("bn254_sign_batch_dev", d_msgs, S(32), d_sks, S(n), d_sigs, d_st))
    print(json.dumps({"config": "3: hash_to_g1 + sign, distinct keys", "n": n, "signs_per_sec": n / dt, "ms": dt * 1e3}))
    d_h = torch.empty(64 * n, dtype=torch.uint8, device="cuda")
    dt = timed(ctx, lambda: ctx.call("bn254_hash_to_g1_batch_dev", d_msgs, S(32), S(n), d_h, d_st))
    print(json.dumps({"config": "3a: hash_to_g1 only", "n": n, "hashes_per_sec": n / dt, "ms": dt * 1e3}))
    sigs = bytes(d_sigs.cpu().numpy().tobytes())
    t = time.perf_counter()
    pks = E.derive_pk_g2_batch(sks, ctx=ctx)
    dt = time.perf_counter() - t
    print(json.dumps({"config": "key derivation G2 * sk (host buffers)", "n": n, "keys_per_sec": n / dt, "ms": dt * 1e3}))
    d_pks = dev(pks)
    o64, o128, st1 = torch.empty(64, dtype=torch.uint8, device="cuda"), torch.empty(128, dtype=torch.uint8, device="cuda"), torch.empty(4, dtype=torch.uint8, device="cuda")
    dt1 = timed(ctx, lambda: ctx.call("bn254_g1_sum_dev", d_sigs, None, S(n), o64, st1))
    dt2 = timed(ctx, lambda: ctx.call("bn254_g2_sum_dev", d_pks, None, S(n), o128, st1))
    print(json.dumps({"config": "4: G1 / G2 aggregation", "n": n, "g1_points_per_sec": n / dt1, "g2_points_per_sec": n / dt2, "g1_ms": dt1 * 1e3, "g2_ms": dt2 * 1e3}))
    msg0 = msgs[:32]
    sk_sig, st = E.sign_batch(msg0 * 4096, 32, sks[:32 * 4096], ctx=ctx)
    t = time.perf_counter()
    v = E.aggregate_verify_same_msg(msg0, sk_sig, pks[:128 * 4096], ctx=ctx)
    print(json.dumps({"config": "4: same-message aggregate verify, 4096 signers, host buffers", "status": v, "ms": (time.perf_counter() - t) * 1e3}))
    m5 = n
    d_f = torch.empty(384, dtype=torch.uint8, device="cuda")
    dt = timed(ctx, lambda: ctx.call("bn254_miller_partial_distinct_dev", d_msgs, S(32), d_pks, S(m5), d_f, st1), reps=1)
    agg, _ = E.g1_sum(sigs[:64 * m5], ctx=ctx)
    v = E.finish_distinct(bytes(d_f.cpu().numpy().tobytes()), agg, ctx=ctx)
    print(json.dumps({"config": "5: distinct-message aggregate verify (Miller partial of %d pairs + shared final exp)" % m5, "pairs_per_sec": m5 / dt, "ms": dt * 1e3, "status": v}))
    # randomised batch verification (additional entry point): all-valid batch, device-resident inputs
    import ctypes
    d_c = dev(synth.rand_bytes(77, 16 * n))
    fast = ctypes.c_int(0)
    for flags, label in ((0, "with the r-torsion test of every key"), (1, "keys vouched for by the caller")):
        dt = timed(ctx, lambda: ctx.call("bn254_verify_batch_rlc_dev", d_msgs, S(32), d_sigs, d_pks, S(n), d_c, I(flags), d_st, fast), reps=2)
        assert fast.value == 1 and not d_st.any().item()
        print(json.dumps({"config": "randomised batch verify (one shared final exponentiation), " + label, "n": n, "verifies_per_sec": n / dt, "ms": dt * 1e3}))
    # the same with ONE forged signature in the batch: its slice (1/64 of a 2^20-triple chunk) is redone by the exact path
    d_bad = d_sigs.clone()
    d_bad[64 * 12345:64 * 12346] = d_sigs[64 * 12346:64 * 12347]
    dt = timed(ctx, lambda: ctx.call("bn254_verify_batch_rlc_dev", d_msgs, S(32), d_bad, d_pks, S(n), d_c, I(1), d_st, fast), reps=2)
    assert fast.value == 0 and int(d_st.count_nonzero().item()) == 1 and int(d_st[12345].item()) == 9
    print(json.dumps({"config": "randomised batch verify, one forged signature among the triples (failing slice redone exactly), keys vouched for", "n": n, "verifies_per_sec": n / dt, "ms": dt * 1e3}))
    m6 = min(n, 1 << 18)
    comp, st = E.g2_compress_batch(pks[:128 * m6], ctx=ctx)
    raw, st = E.g2_decompress_batch(comp, ctx=ctx)  # warm
    assert raw == pks[:128 * m6] and not any(st)
    t = time.perf_counter()
    for _ in range(3):
        E.g2_decompress_batch(comp, ctx=ctx)
    dt = (time.perf_counter() - t) / 3
    print(json.dumps({"config": "ingest: G2 from_compressed incl. Fq2 sqrt and r-torsion check (host buffers)", "n": m6, "keys_per_sec": m6 / dt, "ms": dt * 1e3}))
    t = time.perf_counter()
    for _ in range(3):
        st = E.g2_validate_batch(pks[:128 * m6], ctx=ctx)
    dt = (time.perf_counter() - t) / 3
    assert not any(st)
    print(json.dumps({"config": "ingest: G2 uncompressed validate (curve + r-torsion check, host buffers)", "n": m6, "keys_per_sec": m6 / dt, "ms": dt * 1e3}))

if __name__ == "__main__":
    main()
