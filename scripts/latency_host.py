#!/usr/bin/env python3
"""Latency of ONE host-buffer bn254_verify_batch call at small n (median of 20 after 5 warm-up calls), wall clock around the call.
Prints one JSON line.  Knobs under test come from the environment (BN254_PIPELINE, BN254_COOP12, BN254_LINES_LAT)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from bn254_b200 import engine as E


def main():
    ctx = E.context(0)
    E.set_input_policy(E.INPUTS_TYPED, ctx=ctx)
    n = 1 << 15
    msgs, sks = synth.messages(n, 32, seed=1), synth.secret_keys(n, seed=2)
    sigs, st = E.sign_batch(msgs, 32, sks, ctx=ctx)
    pks = E.derive_pk_g2_batch(sks, ctx=ctx)
    out = {"env": {k: os.environ[k] for k in ("BN254_PIPELINE", "BN254_COOP12", "BN254_LINES_LAT") if k in os.environ}}
    for nl in ([int(a) for a in sys.argv[1:]] or [1, 32, 1000, 4736, 8192]):
        m, s, p = msgs[: 32 * nl], sigs[: 64 * nl], pks[: 128 * nl]
        ts = []
        for rep in range(25):
            t0 = time.perf_counter()
            st = E.verify_batch(m, 32, s, p, ctx=ctx)
            ts.append((time.perf_counter() - t0) * 1e3)
            if any(st):
                print("n", nl, "rep", rep, "WRONG: %d items rejected, first %d" % (sum(1 for b in st if b), next(i for i, b in enumerate(st) if b)),
                      file=sys.stderr, flush=True)
                break
        ts = sorted(ts[5:])
        out["n%d_ms" % nl] = round(ts[len(ts) // 2], 3)
        print("n", nl, out["n%d_ms" % nl], file=sys.stderr, flush=True)
    print(json.dumps(out))


main()
