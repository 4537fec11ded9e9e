#!/usr/bin/env python3
"""Where the time of the small / per-rank workloads goes (one GPU): phases of a verify at n = 1 .. 2^14, and the two halves of a
distinct-message aggregate step at the per-rank size of an 8-GPU run (2^19 pairs): payload (hash, line sets, multi-pairing, product
tree) and finish (cooperative final exponentiation).  CUDA events on the engine's stream.  Prints one JSON line."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import synth
from bn254_b200 import dist as D
from bn254_b200 import engine as E
from bn254_b200._native import I, S


def main():
    ctx = E.context(0)
    E.set_input_policy(E.INPUTS_TYPED, ctx=ctx)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    dev = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    out = {}
    n = 1 << 14
    msgs, sks = synth.messages(n, 32, seed=1), synth.secret_keys(n, seed=2)
    sigs, st = E.sign_batch(msgs, 32, sks, ctx=ctx)
    pks = E.derive_pk_g2_batch(sks, ctx=ctx)
    d_m, d_s, d_p = dev(msgs), dev(sigs), dev(pks)
    d_st = torch.zeros(n, dtype=torch.uint8, device="cuda")
    phase = (ctypes.c_float * 3)()
    for nl in (1, 32, 33, 1024, 4736, 1 << 14):
        ctx.call("bn254_verify_batch_dev", d_m, S(32), d_s, d_p, S(nl), d_st)
        ctx.call("bn254_set_profiling", I(1))
        for _ in range(3):
            ctx.call("bn254_verify_batch_dev", d_m, S(32), d_s, d_p, S(nl), d_st)
        ctx.call("bn254_phase_ms", phase)
        ctx.call("bn254_set_profiling", I(0))
        assert not d_st[:nl].any().item()
        out["verify_n%d" % nl] = {"hash_ms": phase[0] / 3, "lines_ms": phase[1] / 3, "machine_ms": phase[2] / 3}
    log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 19
    n5 = 1 << log2
    m5, k5 = synth.messages(n5, 32, seed=5), synth.secret_keys(n5, seed=6)
    d_m5, d_k5 = dev(m5), dev(k5)
    d_s5 = torch.empty(64 * n5, dtype=torch.uint8, device="cuda")
    d_t5 = torch.empty(n5, dtype=torch.uint8, device="cuda")
    ctx.call("bn254_sign_batch_dev", d_m5, S(32), d_k5, S(n5), d_s5, d_t5)
    d_p5 = dev(E.derive_pk_g2_batch(k5, ctx=ctx))
    payload = torch.zeros(448, dtype=torch.uint8, device="cuda")
    verdict = torch.zeros(1, dtype=torch.uint8, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for rep in range(3):
        ctx.sync()
        ev[0].record(stream)
        ctx.call("bn254_distinct_payload_dev", d_m5, S(32), d_p5, d_s5, S(n5), payload)
        ev[1].record(stream)
        ctx.call("bn254_finish_distinct_dev", payload, S(1), None, verdict)
        ev[2].record(stream)
        ctx.sync()
    assert int(verdict.cpu()[0]) == 0
    out["distinct_2^%d" % log2] = {"payload_ms": ev[0].elapsed_time(ev[1]), "finish_ms": ev[1].elapsed_time(ev[2]),
                                    "ideal_ms_at_1gpu_rate_of_2^22": None}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
