#!/usr/bin/env python3
"""BASELINE.json configs[4]: distinct-message aggregate verify of a 2^22-pair multi-pairing sharded over the GPUs of one box
(strong scaling: the total is fixed, every rank takes 2^22 / world pairs), one shared final exponentiation.

    python scripts/bench_distinct.py [log2_pairs]                                             # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_distinct.py

Per step, on every rank: Miller partial of its slice (hash + line sets + cooperative multi-pairing program + block products)
-> 384 bytes; partial G1 sum of its signatures -> 64 bytes; ONE all-gather of both (NCCL); finish (product of the partials,
Miller value of (sum sig, -G2), final exponentiation, verdict) on every rank.  Inputs are device-resident; time = CUDA events
on the engine's stream around the local work plus the wall time of the exchange and the finish, max over ranks.
Prints one JSON line on rank 0.  Not the bench line (bench.py is): a secondary measurement, like scripts/bench_configs.py."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import synth
from bn254_b200 import engine as E
from bn254_b200._native import S


def main():
    total = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.context(local)
    n = total // world
    dev = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    msgs, sks = synth.messages(n, 32, seed=500 + rank), synth.secret_keys(n, seed=600 + rank)
    d_msgs, d_sks = dev(msgs), dev(sks)
    d_sigs = torch.empty(64 * n, dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.call("bn254_sign_batch_dev", d_msgs, S(32), d_sks, S(n), d_sigs, d_st)
    ctx.sync()
    assert not d_st.any().item()
    d_pks = dev(E.derive_pk_g2_batch(sks, ctx=ctx))
    payload = torch.zeros(456, dtype=torch.uint8, device="cuda")   # Miller partial 384 | partial signature sum 64 | status bytes at 448 and 452
    gathered = [torch.empty_like(payload) for _ in range(world)]

    def step():
        ctx.call("bn254_miller_partial_distinct_dev", d_msgs, S(32), d_pks, S(n), payload[:384], payload[448:449])
        ctx.call("bn254_g1_sum_dev", d_sigs, None, S(n), payload[384:448], payload[452:453])
        ctx.sync()
        if world > 1:
            dist.all_gather(gathered, payload)
            parts = [bytes(g.cpu().numpy().tobytes()) for g in gathered]
        else:
            parts = [bytes(payload.cpu().numpy().tobytes())]
        assert all(p[448] == 0 and p[452] == 0 for p in parts)
        agg, st = E.g1_sum(b"".join(p[384:448] for p in parts), ctx=ctx)
        assert st == 0
        return E.finish_distinct(b"".join(p[:384] for p in parts), agg, ctx=ctx)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    assert step() == 0   # warm-up, and the verdict: the aggregate verifies
    steps = 2
    barrier()
    t = time.perf_counter()
    for _ in range(steps):
        v = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / steps
    if world > 1:
        tt = torch.tensor([dt], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    barrier()
    if rank == 0:
        print(json.dumps({"config": "5: distinct-message aggregate verify, %d pairs over %d GPU(s), shared final exponentiation" % (total, world),
                          "pairs_per_sec": total / dt, "ms_per_step": dt * 1e3, "n_gpus": world, "scaling": "strong", "verdict": v,
                          "exchange_bytes_per_rank": int(payload.numel())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
