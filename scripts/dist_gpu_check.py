#!/usr/bin/env python3
"""Multi-GPU check of bn254_b200/dist.py on real GPUs (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_gpu_check.py
Sharded batch verify (no collective on the data path) and the distinct-message aggregate verify (per-rank Miller partials,
one 385-byte all-gather, shared final exponentiation) against verdicts known by construction."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import synth
from bn254_b200 import dist as D
from bn254_b200 import engine as E


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.context(local)
    n = 4099  # ragged across ranks
    msgs, sks = synth.messages(n, 32, seed=61), synth.secret_keys(n, seed=62)
    sigs, st = E.sign_batch(msgs, 32, sks, ctx=ctx)
    assert not any(st)
    pks = E.derive_pk_g2_batch(sks, ctx=ctx)
    bad = bytearray(sigs)
    for i in (0, n // 2, n - 1):
        bad[64 * i:64 * i + 64] = sigs[64 * ((i + 1) % n):64 * ((i + 1) % n) + 64]
    got = D.verify_batch_sharded(msgs, 32, bytes(bad), pks, verify_fn=lambda m, l, s, p: E.verify_batch(m, l, s, p, ctx=ctx))
    want = bytearray(n)
    for i in (0, n // 2, n - 1):
        want[i] = 9
    assert got == bytes(want), "sharded verify verdicts differ"
    agg, st1 = E.g1_sum(sigs, ctx=ctx)
    assert st1 == 0
    pf = lambda m, l, p: E.miller_partial_distinct(m, l, p, ctx=ctx)
    ff = lambda parts, s: E.finish_distinct(parts, s, ctx=ctx)
    assert D.aggregate_verify_distinct_sharded(msgs, 32, pks, agg, pf, ff) == 0
    wrong, _ = E.g1_sum(sigs[:64 * (n - 1)], ctx=ctx)
    assert D.aggregate_verify_distinct_sharded(msgs, 32, pks, wrong, pf, ff) == 9
    # same verdict as the single-GPU entry point over the whole set
    assert E.aggregate_verify_distinct(msgs, 32, pks, agg, ctx=ctx) == 0
    dist.barrier()
    if rank == 0:
        print("dist_gpu_check ok: world=%d n=%d" % (world, n))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
