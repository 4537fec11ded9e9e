#!/usr/bin/env python3
"""Two-rank NCCL check of bn254_b200/dist.py with the CUDA engine (launched by tests/test_gpu_parity.py through
`python -m torch.distributed.run --nproc-per-node 2`; also runs at any world size).  Every rank builds the SAME seeded inputs,
takes its shard, and the collective paths must give the verdict the oracle gives for the whole input:

  * DistinctAggregate (device-resident: payload -> all_gather_into_tensor -> finish): accept; a forged signature on the last
    rank; an undecodable key on the last rank (first failing record decides);
  * SameMessageAggregate (per-rank sums -> all-gather -> one verify): accept / reject;
  * the byte-string forms (verify_batch_sharded, aggregate_verify_distinct_sharded, aggregate_verify_same_msg_sharded) with
    the engine callbacks, a bad item on the last rank.
Prints DIST_GPU_OK on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import oracle_lib as O
import synth
from bn254_b200 import dist as D
from bn254_b200 import engine as E


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.context(local)
    E.set_input_policy(E.INPUTS_TYPED, ctx=ctx)
    dev = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    n = 4099  # ragged shards
    msgs, sks = synth.messages(n, 32, seed=901), synth.secret_keys(n, seed=902)
    sigs, st = E.sign_batch(msgs, 32, sks, ctx=ctx)
    assert not any(st)
    pks = E.derive_pk_g2_batch(sks, ctx=ctx)
    lo, hi = D.shard_range(n, rank, world)
    last_lo, _ = D.shard_range(n, world - 1, world)
    bad_i = last_lo + 3  # an item of the LAST rank's shard

    # ---- distinct-message aggregate, device-resident
    agg = D.DistinctAggregate(ctx)

    def distinct(sig_bytes, pk_bytes):
        agg.step(dev(msgs[32 * lo:32 * hi]), 32, dev(pk_bytes[128 * lo:128 * hi]), dev(sig_bytes[64 * lo:64 * hi]), hi - lo)
        return agg.status()

    assert distinct(sigs, pks) == 0
    forged = sigs[:64 * bad_i] + sigs[:64] + sigs[64 * bad_i + 64:]
    assert distinct(forged, pks) == O.VERIFICATION_FAILED
    badkey = pks[:128 * bad_i + 127] + bytes([pks[128 * bad_i + 127] ^ 1]) + pks[128 * bad_i + 128:]
    assert distinct(sigs, badkey) == O.INVALID_GROUP_POINT
    # the oracle's verdict for the same aggregate, over a prefix that it can fold in seconds (rank 0 only)
    if rank == 0:
        k = 64
        hs = b"".join(O.hash_to_g1(msgs[32 * i:32 * i + 32])[1] for i in range(k))
        neg_g2 = O.g2_neg(O.derive_pk_g2((1).to_bytes(32, "big"))[1])[1]
        a = E.g1_sum(sigs[:64 * k], ctx=ctx)[0]
        assert O.pairing_check(hs + a, pks[:128 * k] + neg_g2, k + 1)[1] == (1).to_bytes(32, "big") + bytes(352)
    # external aggregate signature instead of local signatures
    a_all = dev(E.g1_sum(sigs, ctx=ctx)[0])
    agg.step(dev(msgs[32 * lo:32 * hi]), 32, dev(pks[128 * lo:128 * hi]), None, hi - lo, agg_sig=a_all)
    assert agg.status() == 0

    # ---- same-message aggregate, device-resident
    msg = b"same message"
    s2, st = E.sign_batch(msg * n, len(msg), sks, ctx=ctx)
    same = D.SameMessageAggregate(ctx)
    same.step(dev(msg), len(msg), dev(s2[64 * lo:64 * hi]), dev(pks[128 * lo:128 * hi]), hi - lo)
    assert same.status() == 0
    s2bad = s2[:64 * bad_i] + s2[:64] + s2[64 * bad_i + 64:]
    same.step(dev(msg), len(msg), dev(s2bad[64 * lo:64 * hi]), dev(pks[128 * lo:128 * hi]), hi - lo)
    assert same.status() == O.VERIFICATION_FAILED

    # ---- byte-string forms with the engine callbacks
    vfy = lambda m, l, s, p: E.verify_batch(m, l, s, p, ctx=ctx)
    got = D.verify_batch_sharded(msgs, 32, forged, pks, verify_fn=vfy)
    assert got == bytes(O.VERIFICATION_FAILED if i == bad_i else 0 for i in range(n))
    part = lambda m, l, p: E.miller_partial_distinct(m, l, p, ctx=ctx)
    fin = lambda parts, sg: E.finish_distinct(parts, sg, ctx=ctx)
    a_host = E.g1_sum(sigs, ctx=ctx)[0]
    assert D.aggregate_verify_distinct_sharded(msgs, 32, pks, a_host, part, fin) == 0
    assert D.aggregate_verify_distinct_sharded(msgs, 32, badkey, a_host, part, fin) == O.INVALID_GROUP_POINT
    s1 = lambda pts: E.g1_sum(pts, ctx=ctx)
    s2f = lambda pts: E.g2_sum(pts, ctx=ctx)
    v1 = lambda m, s, p: E.verify_batch(m, len(m), s, p, ctx=ctx)[0]
    assert D.aggregate_verify_same_msg_sharded(msg, s2[64 * lo:64 * hi], pks[128 * lo:128 * hi], s1, s2f, v1) == 0
    assert D.aggregate_verify_same_msg_sharded(msg, s2bad[64 * lo:64 * hi], pks[128 * lo:128 * hi], s1, s2f, v1) == O.VERIFICATION_FAILED
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
