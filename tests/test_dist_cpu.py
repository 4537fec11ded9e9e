"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic in bn254_b200/dist.py: index sharding, the verdict gather of
the sharded batch verify, and the all-gather + shared final exponentiation of the distinct-message aggregate verify.
The compute callbacks are the oracle here (test infrastructure); on a GPU box the same functions drive the CUDA engine."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import oracle_lib as O
    import synth
    from bn254_b200 import dist as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        n = 7  # ragged: 4 + 3
        msgs, sks = synth.messages(n, 32, seed=21), synth.secret_keys(n, seed=22)
        sigs, st = O.sign_batch(msgs, 32, sks, n)
        assert st == bytes(n)
        pks = O.derive_pk_g2_batch(sks, n)
        bad = bytearray(sigs)
        bad[64 * 5:64 * 6] = sigs[64 * 4:64 * 5]  # item 5 (second shard) carries item 4's signature
        verify = lambda m, l, s, p: O.verify_batch(m, l, s, p, len(s) // 64)
        got = D.verify_batch_sharded(msgs, 32, bytes(bad), pks, verify_fn=verify)
        assert got == bytes([0, 0, 0, 0, 0, 9, 0]), got

        neg_g2 = O.g2_neg(O.derive_pk_g2((1).to_bytes(32, "big"))[1])[1]

        def partial(m, l, p):
            k = len(p) // 128
            hs = b"".join(O.hash_to_g1(m[l * i:l * i + l])[1] for i in range(k))
            return O.miller_product(hs, p, k)[1], 0

        def finish(parts, sig):
            f = O.miller_product(sig, neg_g2, 1)[1]
            for i in range(len(parts) // 384):
                f = O.fq12_op(0, f, parts[384 * i:384 * i + 384])[1]
            one = (1).to_bytes(32, "big") + bytes(352)
            return 0 if O.final_exp(f)[1] == one else 9

        agg = bytes(64)
        for i in range(n):
            agg = O.g1_add(agg, sigs[64 * i:64 * i + 64])[1]
        assert D.aggregate_verify_distinct_sharded(msgs, 32, pks, agg, partial, finish) == 0
        wrong = O.g1_add(agg, sigs[:64])[1]
        assert D.aggregate_verify_distinct_sharded(msgs, 32, pks, wrong, partial, finish) == 9
        # same-message aggregate over sharded keys (config 4, /root/reference/examples/bn254.rs:25-32): per-rank sums, one
        # all-gather of the 64 + 128 byte partial sums, the last additions and one verify on every rank
        msg = b"sample"
        s_same = b"".join(O.sign(msg, sks[32 * i:32 * i + 32])[1] for i in range(n))
        lo, hi = D.shard_range(n, rank, world)
        g1s = lambda pts: O.g1_sum(pts, len(pts) // 64)[::-1]
        g2s = lambda pts: O.g2_sum(pts, len(pts) // 128)[::-1]
        vfy1 = lambda m, s, p: O.verify(m, s, p)
        assert D.aggregate_verify_same_msg_sharded(msg, s_same[64 * lo:64 * hi], pks[128 * lo:128 * hi], g1s, g2s, vfy1) == 0
        bad_same = s_same[:64 * 5] + s_same[:64] + s_same[64 * 6:]  # item 5 lives on rank 1
        assert D.aggregate_verify_same_msg_sharded(msg, bad_same[64 * lo:64 * hi], pks[128 * lo:128 * hi], g1s, g2s, vfy1) == 9
        # the split must not matter: the whole-input sums verify under the oracle as well
        assert O.verify(msg, O.g1_sum(s_same, n)[1], O.g2_sum(pks, n)[1]) == 0
        assert [D.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
        assert [D.shard_range(1, r, 2) for r in range(2)] == [(0, 1), (1, 1)]
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        q.put((rank, "FAIL %s: %s" % (type(e).__name__, e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_verify_and_distinct_aggregate_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
