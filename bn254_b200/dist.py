"""Multi-GPU host logic: one process per GPU (torch.distributed), shards of independent items, and the single exchange
step the path has -- the all-gather of one 384-byte Miller product per rank for the distinct-message aggregate verify
(SURVEY.md 8e).  Everything else (batch verify, sign, hash) is embarrassingly parallel: ranks take index ranges and
no data-path collective runs.

The compute callbacks default to the CUDA engine; the CPU tests (`gloo`, world_size 2) inject the oracle instead, so
that the sharding / exchange logic is covered on a machine without a GPU.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced index range of `rank` among `world` ranks: sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _device_for(group=None):
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def all_gather_bytes(payload, group=None):
    """all-gather of one fixed-size byte string per rank -> list of byte strings, rank order"""
    dev = _device_for(group)
    t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, t, group=group)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def verify_batch_sharded(msgs, msg_len, sigs, pks, verify_fn=None, group=None, gather=True):
    """ECDSA::verify over n triples split across the ranks.  Every rank passes the FULL buffers (or only its own slice
    when gather=False) and verifies its index range; with gather=True the per-item statuses are all-gathered so that
    every rank returns the full verdict vector.  No collective touches the data path."""
    if verify_fn is None:
        from . import engine
        verify_fn = lambda m, l, s, p: engine.verify_batch(m, l, s, p)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = len(sigs) // 64
    lo, hi = shard_range(n, rank, world)
    st = verify_fn(msgs[lo * msg_len:hi * msg_len], msg_len, sigs[64 * lo:64 * hi], pks[128 * lo:128 * hi])
    if not gather:
        return st
    width = -(-n // world) if n else 1  # pad to a common length
    parts = all_gather_bytes(st + bytes(width - len(st)), group)
    out = b""
    for r in range(world):
        a, b = shard_range(n, r, world)
        out += parts[r][:b - a]
    return out


def aggregate_verify_distinct_sharded(msgs, msg_len, pks, agg_sig, partial_fn=None, finish_fn=None, group=None):
    """prod_i e(H(msg_i), pk_i) * e(agg_sig, -G2) == 1 with the pairs sharded over the ranks: each rank folds the Miller
    values of its slice into one Fq12 (384 bytes), the partials are all-gathered, and every rank finishes with the single
    shared final exponentiation (identical bits on every rank: field multiplication is commutative).  Returns the
    status byte (0 = Ok, 9 = VerificationFailed, else the first decode error)."""
    if partial_fn is None or finish_fn is None:
        from . import engine
        partial_fn = partial_fn or (lambda m, l, p: engine.miller_partial_distinct(m, l, p))
        finish_fn = finish_fn or (lambda parts, sig: engine.finish_distinct(parts, sig))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = len(pks) // 128
    lo, hi = shard_range(n, rank, world)
    f, st = partial_fn(msgs[lo * msg_len:hi * msg_len], msg_len, pks[128 * lo:128 * hi])
    parts = all_gather_bytes(f + bytes([st]), group)
    for p in parts:  # the first failing shard decides, as a left fold over the items would
        if p[384]:
            return p[384]
    return finish_fn(b"".join(p[:384] for p in parts), agg_sig)
