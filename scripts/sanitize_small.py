"""Small verify / check_public_keys / hash batches for compute-sanitizer (memcheck, racecheck, synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from bn254_b200 import engine as E
ctx = E.context(0)
n = 70  # three 32-item groups, ragged
msgs, sks = synth.messages(n, 32, seed=3), synth.secret_keys(n, seed=4)
sigs, st = E.sign_batch(msgs, 32, sks, ctx=ctx)
pks = E.derive_pk_g2_batch(sks, ctx=ctx)
bad = bytearray(sigs); bad[64:128] = sigs[:64]
st = E.verify_batch(msgs, 32, bytes(bad), pks, ctx=ctx)
assert st[0] == 0 and st[1] == 9 and sum(1 for s in st if s) == 1, st
# the warp-local layout of the cooperative machine (pairing mode 3), the randomised batch path, key validation and the multi-pairing
from bn254_b200._native import I
ctx.call("bn254_set_pairing_mode", I(3))
assert E.verify_batch(msgs, 32, bytes(bad), pks, ctx=ctx) == st
ctx.call("bn254_set_pairing_mode", I(0))
st2, fast = E.verify_batch_rlc(msgs, 32, sigs, pks, synth.rand_bytes(8, 16 * n), ctx=ctx)
assert fast and st2 == bytes(n)
st2, fast = E.verify_batch_rlc(msgs, 32, bytes(bad), pks, synth.rand_bytes(8, 16 * n), ctx=ctx)
assert not fast and st2 == st
assert not any(E.g2_validate_batch(pks, ctx=ctx))
agg, s1 = E.g1_sum(sigs, ctx=ctx)
assert E.aggregate_verify_distinct(msgs, 32, pks, agg, ctx=ctx) == 0
m = 5000
h, st = E.hash_to_g1_batch(synth.messages(m, 32, seed=9), 32, m, ctx=ctx)
assert not any(st)
# round 2: the four-group block layout (more than two groups per SM), the untrusted policy's validation kernel (this context's
# default), the cooperative pairing check, the formatter, and the device-resident aggregate checks with their cooperative finish
import torch
from bn254_b200 import dist as D
big = 2 * 148 * 32 + 70
bm, bk = synth.messages(big, 32, seed=13), synth.secret_keys(big, seed=14)
bs, st = E.sign_batch(bm, 32, bk, ctx=ctx)
bp = E.derive_pk_g2_batch(bk, ctx=ctx)
assert E.verify_batch(bm, 32, bs, bp, ctx=ctx) == bytes(big)
assert E.pairing_check_batch(sigs, pks, 1, n, ctx=ctx) == bytes([9]) * n
neg_g2 = E.g2_sum(E.derive_pk_g2_batch((1).to_bytes(32, "big"), ctx=ctx), bytes([1]), ctx=ctx)[0]
hs = E.hash_to_g1_batch(msgs, 32, n, ctx=ctx)[0]
g1s = b"".join(hs[64 * i:64 * i + 64] + sigs[64 * i:64 * i + 64] for i in range(n))
g2s = b"".join(pks[128 * i:128 * i + 128] + neg_g2 for i in range(n))
assert E.pairing_check_batch(g1s, g2s, 2, n, ctx=ctx) == bytes(n)
blob, st = E.format_pairing_check_batch(msgs, 32, sigs, pks, False, ctx=ctx)
assert not any(st) and len(blob) == 384 * n
dev = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
da = D.DistinctAggregate(ctx, world=1)
da.step(dev(bm), 32, dev(bp), dev(bs), big)
assert da.status() == 0
E.set_input_policy(E.INPUTS_TYPED, ctx=ctx)
sa = D.SameMessageAggregate(ctx, world=1)
ms, st = E.sign_batch(b"m" * 70, 1, sks, ctx=ctx)
sa.step(dev(b"m"), 1, dev(ms), dev(pks), n)
assert sa.status() == 0
print("sanitize_small ok")
