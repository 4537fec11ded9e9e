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
print("sanitize_small ok")
