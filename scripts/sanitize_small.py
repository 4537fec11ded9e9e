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
m = 5000
h, st = E.hash_to_g1_batch(synth.messages(m, 32, seed=9), 32, m, ctx=ctx)
assert not any(st)
print("sanitize_small ok")
