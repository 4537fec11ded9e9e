import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle_lib as O, synth
from bn254_b200 import engine as E
be=lambda x:x.to_bytes(32,'big')
sks=synth.secret_keys(2000, seed=61)
p1=E.derive_pk_g1_batch(sks)
assert p1==O.derive_pk_g1_batch(sks,2000,8)
for n in [1,2,3,4,5,8,9,16,17,64,128,129,256,1000,1024,1025,2000]:
    r=E.g1_sum(p1[:64*n]); e=O.g1_sum(p1[:64*n],n)[1]
    print('g1_sum',n,r[0]==e,r[1])
p2=E.derive_pk_g2_batch(sks[:32*40])
for n in [1,2,3,9,17,40]:
    r=E.g2_sum(p2[:128*n]); e=O.g2_sum(p2[:128*n],n)[1]
    print('g2_sum',n,r[0]==e,r[1])
# single pair miller
g1=p1[:64]; g2=p2[:128]
f,st=E.miller_loop_batch(g1,g2,1,1)
print('miller1', f==O.miller_product(g1,g2,1)[1], st)
f,st=E.miller_loop_batch(g1*3,g2*3,1,3)
print('miller1x3', [f[384*i:384*i+384]==O.miller_product(g1,g2,1)[1] for i in range(3)])
