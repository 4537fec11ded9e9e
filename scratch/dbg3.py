
import sys, os, ctypes, random, subprocess
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from bn254_b200 import engine as E
import oracle_lib as O
subprocess.check_call(['make','-C','tests/hostsim'],stdout=subprocess.DEVNULL)
hs=ctypes.CDLL('tests/hostsim/libhostsim.so')
be=lambda x:x.to_bytes(32,'big')
g1=O.derive_pk_g1(be(12345))[1]; g2=O.derive_pk_g2(be(6789))[1]
data=g1+g2
for op in [100,101,102,103,104,108,116,132,164,165,166,200]:
    got=E.layer_op_batch(op,data,6,12)
    o=ctypes.create_string_buffer(384)
    hs.hs_layer_op(op,data,6,o,12)
    print(op, o.raw==got, [o.raw[32*k:32*k+32]==got[32*k:32*k+32] for k in range(12)])
print('oracle', O.miller_product(g1,g2,1)[1]==E.layer_op_batch(200,data,6,12))
