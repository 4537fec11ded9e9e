import sys, os, ctypes, random, subprocess
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from bn254_b200 import engine as E
subprocess.check_call(['make','-C','tests/hostsim'],stdout=subprocess.DEVNULL)
hs=ctypes.CDLL('tests/hostsim/libhostsim.so')
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
be=lambda x:x.to_bytes(32,'big')
rng=random.Random(1)
ops={6:(5,3),10:(5,12),11:(5,3),12:(6,3),13:(10,6)}
for op,(ni,no) in ops.items():
    n=200
    data=b''.join(be(rng.randrange(Q)) for _ in range(ni*n))
    got=E.layer_op_batch(op,data,ni,no)
    bad=0
    for i in range(n):
        o=ctypes.create_string_buffer(32*no)
        hs.hs_layer_op(op,data[32*ni*i:32*ni*(i+1)],ni,o,no)
        if o.raw!=got[32*no*i:32*no*(i+1)]:
            bad+=1
            if bad==1 or (bad==2 and op==10):
                for k in range(no):
                    print('  op',op,'item',i,'limb',k, o.raw[32*k:32*k+32]==got[32*no*i+32*k:32*no*i+32*k+32])
    print('op',op,'bad',bad,'/',n)
