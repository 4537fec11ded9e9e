"""Seeded synthetic workloads shared by the GPU parity tests and bench.py.

Keys / signatures are produced by the ENGINE (GPU) and spot-checked against the oracle by the callers; this module
itself only makes random bytes and never touches oracle/."""
import numpy as np

R_ORDER = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def rand_bytes(seed, n):
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8).tobytes()


def messages(n, msg_len=32, seed=1):
    return rand_bytes(seed, n * msg_len)


def secret_keys(n, seed=2):
    """n x 32 random bytes with the top byte cleared to 0x0f..: always in [1, r) after reduction, never zero."""
    a = np.random.default_rng(seed).integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] &= 0x1F
    a[:, 31] |= 1
    return a.tobytes()
