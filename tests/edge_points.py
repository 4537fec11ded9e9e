"""Adversarial twist points for the r-torsion test (TEST INFRASTRUCTURE: built with oracle/pyoracle.py)."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def subgroup_edge_points():
    """Twist points that stress the r-torsion test: points of each small prime order dividing the twist cofactor, their
    sums with a G2 point, a random twist point (all outside G2), and cofactor-cleared points (inside).  -> [(raw128, in_g2)]"""
    import random as _r
    import sys as _s
    _s.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as P
    rng = _r.Random(99)
    raw = lambda p: b"".join(c.to_bytes(32, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))

    def rand_pt():
        while True:
            x = (rng.randrange(P.Q), rng.randrange(P.Q))
            y2 = P.f2_add(P.f2_mul(P.f2_mul(x, x), x), P.B2)
            y = P.f2_sqrt(y2)
            if y is not None and P.f2_mul(y, y) == y2:
                return (x, y)
    h2 = 2 * P.Q - P.R
    out = []
    for f in P.TWIST_COFACTOR_FACTORS:
        s = None
        while s is None:
            s = P.g2_mul(rand_pt(), P.R * h2 // f)
        out.append((raw(s), False))
        out.append((raw(P.g2_add(s, P.g2_mul(P.G2_GEN, rng.randrange(1, P.R)))), False))
    t = rand_pt()
    out.append((raw(t), False))
    out.append((raw(P.g2_mul(t, h2)), True))
    out.append((raw(P.g2_mul(P.G2_GEN, P.R - 1)), True))
    return out


_OUTSIDE = None


def twist_point_outside_g2():
    """128 raw bytes of a point on the twist that is NOT in the r-torsion (a random twist point: the cofactor is ~2^254)."""
    global _OUTSIDE
    if _OUTSIDE is None:
        import random as _r
        import sys as _s
        _s.path.insert(0, os.path.join(ROOT, "oracle"))
        _s.path.insert(0, os.path.join(ROOT, "tests"))
        import pyoracle as P
        import oracle_lib as O
        rng = _r.Random(4242)
        while True:
            x = (rng.randrange(P.Q), rng.randrange(P.Q))
            y2 = P.f2_add(P.f2_mul(P.f2_mul(x, x), x), P.B2)
            y = P.f2_sqrt(y2)
            if y is not None and P.f2_mul(y, y) == y2:
                break
        raw = b"".join(c.to_bytes(32, "big") for c in (x[0], x[1], y[0], y[1]))
        assert O.g2_validate_uncompressed(raw) == O.INVALID_GROUP_POINT
        _OUTSIDE = raw
    return _OUTSIDE
