"""Pure-Python big-integer oracle for the BN254 sign / aggregate / pairing-verify path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bn254_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may touch ``oracle/``.

This is the *second*, deliberately naive oracle: textbook affine group law, the pairing
computed on E(Fq12) after untwisting (degree-12 polynomial arithmetic, plain Miller loop
over the bits of 6u+2, final exponentiation by a plain pow with (q^12-1)/r).  It shares no
code and no algorithmic shortcut with either the C oracle (``oracle/bn254_oracle.c``, which
follows the dependency's tower/optimal-ate structure) or the CUDA engine, so agreement of
the three is meaningful.  It is slow (~1 s per pairing) and is used on small cases only.

Reference semantics restated (file:line are relative to /root/reference):
  * hash_to_try_and_increment        src/hash.rs:29-63   (constant 5q: src/hash.rs:11-14)
  * mod_u256 (strict ">" loop)       src/utils.rs:27-37
  * arbitrary_string_to_g1           src/utils.rs:56-63
  * ECDSA::sign / verify             src/ecdsa.rs:26-35, 49-64
  * check_public_keys                src/ecdsa.rs:78-93
  * G1/G2 codecs                     src/utils.rs:84-194
  * PrivateKey conversions           src/types.rs:13-77  (Fr::from_slice reduces mod r)
  * Error variants                   src/error.rs:5-62
Dependency semantics (crate zeropool-bn 0.5.11, not in the tree) follow SURVEY.md
Appendix A; every rule here is pinned by the reference's own KATs in tests/golden/.
"""
import hashlib

# --------------------------------------------------------------------------- constants
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
U = 4965661367192848881
ATE = 6 * U + 2
FIVE_Q = 0xF1F5883E65F820D099915C908786B9D3F58714D70A38F4C22CA2BC723A70F263  # src/hash.rs:11-14
assert FIVE_Q == 5 * Q and 6 * Q >= 1 << 256  # src/hash_test.rs:33-43

G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)

# status codes: one per Error variant of src/error.rs:6-29, 0 = Ok
OK = 0
HASH_TO_POINT_ERROR = 1
INDEX_OUT_OF_BOUNDS = 2
INVALID_ENCODING = 3
INVALID_GROUP_POINT = 4
INVALID_LENGTH = 5
NOT_MEMBER_ERROR = 6
TO_AFFINE_CONVERSION = 7
POINT_IN_JACOBIAN = 8
VERIFICATION_FAILED = 9
SERIALIZATION_ERROR = 10
HEX_DECODE_FAILED = 11


class Bn254Error(Exception):
    def __init__(self, code):
        super().__init__(code)
        self.code = code


# --------------------------------------------------------------------------- Fq2 = Fq[i]/(i^2+1)
def f2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_neg(a):
    return ((-a[0]) % Q, (-a[1]) % Q)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], Q - 2, Q)
    return (a[0] * n % Q, (-a[1]) * n % Q)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_mul(a, a)
        e >>= 1
    return r


XI = (9, 1)
B2 = f2_mul((3, 0), f2_inv(XI))  # twist coefficient 3/xi


def fq_sqrt(a):
    """Fq::sqrt: a1 = a^((q-3)/4); root = a1*a; reject when a1*root == -1."""
    a %= Q
    a1 = pow(a, (Q - 3) // 4, Q)
    root = a1 * a % Q
    if a1 * root % Q == Q - 1:
        return None
    if root * root % Q != a:  # unreachable for prime q = 3 mod 4; kept as a guard
        return None
    return root


def f2_sqrt(a):
    """Fq2::sqrt, complex method for q = 3 mod 4."""
    a1 = f2_pow(a, (Q - 3) // 4)
    alpha = f2_mul(f2_mul(a1, a1), a)
    a0 = f2_mul((alpha[0], (-alpha[1]) % Q), alpha)  # alpha^q * alpha
    if a0 == (Q - 1, 0):
        return None
    x0 = f2_mul(a1, a)
    if alpha == (Q - 1, 0):
        return f2_mul((0, 1), x0)
    b = f2_pow(f2_add((1, 0), alpha), (Q - 1) // 2)
    return f2_mul(b, x0)


# --------------------------------------------------------------------------- affine groups (None = infinity)
def g1_on_curve(p):
    x, y = p
    return (y * y - x * x * x - 3) % Q == 0


def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, Q - 2, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, Q - 2, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    return (x3, (lam * (x1 - x3) - y1) % Q)


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % Q)


def g1_mul(p, k):
    r = None
    while k:
        if k & 1:
            r = g1_add(r, p)
        p = g1_add(p, p)
        k >>= 1
    return r


def g2_on_curve(p):
    x, y = p
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), B2)) == (0, 0)


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(x1, x1)), f2_inv(f2_add(y1, y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_neg(p):
    return None if p is None else (p[0], f2_neg(p[1]))


def g2_mul(p, k):
    r = None
    while k:
        if k & 1:
            r = g2_add(r, p)
        p = g2_add(p, p)
        k >>= 1
    return r


def g2_in_subgroup(p):
    return g2_mul(p, R) is None


# The engine decides the same predicate with one 63-bit scalar multiplication (bn254_b200/csrc/items.cuh
# g2_in_subgroup): [u+1]P + psi([u]P) + psi^2([u]P) == psi^3([2u]P), psi = untwist-Frobenius-twist.  Restated here so
# that tests/test_oracle.py can prove the two predicates equal on E'(Fq2) (one point of every prime order dividing the
# twist cofactor h2 = 2q - r).
PSI_X = f2_pow(XI, (Q - 1) // 3)
PSI_Y = f2_pow(XI, (Q - 1) // 2)
TWIST_COFACTOR_FACTORS = (10069, 5864401, 1875725156269, 197620364512881247228717050342013327560683201906968909)


def g2_psi(p):
    if p is None:
        return None
    (x0, x1), (y0, y1) = p
    return (f2_mul((x0, (-x1) % Q), PSI_X), f2_mul((y0, (-y1) % Q), PSI_Y))


def g2_in_subgroup_psi(p):
    a = g2_mul(p, U)
    b = g2_psi(a)
    c = g2_psi(b)
    lhs = g2_add(g2_add(g2_add(a, p), b), c)
    return lhs == g2_psi(g2_psi(g2_psi(g2_add(a, a))))


# --------------------------------------------------------------------------- byte formats
def _be32(x):
    return x.to_bytes(32, "big")


def fq_from_slice(b):
    """Fq::from_slice: len != 32 -> InvalidLength, value >= q -> NotMemberError."""
    if len(b) != 32:
        raise Bn254Error(INVALID_LENGTH)
    v = int.from_bytes(b, "big")
    if v >= Q:
        raise Bn254Error(NOT_MEMBER_ERROR)
    return v


def fr_from_slice(b):
    """Fr::from_slice (src/types.rs:37): length check, then reduce mod r, never rejects."""
    if len(b) != 32:
        raise Bn254Error(INVALID_LENGTH)
    return int.from_bytes(b, "big") % R


def fr_to_bytes(k):
    return _be32(k)  # src/utils.rs:66-72


def g1_from_compressed(b):
    """bn::G1::from_compressed."""
    if len(b) != 33:
        raise Bn254Error(INVALID_ENCODING)
    sign = b[0]
    x = fq_from_slice(b[1:])
    y = fq_sqrt((x * x * x + 3) % Q)
    if y is None:
        raise Bn254Error(NOT_MEMBER_ERROR)
    if sign == 2:
        if y & 1:
            y = Q - y
    elif sign == 3:
        if not (y & 1):
            y = (Q - y) % Q
    else:
        raise Bn254Error(INVALID_ENCODING)
    if not g1_on_curve((x, y)):
        raise Bn254Error(NOT_MEMBER_ERROR)
    return (x, y)


def g1_to_compressed(p):
    """src/utils.rs:84-104."""
    if p is None:
        raise Bn254Error(POINT_IN_JACOBIAN)
    return bytes([3 if p[1] & 1 else 2]) + _be32(p[0])


def g1_from_uncompressed(b):
    """src/utils.rs:119-127."""
    if len(b) != 64:
        raise Bn254Error(INVALID_LENGTH)
    x = fq_from_slice(b[:32])
    y = fq_from_slice(b[32:])
    if not g1_on_curve((x, y)):
        raise Bn254Error(INVALID_GROUP_POINT)
    return (x, y)


def g1_to_uncompressed(p):
    """src/utils.rs:182-194."""
    if p is None:
        raise Bn254Error(POINT_IN_JACOBIAN)
    return _be32(p[0]) + _be32(p[1])


def _u512(c):  # to_u512, src/utils.rs:40-45: imaginary*q + real
    return c[1] * Q + c[0]


def g2_from_compressed(b):
    """bn::G2::from_compressed incl. the r-torsion check of AffineG2::new."""
    if len(b) != 65:
        raise Bn254Error(INVALID_ENCODING)
    sign = b[0]
    c1, c0 = divmod(int.from_bytes(b[1:], "big"), Q)
    if c1 >= Q:  # quotient does not fit / not a field element
        raise Bn254Error(NOT_MEMBER_ERROR)
    x = (c0, c1)
    y = f2_sqrt(f2_add(f2_mul(f2_mul(x, x), x), B2))
    if y is None:
        raise Bn254Error(NOT_MEMBER_ERROR)
    ny = f2_neg(y)
    gt = _u512(y) > _u512(ny)
    if sign == 10:
        if gt:
            y = ny
    elif sign == 11:
        if not gt:
            y = ny
    else:
        raise Bn254Error(INVALID_ENCODING)
    p = (x, y)
    if not g2_on_curve(p) or not g2_in_subgroup(p):
        raise Bn254Error(NOT_MEMBER_ERROR)
    return p


def g2_to_compressed(p):
    """src/utils.rs:130-160."""
    if p is None:
        raise Bn254Error(POINT_IN_JACOBIAN)
    x, y = p
    sign = 0x0B if _u512(y) > _u512(f2_neg(y)) else 0x0A
    return bytes([sign]) + _u512(x).to_bytes(64, "big")


def g2_from_uncompressed(b):
    """src/utils.rs:107-116: real first, then imaginary."""
    if len(b) != 128:
        raise Bn254Error(INVALID_LENGTH)
    x = (fq_from_slice(b[0:32]), fq_from_slice(b[32:64]))
    y = (fq_from_slice(b[64:96]), fq_from_slice(b[96:128]))
    p = (x, y)
    if not g2_on_curve(p) or not g2_in_subgroup(p):
        raise Bn254Error(INVALID_GROUP_POINT)
    return p


def g2_to_uncompressed(p):
    """src/utils.rs:162-179."""
    if p is None:
        raise Bn254Error(POINT_IN_JACOBIAN)
    return _be32(p[0][0]) + _be32(p[0][1]) + _be32(p[1][0]) + _be32(p[1][1])


# --------------------------------------------------------------------------- hash to G1
def mod_u256(x, m):
    """src/utils.rs:27-37 -- note the strict '>'."""
    while x > m:
        x -= m
    return x


def hash_to_try_and_increment(msg, want_counter=False):
    """src/hash.rs:29-63."""
    for ctr in range(255):
        h = int.from_bytes(hashlib.sha256(bytes(msg) + bytes([ctr])).digest(), "big")
        if h >= FIVE_Q:
            continue
        x = mod_u256(h, Q)
        try:
            p = g1_from_compressed(b"\x02" + _be32(x))
        except Bn254Error:
            continue
        return (p, ctr) if want_counter else p
    raise Bn254Error(HASH_TO_POINT_ERROR)


# --------------------------------------------------------------------------- textbook pairing on E(Fq12)
# Fq12 = Fq[w]/(w^12 - 18 w^6 + 82); i = w^6 - 9.
_DEG = 12
_MODC = {0: 82, 6: -18}  # w^12 = 18 w^6 - 82


def f12(c):
    return tuple(x % Q for x in c)


F12_ONE = f12([1] + [0] * 11)
F12_ZERO = f12([0] * 12)


def f12_add(a, b):
    return tuple((x + y) % Q for x, y in zip(a, b))


def f12_sub(a, b):
    return tuple((x - y) % Q for x, y in zip(a, b))


def f12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):
        v = t[k]
        if v:
            t[k - 6] += 18 * v
            t[k - 12] -= 82 * v
    return tuple(x % Q for x in t[:12])


def f12_scalar(a, s):
    return tuple(x * s % Q for x in a)


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r


def f12_inv(a):
    # a^(q^12-2) would be far too slow; use norm chain through conjugates: a^-1 = prod_{k=1..11} a^(q^k) / N(a)
    # Cheaper and still independent: solve via extended Euclid on polynomials over Fq.
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0, 1]

    def deg(p):
        d = len(p) - 1
        while d and p[d] % Q == 0:
            d -= 1
        return d

    def poly_rounded_div(x, y):
        dx, dy = deg(x), deg(y)
        temp = list(x)
        o = [0] * len(x)
        for i in range(dx - dy, -1, -1):
            c = temp[dy + i] * pow(y[dy], Q - 2, Q) % Q
            o[i] = (o[i] + c) % Q
            for j in range(dy + 1):
                temp[i + j] = (temp[i + j] - c * y[j]) % Q
        return o[: deg(o) + 1]

    while deg(low):
        r = poly_rounded_div(high, low)
        r += [0] * (13 - len(r))
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] -= lm[i] * r[j]
                new[i + j] -= low[i] * r[j]
        nm = [x % Q for x in nm]
        new = [x % Q for x in new]
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], Q - 2, Q)
    return tuple(x * inv0 % Q for x in lm[:12])


def _f2_to_f12(c):
    # a + b i with i = w^6 - 9  ->  (a - 9b) + b w^6
    return f12([(c[0] - 9 * c[1])] + [0] * 5 + [c[1]] + [0] * 5)


_W = f12([0, 1] + [0] * 10)
_W2 = f12_mul(_W, _W)
_W3 = f12_mul(_W2, _W)


def twist(p):
    """Untwist a G2 point into E(Fq12): (x w^2, y w^3)."""
    if p is None:
        return None
    return (f12_mul(_f2_to_f12(p[0]), _W2), f12_mul(_f2_to_f12(p[1]), _W3))


def _cast_g1(p):
    return (f12([p[0]] + [0] * 11), f12([p[1]] + [0] * 11))


def _e12_double(p):
    x, y = p
    lam = f12_mul(f12_scalar(f12_mul(x, x), 3), f12_inv(f12_scalar(y, 2)))
    nx = f12_sub(f12_mul(lam, lam), f12_scalar(x, 2))
    ny = f12_sub(f12_mul(lam, f12_sub(x, nx)), y)
    return (nx, ny)


def _e12_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2 and y1 == y2:
        return _e12_double(p)
    if x1 == x2:
        return None
    lam = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
    nx = f12_sub(f12_sub(f12_mul(lam, lam), x1), x2)
    ny = f12_sub(f12_mul(lam, f12_sub(x1, nx)), y1)
    return (nx, ny)


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if x1 != x2:
        m = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
        return f12_sub(f12_mul(m, f12_sub(xt, x1)), f12_sub(yt, y1))
    if y1 == y2:
        m = f12_mul(f12_scalar(f12_mul(x1, x1), 3), f12_inv(f12_scalar(y1, 2)))
        return f12_sub(f12_mul(m, f12_sub(xt, x1)), f12_sub(yt, y1))
    return f12_sub(xt, x1)


def miller_loop(q2, p1):
    """Miller value (no final exponentiation) of an affine (G2, G1) pair."""
    qq = twist(q2)
    pp = _cast_g1(p1)
    r = qq
    f = F12_ONE
    for i in range(ATE.bit_length() - 2, -1, -1):
        f = f12_mul(f12_mul(f, f), _linefunc(r, r, pp))
        r = _e12_double(r)
        if (ATE >> i) & 1:
            f = f12_mul(f, _linefunc(r, qq, pp))
            r = _e12_add(r, qq)
    q1 = (f12_pow(qq[0], Q), f12_pow(qq[1], Q))
    nq2 = (f12_pow(q1[0], Q), f12_sub(F12_ZERO, f12_pow(q1[1], Q)))
    f = f12_mul(f, _linefunc(r, q1, pp))
    r = _e12_add(r, q1)
    f = f12_mul(f, _linefunc(r, nq2, pp))
    return f


def final_exponentiation(f):
    return f12_pow(f, (Q ** 12 - 1) // R)


def pairing_batch_is_one(pairs):
    """bn::pairing_batch(pairs) == Gt::one(); pairs with an infinity are skipped."""
    f = F12_ONE
    for p1, q2 in pairs:
        if p1 is None or q2 is None:
            continue
        f = f12_mul(f, miller_loop(q2, p1))
    return final_exponentiation(f) == F12_ONE


# --------------------------------------------------------------------------- scheme
def sign(msg, sk):
    """src/ecdsa.rs:26-35; sk is the canonical integer in [0, r)."""
    return g1_mul(hash_to_try_and_increment(msg), sk)


def verify(msg, sig, pk):
    """src/ecdsa.rs:49-64 -> status code."""
    try:
        h = hash_to_try_and_increment(msg)
    except Bn254Error as e:
        return e.code
    ok = pairing_batch_is_one([(h, pk), (sig, g2_neg(G2_GEN))])
    return OK if ok else VERIFICATION_FAILED


def check_public_keys(pk_g2, pk_g1):
    """src/ecdsa.rs:78-93 -> status code."""
    ok = pairing_batch_is_one([(G1_GEN, pk_g2), (pk_g1, g2_neg(G2_GEN))])
    return OK if ok else VERIFICATION_FAILED


def pk_g2_from_sk(sk):
    return g2_mul(G2_GEN, sk)  # src/types.rs:85-87


def pk_g1_from_sk(sk):
    return g1_mul(G1_GEN, sk)  # src/types.rs:155-157
