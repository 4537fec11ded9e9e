#!/usr/bin/env python3
"""Generates the tables and the carry-chain code of the cooperative pairing machine (bn254_b200/csrc/coop.cuh):

  bn254_b200/csrc/coop_mac.cuh     wide_mac (512-bit accumulate of an 8x8-limb product) and wide_redc (Montgomery
                                   reduction of the accumulator) as PTX carry chains, plus the portable forms
  bn254_b200/csrc/coop_tables.cuh  dot-product plans (which operand triples each of the six warps multiplies),
                                   the Miller-loop and final-exponentiation programs, Frobenius constants

    python scripts/gen_coop.py

Layout facts shared with coop.cuh: an Fq12 value is sum a_k w^k (a_k in Fq2, w^6 = xi = 9 + i); warp k owns a_k.
Shared memory holds the primary value P as six records of six Fq (x0, x1, x0+x1, y0, y1, y0+y1 with y = xi*x) and
the secondary value S as six triples (x0, x1, x0+x1).  A "triple slot" is the index of the first Fq of a triple:
P plain k -> 6k, P xi k -> 6k+3, S k -> 36+3k.
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "bn254_b200", "csrc")
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
RM = (1 << 256) % Q
QL = [(Q >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


# ------------------------------------------------------------------------------------------------ carry chains
def asm_block(lines, outs, ins):
    """one asm statement: outs = list of C lvalues ("+r"), ins = list of C rvalues ("r")"""
    ops = {}
    for i, o in enumerate(outs):
        ops[o] = "%%%d" % i
    for j, o in enumerate(ins):
        ops[o] = "%%%d" % (len(outs) + j)
    body = []
    for l in lines:
        op, args = l
        body.append('"%s %s;\\n\\t"' % (op, ", ".join(ops.get(a, a) for a in args)))
    s = "  asm(" + "\n      ".join(body) + "\n      : " + ", ".join('"+r"(%s)' % o for o in outs)
    s += "\n      : " + ", ".join('"r"(%s)' % o for o in ins) + ");\n"
    return s


def chain64(arr, first_pair, muls, b, carry_to):
    """One carry chain over the four 64-bit accumulators arr[first_pair .. first_pair+3]: muls = the four multiplicands
    (register expressions or immediates), each adding mul*b to one accumulator; the carry out goes to the counter
    `carry_to` (None: known to be zero).  The accumulators are 64-bit variables, unpacked to (lo, hi) inside the asm:
    this pins every (lo, hi) to one aligned register pair for the whole routine."""
    outs = ["%s[%d]" % (arr, first_pair + t) for t in range(4)]
    if carry_to is not None:
        outs.append(carry_to)
    regs = []
    for a in muls:
        if not a.startswith("0x") and a not in regs:
            regs.append(a)
    ins = regs + [b]
    name = {}
    for i, o in enumerate(outs):
        name[o] = "%%%d" % i
    for j, o in enumerate(ins):
        name[o] = "%%%d" % (len(outs) + j)
    body = ['"{\\n\\t.reg .u32 l0, h0, l1, h1, l2, h2, l3, h3;\\n\\t"']
    for t in range(4):
        body.append('"mov.b64 {l%d, h%d}, %s;\\n\\t"' % (t, t, name[outs[t]]))
    for t, a in enumerate(muls):
        an = name.get(a, a)
        body.append('"%s l%d, %s, %s, l%d;\\n\\t"' % ("mad.lo.cc.u32" if t == 0 else "madc.lo.cc.u32", t, an, name[b], t))
        last = t == 3 and carry_to is None
        body.append('"%s h%d, %s, %s, h%d;\\n\\t"' % ("madc.hi.u32" if last else "madc.hi.cc.u32", t, an, name[b], t))
    if carry_to is not None:
        body.append('"addc.u32 %s, %s, 0;\\n\\t"' % (name[carry_to], name[carry_to]))
    for t in range(4):
        body.append('"mov.b64 %s, {l%d, h%d};\\n\\t"' % (name[outs[t]], t, t))
    body.append('"}"')
    s = "  asm(" + "\n      ".join(body) + "\n      : "
    s += ", ".join(('"+l"(%s)' if o.startswith(arr + "[") else '"+r"(%s)') % o for o in outs)
    s += "\n      : " + ", ".join('"r"(%s)' % o for o in ins) + ");\n"
    return s


def rows(o, a_of, b_of, pre=None, post=None):
    """the 8 rows of T += a * b in the E / O split; a_of(j) / b_of(i) give operand expressions"""

    def cnt(arr, word):  # counter for a carry into word `word` of E / index `word` of O
        if word >= 15:
            return None
        assert word in (8, 10, 12, 14)
        return "C[%d]" % ((word - 8) // 2 + (0 if arr == "E" else 4))

    for i in range(8):
        if pre:
            o.append(pre(i))
        b = b_of(i)
        if i % 2 == 0:   # E words i..i+7 (a even), O indices i..i+7 (a odd)
            o.append(chain64("E", i // 2, [a_of(j) for j in (0, 2, 4, 6)], b, cnt("E", i + 8)))
            o.append(chain64("O", i // 2, [a_of(j) for j in (1, 3, 5, 7)], b, cnt("O", i + 8)))
        else:            # E words i+1..i+8 (a odd), O indices i-1..i+6 (a even)
            o.append(chain64("E", (i + 1) // 2, [a_of(j) for j in (1, 3, 5, 7)], b, cnt("E", i + 9)))
            o.append(chain64("O", (i - 1) // 2, [a_of(j) for j in (0, 2, 4, 6)], b, cnt("O", i + 7)))
        if post:
            o.append(post(i))


def word(arr, w):
    return "(uint32_t)%s[%d]" % (arr, w // 2) if w % 2 == 0 else "(uint32_t)(%s[%d] >> 32)" % (arr, w // 2)


def gen_mac():
    o = []
    o.append("// GENERATED by scripts/gen_coop.py -- do not edit.\n")
    o.append("// T = E + (O << 32) is a 512-bit accumulator held as two arrays of 64-bit registers: E[s] covers limbs 2s, 2s+1 and takes\n")
    o.append("// the partial products that start on an even limb, O[s] covers limbs 2s+1, 2s+2 and takes those that start on an odd\n")
    o.append("// limb, so every (mad.lo.cc, madc.hi.cc) pair is one IMAD.WIDE.U32.X on a carry chain of four.  A chain's carry-out\n")
    o.append("// is not rippled upwards: it is counted in C[j] (limb 8+2j of E) / C[4+j] (index 8+2j of O) and folded in once.\n")
    o.append("#pragma once\n#include \"fq.cuh\"\n\nnamespace bn {\n\n")
    o.append("#if defined(__CUDA_ARCH__)\n")
    o.append("BN_FN void wide_mac(uint64_t (&E)[8], uint64_t (&O)[8], uint32_t (&C)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {\n")
    rows(o, lambda j: "a[%d]" % j, lambda i: "b[%d]" % i)
    o.append("}\n\n")
    o.append("// Montgomery reduction of the accumulator: returns a value congruent to T / 2^256, below T / 2^256 + q.  Requires T < 6.2 q^2.\n")
    o.append("// The rows m_i * q are added exactly like product rows (a = q, b = m_i); the true low limb of round i is\n")
    o.append("// E.limb[i] + O.limb[i-1] + c, c being the carry of the limb below (which the round before made zero).\n")
    o.append("BN_FN fq wide_redc(uint64_t (&E)[8], uint64_t (&O)[8], uint32_t (&C)[8]) {\n")
    o.append("  uint32_t c = 0, m;\n")

    def pre(i):
        low = word("E", i) + ((" + " + word("O", i - 1)) if i else "") + " + c"
        return "  m = (%s) * K_QINV_NEG;\n" % low

    def post(i):
        orv = word("E", i) + ((" | " + word("O", i - 1)) if i else "") + " | c"
        return "  c = (%s) != 0 ? 1u : 0u;\n" % orv

    rows(o, lambda j: "0x%08x" % QL[j], lambda i: "m", pre, post)
    # r = E limbs 8..15 + O indices 7..14 + carry counters + c
    o.append("  uint32_t e[8], p[8];\n")
    for k in range(8):
        o.append("  e[%d] = %s;\n  p[%d] = %s;\n" % (k, word("E", 8 + k), k, word("O", 7 + k)))

    def addchain(dst, src):
        lines = []
        for k in range(8):
            op = "add.cc.u32" if k == 0 else ("addc.u32" if k == 7 else "addc.cc.u32")
            lines.append((op, ["%s[%d]" % (dst, k), "%s[%d]" % (dst, k), src(k)]))
        srcs = [x for x in dict.fromkeys(src(k) for k in range(8)) if x != "0"]
        return asm_block(lines, ["%s[%d]" % (dst, k) for k in range(8)], srcs)

    o.append(addchain("e", lambda k: "p[%d]" % k))
    # counters: C[j] -> limb 8+2j of E = e[2j] ; C[4+j] -> index 8+2j of O = p[2j+1] -> e[2j+1]
    o.append("  C[0] += c;  // a counter holds at most a few dozen\n")
    o.append(addchain("e", lambda k: "C[%d]" % (k // 2) if k % 2 == 0 else "C[%d]" % (4 + k // 2)))
    o.append("  fq r;\n")
    o.append("#pragma unroll\n  for (int k = 0; k < 8; k++) r.l[k] = e[k];\n")
    o.append("  return r;  // not yet canonical: the caller subtracts q once or twice\n}\n")
    # ---- the accumulator as two plain 256-bit halves (no reduction): T = L + H * 2^256
    o.append("// The accumulator as two plain halves, T = L + H * 2^256 (H takes the carry counters).  With the halves of the three Karatsuba\n")
    o.append("// components in hand, ONE reduction per Fq2 coefficient is enough: REDC is linear, (T0 - T1) / 2^256 = redc_low(L0 - L1 mod 2^256)\n")
    o.append("// + H0 - H1 - borrow (mod q), and the high halves are only added and subtracted (coop.cuh coop_dot_block).\n")
    o.append("BN_FN void wide_split(const uint64_t (&E)[8], const uint64_t (&O)[8], const uint32_t (&C)[8], uint32_t (&L)[8], uint32_t (&H)[8]) {\n")
    o.append("  uint32_t t[16];\n")
    lines = []
    for w in range(16):
        op = "add.cc.u32" if w == 1 else ("addc.u32" if w == 15 else "addc.cc.u32")
        if w == 0:
            continue
        lines.append((op, ["t[%d]" % w, "x[%d]" % w, "y[%d]" % (w - 1)]))
    o.append("  uint32_t x[16], y[15];\n")
    for w in range(16):
        o.append("  x[%d] = %s;\n" % (w, word("E", w)))
    for u in range(15):
        o.append("  y[%d] = %s;\n" % (u, word("O", u)))
    o.append("  t[0] = x[0];\n")
    body = []
    outs = ["t[%d]" % w for w in range(1, 16)]
    ins_ = ["x[%d]" % w for w in range(1, 16)] + ["y[%d]" % u for u in range(15)]
    ops_ = {}
    for i, oo in enumerate(outs):
        ops_[oo] = "%%%d" % i
    for j, oo in enumerate(ins_):
        ops_[oo] = "%%%d" % (len(outs) + j)
    for op, args in lines:
        body.append('"%s %s;\\n\\t"' % (op, ", ".join(ops_[a] for a in args)))
    o.append("  asm(" + "\n      ".join(body) + "\n      : " + ", ".join('"=&r"(%s)' % oo for oo in outs))
    o.append("\n      : " + ", ".join('"r"(%s)' % oo for oo in ins_) + ");\n")
    # counters into the high half: C[j] -> limb 8+2j, C[4+j] -> limb 9+2j
    body = []
    for k in range(8):
        op = "add.cc.u32" if k == 0 else ("addc.u32" if k == 7 else "addc.cc.u32")
        body.append('"%s %%%d, %%%d, %%%d;\\n\\t"' % (op, k, k, 8 + k))
    o.append("  asm(" + "\n      ".join(body) + "\n      : " + ", ".join('"+r"(t[%d])' % (8 + k) for k in range(8)))
    o.append("\n      : " + ", ".join('"r"(C[%d])' % ((k // 2) if k % 2 == 0 else (4 + k // 2)) for k in range(8)) + ");\n")
    o.append("#pragma unroll\n  for (int k = 0; k < 8; k++) {\n    L[k] = t[k];\n    H[k] = t[8 + k];\n  }\n}\n")
    o.append("#else\n")
    o.append("""// portable form (host simulation): the whole accumulator lives in E as 16 limbs
BN_FN uint32_t wide_limb(const uint64_t (&E)[8], int w) { return (uint32_t)(E[w >> 1] >> (32 * (w & 1))); }
BN_FN void wide_set_limb(uint64_t (&E)[8], int w, uint32_t v) {
  E[w >> 1] = (E[w >> 1] & ~((uint64_t)0xffffffffu << (32 * (w & 1)))) | ((uint64_t)v << (32 * (w & 1)));
}
BN_FN void wide_row(uint64_t (&E)[8], int i, const uint32_t* a, uint32_t b) {
  uint64_t c = 0;
  for (int j = 0; j < 8; j++) {
    c += (uint64_t)a[j] * b + wide_limb(E, i + j);
    wide_set_limb(E, i + j, (uint32_t)c);
    c >>= 32;
  }
  for (int k = i + 8; k < 16; k++) {
    c += wide_limb(E, k);
    wide_set_limb(E, k, (uint32_t)c);
    c >>= 32;
  }
}
BN_FN void wide_mac(uint64_t (&E)[8], uint64_t (&O)[8], uint32_t (&C)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  (void)O;
  (void)C;
  for (int i = 0; i < 8; i++) wide_row(E, i, a, b[i]);
}
BN_FN fq wide_redc(uint64_t (&E)[8], uint64_t (&O)[8], uint32_t (&C)[8]) {
  (void)O;
  (void)C;
  for (int i = 0; i < 8; i++) wide_row(E, i, K_Q, wide_limb(E, i) * K_QINV_NEG);
  fq r;
  for (int k = 0; k < 8; k++) r.l[k] = wide_limb(E, 8 + k);
  return r;
}
""")
    o.append("BN_FN void wide_split(const uint64_t (&E)[8], const uint64_t (&O)[8], const uint32_t (&C)[8], uint32_t (&L)[8], uint32_t (&H)[8]) {\n")
    o.append("  (void)O;\n  (void)C;\n  for (int k = 0; k < 8; k++) {\n    L[k] = wide_limb(E, k);\n    H[k] = wide_limb(E, 8 + k);\n  }\n}\n")
    o.append("#endif\n\n")
    o.append("// (L + m q) / 2^256 for a plain 256-bit L, m = -L / q mod 2^256: an integer in [0, q]\n")
    o.append("BN_FN fq redc_low(const uint32_t (&L)[8]) {\n  uint64_t E[8], O[8];\n  uint32_t C[8];\n")
    o.append("#if defined(__CUDA_ARCH__)\n#pragma unroll\n#endif\n  for (int i = 0; i < 8; i++) {\n    E[i] = i < 4 ? ((uint64_t)L[2 * i] | ((uint64_t)L[2 * i + 1] << 32)) : 0;\n    O[i] = 0;\n    C[i] = 0;\n  }\n")
    o.append("  return wide_redc(E, O, C);\n}\n\n")
    o.append("// one Montgomery product through the accumulator (fq.cuh BN_FQ_MUL_WIDE): T = a b < q^2, REDC leaves less than 1.19 q\n")
    o.append("BN_FN fq fq_mul_wide(const fq& a, const fq& b) {\n  uint64_t E[8], O[8];\n  uint32_t C[8];\n")
    o.append("#if defined(__CUDA_ARCH__)\n#pragma unroll\n#endif\n  for (int i = 0; i < 8; i++) {\n    E[i] = 0;\n    O[i] = 0;\n    C[i] = 0;\n  }\n")
    o.append("  wide_mac(E, O, C, a.l, b.l);\n  return fq_csub(wide_redc(E, O, C));\n}\n\n}  // namespace bn\n")
    return "".join(o)


# ------------------------------------------------------------------------------------------------ plans
P_PLAIN = lambda k: 6 * k
P_XI = lambda k: 6 * k + 3
S_T = lambda k: 36 + 3 * k
DEST_NONE = 15
POST_NONE, POST_CYC_MINUS, POST_CYC_PLUS = 0, 1, 2


def row(entries=(), dbl_upto=0, neg_from=7, dest=DEST_NONE, post=POST_NONE):
    """entries: (X slot, Y slot); the first dbl_upto entries count twice, entries from neg_from on are subtracted
    (the machine doubles / negates the X operand as it loads it)"""
    es = list(entries)
    assert len(es) <= 6
    dbl = sum(1 << i for i in range(len(es)) if i < dbl_upto)
    neg = sum(1 << i for i in range(len(es)) if i >= neg_from)
    weight = len(es) + min(dbl_upto, len(es))
    return dict(n=len(es), dbl=dbl, neg=neg, dest=dest, post=post, e=es, weight=weight)


def plan_mul():
    rows = []
    for k in range(6):
        es = []
        for i in range(6):
            j = (k - i) % 6
            es.append((S_T(i), P_PLAIN(j) if i + j == k else P_XI(j)))
        rows.append(row(es, dest=k))
    return rows


def plan_sqr():
    rows = []
    for k in range(6):
        cross, sq = [], []
        for i in range(6):
            for j in range(i, 6):
                if (i + j) % 6 != k:
                    continue
                y = P_PLAIN(j) if i + j == k else P_XI(j)
                (sq if i == j else cross).append((P_PLAIN(i), y))
        es = cross + sq
        rows.append(row(es, dbl_upto=len(cross), dest=k))
    return rows


def plan_sparse(buf):
    # P * (l0 + l3 w^3 + l4 w^4); the line coefficients sit in S triples (0, 3, 4) [buffer 0] or (1, 2, 5) [buffer 1]:
    # consecutive line sets alternate between the two so that the next one can be fetched while this one is read
    t0, t3, t4 = ((0, 3, 4), (1, 2, 5))[buf]
    rows = []
    for k in range(6):
        es = [(P_PLAIN(k), S_T(t0))]
        for d, t in ((3, t3), (4, t4)):
            i = (k - d) % 6
            es.append((P_PLAIN(i) if i + d == k else P_XI(i), S_T(t)))
        rows.append(row(es, dest=k))
    return rows


def plan_cyclo():
    # Granger-Scott: Fq4 pairs (a0,a3), (a1,a4), (a2,a5); out_k = 3 t -/+ 2 a_k
    rows = [None] * 6
    rows[0] = row([(P_PLAIN(0), P_PLAIN(0)), (P_PLAIN(3), P_XI(3))], dest=0, post=POST_CYC_MINUS)   # t0
    rows[3] = row([(P_PLAIN(0), P_PLAIN(3))], dbl_upto=1, dest=3, post=POST_CYC_PLUS)                 # t1 = 2 a0 a3
    rows[1] = row([(P_PLAIN(2), P_XI(5))], dbl_upto=1, dest=1, post=POST_CYC_PLUS)                    # xi t5 = 2 a2 (xi a5)
    rows[4] = row([(P_PLAIN(2), P_PLAIN(2)), (P_PLAIN(5), P_XI(5))], dest=4, post=POST_CYC_MINUS)   # t4
    rows[2] = row([(P_PLAIN(1), P_PLAIN(1)), (P_PLAIN(4), P_XI(4))], dest=2, post=POST_CYC_MINUS)   # t2
    rows[5] = row([(P_PLAIN(1), P_PLAIN(4))], dbl_upto=1, dest=5, post=POST_CYC_PLUS)                 # t3 = 2 a1 a4
    return rows


def plan_inv_a():
    # N = (n0, n1, n2) in P records 0, 2, 4.  c0 = n0^2 - xi n1 n2 ; c1 = xi n2^2 - n0 n1 ; c2 = n1^2 - n0 n2  -> S records 0..2
    n0, n1, n2 = 0, 2, 4
    rows = [row() for _ in range(6)]
    rows[0] = row([(P_PLAIN(n0), P_PLAIN(n0)), (P_PLAIN(n1), P_XI(n2))], neg_from=1, dest=8 + 0)
    rows[1] = row([(P_PLAIN(n2), P_XI(n2)), (P_PLAIN(n0), P_PLAIN(n1))], neg_from=1, dest=8 + 1)
    rows[2] = row([(P_PLAIN(n1), P_PLAIN(n1)), (P_PLAIN(n0), P_PLAIN(n2))], neg_from=1, dest=8 + 2)
    return rows


def plan_inv_b():
    # t = n0 c0 + xi (n2 c1 + n1 c2) -> S record 3
    rows = [row() for _ in range(6)]
    rows[0] = row([(P_PLAIN(0), S_T(0)), (P_XI(4), S_T(1)), (P_XI(2), S_T(2))], dest=8 + 3)
    return rows


def plan_inv_c():
    # N^-1 = (c0, c1, c2) * t^-1 (t^-1 in S record 3) -> P records 0, 2, 4 ; odd records become zero
    rows = []
    for k in range(6):
        if k % 2 == 0:
            rows.append(row([(S_T(k // 2), S_T(3))], dest=k))
        else:
            rows.append(row([], dest=k))
    return rows


PLANS = [("MUL", plan_mul), ("SQR", plan_sqr), ("SPARSE_A", lambda: plan_sparse(0)), ("SPARSE_B", lambda: plan_sparse(1)), ("CYCLO", plan_cyclo), ("INV_A", plan_inv_a), ("INV_B", plan_inv_b),
         ("INV_C", plan_inv_c)]

# ------------------------------------------------------------------------------------------------ programs
OPS = ["END", "DOT", "LINE", "LOADP", "LOADS", "STORE", "COPY_PS_CONJ", "CONJP", "FROBP", "INVT", "ONE", "CHECK", "LOADF", "STOREF", "XLANE", "LOADFS"]
OPC = {n: i for i, n in enumerate(OPS)}
PLAN_ID = {n: i for i, (n, _) in enumerate(PLANS)}
CONJ_FLAG = 0x80


def ins(op, a=0, b=0):
    assert 0 <= a < 256 and 0 <= b < 65536
    return OPC[op] | (a << 8) | (b << 16)


ATE = [1, 0, 1, 0, 0, 0, -1, 0, -1, 0, 0, 0, -1, 0, 1, 0, -1, 0, 0, -1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, 1, 0,
       0, -1, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, -1, 0, -1, 0, 0, 1, 0, 0, 0, -1, 0, 0, -1, 0, 1, 0, 1, 0, 0, 0]
U = 4965661367192848881


def dot(plan, prefetch=False):
    """DOT instruction; prefetch: fetch the next line set (a running counter in the machine; set m goes to S buffer m & 1)
    while the dot product runs"""
    return ins("DOT", PLAN_ID[plan], 1 if prefetch else 0)


def prog_miller(pairs):
    """f = product over `pairs` line streams: per step the lines of pair 0, pair 1, ... are consumed in order; line sets
    are numbered in exactly this order by the line kernel.  Line set m goes to S buffer m & 1; every arithmetic
    instruction fetches the next line set into the other buffer."""
    ops = []  # ("SQR",) or ("SPARSE", m)
    idx = 0

    def lines():
        nonlocal idx
        for _ in range(pairs):
            ops.append(("SPARSE", idx))
            idx += 1

    for k in range(64):
        if k > 0:
            ops.append(("SQR",))
        lines()
        if ATE[k]:
            lines()
    lines()
    lines()
    total = 87 * pairs
    assert idx == total
    p = [ins("ONE"), ins("LINE", 0, 0)]
    fetched = 0  # highest line set already requested
    for o in ops:
        nxt = fetched + 1 if fetched + 1 < total else None
        if o[0] == "SQR":
            # the set needed next is already in place (requested by the instruction before); fetch nothing new
            p.append(dot("SQR"))
        else:
            m = o[1]
            assert m <= fetched
            pf = False
            if m == fetched and nxt is not None:
                pf = True
                fetched = nxt
            p.append(dot("SPARSE_A" if m % 2 == 0 else "SPARSE_B", pf))
    return p


MULTI_K = 8  # pairs per lane in the multi-pairing program (they share one squaring chain)


def prog_multi_miller(k):
    """Every lane folds k pairs into one Miller value (shared squarings), then the 32 lanes of the block are multiplied
    together by a butterfly: S <- P of lane ^ s, P <- S * P for s = 16, 8, 4, 2, 1.  Every lane ends with the block's product."""
    p = prog_miller(k)
    for sft in (16, 8, 4, 2, 1):
        p.append(ins("XLANE", sft))
        p.append(ins("DOT", PLAN_ID["MUL"]))
    return p


def wnaf(n, w):
    d = []
    while n:
        if n & 1:
            z = n % (1 << (w + 1))
            if z >= (1 << w):
                z -= 1 << (w + 1)
            d.append(z)
            n -= z
        else:
            d.append(0)
        n //= 2
    return d


WROW = 5  # items per warp in the warp-local layout (6 coefficients x 5 items = 30 lanes)
HROW = 16  # items per group in the half-warp layout (a warp = two coefficients x 16 items)
EXP_TMP_SLOT = 6  # global slot that holds base^3 during an exponentiation


def prog_exp_neg_u(p, base_slot):
    """P <- conj(P^u) with P == G[base_slot] on entry.  Signed digits {+-1, +-3} of u (18 non-zero instead of the 28 one
    bits): in the cyclotomic subgroup the inverse is the conjugate, so a negative digit multiplies by the conjugated table
    entry (LOADS with the conjugation flag).  S is reloaded only when the table entry or its sign changes."""
    MUL = ins("DOT", PLAN_ID["MUL"])
    CYC = ins("DOT", PLAN_ID["CYCLO"])
    digs = wnaf(U, 2)
    assert sum(d << i for i, d in enumerate(digs)) == U and set(abs(d) for d in digs if d) <= {1, 3}
    # base^3 = base^2 * base -> G[EXP_TMP_SLOT]
    p.append(CYC)
    p.append(ins("LOADS", base_slot))
    p.append(MUL)
    p.append(ins("STORE", EXP_TMP_SLOT))
    slot_of = {1: base_slot, 3: EXP_TMP_SLOT}
    top = digs[-1]
    assert top > 0
    if top == 3:
        cur = None          # P already holds base^3
    else:
        p.append(ins("LOADP", base_slot))
        cur = None
    in_s = (1, False)       # S holds base (loaded above), not conjugated
    for d in reversed(digs[:-1]):
        p.append(CYC)
        if d:
            want = (abs(d), d < 0)
            if want != in_s:
                p.append(ins("LOADS", slot_of[want[0]] | (CONJ_FLAG if want[1] else 0)))
                in_s = want
            p.append(MUL)
    p.append(ins("CONJP"))


def prog_final_exp():
    """P <- P^((q^12-1)/r), same chain as pairing.cuh final_exponentiation(); global slots 0..5 hold intermediates"""
    p = []
    MUL = ins("DOT", PLAN_ID["MUL"])
    CYC = ins("DOT", PLAN_ID["CYCLO"])
    # easy part
    p.append(ins("STORE", 0))                      # G0 = f
    p.append(ins("COPY_PS_CONJ"))                  # S = conj(f)
    p.append(MUL)                                  # P = f conj(f) = N (even coefficients)
    p.append(ins("DOT", PLAN_ID["INV_A"]))
    p.append(ins("DOT", PLAN_ID["INV_B"]))
    p.append(ins("INVT"))
    p.append(ins("DOT", PLAN_ID["INV_C"]))         # P = N^-1
    p.append(ins("LOADS", 0 | CONJ_FLAG))          # S = conj(f)
    p.append(MUL)                                  # P = f^-1
    p.append(MUL)                                  # P = conj(f) f^-1 = f^(q^6-1) = t
    p.append(ins("STORE", 1))
    p.append(ins("FROBP", 2))
    p.append(ins("LOADS", 1))
    p.append(MUL)                                  # P = e
    p.append(ins("STORE", 0))                      # G0 = e
    # hard part
    prog_exp_neg_u(p, 0)                           # A
    p.append(CYC)                                  # B
    p.append(ins("STORE", 2))                      # G2 = B
    p.append(CYC)                                  # C
    p.append(ins("LOADS", 2))
    p.append(MUL)                                  # D = C B
    p.append(ins("STORE", 3))                      # G3 = D
    prog_exp_neg_u(p, 3)                           # E
    p.append(ins("STORE", 4))                      # G4 = E
    p.append(CYC)                                  # F
    p.append(ins("STORE", 5))                      # G5 = F
    prog_exp_neg_u(p, 5)                           # G
    p.append(ins("CONJP"))                         # I = conj(G)
    p.append(ins("LOADS", 4))
    p.append(MUL)                                  # J = I E
    p.append(ins("LOADS", 3 | CONJ_FLAG))
    p.append(MUL)                                  # K = J conj(D)
    p.append(ins("STORE", 5))                      # G5 = K
    p.append(ins("LOADS", 2))
    p.append(MUL)                                  # L = K B
    p.append(ins("STORE", 2))                      # G2 = L
    p.append(ins("LOADP", 5))                      # P = K
    p.append(ins("LOADS", 4))
    p.append(MUL)                                  # M = K E
    p.append(ins("LOADS", 0))
    p.append(MUL)                                  # N = M e
    p.append(ins("STORE", 4))                      # G4 = N
    p.append(ins("LOADP", 2))
    p.append(ins("FROBP", 1))                      # O = frob(L)
    p.append(ins("LOADS", 4))
    p.append(MUL)                                  # P' = O N
    p.append(ins("STORE", 4))
    p.append(ins("LOADP", 5))
    p.append(ins("FROBP", 2))                      # Q = frob2(K)
    p.append(ins("LOADS", 4))
    p.append(MUL)                                  # R = Q P'
    p.append(ins("STORE", 4))
    p.append(ins("LOADP", 2))                      # L
    p.append(ins("LOADS", 0 | CONJ_FLAG))          # conj(e)
    p.append(MUL)                                  # T = conj(e) L
    p.append(ins("FROBP", 3))                      # U = frob3(T)
    p.append(ins("LOADS", 4))
    p.append(MUL)                                  # result = U R
    return p


def plan_info():
    """per plan: set of P records whose xi variant is read, and whether the plan overwrites P"""
    info = {}
    for pid, (name, fn) in enumerate(PLANS):
        rows_ = fn()
        xi = set()
        for r in rows_:
            for e in r["e"]:
                for slot in e:
                    if slot < 36 and slot % 6 == 3:
                        xi.add(slot // 6)
        writes_p = any(r["dest"] < 6 for r in rows_)
        info[pid] = (xi, writes_p)
    return info


def add_xi_skip(prog):
    """DOT instructions that write P get, in bits 9..14 of arg2, the records whose xi variant nobody reads before P is
    rewritten -- the machine then skips computing them (ten modular additions per record)."""
    info = plan_info()
    rewrites = {OPC["LOADP"], OPC["LOADF"], OPC["ONE"]}
    out = list(prog)
    for t, w in enumerate(prog):
        if (w & 0xff) != OPC["DOT"] or not info[(w >> 8) & 0xff][1]:
            continue
        need = set()
        for u in prog[t + 1:]:
            op = u & 0xff
            if op == OPC["DOT"]:
                xi, wp = info[(u >> 8) & 0xff]
                need |= xi
                if wp:
                    break
            elif op in rewrites:
                break
            elif op == OPC["FROBP"]:
                break  # rewrites every record (with all variants) from the plain parts; CONJP only rewrites the odd ones
        skip = 0
        for k in range(6):
            if k not in need:
                skip |= 1 << k
        out[t] = w | (skip << (16 + 9))
    return out


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_mul(a, a)
        e >>= 1
    return r


def limbs(x):
    return ", ".join("0x%08xu" % ((x >> (32 * i)) & 0xFFFFFFFF) for i in range(8))


def gen_tables():
    o = ["// GENERATED by scripts/gen_coop.py -- do not edit.", "#pragma once", "", "namespace bn {", ""]
    o.append("enum { " + ", ".join("COP_%s = %d" % (n, i) for i, n in enumerate(OPS)) + " };")
    o.append("enum { " + ", ".join("CPLAN_%s = %d" % (n, i) for i, (n, _) in enumerate(PLANS)) + ", CPLAN_COUNT = %d };" % len(PLANS))
    o.append("#define COOP_CONJ_FLAG 0x%02x" % CONJ_FLAG)
    o.append("#define COOP_GSLOTS %d" % (EXP_TMP_SLOT + 1))
    o.append("#define COOP_MULTI_K %d" % MULTI_K)
    o.append("#define COOP_DEST_NONE %d" % DEST_NONE)
    o.append("// plan row: word 0 = n | double-X mask << 4 | negate-X mask << 10 | dest << 16 | post << 20 | (weight <= 3) << 24 ;")
    o.append("// words 1..6 = byte offset of the X triple | byte offset of the Y triple << 16 (slot * slot bytes: 1024 in the")
    o.append("// block layout = 32 items per row, %d in the warp-local layout = %d items per row, %d in the half-warp layout)" % (32 * WROW, WROW, 32 * HROW))
    o.append("#define COOPW_ROW %d" % WROW)
    o.append("#define COOPH_ROW %d" % HROW)
    for tname, slot_bytes in (("K_COOP_PLANS", 1024), ("K_COOP_PLANS_W", 32 * WROW), ("K_COOP_PLANS_H", 32 * HROW)):
        o.append("BN_CONST uint32_t %s[CPLAN_COUNT][6][7] = {" % tname)
        for name, fn in PLANS:
            rows = fn()
            o.append("  {  // %s" % name)
            for r in rows:
                w0 = r["n"] | (r["dbl"] << 4) | (r["neg"] << 10) | (r["dest"] << 16) | (r["post"] << 20) | ((1 if r["weight"] <= 3 else 0) << 24)
                es = ["0x%08x" % ((x * slot_bytes) | ((y * slot_bytes) << 16)) for x, y in r["e"]] + ["0"] * (6 - len(r["e"]))
                o.append("    {0x%05x, %s}," % (w0, ", ".join(es)))
            o.append("  },")
        o.append("};")
    for name, prog0 in (("VERIFY", prog_miller(2) + prog_final_exp() + [ins("CHECK"), ins("END")]),
                       ("MILLER1", prog_miller(1) + [ins("STOREF"), ins("END")]),
                       # one full pairing check per item: Miller loop of ONE line stream, final exponentiation, verdict
                       ("PAIRING1", prog_miller(1) + prog_final_exp() + [ins("CHECK"), ins("END")]),
                       ("MILLER2", prog_miller(2) + [ins("STOREF"), ins("END")]),
                       ("MULTI", prog_multi_miller(MULTI_K) + [ins("STOREF"), ins("END")]),
                       # the same with fewer pairs per lane: the last, partial wave of a multi-pairing is re-cut into more blocks
                       ("MULTI4", prog_multi_miller(4) + [ins("STOREF"), ins("END")]),
                       ("MULTI2", prog_multi_miller(2) + [ins("STOREF"), ins("END")]),
                       ("MULTI1", prog_multi_miller(1) + [ins("STOREF"), ins("END")]),
                       # finish of an aggregate check: Miller value of ONE line stream (sum of signatures, -G2) times the value in
                       # fio (the product of the exchanged partials), final exponentiation, verdict
                       ("FINISH", prog_miller(1) + [ins("LOADFS"), ins("DOT", PLAN_ID["MUL"])] + prog_final_exp() + [ins("CHECK"), ins("END")]),
                       ("FINALEXP", [ins("LOADF")] + prog_final_exp() + [ins("STOREF"), ins("CHECK"), ins("END")])):
        prog = add_xi_skip(prog0)
        o.append("#define K_COOP_PROG_%s_LEN %d" % (name, len(prog)))
        o.append("BN_CONST uint32_t K_COOP_PROG_%s[%d] = {" % (name, len(prog)))
        for i in range(0, len(prog), 12):
            o.append("    " + ", ".join("0x%08x" % w for w in prog[i:i + 12]) + ",")
        o.append("};")
    # Frobenius: a_k -> conj^n(a_k) * xi^(k (q^n - 1) / 6)
    o.append("// K_COOP_FROB[n-1][k] = xi^(k (q^n - 1) / 6) as (re, im), Montgomery form")
    o.append("BN_CONST uint32_t K_COOP_FROB[3][6][16] = {")
    for n in (1, 2, 3):
        o.append("  {")
        for k in range(6):
            g = f2_pow((9, 1), k * (Q ** n - 1) // 6)
            o.append("    {%s,\n     %s}," % (limbs(g[0] * RM % Q), limbs(g[1] * RM % Q)))
        o.append("  },")
    o.append("};")
    o.append("")
    o.append("}  // namespace bn")
    return "\n".join(o) + "\n"


def main():
    with open(os.path.join(CSRC, "coop_mac.cuh"), "w") as f:
        f.write(gen_mac())
    with open(os.path.join(CSRC, "coop_tables.cuh"), "w") as f:
        f.write(gen_tables())
    print("wrote coop_mac.cuh, coop_tables.cuh")


if __name__ == "__main__":
    main()
