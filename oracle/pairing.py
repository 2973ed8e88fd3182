"""ORACLE — TEST INFRASTRUCTURE ONLY.  BN254 optimal-ate pairing in plain Python big integers.

Textbook construction (Fp12 = Fp[w] / (w^12 - 18 w^6 + 82), sextic twist, Miller loop over
6u+2 with the two Frobenius corrections, final exponentiation by (p^12 - 1) / r).  Used by the
verifier restatement to check the KZG opening equation e(W, [s]G2) = e(P, G2) without trusting the
prover's own arithmetic.  Slow (seconds per pairing) and only used in tests.
"""
P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE = 63
FQ12_MOD = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]   # w^12 = 18 w^6 - 82

G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


class FQP:
    """element of Fp[x]/(modulus), coefficients low-to-high"""
    __slots__ = ("c",)
    deg = 12
    mod = FQ12_MOD

    def __init__(self, c):
        self.c = [v % P for v in c]

    @classmethod
    def one(cls):
        return cls([1] + [0] * (cls.deg - 1))

    @classmethod
    def zero(cls):
        return cls([0] * cls.deg)

    def __add__(self, o):
        return type(self)([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return type(self)([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return type(self)([-a for a in self.c])

    def __eq__(self, o):
        return self.c == o.c

    def scale(self, k):
        return type(self)([a * k for a in self.c])

    def __mul__(self, o):
        d = self.deg
        b = [0] * (2 * d - 1)
        for i, x in enumerate(self.c):
            if x:
                for j, y in enumerate(o.c):
                    b[i + j] += x * y
        for top in range(2 * d - 2, d - 1, -1):
            t = b[top]
            if t:
                for i, m in enumerate(self.mod):
                    if m:
                        b[top - d + i] -= t * m
        return type(self)(b[:d])

    def inv(self):
        # extended Euclid over Fp[x]
        d = self.deg
        lm, hm = [1] + [0] * d, [0] * (d + 1)
        low, high = self.c + [0], list(self.mod) + [1]

        def degree(p):
            i = len(p) - 1
            while i and p[i] % P == 0:
                i -= 1
            return i

        def poly_div(a, b):
            dega, degb = degree(a), degree(b)
            temp, o = list(a), [0] * len(a)
            for i in range(dega - degb, -1, -1):
                o[i] = (o[i] + temp[degb + i] * pow(b[degb], -1, P)) % P
                for c in range(degb + 1):
                    temp[c + i] = (temp[c + i] - o[c] * 0) % P
                for c in range(degb + 1):
                    temp[c + i] = (temp[c + i] - o[i] * b[c]) % P
            return o[:degree(o) + 1]
        while degree(low):
            r = poly_div(high, low)
            r += [0] * (d + 1 - len(r))
            nm, new = list(hm), list(high)
            for i in range(d + 1):
                for jx in range(d + 1 - i):
                    nm[i + jx] -= lm[i] * r[jx]
                    new[i + jx] -= low[i] * r[jx]
            nm = [x % P for x in nm]
            new = [x % P for x in new]
            lm, low, hm, high = nm, new, lm, low
        return type(self)(lm[:d]).scale(pow(low[0], -1, P))

    def pow(self, e):
        result, base = type(self).one(), self
        while e:
            if e & 1:
                result = result * base
            base = base * base
            e >>= 1
        return result


class FQ2(FQP):
    deg = 2
    mod = [1, 0]        # i^2 = -1


class FQ12(FQP):
    pass


W = FQ12([0, 1] + [0] * 10)
W2 = W * W
W3 = W2 * W


def _cast_g1(pt):
    x, y = pt
    return (FQ12([x] + [0] * 11), FQ12([y] + [0] * 11))


def twist(pt):
    (x0, x1), (y0, y1) = pt
    nx = FQ12([x0 - 9 * x1] + [0] * 5 + [x1] + [0] * 5)
    ny = FQ12([y0 - 9 * y1] + [0] * 5 + [y1] + [0] * 5)
    return (nx * W2, ny * W3)


def _double(pt):
    x, y = pt
    m = (x * x).scale(3) * (y.scale(2)).inv()
    nx = m * m - x.scale(2)
    return (nx, m * (x - nx) - y)


def _add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if y1 == y2:
            return _double(p1)
        return None
    m = (y2 - y1) * (x2 - x1).inv()
    nx = m * m - x1 - x2
    return (nx, m * (x1 - nx) - y1)


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if not (x1 == x2):
        m = (y2 - y1) * (x2 - x1).inv()
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1).scale(3) * (y1.scale(2)).inv()
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q_twisted, p_cast):
    r_pt, f = q_twisted, FQ12.one()
    for i in range(LOG_ATE, -1, -1):
        f = f * f * _linefunc(r_pt, r_pt, p_cast)
        r_pt = _double(r_pt)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _linefunc(r_pt, q_twisted, p_cast)
            r_pt = _add(r_pt, q_twisted)
    q1 = (q_twisted[0].pow(P), q_twisted[1].pow(P))
    nq2 = (q1[0].pow(P), -(q1[1].pow(P)))
    f = f * _linefunc(r_pt, q1, p_cast)
    r_pt = _add(r_pt, q1)
    f = f * _linefunc(r_pt, nq2, p_cast)
    return f


def final_exponentiate(f):
    return f.pow((P ** 12 - 1) // R)


def pairing_product_is_one(pairs):
    """pairs: [(G1 affine (x, y) ints or None, G2 affine ((x0,x1),(y0,y1)))]: checks prod e(P_i, Q_i) == 1"""
    f = FQ12.one()
    for g1, g2 in pairs:
        if g1 is None or g2 is None:
            continue
        f = f * miller_loop(twist(g2), _cast_g1(g1))
    return final_exponentiate(f) == FQ12.one()


# ---- G2 arithmetic over FQ2 (affine) for building [s]G2 in tests ---------------------------------------
def _fq2(v):
    return FQ2(list(v))


def g2_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = _fq2(p1[0]), _fq2(p1[1])
    x2, y2 = _fq2(p2[0]), _fq2(p2[1])
    if x1 == x2:
        if not (y1 == y2):
            return None
        m = (x1 * x1).scale(3) * (y1.scale(2)).inv()
    else:
        m = (y2 - y1) * (x2 - x1).inv()
    nx = m * m - x1 - x2
    ny = m * (x1 - nx) - y1
    return (tuple(nx.c), tuple(ny.c))


def g2_mul(pt, k):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, pt)
        pt = g2_add(pt, pt)
        k >>= 1
    return acc
