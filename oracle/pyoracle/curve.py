"""ORACLE (test infrastructure only).

BLS12-381 G1 (y^2 = x^3 + 4 over Fq), G2 and the optimal-ate pairing on Python
integers.  Restates what the reference gets from ark-ec / ark-bls12-381 0.3.0
(un-vendored): `AffineCurve::mul`, `into_projective`, `Sum`, `Neg`
(kzg/src/lib.rs:46-53,110-158; kzg/src/srs.rs:15-29) and `Bls12_381::pairing`
(kzg/src/lib.rs:78-79).

Affine points are `(x, y)` tuples of canonical integers, `None` = point at
infinity (arkworks: `infinity: bool`).  Group results are unique, so any correct
algorithm is bit-exact with arkworks' double-and-add.
"""
from .fields import Q_MOD, R_MOD

P = Q_MOD
B1 = 4

G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)


def g1_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B1) % P == 0


# --- Jacobian arithmetic (X/Z^2, Y/Z^3); identity has Z == 0 -------------------

def _jac_double(pt):
    X, Y, Z = pt
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    A = X * X % P
    Bv = Y * Y % P
    C = Bv * Bv % P
    D = 2 * ((X + Bv) * (X + Bv) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def _jac_add(p1, p2):
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    if Z1 == 0:
        return p2
    if Z2 == 0:
        return p1
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        if S1 == S2:
            return _jac_double(p1)
        return (1, 1, 0)
    H = (U2 - U1) % P
    I = 4 * H * H % P
    J = H * I % P
    rr = 2 * (S2 - S1) % P
    V = U1 * I % P
    X3 = (rr * rr - J - 2 * V) % P
    Y3 = (rr * (V - X3) - 2 * S1 * J) % P
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % P
    return (X3, Y3, Z3)


def _to_jac(pt):
    if pt is None:
        return (1, 1, 0)
    return (pt[0], pt[1], 1)


def _to_affine(j):
    X, Y, Z = j
    if Z == 0:
        return None
    zi = pow(Z, -1, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_add(a, b):
    return _to_affine(_jac_add(_to_jac(a), _to_jac(b)))


def g1_sub(a, b):
    return g1_add(a, g1_neg(b))


def g1_mul(pt, k: int):
    """`AffineCurve::mul(Fr)`: scalar is the canonical integer of the Fr element."""
    k %= R_MOD
    acc = (1, 1, 0)
    base = _to_jac(pt)
    for bit in bin(k)[2:] if k else "":
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add(acc, base)
    return _to_affine(acc)


def g1_sum(points):
    acc = (1, 1, 0)
    for p_ in points:
        acc = _jac_add(acc, _to_jac(p_))
    return _to_affine(acc)


def g1_msm(points, scalars):
    """sum_i scalars[i] * points[i], the quantity kzg/src/lib.rs:46-53 computes by
    n independent double-and-add multiplications.  Here: a windowed bucket method
    (result identical because the affine answer is unique)."""
    n = min(len(points), len(scalars))
    if n == 0:
        return None
    c = 4 if n < 64 else 8
    nwin = (255 + c - 1) // c
    total = (1, 1, 0)
    jp = [_to_jac(p_) for p_ in points[:n]]
    sc = [s % R_MOD for s in scalars[:n]]
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = _jac_double(total)
        buckets = [None] * (1 << c)
        for p_, s in zip(jp, sc):
            d = (s >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d] = p_ if buckets[d] is None else _jac_add(buckets[d], p_)
        run = (1, 1, 0)
        acc = (1, 1, 0)
        for d in range((1 << c) - 1, 0, -1):
            if buckets[d] is not None:
                run = _jac_add(run, buckets[d])
            acc = _jac_add(acc, run)
        total = _jac_add(total, acc)
    return _to_affine(total)


# --- ark-serialize 0.3 uncompressed encoding (plonk/src/proof/challenges.rs:17-22)

def g1_serialize_unchecked(pt) -> bytes:
    """`G1Affine::serialize_unchecked` = uncompressed: x (48 B LE canonical) then y
    with SWFlags in the top bits of the last byte.  ark-serialize 0.3:
    SWFlags::default() = NegativeY -> mask 0; Infinity -> 1<<6 with (x, y) = (0, 1)."""
    if pt is None:
        out = bytearray((0).to_bytes(48, "little") + (1).to_bytes(48, "little"))
        out[-1] |= 1 << 6
        return bytes(out)
    return pt[0].to_bytes(48, "little") + pt[1].to_bytes(48, "little")


def g1_deserialize_unchecked(b: bytes):
    assert len(b) == 96
    if b[-1] & (1 << 6):
        return None
    y = bytearray(b[48:])
    y[-1] &= 0x3F
    return (int.from_bytes(b[:48], "little"), int.from_bytes(bytes(y), "little"))


# =============================================================================
# G2 + pairing (only `verify` needs these; kzg/src/lib.rs:66-81).
# Fq12 is represented as Fq[w]/(w^12 - 2 w^6 + 2)  (w^6 = u + 1, u^2 = -1).
# =============================================================================

class Fq2:
    __slots__ = ("a", "b")

    def __init__(self, a, b=0):
        self.a = a % P
        self.b = b % P

    def __add__(self, o):
        return Fq2(self.a + o.a, self.b + o.b)

    def __sub__(self, o):
        return Fq2(self.a - o.a, self.b - o.b)

    def __neg__(self):
        return Fq2(-self.a, -self.b)

    def __mul__(self, o):
        if isinstance(o, int):
            return Fq2(self.a * o, self.b * o)
        return Fq2(self.a * o.a - self.b * o.b, self.a * o.b + self.b * o.a)

    def __eq__(self, o):
        return self.a == o.a and self.b == o.b

    def inv(self):
        d = pow(self.a * self.a + self.b * self.b, -1, P)
        return Fq2(self.a * d, -self.b * d)

    def is_zero(self):
        return self.a == 0 and self.b == 0


G2_GEN = (
    Fq2(0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
        0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
    Fq2(0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
        0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
)
B2 = Fq2(4, 4)


def g2_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return y * y - x * x * x == B2


def g2_double(pt):
    if pt is None:
        return None
    x, y = pt
    if y.is_zero():
        return None
    m = (x * x * 3) * (y * 2).inv()
    nx = m * m - x * 2
    ny = m * (x - nx) - y
    return (nx, ny)


def g2_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if y1 == y2:
            return g2_double(p1)
        return None
    m = (y2 - y1) * (x2 - x1).inv()
    nx = m * m - x1 - x2
    ny = m * (x1 - nx) - y1
    return (nx, ny)


def g2_neg(pt):
    if pt is None:
        return None
    return (pt[0], -pt[1])


def g2_mul(pt, k):
    k %= R_MOD
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_double(acc)
        if bit == "1":
            acc = g2_add(acc, pt)
    return acc


# ---- Fq12 as polynomials of degree < 12 in w --------------------------------

_FQ12_MOD = (2, 0, 0, 0, 0, 0, -2, 0, 0, 0, 0, 0)  # w^12 = 2 w^6 - 2


class Fq12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [x % P for x in c]

    @staticmethod
    def one():
        return Fq12([1] + [0] * 11)

    def __mul__(self, o):
        a, b = self.c, o.c
        t = [0] * 23
        for i in range(12):
            ai = a[i]
            if ai:
                for j in range(12):
                    t[i + j] += ai * b[j]
        for k in range(22, 11, -1):
            top = t[k]
            if top:
                # w^k = w^(k-12) * (2 w^6 - 2)
                t[k - 6] += 2 * top
                t[k - 12] -= 2 * top
        return Fq12(t[:12])

    def __eq__(self, o):
        return self.c == o.c

    def __sub__(self, o):
        return Fq12([x - y for x, y in zip(self.c, o.c)])

    def __add__(self, o):
        return Fq12([x + y for x, y in zip(self.c, o.c)])

    def scale(self, k):
        return Fq12([x * k for x in self.c])

    def pow(self, e):
        result = Fq12.one()
        base = self
        while e:
            if e & 1:
                result = result * base
            base = base * base
            e >>= 1
        return result

    def inv(self):
        # extended Euclid over Fq[w] (as py_ecc does)
        lm, hm = [1] + [0] * 12, [0] * 13
        low = self.c + [0]
        high = [(-x) % P for x in _FQ12_MOD] + [1]
        high = [2, 0, 0, 0, 0, 0, (-2) % P, 0, 0, 0, 0, 0, 1]

        def deg(p_):
            d = len(p_) - 1
            while d and p_[d] == 0:
                d -= 1
            return d

        def poly_rounded_div(a, b):
            dega, degb = deg(a), deg(b)
            temp = list(a)
            o = [0] * len(a)
            binv = pow(b[degb], -1, P)
            for i in range(dega - degb, -1, -1):
                o[i] = (o[i] + temp[degb + i] * binv) % P
                for c_ in range(degb + 1):
                    temp[c_ + i] = (temp[c_ + i] - o[i] * b[c_]) % P
            return o[: deg(o) + 1]

        while deg(low):
            r_ = poly_rounded_div(high, low)
            r_ += [0] * (13 - len(r_))
            nm = list(hm)
            new = list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] = (nm[i + j] - lm[i] * r_[j]) % P
                    new[i + j] = (new[i + j] - low[i] * r_[j]) % P
            lm, low, hm, high = nm, new, lm, low
        li = pow(low[0], -1, P)
        return Fq12([x * li for x in lm[:12]])


def _fq2_to_fq12_coeffs(x: Fq2):
    # u = w^6 - 1  ->  a + b u = (a - b) + b w^6
    c = [0] * 12
    c[0] = (x.a - x.b) % P
    c[6] = x.b
    return c


def _twist(pt):
    """Map a G2 point on the twist to the curve over Fq12: (x / w^2, y / w^3)."""
    x, y = pt
    nx = Fq12(_fq2_to_fq12_coeffs(x))
    ny = Fq12(_fq2_to_fq12_coeffs(y))
    w = Fq12([0, 1] + [0] * 10)
    w2i = (w * w).inv()
    w3i = (w * w * w).inv()
    return (nx * w2i, ny * w3i)


def _cast_g1(pt):
    return (Fq12([pt[0]] + [0] * 11), Fq12([pt[1]] + [0] * 11))


def _f12_double(pt):
    x, y = pt
    m = (x * x).scale(3) * (y.scale(2)).inv()
    nx = m * m - x.scale(2)
    ny = m * (x - nx) - y
    return (nx, ny)


def _f12_add(p1, p2):
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2 and y1 == y2:
        return _f12_double(p1)
    m = (y2 - y1) * (x2 - x1).inv()
    nx = m * m - x1 - x2
    ny = m * (x1 - nx) - y1
    return (nx, ny)


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


ATE_LOOP_COUNT = 15132376222941642752  # |x| for BLS12-381
_FINAL_EXP = (P ** 12 - 1) // R_MOD


def miller_loop(q_g2, p_g1) -> Fq12:
    """Unreduced pairing value f_{|x|,Q}(P); inputs affine, neither infinity."""
    if q_g2 is None or p_g1 is None:
        return Fq12.one()
    Q = _twist(q_g2)
    Pt = _cast_g1(p_g1)
    R_ = Q
    f = Fq12.one()
    for bit in bin(ATE_LOOP_COUNT)[3:]:
        f = f * f * _linefunc(R_, R_, Pt)
        R_ = _f12_double(R_)
        if bit == "1":
            f = f * _linefunc(R_, Q, Pt)
            R_ = _f12_add(R_, Q)
    return f


def final_exponentiation(f: Fq12) -> Fq12:
    return f.pow(_FINAL_EXP)


def pairing(p_g1, q_g2) -> Fq12:
    """e(P, Q) up to the fixed power arkworks' representation differs by; only
    equalities between pairings are used (kzg/src/lib.rs:78-80), which are
    representation-independent."""
    return final_exponentiation(miller_loop(q_g2, p_g1))


def pairing_product_is_one(pairs) -> bool:
    """prod e(P_i, Q_i) == 1 with a single final exponentiation."""
    f = Fq12.one()
    for p_g1, q_g2 in pairs:
        f = f * miller_loop(q_g2, p_g1)
    return final_exponentiation(f) == Fq12.one()
