"""ORACLE (test infrastructure only).

ark-poly 0.3.0 behaviours the reference calls (not vendored): radix-2
`GeneralEvaluationDomain` (fft / ifft, elements, vanishing polynomial) and
`DensePolynomial` (trailing-zero stripping, Horner `evaluate`, `naive_mul`,
long division).  Call sites: plonk/src/proof.rs:50,106,115,125,128,317-373,416;
plonk/src/builder.rs:70,85; permutation/src/lib.rs:106-107,142-147,171,188;
kzg/src/lib.rs:57-61; plonk/src/utils.rs:150-159.

Polynomials are Python lists of canonical Fr integers, lowest degree first, with
trailing zeros stripped (the zero polynomial is []), as `from_coefficients_vec`
does.
"""
from .fields import R_MOD, fr_inv, root_of_unity

M = R_MOD


class Domain:
    """`GeneralEvaluationDomain::<Fr>::new(n)`: radix-2, size = n.next_power_of_two()."""

    def __init__(self, n: int):
        size = 1
        log = 0
        while size < n:
            size <<= 1
            log += 1
        self.size = size
        self.log_size = log
        self.group_gen = root_of_unity(size)
        self.group_gen_inv = fr_inv(self.group_gen)
        self.size_inv = fr_inv(size % M)

    def element(self, i: int) -> int:
        return pow(self.group_gen, i, M)

    def elements(self):
        out = [1] * self.size
        for i in range(1, self.size):
            out[i] = out[i - 1] * self.group_gen % M
        return out

    def evaluate_vanishing_polynomial(self, x: int) -> int:
        return (pow(x, self.size, M) - 1) % M

    def _transform(self, a, w):
        n = self.size
        a = list(a) + [0] * (n - len(a))
        assert len(a) == n
        # bit reversal then DIT
        j = 0
        for i in range(1, n):
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                a[i], a[j] = a[j], a[i]
        length = 2
        while length <= n:
            wl = pow(w, n // length, M)
            half = length >> 1
            tw = [1] * half
            for k in range(1, half):
                tw[k] = tw[k - 1] * wl % M
            for start in range(0, n, length):
                for k in range(half):
                    u = a[start + k]
                    v = a[start + k + half] * tw[k] % M
                    a[start + k] = (u + v) % M
                    a[start + k + half] = (u - v) % M
            length <<= 1
        return a

    def fft(self, coeffs):
        """`evaluate_over_domain` / `fft`: natural order in and out."""
        return self._transform(coeffs, self.group_gen)

    def ifft(self, evals):
        out = self._transform(evals, self.group_gen_inv)
        return [x * self.size_inv % M for x in out]

    def coset_fft(self, coeffs, g: int):
        """Evaluations over g*H (used only by the NTT sweep)."""
        scaled = []
        gp = 1
        for c in coeffs:
            scaled.append(c * gp % M)
            gp = gp * g % M
        return self.fft(scaled)


# ---- DensePolynomial ----------------------------------------------------------

def strip(c):
    c = list(c)
    while c and c[-1] % M == 0:
        c.pop()
    return [x % M for x in c]


def interpolate(evals, domain: Domain):
    """`Evaluations::from_vec_and_domain(evals, domain).interpolate()`."""
    return strip(domain.ifft(evals))


def degree(p):
    return len(p) - 1 if p else 0


def evaluate(p, x):
    acc = 0
    for c in reversed(p):
        acc = (acc * x + c) % M
    return acc


def add(a, b):
    n = max(len(a), len(b))
    return strip([((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % M for i in range(n)])


def sub(a, b):
    n = max(len(a), len(b))
    return strip([((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % M for i in range(n)])


def neg(a):
    return [(-x) % M for x in a]


def scale(a, k):
    """`&DensePolynomial * F` (zero if either is zero)."""
    k %= M
    return strip([x * k % M for x in a])


def naive_mul(a, b):
    """`DensePolynomial::naive_mul` (schoolbook), plonk/src/proof.rs:317-359."""
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] = (out[i + j] + x * y) % M
    return strip(out)


def fast_mul(a, b):
    """Same product via NTT (any correct algorithm gives the same coefficients)."""
    if not a or not b:
        return []
    n = len(a) + len(b) - 1
    d = Domain(n)
    fa = d.fft(a)
    fb = d.fft(b)
    return strip(d.ifft([x * y % M for x, y in zip(fa, fb)]))


def divide_with_q_and_r(num, den):
    """Exact long division (`DenseOrSparsePolynomial::divide_with_q_and_r`)."""
    den = strip(den)
    assert den, "division by zero polynomial"
    num = strip(num)
    if len(num) < len(den):
        return [], num
    q = [0] * (len(num) - len(den) + 1)
    rem = list(num)
    dinv = fr_inv(den[-1])
    for i in range(len(q) - 1, -1, -1):
        coef = rem[i + len(den) - 1] * dinv % M
        q[i] = coef
        if coef:
            for j, dcoef in enumerate(den):
                rem[i + j] = (rem[i + j] - coef * dcoef) % M
    return strip(q), strip(rem[: len(den) - 1])


def divide_by_vanishing_poly(p, n: int):
    """`DensePolynomial::divide_by_vanishing_poly(domain)` -> (quotient, remainder)
    for Z_H = X^n - 1 (plonk/src/proof.rs:373,505)."""
    p = strip(p)
    if len(p) <= n:
        return [], p
    q = [0] * (len(p) - n)
    for k in range(len(q) - 1, -1, -1):
        q[k] = (p[k + n] + (q[k + n] if k + n < len(q) else 0)) % M
    rem = [(p[k] + (q[k] if k < len(q) else 0)) % M for k in range(n)]
    return strip(q), strip(rem)


def divide_by_linear(p, z):
    """(p - p(z)) / (X - z) by synthetic division; caller has already subtracted
    p(z) from coeff 0 (kzg/src/lib.rs:57-61)."""
    q, rem = divide_with_q_and_r(p, [(-z) % M, 1])
    return q, rem
