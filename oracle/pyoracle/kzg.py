"""ORACLE (test infrastructure only).

Restatement of the reference's `kzg` crate: kzg/src/srs.rs:8-52 (Srs) and
kzg/src/lib.rs:33-86 (KzgScheme::{commit, open, verify, identity}) plus the
operator impls kzg/src/lib.rs:110-158 used by the verifier.
"""
from . import poly
from .curve import (G1_GEN, G2_GEN, g1_add, g1_msm, g1_mul, g1_neg, g1_sub, g2_add,
                    g2_mul, g2_neg, pairing_product_is_one)
from .fields import R_MOD


class Srs:
    """kzg/src/srs.rs:8-13."""

    def __init__(self, g1, g2, g2s, secret=None):
        self.g1 = g1
        self.g2 = g2
        self.g2s = g2s
        self.secret = secret  # kept only so tests can use the trapdoor check

    @staticmethod
    def g1_powers(s: int, length: int):
        """srs.rs:15-24: [G, sG, s^2 G, ...] (length points)."""
        out = []
        power = 1
        for _ in range(length):
            out.append(g1_mul(G1_GEN, power))
            power = power * s % R_MOD
        return out

    @classmethod
    def from_secret(cls, s: int, gates: int) -> "Srs":
        """srs.rs:30-34: gates + 3 G1 powers, (G2, s*G2)."""
        s %= R_MOD
        return cls(cls.g1_powers(s, gates + 3), G2_GEN, g2_mul(G2_GEN, s), secret=s)


def commit(srs: Srs, p):
    """kzg/src/lib.rs:37-54: sum_i coeff_i * srs[i]; asserts srs.len() > degree."""
    assert len(srs.g1) > poly.degree(p), "srs too short"
    return g1_msm(srs.g1[: len(p)], p)


def open_at(srs: Srs, p, z: int):
    """kzg/src/lib.rs:55-64 -> (witness point, evaluation)."""
    z %= R_MOD
    y = poly.evaluate(p, z)
    if not p:
        raise IndexError("at least 1")  # `.expect("at least 1")`, lib.rs:58
    shifted = list(p)
    shifted[0] = (shifted[0] - y) % R_MOD
    # NB: the reference mutates the Vec in place without re-stripping; the division
    # result is the same either way.
    q, _ = poly.divide_with_q_and_r(poly.strip(shifted), [(-z) % R_MOD, 1])
    return g1_msm(srs.g1[: len(q)], q) if q else None, y


def verify(srs: Srs, commitment, opening, z: int) -> bool:
    """kzg/src/lib.rs:66-81: e(W, [s]G2 - z G2) == e(C - y G, G2)."""
    w_pt, y = opening
    a = g2_add(srs.g2s, g2_neg(g2_mul(srs.g2, z)))
    b = g1_sub(commitment, g1_mul(G1_GEN, y))
    # e(W, a) == e(b, g2)  <=>  e(W, a) * e(-b, g2) == 1
    return pairing_product_is_one([(w_pt, a), (g1_neg(b), srs.g2)])


def verify_trapdoor(srs: Srs, commitment, opening, z: int) -> bool:
    """Same relation checked in G1 with the known secret s (fast path for large
    tests): C - y G == (s - z) W.  Equivalent to `verify` when s is known."""
    assert srs.secret is not None
    w_pt, y = opening
    lhs = g1_sub(commitment, g1_mul(G1_GEN, y))
    rhs = g1_mul(w_pt, (srs.secret - z) % R_MOD)
    return lhs == rhs


def identity(srs: Srs):
    """kzg/src/lib.rs:82-85: commit(1)."""
    return commit(srs, [1])


# KzgCommitment operators (kzg/src/lib.rs:110-158)
commitment_add = g1_add
commitment_sub = g1_sub
commitment_neg = g1_neg


def commitment_mul(c, k: int):
    return g1_mul(c, k)
