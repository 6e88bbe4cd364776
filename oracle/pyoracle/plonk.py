"""ORACLE (test infrastructure only).

Restatement of the reference's prover and verifier, plonk/src/proof.rs (all) and
plonk/src/utils.rs:13-126,150-159, with the blinding scalars passed in (the
reference draws them from thread_rng, proof.rs:42-48).

`literal=True` follows proof.rs:317-373 step by step with schoolbook products
(`naive_mul`); `literal=False` uses NTT products.  Both give the same
coefficients (polynomial arithmetic is exact), which tests assert at small n.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from . import kzg, poly
from .builder import CompiledCircuit, witness_columns
from .curve import g1_add, g1_mul, g1_neg, g1_serialize_unchecked, g1_sub
from .fields import R_MOD, fr_inv
from .rng import generate_challenges

M = R_MOD


class GateUnsatisfied(AssertionError):
    """The reference's `vanishes` assert (proof.rs:321,361,504-508) fired."""


@dataclass
class Proof:
    """proof.rs:65-95.  Points are affine (x, y) or None; scalars canonical ints."""
    a: Tuple  # (commitment, (W, y))
    b: Tuple
    c: Tuple
    z_commitment: Optional[Tuple]
    z: Tuple  # (W, y)
    zw: Tuple
    evaluation_point: int
    t: List
    r: Tuple
    public_inputs: List[int] = field(default_factory=list)

    def to_bytes(self) -> bytes:
        """Canonical proof bytes (the reference has no serialiser):
        field order of `Proof`, ark-serialize 0.3 uncompressed conventions."""
        out = b""
        for com, (w, y) in (self.a, self.b, self.c):
            out += g1_serialize_unchecked(com) + g1_serialize_unchecked(w) + y.to_bytes(32, "little")
        out += g1_serialize_unchecked(self.z_commitment)
        out += g1_serialize_unchecked(self.z[0]) + self.z[1].to_bytes(32, "little")
        out += g1_serialize_unchecked(self.zw[0]) + self.zw[1].to_bytes(32, "little")
        out += self.evaluation_point.to_bytes(32, "little")
        for tc in self.t:
            out += g1_serialize_unchecked(tc)
        out += g1_serialize_unchecked(self.r[0]) + self.r[1].to_bytes(32, "little")
        out += len(self.public_inputs).to_bytes(8, "little")
        for v in self.public_inputs:
            out += v.to_bytes(32, "little")
        return out


def l0_poly(n: int):
    """utils.rs:150-159: (X^n - 1) / (n (X - 1))."""
    den = poly.scale([(-1) % M, 1], n % M)
    vanish = [(-1) % M] + [0] * (n - 1) + [1]
    return poly.divide_with_q_and_r(vanish, den)[0]


def add_to_poly(p, number):
    """utils.rs:13-20."""
    if not p:
        return poly.strip([number])
    out = list(p)
    out[0] = (out[0] + number) % M
    return out  # NB: the reference does not re-strip here either


def _vanishes(p, n):
    _, rest = poly.divide_by_vanishing_poly(p, n)
    if rest:
        raise GateUnsatisfied("constraint polynomial does not vanish on the domain")


def transcript(points) -> bytes:
    return b"".join(g1_serialize_unchecked(p) for p in points)


def quotient_polynomial(circuit: CompiledCircuit, advice, acc, challenges, public_inputs_poly, literal):
    """proof.rs:292-375 -> 3 slices of at most n coefficients."""
    mul = poly.naive_mul if literal else poly.fast_mul
    n = circuit.domain.size
    w = circuit.domain.element(1)
    a, b, c = advice
    alpha, beta, gamma = challenges
    perm = circuit.copy_constrains
    ks = perm.cosets

    line1 = poly.add(poly.add(poly.add(
        poly.sub(poly.add(mul(circuit.q_l, a), mul(circuit.q_r, b)), mul(circuit.q_o, c)),
        mul(mul(circuit.q_m, a), b)), circuit.q_c), public_inputs_poly)
    _vanishes(line1, n)

    def reduce_mul(polys):
        out = polys[0]
        for p_ in polys[1:]:
            out = mul(out, p_)
        return out

    line2 = reduce_mul([poly.add(adv, poly.strip([gamma, k * beta % M])) for adv, k in zip((a, b, c), ks)])
    line2_eval = poly.evaluate(line2, w)
    line2 = mul(line2, acc[0])
    sigmas = perm.sigma_polys(circuit.domain)
    line3 = reduce_mul([poly.add(poly.add(adv, poly.scale(s, beta)), poly.strip([gamma]))
                        for adv, s in zip((a, b, c), sigmas)])
    line3_eval = poly.evaluate(line3, w)
    # proof.rs:350-353
    assert (line2_eval * poly.evaluate(acc[0], w) - line3_eval * poly.evaluate(acc[0], w * w % M)) % M == 0
    line3 = mul(line3, acc[1])
    zm1 = list(acc[0])
    zm1[0] = (zm1[0] - 1) % M
    line4 = mul(poly.strip(zm1), l0_poly(n))
    _vanishes(line4, n)

    target = poly.add(poly.add(poly.add(line1, poly.scale(line2, alpha)),
                               poly.scale(poly.neg(line3), alpha)),
                      poly.scale(line4, alpha * alpha % M))
    q, _rem = poly.divide_by_vanishing_poly(target, n)  # remainder discarded (proof.rs:373)
    # SlicedPoly::from_poly(target, n)   utils.rs:31-43
    assert poly.degree(q) // 3 <= n
    slices = [[], [], []]
    for idx in range(0, len(q), n):
        slices[idx // n] = poly.strip(q[idx: idx + n])
    return slices


def compact(slices, n, point):
    """SlicedPoly::compact (utils.rs:96-109): sum_i slice_i * point^(n i)."""
    out = []
    for i, s in enumerate(slices):
        out = poly.add(out, poly.scale(s, pow(point, n * i, M)))
    return out


def linearisation_poly(circuit, advice_evals, acc_evals, acc, challenges, eval_point, t_slices, public_eval):
    """proof.rs:376-439."""
    n = circuit.domain.size
    a, b, c = advice_evals
    alpha, beta, gamma = challenges
    perm = circuit.copy_constrains
    line1 = poly.add(poly.add(poly.scale(circuit.q_l, a),
                              poly.sub(poly.scale(circuit.q_r, b), poly.scale(circuit.q_o, c))),
                     poly.add(poly.scale(circuit.q_m, a * b % M), circuit.q_c))
    line1 = add_to_poly(line1, public_eval)
    l2 = 1
    for k, ev in zip(perm.cosets, advice_evals):
        l2 = l2 * ((ev + k * beta % M * eval_point + gamma) % M) % M
    line2 = poly.scale(acc, l2)
    sigma_polys = perm.sigma_polys(circuit.domain)
    sigma_evals = [poly.evaluate(p_, eval_point) for p_ in sigma_polys]
    perm_ab = (a + beta * sigma_evals[0] + gamma) % M * ((b + beta * sigma_evals[1] + gamma) % M) % M
    perm_c = add_to_poly(poly.scale(sigma_polys[2], beta), (gamma + c) % M)
    line3 = poly.scale(poly.scale(perm_c, perm_ab), acc_evals[1])
    copy_constrain = poly.sub(line2, line3)
    l0_eval = poly.evaluate(l0_poly(n), eval_point)
    line4 = poly.scale(add_to_poly(acc, (-1) % M), l0_eval)
    line5 = poly.scale(compact(t_slices, n, eval_point), circuit.domain.evaluate_vanishing_polynomial(eval_point))
    return poly.sub(poly.add(poly.add(line1, poly.scale(copy_constrain, alpha)),
                             poly.scale(line4, alpha * alpha % M)), line5)


def prove_columns(circuit: CompiledCircuit, columns, public_inputs, literal=False) -> Proof:
    """proof.rs:50-57 + 96-194 starting from the three evaluation columns (which
    already include the blinders)."""
    domain = circuit.domain
    n = domain.size
    srs = circuit.srs
    w = domain.element(1)
    advice = [poly.interpolate(col, domain) for col in columns]
    public_inputs = [v % M for v in public_inputs] + [0] * (n - len(public_inputs))
    public_inputs = public_inputs[:n]
    pi_poly = poly.interpolate(public_inputs, domain)

    commitments = [kzg.commit(srs, p_) for p_ in advice]  # round1, proof.rs:283-290
    beta, gamma = generate_challenges(transcript(commitments), 2)
    values = [domain.fft(p_) for p_ in advice]
    evals = circuit.copy_constrains.prove(values, beta, gamma)
    evals.pop()
    acc_shifted = poly.interpolate(evals[1:] + evals[:1], domain)
    acc = poly.interpolate(evals, domain)
    acc_commitment = kzg.commit(srs, acc)
    alpha, evaluation_point = generate_challenges(transcript(commitments + [acc_commitment]), 2)

    public_eval = poly.evaluate(pi_poly, evaluation_point)
    t_slices = quotient_polynomial(circuit, advice, (acc, acc_shifted), (alpha, beta, gamma), pi_poly, literal)
    openings = [kzg.open_at(srs, p_, evaluation_point) for p_ in advice]
    advice_evals = [o[1] for o in openings]
    z_open = kzg.open_at(srs, acc, evaluation_point)
    zw_open = kzg.open_at(srs, acc, evaluation_point * w % M)
    lin = linearisation_poly(circuit, advice_evals, (z_open[1], zw_open[1]), acc, (alpha, beta, gamma),
                             evaluation_point, t_slices, public_eval)
    r_open = kzg.open_at(srs, poly.strip(lin), evaluation_point)
    t_commit = [kzg.commit(srs, s) for s in t_slices]
    return Proof(
        a=(commitments[0], openings[0]), b=(commitments[1], openings[1]), c=(commitments[2], openings[2]),
        z_commitment=acc_commitment, z=z_open, zw=zw_open, evaluation_point=evaluation_point,
        t=t_commit, r=r_open, public_inputs=public_inputs)


def prove(circuit: CompiledCircuit, inputs, public_inputs, blinders, literal=False) -> Proof:
    """`CompiledCircuit::prove` (proof.rs:26-57)."""
    cols = witness_columns(circuit.run, inputs, circuit.rows, blinders)
    return prove_columns(circuit, cols, public_inputs, literal)


def _verify_challenges(proof: Proof):
    coms = [proof.a[0], proof.b[0], proof.c[0]]
    beta, gamma = generate_challenges(transcript(coms), 2)
    alpha, point = generate_challenges(transcript(coms + [proof.z_commitment]), 2)
    return alpha, beta, gamma, point


def linearisation_commitment(circuit, advice_evals, acc, acc_evals, eval_point, quotient, challenges, public_eval):
    """proof.rs:441-503."""
    srs = circuit.srs
    n = circuit.domain.size
    perm = circuit.copy_constrains
    sigma_evals = perm.sigma_evals(eval_point, circuit.domain)
    sigma_commitments = perm.sigma_commitments(srs, circuit.domain)
    alpha, beta, gamma = challenges
    a, b, c = advice_evals
    q_l, q_r, q_o, q_m, q_c = circuit.fixed_commitments
    line1 = g1_add(g1_add(g1_sub(g1_add(g1_mul(q_l, a), g1_mul(q_r, b)), g1_mul(q_o, c)),
                          g1_mul(g1_mul(q_m, a), b)), q_c)
    l2 = 1
    for k, ev in zip(perm.cosets, advice_evals):
        l2 = l2 * ((ev + beta * k % M * eval_point + gamma) % M) % M
    l0_eval = poly.evaluate(l0_poly(n), eval_point)
    line2 = g1_mul(acc, (l2 * alpha + l0_eval * alpha * alpha) % M)
    l3 = (a + beta * sigma_evals[0] + gamma) % M * ((b + beta * sigma_evals[1] + gamma) % M) % M
    line3 = g1_mul(g1_mul(g1_mul(g1_mul(sigma_commitments[2], l3), alpha), beta), acc_evals[1])
    q_com = None
    for i, tc in enumerate(quotient):  # SlicedPoly::compact_commitment, utils.rs:110-125
        q_com = g1_add(q_com, g1_mul(tc, pow(eval_point, n * i, M)))
    line5 = g1_mul(q_com, circuit.domain.evaluate_vanishing_polynomial(eval_point))
    constant_perm = l3 * ((c + gamma) % M) % M * acc_evals[1] % M
    constant = (alpha * constant_perm + l0_eval * alpha * alpha + public_eval) % M
    ident = g1_mul(kzg.identity(srs), constant)
    return g1_sub(g1_add(line1, g1_sub(line2, g1_add(line3, ident))), line5)


def verify(circuit: CompiledCircuit, proof: Proof, use_trapdoor=False) -> bool:
    """`CompiledCircuit::verify` (proof.rs:59-62, 195-281).  `use_trapdoor` swaps the
    pairing check for the equivalent G1 check with the known SRS secret."""
    check = kzg.verify_trapdoor if use_trapdoor else kzg.verify
    srs = circuit.srs
    domain = circuit.domain
    alpha, beta, gamma, point = _verify_challenges(proof)
    pis = list(proof.public_inputs) + [0] * (circuit.rows - len(proof.public_inputs))
    public_eval = poly.evaluate(poly.interpolate(pis[: circuit.rows], domain), point)
    if proof.evaluation_point != point:
        return False
    w = domain.element(1)
    for com, opening in (proof.a, proof.b, proof.c):
        if not check(srs, com, opening, proof.evaluation_point):
            return False
    if not check(srs, proof.z_commitment, proof.z, proof.evaluation_point):
        return False
    if not check(srs, proof.z_commitment, proof.zw, proof.evaluation_point * w % M):
        return False
    advice_evals = [proof.a[1][1], proof.b[1][1], proof.c[1][1]]
    r_com = linearisation_commitment(circuit, advice_evals, proof.z_commitment, (proof.z[1], proof.zw[1]),
                                     proof.evaluation_point, proof.t, (alpha, beta, gamma), public_eval)
    return check(srs, r_com, proof.r, proof.evaluation_point) and proof.r[1] == 0
