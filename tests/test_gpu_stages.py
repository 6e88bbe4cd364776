"""GPU, stage by stage: the intermediate polynomials of a proof (tp_circuit_read_poly) against the reference's own
functions as the oracle restates them -- `CompiledPermutation::prove` (permutation/src/proving.rs:7-31: one inversion
per cell), `quotient_polynomial` with the reference's schoolbook `naive_mul` and long division by X^n - 1, remainder
discarded (plonk/src/proof.rs:292-375), `linearisation_poly` (:376-439) -- instead of only through the final proof
bytes.  The challenges are re-derived from the GPU proof, so each stage is checked on its own inputs."""
import pytest

from oracle.pyoracle import builder as obuilder, fields, plonk as oplonk, poly, rng
from typlonk_b200 import field as F, ffi
from typlonk_b200.plonk import CircuitDescription

pytestmark = pytest.mark.gpu
M = fields.R_MOD
TAU = rng.fr_rand_stream(1, 1)[0]
BLINDERS = rng.fr_rand_stream(2, 9)


class Pythagoras(CircuitDescription):
    INPUTS = 3

    @staticmethod
    def run(inputs):
        a, b, c = inputs
        a = a.clone() * a
        b = b.clone() * b
        c = c.clone() * c
        d = a + b
        d.assert_eq(c)


def _mul_chain(gates):
    class MulChain(CircuitDescription):
        INPUTS = 2

        @staticmethod
        def run(inputs):
            x, y = inputs
            for _ in range(gates):
                x = x * y.clone()
    return MulChain


def _pad(p, n):
    return list(p) + [0] * (n - len(p))


@pytest.mark.parametrize("name,desc,orun,nin,inputs", [
    ("pythagoras", Pythagoras, obuilder.circuit_pythagoras, 3, [3, 4, 5]),
    ("pythagoras_broken_copy", Pythagoras, obuilder.circuit_pythagoras, 3, [3, 4, 6]),   # floor quotient, all four cosets
    ("mulchain_29", _mul_chain(29), obuilder.make_mul_chain(29), 2, [3, 5]),
    ("mulchain_125", _mul_chain(125), obuilder.make_mul_chain(125), 2, [3, 5]),
])
def test_each_stage_against_the_reference_functions(ctx, name, desc, orun, nin, inputs):
    circuit = desc.build(ctx, TAU)
    oc = obuilder.compile_circuit(orun, nin, TAU)
    n = circuit.rows
    proof = circuit.prove(inputs, [0], BLINDERS)
    h = circuit.handle
    alpha, beta, gamma, zeta = [F.fr_from_bytes(x) for x in ffi.proof_challenges(proof.fixed)]
    # witness polynomials
    cols = obuilder.witness_columns(orun, inputs, n, BLINDERS)
    advice = [poly.interpolate(col, oc.domain) for col in cols]
    for k, which in enumerate((h.POLY_A, h.POLY_B, h.POLY_C)):
        assert F.fr_vec_from_bytes(h.read_poly(which)) == _pad(advice[k], n), "witness polynomial %d" % k
    # grand product: the reference's loop with one inversion per cell
    evals = oc.copy_constrains.prove(cols, beta, gamma)
    assert F.fr_vec_from_bytes(h.read_poly(h.POLY_Z_EVALS)) == evals, "grand product"
    evals.pop()
    acc = poly.interpolate(evals, oc.domain)
    acc_shifted = poly.interpolate(evals[1:] + evals[:1], oc.domain)
    assert F.fr_vec_from_bytes(h.read_poly(h.POLY_Z)) == _pad(acc, n)
    # quotient: schoolbook products and long division as the reference does (literal=True), remainder discarded
    pi_poly = []
    t_slices = oplonk.quotient_polynomial(oc, advice, (acc, acc_shifted), (alpha, beta, gamma), pi_poly, True)
    got_t = F.fr_vec_from_bytes(h.read_poly(h.POLY_QUOTIENT))
    for i in range(3):
        assert got_t[i * n:(i + 1) * n] == _pad(t_slices[i], n), "quotient slice %d" % i
    # linearisation polynomial from the proof's evaluations
    raw = proof.fixed

    def fr_at(off):
        return int.from_bytes(raw[off:off + 32], "little")
    a_bar, b_bar, c_bar = fr_at(192), fr_at(192 + 224), fr_at(192 + 448)
    z_bar, zw_bar = fr_at(672 + 192), fr_at(672 + 224 + 96)
    assert fr_at(672 + 224 + 128) == zeta
    lin = oplonk.linearisation_poly(oc, (a_bar, b_bar, c_bar), (z_bar, zw_bar), acc, (alpha, beta, gamma), zeta, t_slices, 0)
    assert F.fr_vec_from_bytes(h.read_poly(h.POLY_LINEARISATION)) == _pad(poly.strip(lin), n), "linearisation polynomial"
    assert poly.evaluate(lin, zeta) == int.from_bytes(raw[1440:1472], "little")     # r(zeta), the proof's last scalar
    h.destroy()
    circuit.srs.handle.destroy()
