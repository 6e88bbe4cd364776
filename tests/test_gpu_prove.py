"""GPU parity tests of the full prover (tp_circuit_compile + tp_prove through the Python mirror of
the reference API) against the CPU oracle: every proof byte equal under fixed tau / blinders, and
verify() (oracle, pairings) accepts."""
import os

import pytest

from oracle.pyoracle import builder as obuilder, plonk as oplonk, rng
from typlonk_b200.ffi import GateUnsatisfied
import py_tracer
from typlonk_b200.plonk import CircuitDescription

pytestmark = pytest.mark.gpu

TAU = rng.fr_rand_stream(1, 1)[0]
BLINDERS = rng.fr_rand_stream(2, 9)


class Pythagoras(CircuitDescription):  # README.md:16-27, builder/test.rs:12-23
    INPUTS = 3

    @staticmethod
    def run(inputs):
        a, b, c = inputs
        a = a.clone() * a
        b = b.clone() * b
        c = c.clone() * c
        d = a + b
        d.assert_eq(c)


class Additive(CircuitDescription):  # builder/test.rs:3-11
    INPUTS = 5

    @staticmethod
    def run(inputs):
        a, b, c, d, e = inputs
        x = (c + d) + e
        a = a + b
        a.assert_eq(x)


def mul_chain(gates):
    class MulChain(CircuitDescription):
        INPUTS = 2

        @staticmethod
        def run(inputs):
            x, y = inputs
            for _ in range(gates):
                x = x * y.clone()
    return MulChain


def _oracle_circuit(run, n_inputs):
    return obuilder.compile_circuit(run, n_inputs, TAU)


def test_readme_circuit_bytes_and_verify(ctx):
    """builder/test.rs:25-30 `circuit2_test` / README doctest."""
    circuit = Pythagoras.build(ctx, TAU)
    oc = _oracle_circuit(obuilder.circuit_pythagoras, 3)
    assert circuit.rows == oc.rows == 8
    assert circuit.fixed_commitments == oc.fixed_commitments
    proof = circuit.prove([3, 4, 5], [0], BLINDERS)
    oproof = oplonk.prove(oc, [3, 4, 5], [0], BLINDERS)
    assert proof.to_bytes() == oproof.to_bytes()
    assert oplonk.verify(oc, oproof)  # pairings


def test_readme_circuit_bad_inputs(ctx):
    """builder/test.rs:31-37 `circuit2_test_bad_inputs`: a proof is produced, verify is false."""
    circuit = Pythagoras.build(ctx, TAU)
    oc = _oracle_circuit(obuilder.circuit_pythagoras, 3)
    proof = circuit.prove([3, 4, 6], [0], BLINDERS)
    oproof = oplonk.prove(oc, [3, 4, 6], [0], BLINDERS)
    assert proof.to_bytes() == oproof.to_bytes()
    assert not oplonk.verify(oc, oproof, use_trapdoor=True)


def test_additive_circuit(ctx):
    """builder/test.rs:39-44 `circuit1_test`."""
    circuit = Additive.build(ctx, TAU)
    oc = _oracle_circuit(obuilder.circuit_additive, 5)
    proof = circuit.prove([2, 7, 2, 3, 4], [0], BLINDERS)
    oproof = oplonk.prove(oc, [2, 7, 2, 3, 4], [0], BLINDERS)
    assert proof.to_bytes() == oproof.to_bytes()
    assert oplonk.verify(oc, oproof, use_trapdoor=True)


def test_nonzero_public_input_panics_like_reference(ctx):
    """a non-zero public input makes `vanishes(line1)` fire."""
    circuit = Pythagoras.build(ctx, TAU)
    with pytest.raises(GateUnsatisfied):
        circuit.prove([3, 4, 5], [1], BLINDERS)
    oc = _oracle_circuit(obuilder.circuit_pythagoras, 3)
    with pytest.raises(oplonk.GateUnsatisfied):
        oplonk.prove(oc, [3, 4, 5], [1], BLINDERS)


@pytest.mark.parametrize("gates", [5, 13, 61, 253])
def test_mul_chain_bytes(ctx, gates):
    circuit = mul_chain(gates).build(ctx, TAU)
    oc = _oracle_circuit(obuilder.make_mul_chain(gates), 2)
    assert circuit.rows == oc.rows == gates + 3
    proof = circuit.prove([3, 5], [0], BLINDERS)
    oproof = oplonk.prove(oc, [3, 5], [0], BLINDERS)
    assert proof.to_bytes() == oproof.to_bytes()
    assert oplonk.verify(oc, oproof, use_trapdoor=True)


# ---- BASELINE.json configs[1] / configs[2]: 2^16 byte parity with the C++ oracle, 2^20 verify ------

def test_mul_chain_2_16_bytes_vs_c_oracle(ctx):
    from oracle import coracle
    from typlonk_b200 import field as F, synthetic
    log_n = 16
    n = 1 << log_n
    circuit = synthetic.mul_chain_direct(ctx, log_n)
    cols = synthetic.mul_chain_witness(n - 3, n)
    proof = circuit.handle.prove([F.fr_vec_to_bytes(c) for c in cols], bytes(32 * n))
    tau_b, sel, perm, ocols, pi = coracle.mul_chain_inputs(log_n)
    assert perm == circuit.perm.perm
    oc = coracle.Circuit(tau_b, sel, perm, n)
    assert [F.g1_from_abi(c) for c in oc.fixed_commitments()] == circuit.fixed_commitments
    assert oc.prove(ocols, pi) == proof
    oc.close()


def test_mul_chain_2_20_bytes_vs_c_oracle_live_and_golden(ctx):
    """BASELINE.json configs[2]: the 2^20-gate proof, ALL bytes, against (a) the C++ oracle run here on the same inputs
    (about a minute of host time) and (b) the committed golden proof (tools/make_golden_big.py); both through tp_prove
    (host buffers) and on a 2-rank device group (sharded MSM / quotient / upload)."""
    import json
    from oracle import coracle
    from typlonk_b200 import field as F, synthetic
    from typlonk_b200.ffi import Context
    log_n = 20
    n = 1 << log_n
    circuit = synthetic.mul_chain_direct(ctx, log_n)
    cols = [F.fr_vec_to_bytes(c) for c in synthetic.mul_chain_witness(n - 3, n)]
    proof = circuit.handle.prove(cols, bytes(32 * n))
    circuit.handle.destroy()
    circuit.srs.handle.destroy()
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mulchain_big.json")))["2^20"]
    assert proof.hex() == gold["proof_hex"]
    tau_b, sel, perm, ocols, pi = coracle.mul_chain_inputs(log_n)
    oc = coracle.Circuit(tau_b, sel, perm, n)
    assert oc.prove(ocols, pi) == proof
    oc.close()
    group = Context.multi([0, 0])
    gc = synthetic.mul_chain_direct(group, log_n)
    assert gc.handle.prove_inputs(cols, bytes(32)) == proof
    group.close()


def test_quotient_coset_skip_equals_all_four_cosets(ctx):
    """An honest witness lets the prover skip coset 0 of the 4n domain (the numerator vanishes on H); forcing all four
    cosets must give the same bytes.  The broken-copy witness [3, 4, 6] always takes four (floor quotient)."""
    from typlonk_b200 import field as F, synthetic
    for log_n in (3, 9, 12):
        n = 1 << log_n
        circuit = synthetic.mul_chain_direct(ctx, log_n)
        cols = [F.fr_vec_to_bytes(c) for c in synthetic.mul_chain_witness(n - 3, n)]
        fast = circuit.handle.prove_inputs(cols, bytes(32))
        ctx.set_option("quotient_all_cosets", 1)
        try:
            assert circuit.handle.prove_inputs(cols, bytes(32)) == fast
        finally:
            ctx.set_option("quotient_all_cosets", 0)
        circuit.handle.destroy()
        circuit.srs.handle.destroy()


def test_golden_big_proofs_2_16_and_2_18(ctx):
    import json
    from typlonk_b200 import field as F, synthetic
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mulchain_big.json")))
    for log_n in (16, 18):
        n = 1 << log_n
        circuit = synthetic.mul_chain_direct(ctx, log_n)
        cols = [F.fr_vec_to_bytes(c) for c in synthetic.mul_chain_witness(n - 3, n)]
        assert circuit.handle.prove_inputs(cols, bytes(32)).hex() == gold["2^%d" % log_n]["proof_hex"]
        circuit.handle.destroy()
        circuit.srs.handle.destroy()


@pytest.mark.parametrize("log_n", [10, 20])
def test_large_proof_verifies_with_trapdoor_verifier(ctx, log_n):
    """verify() at BASELINE scale: the O(n) trapdoor verifier of the oracle accepts the GPU proof
    (and rejects it after a one-byte corruption)."""
    from oracle.pyoracle import fastverify
    from typlonk_b200 import field as F, synthetic
    n = 1 << log_n
    circuit = synthetic.mul_chain_direct(ctx, log_n)
    cols = synthetic.mul_chain_witness(n - 3, n)
    proof = circuit.handle.prove([F.fr_vec_to_bytes(c) for c in cols], bytes(32 * n))
    sig = [F.g1_from_abi(s) for s in circuit.handle.sigma_commitments()]
    ok = fastverify.verify_trapdoor(proof, n, synthetic.tau(), circuit.perm.perm, circuit.fixed_commitments, sig)
    assert ok
    if log_n == 10:
        bad = bytearray(proof)
        bad[96 * 2 + 5] ^= 1  # a.y
        assert not fastverify.verify_trapdoor(bytes(bad), n, synthetic.tau(), circuit.perm.perm,
                                              circuit.fixed_commitments, sig)
    circuit.handle.destroy()
    circuit.srs.handle.destroy()


def test_nonzero_public_inputs_general_path_alternating_with_the_zero_fast_path(ctx):
    """A zero public-input vector skips that polynomial's transforms (api.cu); a non-zero one that the witness
    compensates (c = a b + PI on a Mul row, so the gate line still vanishes) takes the general path.  Both must match
    the oracle byte for byte, in any order on the same circuit, with tp_verify in between (it reuses the buffers)."""
    from typlonk_b200 import field as F
    gates = 13
    circuit = mul_chain(gates).build(ctx, TAU)
    oc = _oracle_circuit(obuilder.make_mul_chain(gates), 2)
    n = circuit.rows
    cols = py_tracer.witness(circuit.desc, circuit.rows, [3, 5], BLINDERS)
    zero_pis = [0] * n
    want_zero = oplonk.prove_columns(oc, [list(c) for c in cols], list(zero_pis)).to_bytes()[:1472]
    pis = list(zero_pis)
    pis[2], pis[5] = 7, F.R_MOD - 1
    cols2 = [list(c) for c in cols]
    for j in (2, 5):
        cols2[2][j] = (cols2[0][j] * cols2[1][j] + pis[j]) % F.R_MOD
    want_nz = oplonk.prove_columns(oc, [list(c) for c in cols2], list(pis)).to_bytes()[:1472]
    assert want_nz != want_zero

    def run(c, p):
        return circuit.handle.prove([F.fr_vec_to_bytes(x) for x in c], F.fr_vec_to_bytes(p))
    assert run(cols, zero_pis) == want_zero
    assert run(cols2, pis) == want_nz
    assert run(cols, zero_pis) == want_zero
    assert run(cols, zero_pis) == want_zero
    assert circuit.handle.verify(want_zero, F.fr_vec_to_bytes(zero_pis))
    assert run(cols, zero_pis) == want_zero
    assert not circuit.handle.verify(want_nz, F.fr_vec_to_bytes(pis))   # copy constraints are broken by construction
    assert run(cols2, pis) == want_nz
    assert run(cols, zero_pis) == want_zero
    # a non-zero public input the witness does not compensate still trips the gate check
    with pytest.raises(GateUnsatisfied):
        run(cols, pis)
