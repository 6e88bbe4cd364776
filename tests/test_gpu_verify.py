"""GPU tests of the verifier (tp_verify, tp_kzg_verify, tp_srs_g2 through the Python mirror of the reference API):
the reference's own verify assertions (plonk/src/builder/test.rs:25-44, kzg/src/lib.rs:95-110) and the same verdicts
as the oracle on honest, dishonest and corrupted proofs."""
import pytest

from oracle.pyoracle import builder as obuilder, curve, kzg as okzg, plonk as oplonk, rng
from typlonk_b200 import field as F, synthetic
from typlonk_b200.ffi import TyplonkError
from typlonk_b200.kzg import KzgScheme, Srs
from typlonk_b200.plonk import Proof

from test_gpu_prove import Additive, Pythagoras, mul_chain, BLINDERS, TAU

pytestmark = pytest.mark.gpu


def test_srs_g2_matches_oracle(ctx):
    srs = Srs.from_secret(ctx, TAU, 5)
    osrs = okzg.Srs.from_secret(TAU, 5)
    flat = lambda p: ((p[0].a, p[0].b), (p[1].a, p[1].b))  # noqa: E731
    assert srs.g2_ref() == flat(osrs.g2) == flat(curve.G2_GEN)
    assert srs.g2s_ref() == flat(osrs.g2s)
    srs.handle.destroy()


def test_kzg_commit_open_verify(ctx):
    """kzg/src/lib.rs:95-110 `commit`: tau = 2, p = 1 + 2X + 3X^2, opened at 1."""
    srs = Srs.from_secret(ctx, 2, 10)
    scheme = KzgScheme(srs)
    com = scheme.commit([1, 2, 3])
    assert com == curve.g1_mul(curve.G1_GEN, 17)
    opening = scheme.open([1, 2, 3], 1)
    assert opening[1] == 6
    assert scheme.verify(com, opening, 1)
    assert not scheme.verify(com, (opening[0], 5), 1)
    assert not scheme.verify(com, opening, 3)
    # an uploaded SRS has no G2 points until they are supplied
    up = Srs.from_points(ctx, srs.g1_ref())
    with pytest.raises(TyplonkError):
        up.handle.g2()
    up.handle.set_g2(*srs.handle.g2())
    assert KzgScheme(up).verify(com, opening, 1)
    up.handle.destroy()
    srs.handle.destroy()


def test_readme_circuit_verify(ctx):
    """builder/test.rs:25-37: [3,4,5] verifies, [3,4,6] yields a proof that does not."""
    circuit = Pythagoras.build(ctx, TAU)
    oc = obuilder.compile_circuit(obuilder.circuit_pythagoras, 3, TAU)
    good = circuit.prove([3, 4, 5], [0], BLINDERS)
    assert circuit.verify(good) and oplonk.verify(oc, oplonk.prove(oc, [3, 4, 5], [0], BLINDERS))
    bad = circuit.prove([3, 4, 6], [0], BLINDERS)
    assert not circuit.verify(bad)
    assert not oplonk.verify(oc, oplonk.prove(oc, [3, 4, 6], [0], BLINDERS), use_trapdoor=True)
    # verify twice (commitments cached after the first call), then with wrong public inputs
    assert circuit.verify(good)
    assert not circuit.verify(Proof(good.fixed, [1]))
    assert circuit.verify(Proof(good.fixed, []))           # resized with zeros like proof.rs:204-205
    with pytest.raises(TyplonkError):
        circuit.handle.verify(good.fixed[:100], b"")


def test_additive_circuit_verify(ctx):
    """builder/test.rs:39-44."""
    circuit = Additive.build(ctx, TAU)
    assert circuit.verify(circuit.prove([2, 7, 2, 3, 4], [0], BLINDERS))
    assert not circuit.verify(circuit.prove([2, 7, 2, 3, 5], [0], BLINDERS))


@pytest.mark.parametrize("gates", [13, 253])
def test_mul_chain_verify_and_corruption(ctx, gates):
    circuit = mul_chain(gates).build(ctx, TAU)
    proof = circuit.prove([3, 5], [0], BLINDERS)
    assert circuit.verify(proof)
    for off in (5, 96 + 7, 192, 672 + 3, 1024 + 1, 1056 + 9, 1344 + 2, 1440):
        bad = bytearray(proof.fixed)
        bad[off] ^= 1
        assert not circuit.verify(Proof(bytes(bad), proof.public_inputs)), off


@pytest.mark.parametrize("log_n", [16, 20])
def test_large_proof_verifies(ctx, log_n):
    """BASELINE.json configs[1], configs[2]: prove + verify on the device path; the corrupted proof is rejected."""
    n = 1 << log_n
    circuit = synthetic.mul_chain_direct(ctx, log_n)
    cols = synthetic.mul_chain_witness(n - 3, n)
    fixed = circuit.handle.prove([F.fr_vec_to_bytes(c) for c in cols], bytes(32 * n))
    assert circuit.handle.verify(fixed, bytes(32))
    bad = bytearray(fixed)
    bad[224 + 192] ^= 1  # b's evaluation
    assert not circuit.handle.verify(bytes(bad), bytes(32))
    circuit.handle.destroy()
    circuit.srs.handle.destroy()
