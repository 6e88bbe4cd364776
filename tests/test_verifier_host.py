"""CPU: the library's host-side verifier arithmetic (csrc/pairing.h, csrc/verify.cu) against the oracle -- pairing
product checks, KzgScheme::verify (kzg/src/lib.rs:66-81) and the host half of CompiledCircuit::verify
(plonk/src/proof.rs:195-281, 441-503) on the golden proofs.  These entry points take no context and touch no
device; the circuit-sized half of tp_verify is covered by the GPU tests."""
import json
import os
import subprocess

import pytest

from oracle.pyoracle import builder, curve, fields, kzg as okzg, plonk as oplonk, poly, rng
from typlonk_b200 import field as F, ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PROOFS = json.load(open(os.path.join(GOLD, "proofs.json")))
R = fields.R_MOD
TAU = rng.fr_rand_stream(1, 1)[0]
BLINDERS = rng.fr_rand_stream(2, 9)


def g2_abi(pt):
    return F.g2_to_abi(None if pt is None else ((pt[0].a, pt[0].b), (pt[1].a, pt[1].b)))


def test_tower_and_pairing_unit_tests(tmp_path):
    exe = tmp_path / "pairing_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tools", "host_tests", "pairing_test.cpp")],
                   check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0 and "FAIL" not in res.stdout, res.stdout


def test_pairing_check_is_bilinear_on_oracle_points():
    a, b = rng.fr_rand_stream(11, 2)
    pa, qb = curve.g1_mul(curve.G1_GEN, a), curve.g2_mul(curve.G2_GEN, b)
    pab = curve.g1_mul(curve.G1_GEN, a * b % R)
    g1 = lambda p: F.g1_to_abi(p)  # noqa: E731
    assert ffi.pairing_check([g1(pa), g1(curve.g1_neg(pab))], [g2_abi(qb), g2_abi(curve.G2_GEN)])
    assert not ffi.pairing_check([g1(pa), g1(pab)], [g2_abi(qb), g2_abi(curve.G2_GEN)])
    assert not ffi.pairing_check([g1(curve.G1_GEN)], [g2_abi(curve.G2_GEN)])          # non-degenerate
    assert ffi.pairing_check([g1(None), g1(pa)], [g2_abi(qb), g2_abi(None)])          # infinity on either side -> 1
    assert ffi.pairing_check([], [])
    off = (pa[0], (pa[1] + 1) % fields.Q_MOD)
    assert not ffi.pairing_check([g1(off)], [g2_abi(qb)])                              # off-curve point rejected
    # the same verdicts as the oracle's pairing
    assert curve.pairing_product_is_one([(pa, qb), (curve.g1_neg(pab), curve.G2_GEN)])


def test_kzg_verify_matches_oracle():
    """The reference's own test (kzg/src/lib.rs:95-110): p = 1 + 2X + 3X^2 under tau = 2, opened at 1 -> 6."""
    srs = okzg.Srs.from_secret(2, 10)
    p = [1, 2, 3]
    com = okzg.commit(srs, p)
    assert com == curve.g1_mul(curve.G1_GEN, 17)
    opening = okzg.open_at(srs, p, 1)
    assert opening[1] == 6
    g2, g2s = g2_abi(srs.g2), g2_abi(srs.g2s)
    args = lambda c, o, z: (g2, g2s, F.g1_to_abi(c), F.g1_to_abi(o[0]), F.fr_to_bytes(o[1]), F.fr_to_bytes(z))  # noqa: E731
    assert ffi.kzg_verify(*args(com, opening, 1)) and okzg.verify(srs, com, opening, 1)
    assert not ffi.kzg_verify(*args(com, (opening[0], 7), 1))
    assert not ffi.kzg_verify(*args(com, opening, 2))
    assert not ffi.kzg_verify(*args(curve.g1_mul(curve.G1_GEN, 18), opening, 1))
    # random polynomial / point, secret from the SURVEY 8(d) stream; constant polynomial (witness = infinity)
    srs = okzg.Srs.from_secret(TAU, 13)
    g2, g2s = g2_abi(srs.g2), g2_abi(srs.g2s)
    p = rng.fr_rand_stream(5, 16)
    z = rng.fr_rand_stream(6, 1)[0]
    com, opening = okzg.commit(srs, p), okzg.open_at(srs, p, z)
    assert ffi.kzg_verify(*args(com, opening, z))
    assert not ffi.kzg_verify(*args(com, (opening[0], (opening[1] + 1) % R), z))
    com, opening = okzg.commit(srs, [9]), okzg.open_at(srs, [9], z)
    assert opening[0] is None and ffi.kzg_verify(*args(com, opening, z))
    with pytest.raises(ffi.TyplonkError):                                              # G2 point off the twist
        ffi.kzg_verify(g2, F.g2_to_abi(((1, 2), (3, 4))), *args(com, opening, z)[2:])


CASES = [("readme_pythagoras_3_4_5", builder.circuit_pythagoras, 3, [3, 4, 5]),
         ("additive_2_7_2_3_4", builder.circuit_additive, 5, [2, 7, 2, 3, 4]),
         ("mulchain_13_gates", builder.make_mul_chain(13), 2, [3, 5])]


def _verifier_key(circuit, point, public_inputs):
    """What tp_verify computes on the device, here from the oracle."""
    dom, perm, srs = circuit.domain, circuit.copy_constrains, circuit.srs
    pis = (list(public_inputs) + [0] * circuit.rows)[: circuit.rows]
    return dict(fixed=[F.g1_to_abi(c) for c in circuit.fixed_commitments],
                sigma=[F.g1_to_abi(c) for c in perm.sigma_commitments(srs, dom)],
                identity=F.g1_to_abi(okzg.identity(srs)), g2=g2_abi(srs.g2), g2s=g2_abi(srs.g2s),
                cosets_mont=[F.fr_to_bytes(k) for k in perm.cosets],
                sigma_evals_mont=[F.fr_to_bytes(s) for s in perm.sigma_evals(point, dom)[:2]],
                public_eval_mont=F.fr_to_bytes(poly.evaluate(poly.interpolate(pis, dom), point)), n=circuit.rows)


@pytest.mark.parametrize("name,run,nin,inputs", CASES)
def test_verify_prepared_accepts_golden_and_rejects_tampering(name, run, nin, inputs):
    circuit = builder.compile_circuit(run, nin, TAU)
    proof = oplonk.prove(circuit, inputs, [0], BLINDERS)
    raw = proof.to_bytes()
    assert raw.hex() == PROOFS[name]["proof_hex"]
    fixed = raw[:1472]
    alpha, beta, gamma, point = (F.fr_from_bytes(b) for b in ffi.proof_challenges(fixed))
    assert (alpha, beta, gamma, point) == oplonk._verify_challenges(proof)
    assert point == proof.evaluation_point
    key = _verifier_key(circuit, point, proof.public_inputs)
    assert ffi.verify_prepared(proof_fixed=fixed, **key)
    if name == "readme_pythagoras_3_4_5":
        assert oplonk.verify(circuit, proof)                       # the oracle's own pairing verifier agrees
    # any single corrupted field is rejected (each offset is the low byte of a scalar / coordinate)
    for off in (192, 192 + 224, 192 + 448, 672 + 192, 896 + 96, 1024, 1440):
        bad = bytearray(fixed)
        bad[off] ^= 1
        assert not ffi.verify_prepared(proof_fixed=bytes(bad), **key), off
    # swapped commitments: still valid points, wrong statement
    bad = fixed[224:448] + fixed[:224] + fixed[448:]
    assert not ffi.verify_prepared(proof_fixed=bad, **key)
    # wrong public input evaluation / sigma evaluation / verification key
    assert not ffi.verify_prepared(proof_fixed=fixed, **dict(key, public_eval_mont=F.fr_to_bytes(1)))
    assert not ffi.verify_prepared(proof_fixed=fixed, **dict(key, sigma_evals_mont=key["sigma_evals_mont"][::-1]))
    assert not ffi.verify_prepared(proof_fixed=fixed, **dict(key, fixed=key["fixed"][::-1]))
    # non-canonical scalar (>= r) is a malformed proof, not a rejected one
    bad = bytearray(fixed)
    bad[192:224] = b"\xff" * 32
    with pytest.raises(ffi.TyplonkError):
        ffi.verify_prepared(proof_fixed=bytes(bad), **key)


def test_verify_prepared_rejects_nonzero_r_evaluation():
    """proof.rs:232: `open_valid && r_opening.1.is_zero()` -- an honest opening of a polynomial with r(zeta) != 0."""
    circuit = builder.compile_circuit(builder.circuit_pythagoras, 3, TAU)
    good = oplonk.prove(circuit, [3, 4, 5], [0], BLINDERS)
    fixed = good.to_bytes()[:1472]
    key = _verifier_key(circuit, good.evaluation_point, [0])
    assert ffi.verify_prepared(proof_fixed=fixed, **key)
    # public input that does not match the proof: PI(zeta) changes, so the r commitment no longer opens to 0
    key_bad = _verifier_key(circuit, good.evaluation_point, [1])
    assert not ffi.verify_prepared(proof_fixed=fixed, **key_bad)
