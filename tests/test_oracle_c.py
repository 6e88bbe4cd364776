"""CPU: the C++ oracle port (oracle/c/oracle.cpp) against the Python spec oracle and the golden
fixtures -- primitives and whole proofs, byte for byte."""
import json
import os

import pytest

from oracle import coracle
from oracle.pyoracle import builder, curve, fields, plonk, poly, rng

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PROOFS = json.load(open(os.path.join(GOLD, "proofs.json")))
R = fields.R_MOD


@pytest.fixture(scope="module", autouse=True)
def _built():
    coracle.build()


def test_fr_rand_stream_matches():
    assert coracle.fr_rand_stream(7, 40) == fields.fr_vec_to_mont_bytes(rng.fr_rand_stream(7, 40))


@pytest.mark.parametrize("log_n", [1, 2, 5, 10])
def test_ntt_matches(log_n):
    n = 1 << log_n
    data = rng.fr_rand_stream(4, n)
    raw = fields.fr_vec_to_mont_bytes(data)
    assert fields.fr_vec_from_mont_bytes(coracle.ntt(raw, log_n)) == poly.Domain(n).fft(data)
    assert fields.fr_vec_from_mont_bytes(coracle.ntt(raw, log_n, True)) == poly.Domain(n).ifft(data)


def test_srs_and_msm_match():
    tau = rng.fr_rand_stream(1, 1)[0]
    n = 300
    s = coracle.srs(fields.fr_mont_bytes(tau), n)
    pts = [(fields.fq_from_mont_bytes(s[i * 96:i * 96 + 48]), fields.fq_from_mont_bytes(s[i * 96 + 48:i * 96 + 96]))
           for i in range(n)]
    for i in (0, 1, 2, 17, 255, 256, 299):
        assert pts[i] == curve.g1_mul(curve.G1_GEN, pow(tau, i, R))
    for cnt in (1, 31, 32, 300):
        sc = rng.fr_rand_stream(3, cnt)
        m = coracle.msm(s[: 96 * cnt], fields.fr_vec_to_mont_bytes(sc))
        got = None if m[96] else (fields.fq_from_mont_bytes(m[:48]), fields.fq_from_mont_bytes(m[48:96]))
        assert got == curve.g1_msm(pts[:cnt], sc)


@pytest.mark.parametrize("log_n,golden", [(4, "mulchain_13_gates"), (6, "mulchain_61_gates")])
def test_mul_chain_proof_matches_golden(log_n, golden):
    tau, sel, perm, cols, pi = coracle.mul_chain_inputs(log_n)
    c = coracle.Circuit(tau, sel, perm, 1 << log_n)
    proof = c.prove(cols, pi)
    assert proof.hex() == PROOFS[golden]["proof_hex"][: 2 * 1472]
    fixed = c.fixed_commitments()
    ser = []
    for f in fixed:
        pt = None if f[96] else (fields.fq_from_mont_bytes(f[:48]), fields.fq_from_mont_bytes(f[48:96]))
        ser.append(curve.g1_serialize_unchecked(pt).hex())
    assert ser == PROOFS[golden]["fixed_commitments"]
    c.close()


def test_readme_circuit_through_c_oracle():
    """The C++ port is driven by (selector evals, perm, columns): feed it the README circuit traced by
    the Python oracle and compare whole proofs."""
    tau = rng.fr_rand_stream(1, 1)[0]
    blinders = rng.fr_rand_stream(2, 9)
    for run, nin, inputs, golden in ((builder.circuit_pythagoras, 3, [3, 4, 5], "readme_pythagoras_3_4_5"),
                                     (builder.circuit_pythagoras, 3, [3, 4, 6], "readme_pythagoras_bad_3_4_6"),
                                     (builder.circuit_additive, 5, [2, 7, 2, 3, 4], "additive_2_7_2_3_4")):
        gates, perm = builder.trace(run, nin)
        n = len(gates)
        sel = [fields.fr_vec_to_mont_bytes([builder.GATE_ROWS[g][k] for g in gates]) for k in range(5)]
        cols = builder.witness_columns(run, inputs, n, blinders)
        c = coracle.Circuit(fields.fr_mont_bytes(tau), sel, perm.perm, n)
        proof = c.prove([fields.fr_vec_to_mont_bytes(col) for col in cols], bytes(32 * n))
        assert proof.hex() == PROOFS[golden]["proof_hex"][: 2 * 1472], golden
        c.close()


def test_gate_violation_is_reported_like_the_reference_panic():
    tau, sel, perm, cols, pi = coracle.mul_chain_inputs(3)
    c = coracle.Circuit(tau, sel, perm, 8)
    bad_pi = fields.fr_mont_bytes(1) + bytes(32 * 7)
    with pytest.raises(AssertionError):
        c.prove(cols, bad_pi)
    c.close()


def test_as_written_pieces_agree():
    """The "reference as written" kernels used for the B0 timing give the same results as the fast
    ones: per-point double-and-add commit, schoolbook product."""
    tau, sel, perm, cols, pi = coracle.mul_chain_inputs(4)
    c = coracle.Circuit(tau, sel, perm, 16)
    coeffs = coracle.ntt(cols[0], 4, True)
    srs = coracle.srs(tau, 16)
    assert c.commit_as_written(coeffs) == coracle.msm(srs, coeffs)
    a = rng.fr_rand_stream(5, 9)
    b = rng.fr_rand_stream(6, 5)
    got = fields.fr_vec_from_mont_bytes(coracle.naive_mul(fields.fr_vec_to_mont_bytes(a), fields.fr_vec_to_mont_bytes(b)))
    assert poly.strip(got) == poly.naive_mul(a, b)
    c.close()
