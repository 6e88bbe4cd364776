"""Loader for arkworks-generated vectors (tools/arkworks_dump, needs a Rust toolchain the build image does not have).
Skipped until tests/golden/arkworks_vectors.json exists; the first box with cargo turns the oracle's parity with real
arkworks from "pinned by the crates' published constants" into "pinned by arkworks' own output", byte for byte:
serialize_unchecked, StdRng::seed_from_u64, Fr::rand, the challenge chain, and two whole proofs of the reference."""
import json
import os

import pytest

PATH = os.path.join(os.path.dirname(__file__), "golden", "arkworks_vectors.json")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="no arkworks vectors yet: run tools/arkworks_dump/run.sh where cargo exists")


@pytest.fixture(scope="module")
def vec():
    return json.load(open(PATH))


def test_serialize_unchecked(vec):
    from oracle.pyoracle import curve
    assert curve.g1_serialize_unchecked(curve.G1_GEN).hex() == vec["g1_generator_unchecked"]
    assert curve.g1_serialize_unchecked(curve.g1_mul(curve.G1_GEN, 17)).hex() == vec["g1_17_unchecked"]
    assert curve.g1_serialize_unchecked(None).hex() == vec["g1_zero_unchecked"]


def test_stdrng_and_fr_rand(vec):
    from oracle.pyoracle import fields, rng
    from oracle import coracle
    from typlonk_b200 import ffi
    for k in (0, 1, 2, 13037422643194131432):
        r = rng.StdRng.seed_from_u64(k)
        assert [r.next_u64() for _ in range(4)] == vec["stdrng_seed_from_u64_%d" % k]
        w = ffi.stdrng_words(8, seed_u64=k)
        assert [w[2 * i] | (w[2 * i + 1] << 32) for i in range(4)] == vec["stdrng_seed_from_u64_%d" % k]
    for k in (1, 2, 3, 4):
        want = "".join(vec["fr_rand_montgomery_seed_%d" % k])
        got = rng.fr_rand_stream(k, 9)
        assert b"".join(fields.fr_to_mont(x).to_bytes(32, "little") for x in got).hex() == want
        assert [x.to_bytes(32, "little").hex() for x in got] == vec["fr_rand_canonical_seed_%d" % k]
        assert coracle.fr_rand_stream(k, 9).hex() == want          # the C++ oracle
        assert ffi.fr_rand_stream(k, 9).hex() == want              # the library's transcript generator


def test_challenge_chain(vec):
    from oracle.pyoracle import curve, fields, rng
    t = curve.g1_serialize_unchecked(curve.g1_mul(curve.G1_GEN, 17)) * 3
    assert rng.challenge_seed(t) == vec["challenge_chain_17G_x3"]["seed"]
    got = [fields.fr_to_mont(x).to_bytes(32, "little").hex() for x in rng.generate_challenges(t, 2)]
    assert got == vec["challenge_chain_17G_x3"]["challenges_montgomery"]


def test_domain(vec):
    from oracle.pyoracle import fields
    assert fields.root_of_unity(8).to_bytes(32, "little").hex() == vec["omega_8_canonical"]
    assert fields.FR_ROOT_OF_UNITY.to_bytes(32, "little").hex() == vec["two_adic_root_canonical"]


def test_whole_proofs_of_the_reference(vec):
    """The reference's own prover under tau = seed 1, blinders = seed 2 against the committed golden proofs (which the
    GPU path reproduces byte for byte in tests/test_gpu_prove.py)."""
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "proofs.json")))
    assert vec["readme_pythagoras_3_4_5_proof"] == gold["readme_pythagoras_3_4_5"]["proof_hex"]
    assert vec["mulchain_13_gates_proof"] == gold["mulchain_13_gates"]["proof_hex"]
