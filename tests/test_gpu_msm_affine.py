"""GPU parity of the batch-affine MSM variants (tp_ctx_set_option "msm_affine_rounds" = pair
rounds, "msm_affine_chains" = lockstep affine chains): the commitment must not depend on the
accumulation strategy -- same oracle, same edge cases as the XYZZ accumulation (doubling,
cancellation, identity operands, skewed buckets)."""
import pytest

from oracle.pyoracle import curve, fields, rng
from typlonk_b200.ffi import Context
from typlonk_b200.kzg import KzgScheme, Srs

pytestmark = pytest.mark.gpu
R = fields.R_MOD


@pytest.fixture(scope="module", params=["rounds1", "rounds3", "chains", "xyzz"])
def actx(request):
    c = Context(0)
    c.set_option("msm_affine_chains", 1 if request.param == "chains" else 0)
    c.set_option("msm_affine_rounds", {"rounds1": 1, "rounds3": 3}.get(request.param, 0))
    yield c
    c.close()


def test_option_validation(actx):
    from typlonk_b200.ffi import TyplonkError
    with pytest.raises(TyplonkError):
        actx.set_option("msm_affine_rounds", 99)
    with pytest.raises(TyplonkError):
        actx.set_option("msm_affine_chains", 2)
    with pytest.raises(TyplonkError):
        actx.set_option("no_such_option", 1)


@pytest.mark.parametrize("n", [1, 2, 33, 257, 1024])
def test_commit_matches_oracle(actx, n):
    tau = rng.fr_rand_stream(1, 1)[0]
    srs = Srs.from_secret(actx, tau, 1021)
    pts = srs.g1_ref()
    scalars = rng.fr_rand_stream(3, n)
    scalars[-1] = scalars[-1] or 1
    assert KzgScheme(srs).commit(scalars) == curve.g1_msm(pts[:n], scalars)
    for sc in ([5] * 200, [R - 1] * 200, [1, R - 1] + [0] * 5 + [9], [i % 7 for i in range(199)] + [3]):
        assert KzgScheme(srs).commit(sc) == curve.g1_msm(pts[:len(sc)], sc)


def test_repeated_opposite_and_identity_points(actx):
    g = curve.G1_GEN
    p2 = curve.g1_mul(g, 2)
    pts = [g, g, curve.g1_neg(g), p2, None, g, p2, p2] * 8
    srs = Srs.from_points(actx, pts)
    for sc in ([1] * 64, [3] * 64, list(range(1, 65)), [R - 1] * 64, [2, 2, 2, 1] * 16):
        assert KzgScheme(srs).commit(sc) == curve.g1_msm(pts, sc)


@pytest.mark.parametrize("kind", ["random", "all_equal", "tiny_range", "sparse"])
def test_large_vs_trapdoor(actx, kind):
    tau = rng.fr_rand_stream(1, 1)[0]
    n = 1 << 16
    srs = Srs.from_secret(actx, tau, n - 3)
    if kind == "random":
        seed = rng.fr_rand_stream(3, 64)
        scalars = [(seed[i % 64] * (i + 1) + i * i) % R for i in range(n)]
    elif kind == "all_equal":
        scalars = [rng.fr_rand_stream(9, 1)[0]] * n
    elif kind == "tiny_range":
        scalars = [(i * 2654435761) % 5 for i in range(n)]
        scalars[-1] = 1
    else:
        scalars = [0] * n
        for i in range(0, n, 1000):
            scalars[i] = R - 1 - i
        scalars[-1] = 7
    acc = 0
    for s in reversed(scalars):
        acc = (acc * tau + s) % R
    assert KzgScheme(srs).commit(scalars) == curve.g1_mul(curve.G1_GEN, acc)


def test_proof_bytes_do_not_depend_on_the_strategy(actx):
    """Full prove of a 2^12-row mul-chain circuit: byte-identical to the C++ oracle's proof whatever
    the bucket accumulation does."""
    from oracle import coracle
    from typlonk_b200 import field as F, synthetic
    log_n = 12
    n = 1 << log_n
    circuit = synthetic.mul_chain_direct(actx, log_n)
    cols = synthetic.mul_chain_witness(n - 3, n)
    proof = circuit.handle.prove([F.fr_vec_to_bytes(c) for c in cols], bytes(32 * n))
    tau_b, sel, perm, ocols, pi = coracle.mul_chain_inputs(log_n)
    oc = coracle.Circuit(tau_b, sel, perm, n)
    assert oc.prove(ocols, pi) == proof
    oc.close()
