"""GPU parity of the opt-in batch-affine MSM rounds (tp_ctx_set_option "msm_affine_rounds"):
the commitment must not depend on the number of rounds -- same oracle, same edge cases as the
default XYZZ accumulation (doubling, cancellation, identity operands, skewed buckets)."""
import pytest

from oracle.pyoracle import curve, fields, rng
from typlonk_b200.ffi import Context
from typlonk_b200.kzg import KzgScheme, Srs

pytestmark = pytest.mark.gpu
R = fields.R_MOD


@pytest.fixture(scope="module", params=[1, 3])
def actx(request):
    c = Context(0)
    c.set_option("msm_affine_rounds", request.param)
    yield c
    c.close()


def test_option_validation(actx):
    from typlonk_b200.ffi import TyplonkError
    with pytest.raises(TyplonkError):
        actx.set_option("msm_affine_rounds", 99)
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
