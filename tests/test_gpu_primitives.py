"""GPU parity tests (through the C ABI) of the hot-path primitives against the CPU oracle:
field/curve self-test, NTT / iNTT / coset NTT, SRS generation, KZG commit (MSM) and open,
permutation grand product.  Bit-exact (integer work)."""
import random

import pytest

from oracle.pyoracle import curve, fields, kzg as okzg, permutation as operm, poly, rng
from typlonk_b200 import field as F
from typlonk_b200.ffi import TyplonkError
from typlonk_b200.kzg import KzgScheme, Srs

pytestmark = pytest.mark.gpu
NTT_RADIX_DEFAULT = 2   # tp_ctx option "ntt_radix_log" as the library ships it
R = fields.R_MOD


def test_device_selftest(ctx):
    assert ctx.selftest() == 0


def test_imad_peak_reported(ctx):
    imad, wide = ctx.measure_imad_peak()
    print("IMAD/s %.3e  IMAD.WIDE/s %.3e" % (imad, wide))
    assert imad > 1e12 and wide > 1e11


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 7, 10, 11, 12, 13, 16])
def test_ntt_matches_oracle(ctx, log_n):
    n = 1 << log_n
    data = rng.fr_rand_stream(4, n) if n <= 4096 else [pow(3, i, R) * 12345 % R for i in range(n)]
    dom = poly.Domain(n)
    fwd = F.fr_vec_from_bytes(ctx.ntt(F.fr_vec_to_bytes(data), log_n))
    assert fwd == dom.fft(data)
    inv = F.fr_vec_from_bytes(ctx.ntt(F.fr_vec_to_bytes(data), log_n, inverse=True))
    assert inv == dom.ifft(data)


@pytest.mark.parametrize("log_n", [3, 10, 13])
def test_coset_ntt_round_trip(ctx, log_n):
    n = 1 << log_n
    data = rng.fr_rand_stream(4, n)
    g = F.fr_to_bytes(7)
    ev = ctx.ntt(F.fr_vec_to_bytes(data), log_n, coset_mont=g)
    assert F.fr_vec_from_bytes(ev) == poly.Domain(n).coset_fft(data, 7)
    back = ctx.ntt(ev, log_n, inverse=True, coset_mont=g)
    assert F.fr_vec_from_bytes(back) == data


@pytest.mark.parametrize("radix_log", [3, 2])
@pytest.mark.parametrize("log_n", list(range(1, 23)))
def test_ntt_every_pass_geometry_vs_c_oracle(ctx, log_n, radix_log):
    """Every pass plan of the register-round kernel -- eight elements per thread (rounds of 3 stages, remainders 0/1/2)
    and four (rounds of 2, remainders 0/1); 1-3 passes -- against the C++ oracle's radix-2 NTT, forward, inverse and
    on a coset, bit-exact."""
    import numpy as np
    from oracle import coracle
    n = 1 << log_n
    rs = np.random.RandomState(100 + log_n)
    raw = rs.randint(0, 2**32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    raw[:, 7] &= 0x3FFFFFFF
    data = raw.tobytes()
    g = F.fr_to_bytes(7)
    ctx.set_option("ntt_radix_log", 3)
    want_coset = ctx.ntt(data, log_n, coset_mont=g) if radix_log == 2 else None
    ctx.set_option("ntt_radix_log", radix_log)
    try:
        assert ctx.ntt(data, log_n) == coracle.ntt(data, log_n)
        assert ctx.ntt(data, log_n, inverse=True) == coracle.ntt(data, log_n, inverse=True)
        if want_coset is not None:   # both round shapes agree on a coset, and the inverse coset transform undoes it
            assert ctx.ntt(data, log_n, coset_mont=g) == want_coset
            assert ctx.ntt(want_coset, log_n, inverse=True, coset_mont=g) == data
    finally:
        ctx.set_option("ntt_radix_log", NTT_RADIX_DEFAULT)


@pytest.mark.parametrize("log_n", [1, 2, 5, 12, 15, 19])
def test_coset_ntt_vs_scaled_plain(ctx, log_n):
    """coset NTT(x)_i = NTT(x_j g^j)_i, and the inverse coset transform undoes it (multi-pass sizes)."""
    import numpy as np
    n = 1 << log_n
    rs = np.random.RandomState(200 + log_n)
    raw = rs.randint(0, 2**32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    raw[:, 7] &= 0x3FFFFFFF
    data = raw.tobytes()
    g = F.fr_to_bytes(7)
    ev = ctx.ntt(data, log_n, coset_mont=g)
    vals = F.fr_vec_from_bytes(data)
    gp, scaled = 1, []
    for v in vals:
        scaled.append(v * gp % R)
        gp = gp * 7 % R
    assert ev == ctx.ntt(F.fr_vec_to_bytes(scaled), log_n)
    assert ctx.ntt(ev, log_n, inverse=True, coset_mont=g) == data


def test_ntt_linearity_large(ctx):
    """Size-independent property at a BASELINE-scale size (2^20): NTT(a + b) = NTT(a) + NTT(b),
    iNTT(NTT(a)) = a."""
    import numpy as np
    log_n = 20
    n = 1 << log_n
    rs = np.random.RandomState(7)
    def rand_vec():
        raw = rs.randint(0, 2**32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
        raw[:, 7] &= 0x3FFFFFFF  # < 2^254 < r : valid Montgomery representatives
        return raw
    a, b = rand_vec(), rand_vec()
    def to_int_rows(x):
        return x
    fa = np.frombuffer(ctx.ntt(a.tobytes(), log_n), dtype=np.uint32).reshape(n, 8)
    back = np.frombuffer(ctx.ntt(fa.tobytes(), log_n, inverse=True), dtype=np.uint32).reshape(n, 8)
    assert (back == a).all()
    # linearity on a sample of rows (big-int add mod r on the host)
    fb = np.frombuffer(ctx.ntt(b.tobytes(), log_n), dtype=np.uint32).reshape(n, 8)
    def row_int(x, i):
        return int.from_bytes(x[i].tobytes(), "little")
    s = np.zeros((n, 8), dtype=np.uint32)
    ab = [(row_int(a, i) + row_int(b, i)) % R for i in range(n)]
    s = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in ab), dtype=np.uint32).reshape(n, 8)
    fs = np.frombuffer(ctx.ntt(s.tobytes(), log_n), dtype=np.uint32).reshape(n, 8)
    for i in list(range(0, n, n // 64)) + [1, n - 1]:
        assert row_int(fs, i) == (row_int(fa, i) + row_int(fb, i)) % R


def test_srs_from_secret_and_reference_commit_test(ctx):
    """kzg/src/lib.rs:95-109 (`commit` test): Srs::from_secret(2, 10), commit(1 + 2X + 3X^2) == 17 G,
    p(1) == 6."""
    srs = Srs.from_secret(ctx, 2, 10)
    assert len(srs) == 13
    expect = okzg.Srs.from_secret(2, 10).g1
    assert srs.g1_ref() == expect
    scheme = KzgScheme(srs)
    com = scheme.commit([1, 2, 3])
    assert com == curve.g1_mul(curve.G1_GEN, 17)
    w, y = scheme.open([1, 2, 3], 1)
    assert y == 6
    assert (w, y) == okzg.open_at(okzg.Srs(expect, None, None), [1, 2, 3], 1)


def test_srs_random_tau_sampled(ctx):
    tau = rng.fr_rand_stream(1, 1)[0]
    srs = Srs.from_secret(ctx, tau, 1000)
    pts = srs.g1_ref()
    assert len(pts) == 1003
    for i in (0, 1, 2, 7, 8, 9, 500, 1002):
        assert pts[i] == curve.g1_mul(curve.G1_GEN, pow(tau, i, R)), i
    assert all(curve.g1_is_on_curve(p) for p in pts[::37])


@pytest.fixture(scope="module")
def srs1k(ctx):
    tau = rng.fr_rand_stream(1, 1)[0]
    srs = Srs.from_secret(ctx, tau, 1021)  # 1024 points
    return srs, srs.g1_ref()


@pytest.mark.parametrize("n", [0, 1, 2, 8, 33, 257, 1024])
def test_commit_matches_oracle(ctx, srs1k, n):
    srs, pts = srs1k
    scalars = rng.fr_rand_stream(3, n)
    if n:
        scalars[-1] = scalars[-1] or 1
    got = KzgScheme(srs).commit(scalars)
    assert got == curve.g1_msm(pts[:n], scalars)


def test_commit_edge_scalars(ctx, srs1k):
    srs, pts = srs1k
    n = 200
    cases = {
        "zeros_then_one": [0] * (n - 1) + [1],
        "all_r_minus_1": [R - 1] * n,
        "all_equal_small": [5] * n,
        "small": [i % 7 for i in range(n - 1)] + [3],
        "powers_of_two": [pow(2, i, R) for i in range(n)],
        "cancel": [1, R - 1] + [0] * 5 + [9],
    }
    for name, sc in cases.items():
        got = KzgScheme(srs).commit(sc)
        assert got == curve.g1_msm(pts[:len(sc)], sc), name
    # result at infinity: s*P0 + (-s*tau^-1)... simpler: scalars (tau, -1) on (G, tau G)
    tau = rng.fr_rand_stream(1, 1)[0]
    assert KzgScheme(srs).commit([tau, R - 1]) is None
    assert KzgScheme(srs).commit([]) is None


def test_commit_repeated_points(ctx):
    """SRS with repeated / opposite / infinity points exercises the doubling and cancellation
    branches of the bucket accumulation."""
    g = curve.G1_GEN
    p2 = curve.g1_mul(g, 2)
    pts = [g, g, curve.g1_neg(g), p2, None, g, p2, p2] * 4
    srs = Srs.from_points(ctx, pts)
    for sc in ([1] * 32, [3] * 32, list(range(1, 33)), [R - 1] * 32):
        assert KzgScheme(srs).commit(sc) == curve.g1_msm(pts, sc)


def test_commit_srs_too_short(ctx, srs1k):
    srs, _ = srs1k
    with pytest.raises(TyplonkError) as e:
        KzgScheme(srs).commit([1] * 1025)
    assert e.value.code == 3


def test_commit_homomorphism(ctx, srs1k):
    """kzg/src/lib.rs:160-171 (`scalar_mul` test): commit(9 p) == 9 commit(p)."""
    srs, _ = srs1k
    p = [1, 2, 3, 4, 5]
    c1 = KzgScheme(srs).commit(p)
    c2 = KzgScheme(srs).commit([9 * x for x in p])
    assert curve.g1_mul(c1, 9) == c2


@pytest.mark.parametrize("n", [1, 2, 5, 64, 65, 1000])
def test_open_matches_oracle(ctx, srs1k, n):
    srs, pts = srs1k
    p = rng.fr_rand_stream(5, n)
    p[-1] = p[-1] or 1
    z = rng.fr_rand_stream(6, 1)[0]
    w, y = KzgScheme(srs).open(p, z)
    ow, oy = okzg.open_at(okzg.Srs(pts, None, None), p, z)
    assert y == oy == poly.evaluate(p, z)
    assert w == ow


def test_open_empty_poly_errors(ctx, srs1k):
    srs, _ = srs1k
    with pytest.raises(TyplonkError) as e:
        KzgScheme(srs).open([], 3)
    assert e.value.code == 4


@pytest.mark.parametrize("n", [4, 8, 64, 2048, 4096 + 0])
def test_perm_prove_matches_oracle(ctx, n):
    from typlonk_b200.permutation import CompiledPermutation
    rnd = random.Random(n)
    perm = list(range(3 * n))
    rnd.shuffle(perm)
    compiled = operm.Permutation(perm).compile()
    values = [rng.fr_rand_stream(10 + i, n) for i in range(3)]
    beta, gamma = rng.fr_rand_stream(11, 2)
    expect = compiled.prove(values, beta, gamma)
    ids = [[c[0] for c in col] for col in compiled.cols]
    sgs = [[c[1] for c in col] for col in compiled.cols]
    got = CompiledPermutation(ctx, ids, sgs, compiled.cosets).prove(values, beta, gamma)
    assert got == expect


@pytest.mark.parametrize("n", [4, 64, 4096])
def test_permutation_compile_matches_oracle(ctx, n):
    """Permutation::compile (permutation/src/lib.rs:101-154): id / sigma columns and coset representatives."""
    from typlonk_b200.permutation import Permutation
    rnd = random.Random(7 * n)
    perm = list(range(3 * n))
    rnd.shuffle(perm)
    want = operm.Permutation(perm).compile()
    got = Permutation(perm).compile(ctx)
    assert got.cosets == want.cosets == [2, 3, 4]
    assert got.ids == [[c[0] for c in col] for col in want.cols]
    assert got.sigmas == [[c[1] for c in col] for col in want.cols]
    values = [rng.fr_rand_stream(20 + i, n) for i in range(3)]
    beta, gamma = rng.fr_rand_stream(21, 2)
    assert got.prove(values, beta, gamma) == want.prove(values, beta, gamma)
    with pytest.raises(TyplonkError):
        Permutation(list(range(3 * n - 1)) + [3 * n]).compile(ctx)     # index out of range


def test_perm_prove_zero_denominator(ctx):
    from typlonk_b200.permutation import CompiledPermutation
    n = 8
    compiled = operm.Permutation(list(range(3 * n))).compile()
    ids = [[c[0] for c in col] for col in compiled.cols]
    sgs = [[c[1] for c in col] for col in compiled.cols]
    beta, gamma = 5, 9
    values = [[1] * n for _ in range(3)]
    values[1][3] = (-(beta * sgs[1][3] + gamma)) % R
    with pytest.raises(TyplonkError) as e:
        CompiledPermutation(ctx, ids, sgs, compiled.cosets).prove(values, beta, gamma)
    assert e.value.code == 5
    with pytest.raises(ZeroDivisionError):
        compiled.prove(values, beta, gamma)


# ---- BASELINE-scale MSM checks through the SRS trapdoor: commit(p) == p(tau) * G ----------------

@pytest.fixture(scope="module")
def srs_big(ctx):
    tau = rng.fr_rand_stream(1, 1)[0]
    return tau, Srs.from_secret(ctx, tau, (1 << 18) - 3)


def _expect_from_trapdoor(tau, scalars):
    acc = 0
    for s in reversed(scalars):
        acc = (acc * tau + s) % R
    return curve.g1_mul(curve.G1_GEN, acc)


def test_commit_large_random_vs_trapdoor(ctx, srs_big):
    tau, srs = srs_big
    n = 1 << 18
    seed = rng.fr_rand_stream(3, 64)
    scalars = [(seed[i % 64] * (i + 1) + i * i) % R for i in range(n)]
    assert KzgScheme(srs).commit(scalars) == _expect_from_trapdoor(tau, scalars)


@pytest.mark.parametrize("kind", ["all_equal", "two_values", "tiny_range", "top_bits_only", "sparse"])
def test_commit_skewed_buckets_vs_trapdoor(ctx, srs_big, kind):
    """Degenerate digit distributions: every window funnels its points into one or a few buckets
    (bucket sizes up to 2^17), which the merge must still reduce in parallel."""
    tau, srs = srs_big
    n = 1 << 17
    if kind == "all_equal":
        scalars = [rng.fr_rand_stream(9, 1)[0]] * n
    elif kind == "two_values":
        a, b = rng.fr_rand_stream(9, 2)
        scalars = [a if (i * 7) % 3 else b for i in range(n)]
    elif kind == "tiny_range":
        scalars = [(i * 2654435761) % 5 for i in range(n)]
        scalars[-1] = 1
    elif kind == "top_bits_only":
        scalars = [((i % 3) + 1) << 252 for i in range(n)]
    else:
        scalars = [0] * n
        for i in range(0, n, 1000):
            scalars[i] = R - 1 - i
        scalars[-1] = 7
    assert KzgScheme(srs).commit(scalars) == _expect_from_trapdoor(tau, scalars)
