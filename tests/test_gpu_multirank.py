"""GPU, several ranks on ONE device: the multi-GPU plan of SURVEY.md 8(e) exactly as the library runs it on a box of
B200s -- every MSM sharded by bucket (rank r owns the buckets b with b mod world == r) with the per-rank reduction
outputs all-gathered and combined on the device, the quotient sharded by coset of the 4n domain with one device
broadcast per coset, the witness upload sharded by rows -- must produce the bytes of the single-rank prover and of the
C++ oracle.  `Context.multi([0] * world)` is the single-process device group of tp_ctx_create_multi with every rank
on device 0; the ranks then exchange through peer copies instead of NCCL, everything else is the code the 8-GPU run
executes.  world = 2, 3, 4, 8 cover 2 / uneven / 1 / 0-or-1 cosets per rank and a non-power-of-two bucket split."""
import os
import random

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,log_n", [(2, 10), (3, 8), (4, 12), (8, 11), (5, 3)])
def test_sharded_prover_matches_single_rank_and_oracle(ctx, world, log_n):
    from oracle import coracle
    from typlonk_b200 import field as F, synthetic
    from typlonk_b200.ffi import Context
    n = 1 << log_n
    circuit = synthetic.mul_chain_direct(ctx, log_n)
    cols = [F.fr_vec_to_bytes(c) for c in synthetic.mul_chain_witness(n - 3, n)]
    single = circuit.handle.prove(cols, bytes(32 * n))
    tau_b, sel, perm, ocols, pi = coracle.mul_chain_inputs(log_n)
    oc = coracle.Circuit(tau_b, sel, perm, n)
    assert oc.prove(ocols, pi) == single
    oc.close()

    group = Context.multi([0] * world)
    assert group.group_size() == (world, False)
    gc = synthetic.mul_chain_direct(group, log_n)
    assert gc.handle.prove(cols, bytes(32 * n)) == single
    # the short public-input vector of prove()'s caller: three sliced columns + one whole 32-byte vector
    assert gc.handle.prove_inputs(cols, bytes(32)) == single
    assert gc.handle.verify(single, bytes(32))
    bad = bytearray(single)
    bad[192] ^= 1
    assert not gc.handle.verify(bytes(bad), bytes(32))
    assert group.launch_count() > 0
    group.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_commit_edge_scalars(ctx, world):
    """tp_commit / tp_open on a device group against the single-device result: uniform scalars, scalars below 2^16
    (every digit in the lowest window: with contiguous bucket ranges one rank would get all of them), one repeated
    scalar (a single bucket: all the work of a window on one rank), zeros, and r - 1 (top digit + signed carries)."""
    from typlonk_b200 import field as F
    from typlonk_b200.ffi import Context
    from oracle.pyoracle import rng
    n = 3000
    tau = F.fr_to_bytes(rng.fr_rand_stream(1, 1)[0])
    srs = ctx.srs_from_secret(tau, n - 3)
    group = Context.multi([0] * world)
    gsrs = group.srs_from_secret(tau, n - 3)
    assert gsrs.download(0, 5) == srs.download(0, 5) and len(gsrs) == len(srs)
    rnd = random.Random(7)
    cases = {
        "uniform": [rnd.randrange(F.R_MOD) for _ in range(n)],
        "below_2^16": [rnd.randrange(1 << 16) for _ in range(n)],
        "all_equal": [5] * n,
        "zeros": [0] * n,
        "r_minus_1_and_sparse": [F.R_MOD - 1 if i % 7 == 0 else 0 for i in range(n)],
        "short": [rnd.randrange(F.R_MOD) for _ in range(3)],
    }
    for name, sc in cases.items():
        raw = F.fr_vec_to_bytes(sc)
        assert group.commit(gsrs, raw) == ctx.commit(srs, raw), name
    z = F.fr_to_bytes(12345)
    assert group.open(gsrs, F.fr_vec_to_bytes(cases["uniform"]), z) == ctx.open(srs, F.fr_vec_to_bytes(cases["uniform"]), z)
    assert group.commit(gsrs, b"") == ctx.commit(srs, b"")
    gsrs.destroy()
    group.close()
    srs.destroy()


@pytest.mark.parametrize("n", [40, 3000, 70000])
def test_bucket_reduction_paths_agree(ctx, n):
    """The bucket reduction has two shapes -- with and without the level of running sums in front of the tree sums --
    chosen by the size of the bucket set; forced either way, on one device and on a 3-rank group, a commitment must
    not change (and equals the oracle's on the small size)."""
    from typlonk_b200 import field as F
    from typlonk_b200.ffi import Context
    from oracle.pyoracle import rng
    tau = rng.fr_rand_stream(1, 1)[0]
    srs = ctx.srs_from_secret(F.fr_to_bytes(tau), n - 3)
    rnd = random.Random(n)
    raw = F.fr_vec_to_bytes([rnd.randrange(F.R_MOD) for _ in range(n)])
    want = ctx.commit(srs, raw)
    if n <= 40:
        from oracle.pyoracle import curve
        pts, acc = [], curve.G1_GEN
        for _ in range(n):
            pts.append(acc)
            acc = curve.g1_mul(acc, tau)
        assert F.g1_from_abi(want) == curve.g1_msm(pts, F.fr_vec_from_bytes(raw))
    for mode in (1, 2):
        ctx.set_option("msm_reduce_l1", mode)
        try:
            assert ctx.commit(srs, raw) == want, "single device, msm_reduce_l1=%d" % mode
        finally:
            ctx.set_option("msm_reduce_l1", 0)
    group = Context.multi([0, 0, 0])
    gsrs = group.srs_from_secret(F.fr_to_bytes(tau), n - 3)
    for mode in (0, 1, 2):
        group.set_option("msm_reduce_l1", mode)
        assert group.commit(gsrs, raw) == want, "3-rank group, msm_reduce_l1=%d" % mode
    gsrs.destroy()
    group.close()
    srs.destroy()


def test_group_refuses_device_pointer_entry_points(ctx):
    from typlonk_b200.ffi import Context, TyplonkError
    group = Context.multi([0, 0])
    with pytest.raises(TyplonkError):
        group.ntt_dev(0x1000, 4)
    group.close()


def test_comm_init_rank_argument_checks():
    """One process per GPU: world = 1 needs no communicator; world > 1 needs rank 0's id; a device group owns its own."""
    from typlonk_b200.ffi import Context, TyplonkError
    c = Context(0)
    c.comm_init_rank(0, 1)
    assert c.group_size() == (1, False)
    with pytest.raises(TyplonkError):
        c.comm_init_rank(0, 2, None)
    with pytest.raises(TyplonkError):
        c.comm_init_rank(2, 2, bytes(128))
    c.close()
    g = Context.multi([0, 0])
    with pytest.raises(TyplonkError):
        g.comm_init_rank(0, 1)
    g.close()


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_msm_pipeline_same_bytes(ctx, world):
    """The MSM pipe (msm.cu "pipeline": sub-batches on three streams, lanes of scratch buffers, one exchange and one
    read-back per batch) against the one-stream path: same proof bytes, same commitments on the edge-case scalar
    vectors, with the staged accumulation kernel and without.  `msm_pipe_min_log` = 0 brings small circuits onto the
    path that only large ones take by default."""
    from typlonk_b200 import field as F, synthetic
    from typlonk_b200.ffi import Context
    from oracle.pyoracle import rng
    log_n = 11
    n = 1 << log_n
    plain = synthetic.mul_chain_direct(ctx, log_n)
    cols = [F.fr_vec_to_bytes(c) for c in synthetic.mul_chain_witness(n - 3, n)]
    want = plain.handle.prove(cols, bytes(32 * n))
    g = ctx if world == 1 else Context.multi([0] * world)
    try:
        circuit = plain if world == 1 else synthetic.mul_chain_direct(g, log_n)
        for staged in (0, 1):
            g.set_option("msm_acc_staged", staged)
            for mode in (0, 2):
                g.set_option("msm_pipeline", mode)
                g.set_option("msm_pipe_min_log", 0)
                for _ in range(2):   # twice: the lanes are reused
                    assert circuit.handle.prove(cols, bytes(32 * n)) == want, (world, staged, mode)
                assert circuit.handle.prove_inputs(cols, bytes(32)) == want
        # commitments: zeros (a job with no entries at all), one repeated scalar (one bucket), r - 1
        tau = F.fr_to_bytes(rng.fr_rand_stream(1, 1)[0])
        m = 3000
        srs = g.srs_from_secret(tau, m - 3)
        ref = ctx.srs_from_secret(tau, m - 3)
        rnd = random.Random(11)
        vecs = {"zeros": [0] * m, "same": [12345] * m, "top": [F.R_MOD - 1] * m,
                "uniform": [rnd.randrange(F.R_MOD) for _ in range(m)]}
        expect = {name: ctx.commit(ref, F.fr_vec_to_bytes(vec)) for name, vec in vecs.items()}
        g.set_option("msm_pipeline", 2)
        g.set_option("msm_pipe_min_log", 0)
        for name, vec in vecs.items():
            assert g.commit(srs, F.fr_vec_to_bytes(vec)) == expect[name], name
    finally:
        g.set_option("msm_pipeline", 0)
        g.set_option("msm_pipe_min_log", 15)
        g.set_option("msm_acc_staged", 0)
        if world != 1:
            g.close()
