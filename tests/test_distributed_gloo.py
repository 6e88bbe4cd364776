"""CPU, world_size 2 over gloo: the multi-GPU plan of SURVEY.md 8(e) -- every rank computes the MSM
of its contiguous point range, partial points are all-gathered (144 B each) and summed -- checked
with the oracle standing in for the per-rank device MSM.  Exercises the same allgather callback
shape the C ABI uses (bytes in, world x bytes out) and the shard arithmetic of msm_dev."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard(length, rank, world):
    per = (length + world - 1) // world          # typlonk_b200/csrc/msm.cu: msm_dev
    first = per * rank
    cnt = 0 if first >= length else min(per, length - first)
    return first, cnt


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import curve, fields, rng
    tau = rng.fr_rand_stream(1, 1)[0]
    pts = []
    acc = curve.G1_GEN
    for _ in range(n):
        pts.append(acc)
        acc = curve.g1_mul(acc, tau)
    scalars = rng.fr_rand_stream(3, n)
    first, cnt = _shard(n, rank, world)
    part = curve.g1_msm(pts[first:first + cnt], scalars[first:first + cnt])

    def allgather(data: bytes) -> bytes:            # the callback shape of Context.set_shard
        send = torch.frombuffer(bytearray(data), dtype=torch.uint8)
        recv = torch.empty(world * len(data), dtype=torch.uint8)
        dist.all_gather_into_tensor(recv, send)
        return recv.numpy().tobytes()

    # partial point as 144-byte Jacobian (x, y, z) Montgomery, z = 0 for the identity
    if part is None:
        send = fields.fq_mont_bytes(1) * 2 + bytes(48)
    else:
        send = fields.fq_mont_bytes(part[0]) + fields.fq_mont_bytes(part[1]) + fields.fq_mont_bytes(1)
    recv = allgather(send)
    total = None
    for r in range(world):
        blob = recv[144 * r:144 * (r + 1)]
        if blob[96:] == bytes(48):
            continue
        total = curve.g1_add(total, (fields.fq_from_mont_bytes(blob[:48]), fields.fq_from_mont_bytes(blob[48:96])))
    full = curve.g1_msm(pts, scalars)
    q.put((rank, total == full))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 64])
def test_sharded_msm_combines_to_full_msm(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + n
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_shard_ranges_cover_exactly():
    for length in (0, 1, 7, 8, 9, 1 << 20, (1 << 20) - 1):
        for world in (1, 2, 4, 8):
            spans = [_shard(length, r, world) for r in range(world)]
            covered = sum(c for _, c in spans)
            assert covered == length
            pos = 0
            for first, cnt in spans:
                if cnt:
                    assert first == pos
                    pos += cnt


# ---- quotient sharded by coset (api.cu prove_resident, tp_ctx_set_broadcast) ------------------------------
def _coset_owner(k, world):
    return k % world if world < 4 else k * (world // 4)   # typlonk_b200/csrc/api.cu: owner(k)


def _coset_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    from oracle.pyoracle import fields, poly
    M = fields.R_MOD
    rnd = random.Random(5)                      # same numerator polynomial on every rank
    N = [rnd.randrange(M) for _ in range(4 * n - 1)]
    dom = poly.Domain(n)
    w4n = fields.root_of_unity(4 * n)
    slots = [torch.zeros(n * 32, dtype=torch.uint8) for _ in range(4)]
    for k in range(4):
        if _coset_owner(k, world) != rank:
            continue
        g = pow(w4n, k, M)
        evals = [poly.evaluate(N, g * dom.element(i) % M) for i in range(n)]
        coeffs = dom.ifft(evals)
        ginv = pow(g, M - 2, M)
        ck = [coeffs[i] * pow(ginv, i, M) % M for i in range(n)]
        slots[k] = torch.frombuffer(bytearray(b"".join(v.to_bytes(32, "little") for v in ck)), dtype=torch.uint8)
    for k in range(4):                          # the broadcast callback shape: (buffer, bytes, root)
        dist.broadcast(slots[k], src=_coset_owner(k, world))
    C = [[int.from_bytes(slots[k].numpy().tobytes()[32 * i:32 * i + 32], "little") for i in range(n)] for k in range(4)]
    iota = pow(w4n, n, M)
    s, quarter = (M - iota) % M, pow(4, M - 2, M)
    Nj = [[quarter * sum(pow(s, k * j, M) * C[k][i] for k in range(4)) % M for i in range(n)] for j in range(4)]
    t2 = Nj[3]
    t1 = [(a + b) % M for a, b in zip(Nj[2], t2)]
    t0 = [(a + b) % M for a, b in zip(Nj[1], t1)]
    quo, _ = poly.divide_by_vanishing_poly(N, n)
    q.put((rank, poly.strip(t0 + t1 + t2) == poly.strip(list(quo))))
    dist.destroy_process_group()


def test_coset_sharded_quotient_recombines():
    world, n = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_coset_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_coset_owner_map():
    for world in (1, 2, 3, 4, 8):
        owners = [_coset_owner(k, world) for k in range(4)]
        assert all(0 <= o < world for o in owners)
        if world >= 4:
            assert len(set(owners)) == 4        # one coset per rank, the other ranks idle in this phase
