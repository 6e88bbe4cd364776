"""CPU, world_size 2 over gloo: the multi-GPU plan of SURVEY.md 8(e) as csrc/msm.cu runs it -- every rank extracts the
signed window digits of ALL scalars, keeps the buckets it owns (bucket b belongs to rank b % world as local bucket
b // world), reduces them to P = sum_j B_j and F = sum_j j B_j, the ranks' (P, F) are all-gathered, and every rank
finishes with  sum_r [ world * F_r + (r + 1) * P_r ]  -- checked with the oracle's curve arithmetic standing in for the
device kernels, and the quotient's coset split with its one broadcast per coset.  The arithmetic identities the CUDA
path relies on are what is under test here; the kernels themselves are compared with these results on the GPU
(tests/test_gpu_multirank.py)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _signed_digits(s, c, nwin):
    """typlonk_b200/csrc/msm.cu k_msm_digits: window digits in [-2^(c-1), 2^(c-1)] with carry."""
    out, carry = [], 0
    for w in range(nwin):
        raw = ((s >> (w * c)) & ((1 << c) - 1)) + carry
        carry = 0
        if raw > (1 << (c - 1)):
            raw -= 1 << c
            carry = 1
        out.append(raw)
    assert carry == 0
    return out


def _windows_for(c):
    nwin = (255 + c - 1) // c
    if 255 - (nwin - 1) * c == c:
        nwin += 1
    return nwin


def _rank_partial(pts, scalars, c, rank, world):
    """(P, F) of the buckets rank `rank` owns; the fixed-base table level 2^(c w) P_i is a scalar multiple here."""
    from oracle.pyoracle import curve
    nwin = _windows_for(c)
    buckets = {}
    for p, s in zip(pts, scalars):
        for w, d in enumerate(_signed_digits(s, c, nwin)):
            if d == 0:
                continue
            b = abs(d) - 1
            if b % world != rank:
                continue
            t = curve.g1_mul(p, 1 << (c * w))          # reduced modulo r inside (the points have order r)
            if d < 0:
                t = curve.g1_neg(t)
            buckets[b // world] = curve.g1_add(buckets.get(b // world), t)
    P = F = None
    for j, pt in buckets.items():
        P = curve.g1_add(P, pt)
        F = curve.g1_add(F, curve.g1_mul(pt, j) if j else None)
    return P, F


def _worker(rank, world, port, n, c, q):
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
    scalars[0] = fields.R_MOD - 1                     # top window + carries
    if n > 3:
        scalars[1], scalars[2], scalars[3] = 0, 5, 5  # zero digit rows, a repeated bucket
    P, F = _rank_partial(pts, scalars, c, rank, world)

    def enc(pt):   # 144-byte Jacobian (x, y, z) Montgomery, z = 0 for the identity
        if pt is None:
            return fields.fq_mont_bytes(1) * 2 + bytes(48)
        return fields.fq_mont_bytes(pt[0]) + fields.fq_mont_bytes(pt[1]) + fields.fq_mont_bytes(1)

    def dec(blob):
        if blob[96:] == bytes(48):
            return None
        return (fields.fq_from_mont_bytes(blob[:48]), fields.fq_from_mont_bytes(blob[48:96]))

    send = torch.frombuffer(bytearray(enc(P) + enc(F)), dtype=torch.uint8)
    recv = torch.empty(world * 288, dtype=torch.uint8)
    dist.all_gather_into_tensor(recv, send)
    raw = recv.numpy().tobytes()
    # the host tail of msm_local: F slots summed over ranks, the P points weighted by running sums
    fsum = None
    for r in range(world):
        fsum = curve.g1_add(fsum, dec(raw[288 * r + 144:288 * r + 288]))
    run = pw = None
    for r in reversed(range(world)):
        run = curve.g1_add(run, dec(raw[288 * r:288 * r + 144]))
        pw = curve.g1_add(pw, run)
    total = curve.g1_add(curve.g1_mul(fsum, world) if fsum else None, pw)
    q.put((rank, total == curve.g1_msm(pts, scalars)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,c", [(5, 4), (24, 7)])
def test_bucket_sharded_msm_combines_to_full_msm(n, c):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + n
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, c, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_bucket_ownership_covers_every_digit_once():
    import random
    rnd = random.Random(3)
    for c in (3, 8, 20):
        nwin = _windows_for(c)
        for world in (1, 2, 3, 8):
            for _ in range(50):
                s = rnd.randrange(0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001)
                digits = _signed_digits(s, c, nwin)
                assert sum(d << (c * w) for w, d in enumerate(digits)) == s
                owners = [(abs(d) - 1) % world for d in digits if d]
                assert all(0 <= o < world for o in owners)
                # local index and weight: global bucket b counts b + 1 = world * (b // world) + (b % world) + 1
                for d in digits:
                    if d:
                        b = abs(d) - 1
                        assert world * (b // world) + (b % world) + 1 == abs(d)


def _sharded_digit_pass(s, c, nwin, world, rank):
    """k_msm_digits_sharded (msm.cu) restated: ONE walk over the windows records which digits this rank owns and the
    carry that entered each of them (two bit masks), the emit pass revisits the owned windows only and recomputes
    each digit from its c-bit field and the saved carry.  Ownership: bucket b = |digit| - 1 belongs to rank b % world as
    local bucket b // world -- a mask and a shift when world is a power of two.  Returns [(window, local bucket, neg)]."""
    half, full = 1 << (c - 1), 1 << c
    pow2 = world & (world - 1) == 0
    shift = world.bit_length() - 1

    def field(w):
        return (s >> (w * c)) & (full - 1) if w * c < 256 else 0

    def mine(b):
        return (b & (world - 1)) == rank if pow2 else b % world == rank

    own_mask = carry_mask = carry = 0
    for w in range(nwin):
        raw, cin = field(w) + carry, carry
        carry = 1 if raw > half else 0
        if raw == 0 or raw == full:
            continue
        mag = full - raw if carry else raw
        if mine(mag - 1):
            own_mask |= 1 << w
            carry_mask |= cin << w
    out = []
    while own_mask:
        w = (own_mask & -own_mask).bit_length() - 1
        own_mask &= own_mask - 1
        raw = field(w) + ((carry_mask >> w) & 1)
        neg = raw > half
        mag = full - raw if neg else raw
        out.append((w, (mag - 1) >> shift if pow2 else (mag - 1) // world, neg))
    return out


def test_sharded_digit_pass_partitions_the_signed_digits():
    """Every non-zero signed digit of a scalar is emitted by exactly one rank, with the sign and the local bucket that
    map back to it; power-of-two and general rank counts, window sizes with and without a carry window on top."""
    import random
    rnd = random.Random(9)
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    for c in (8, 15, 16, 20):
        nwin = _windows_for(c)
        assert nwin <= 32   # the masks of the kernel's fast path
        for world in (2, 3, 4, 5, 8):
            for s in [0, 1, R - 1, (1 << 255) - 1 & (R - 1)] + [rnd.randrange(R) for _ in range(40)]:
                want = {w: d for w, d in enumerate(_signed_digits(s, c, nwin)) if d}
                got = {}
                for rank in range(world):
                    for w, local, neg in _sharded_digit_pass(s, c, nwin, world, rank):
                        assert w not in got, "a digit emitted by two ranks"
                        mag = local * world + rank + 1
                        got[w] = -mag if neg else mag
                assert got == want, (c, world, hex(s))


# ---- quotient sharded by coset (api.cu prove_resident, comm_bcast) ------------------------------
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
