"""ORACLE (test infrastructure only): O(n) verifier for large proofs.

Same checks as `verify` (plonk/src/proof.rs:195-236, 441-503) with two substitutions that keep it
linear-time in pure Python: (1) every pairing check e(W, [s - z]G2) == e(C - yG, G2) is replaced by
the equivalent G1 check with the known SRS secret, C - yG == (s - z) W; (2) the sigma evaluations
the reference obtains by interpolating (permutation/src/lib.rs:165-177) are computed with the
barycentric formula from the sigma table, and the sigma / selector commitments are passed in (the
reference recomputes the sigma commitments with three size-n MSMs, lib.rs:178-194)."""
from .curve import G1_GEN, g1_add, g1_deserialize_unchecked, g1_mul, g1_sub
from .fields import R_MOD, fr_inv, root_of_unity
from .permutation import cosets
from .rng import generate_challenges

M = R_MOD


def parse_fixed(raw: bytes):
    """1472-byte block -> dict of points / canonical scalars (layout: include/typlonk_b200.h)."""
    pos = 0

    def g1():
        nonlocal pos
        p = g1_deserialize_unchecked(raw[pos:pos + 96])
        pos += 96
        return p

    def fr():
        nonlocal pos
        v = int.from_bytes(raw[pos:pos + 32], "little")
        pos += 32
        return v
    out = {}
    for name in "abc":
        out[name] = (g1(), (g1(), fr()))
    out["z_commitment"] = g1()
    out["z"] = (g1(), fr())
    out["zw"] = (g1(), fr())
    out["evaluation_point"] = fr()
    out["t"] = [g1(), g1(), g1()]
    out["r"] = (g1(), fr())
    assert pos == 1472
    return out


def barycentric_eval(evals, n, omega, point):
    """p(point) for deg p < n from its evaluations on <omega>:  (x^n - 1)/n * sum e_j w^j / (x - w^j)."""
    roots = [1] * n
    for j in range(1, n):
        roots[j] = roots[j - 1] * omega % M
    dens = [(point - w) % M for w in roots]
    pre = [1] * n
    acc = 1
    for j in range(n):
        pre[j] = acc
        acc = acc * dens[j] % M
    inv = fr_inv(acc)
    total = 0
    for j in range(n - 1, -1, -1):
        di = inv * pre[j] % M
        inv = inv * dens[j] % M
        total = (total + evals[j] * roots[j] % M * di) % M
    return total * ((pow(point, n, M) - 1) % M) % M * fr_inv(n % M) % M


def verify_trapdoor(raw_fixed: bytes, n: int, tau: int, perm, fixed_commitments, sigma_commitments,
                    public_eval: int = 0) -> bool:
    """perm: flat permutation (index = j + i n); commitments: affine points."""
    p = parse_fixed(raw_fixed)
    omega = root_of_unity(n)
    ks = cosets(n)
    from .curve import g1_serialize_unchecked
    tr = b"".join(g1_serialize_unchecked(p[k][0]) for k in "abc")
    beta, gamma = generate_challenges(tr, 2)
    alpha, point = generate_challenges(tr + g1_serialize_unchecked(p["z_commitment"]), 2)
    if p["evaluation_point"] != point:
        return False

    def check(com, opening, z):
        w, y = opening
        return g1_sub(com, g1_mul(G1_GEN, y)) == g1_mul(w, (tau - z) % M)

    for k in "abc":
        if not check(p[k][0], p[k][1], point):
            return False
    if not check(p["z_commitment"], p["z"], point):
        return False
    if not check(p["z_commitment"], p["zw"], point * omega % M):
        return False
    roots = [1] * n
    for j in range(1, n):
        roots[j] = roots[j - 1] * omega % M
    sig_evals = []
    for i in range(2):
        col = [ks[perm[i * n + j] // n] * roots[perm[i * n + j] % n] % M for j in range(n)]
        sig_evals.append(barycentric_eval(col, n, omega, point))
    a, b, c = (p[k][1][1] for k in "abc")
    z_w = p["zw"][1]
    q_l, q_r, q_o, q_m, q_c = fixed_commitments
    line1 = g1_add(g1_add(g1_sub(g1_add(g1_mul(q_l, a), g1_mul(q_r, b)), g1_mul(q_o, c)), g1_mul(q_m, a * b % M)), q_c)
    l2 = 1
    for k, ev in zip(ks, (a, b, c)):
        l2 = l2 * ((ev + beta * k % M * point + gamma) % M) % M
    zn = pow(point, n, M)
    l0 = (zn - 1) * fr_inv(n * (point - 1) % M) % M
    line2 = g1_mul(p["z_commitment"], (l2 * alpha + l0 * alpha * alpha) % M)
    l3 = (a + beta * sig_evals[0] + gamma) % M * ((b + beta * sig_evals[1] + gamma) % M) % M
    line3 = g1_mul(sigma_commitments[2], l3 * alpha % M * beta % M * z_w % M)
    q_com = None
    for i, tc in enumerate(p["t"]):
        q_com = g1_add(q_com, g1_mul(tc, pow(point, n * i, M)))
    line5 = g1_mul(q_com, (zn - 1) % M)
    constant = (alpha * l3 % M * ((c + gamma) % M) % M * z_w + l0 * alpha * alpha + public_eval) % M
    r_com = g1_sub(g1_add(line1, g1_sub(line2, g1_add(line3, g1_mul(G1_GEN, constant)))), line5)
    return check(r_com, p["r"], point) and p["r"][1] == 0
