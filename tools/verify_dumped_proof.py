#!/usr/bin/env python3
"""Verify a proof dumped by `bench.py --dump-proof` (mul-chain circuit, n = 2^log_n) on the CPU with the
oracle's O(n) trapdoor verifier.  The selector / sigma commitments the verifier needs are rebuilt from the
known SRS secret (commit(p) = p(tau) G, p(tau) by the barycentric formula), so no GPU is involved:

    python tools/verify_dumped_proof.py gpurun_out/proof_2p22_n8.bin 22
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import fastverify, permutation as operm, rng  # noqa: E402
from oracle.pyoracle.curve import G1_GEN, g1_mul  # noqa: E402
from oracle.pyoracle.fields import R_MOD, root_of_unity  # noqa: E402

path, log_n = sys.argv[1], int(sys.argv[2])
n = 1 << log_n
gates = n - 3
t0 = time.time()
proof = open(path, "rb").read()
tau = rng.fr_rand_stream(1, 1)[0]
pb = operm.PermutationBuilder.with_rows(gates)
for j in range(1, gates):
    pb.add_constrain((2, j - 1), (0, j))
    pb.add_constrain((1, 0), (1, j))
perm = pb.build(n).perm
omega = root_of_unity(n)
ks = operm.cosets(n)
print("structure built in %.0f s" % (time.time() - t0), flush=True)
sel_on = [1] * gates + [0] * (n - gates)
q_on = g1_mul(G1_GEN, fastverify.barycentric_eval(sel_on, n, omega, tau))
fixed = [None, None, q_on, q_on, None]           # Mul rows = [0, 0, 1, 1, 0]
roots = [1] * n
for j in range(1, n):
    roots[j] = roots[j - 1] * omega % R_MOD
sig3 = [ks[perm[2 * n + j] // n] * roots[perm[2 * n + j] % n] % R_MOD for j in range(n)]
sigma = [None, None, g1_mul(G1_GEN, fastverify.barycentric_eval(sig3, n, omega, tau))]
print("commitments rebuilt in %.0f s" % (time.time() - t0), flush=True)
ok = fastverify.verify_trapdoor(proof, n, tau, perm, fixed, sigma)
bad = bytearray(proof)
bad[96 * 2 + 5] ^= 1
rejected = not fastverify.verify_trapdoor(bytes(bad), n, tau, perm, fixed, sigma) if log_n <= 16 else None
print("verify(%s, n=2^%d) = %s   corrupted copy rejected = %s   (%.0f s)" % (path, log_n, ok, rejected, time.time() - t0))
sys.exit(0 if ok else 1)
