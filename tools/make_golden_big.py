#!/usr/bin/env python3
"""Golden proofs of the bench workload (mul-chain circuit, tau = seed 1, blinders = seed 2,
inputs [3, 5], public inputs 0) at n = 2^16 ... 2^22, computed by the CPU oracle's C++ port (oracle/c) alone --
no product code runs here.  bench.py and tests/test_gpu_prove.py compare the GPU prover's bytes with these.

    python tools/make_golden_big.py [log_n ...]        (2^20: minutes, 2^22: tens of minutes on 8 cores)
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import coracle  # noqa: E402

path = os.path.join(ROOT, "tests", "golden", "mulchain_big.json")
have = json.load(open(path)) if os.path.exists(path) else {}
for log_n in [int(a) for a in sys.argv[1:]] or [16, 18, 20, 22]:
    t0 = time.time()
    tau, sel, perm, cols, pi = coracle.mul_chain_inputs(log_n)
    c = coracle.Circuit(tau, sel, perm, 1 << log_n)
    t1 = time.time()
    proof = c.prove(cols, pi)
    c.close()
    have["2^%d" % log_n] = {"log_n": log_n, "gates": (1 << log_n) - 3, "proof_hex": proof.hex(),
                            "sha256": hashlib.sha256(proof).hexdigest(), "by": "oracle/c/oracle.cpp",
                            "threads": coracle.threads(), "setup_s": round(t1 - t0, 1), "prove_s": round(time.time() - t1, 1)}
    json.dump(have, open(path, "w"), indent=1, sort_keys=True)
    print("2^%d: %s  (setup %.0f s, prove %.0f s)" % (log_n, have["2^%d" % log_n]["sha256"][:16], t1 - t0, time.time() - t1), flush=True)
