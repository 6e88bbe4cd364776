#!/usr/bin/env python3
"""Generate tests/golden/*.json from the Python spec oracle (the reference ships no golden vectors
and cannot be built here -- no Rust toolchain -- so the goldens pin OUR oracle: any later change
to the oracle, the C++ port or the CUDA path that alters a byte shows up against these files).
    python tools/make_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import builder, curve, fields, kzg, plonk, poly, rng  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
tau = rng.fr_rand_stream(1, 1)[0]
blinders = rng.fr_rand_stream(2, 9)

known = {
    "tau_seed1_mont": hex(fields.fr_to_mont(tau)),
    "blinders_seed2_mont": [hex(fields.fr_to_mont(b)) for b in blinders],
    "g1_17": [hex(c) for c in curve.g1_mul(curve.G1_GEN, 17)],
    "chacha12_zero_block0": bytes(b for w in rng.chacha_block(bytes(32), 0, 0, 12) for b in w.to_bytes(4, "little")).hex(),
    "seed_from_u64_0_key": rng.seed_from_u64_key(0).hex(),
    "omega_8": hex(fields.root_of_unity(8)),
    "root_of_unity_2_32": hex(fields.FR_ROOT_OF_UNITY),
    "challenge_chain_17G_x3": {
        "seed": rng.challenge_seed(curve.g1_serialize_unchecked(curve.g1_mul(curve.G1_GEN, 17)) * 3),
        "challenges_mont": [hex(fields.fr_to_mont(x)) for x in
                            rng.generate_challenges(curve.g1_serialize_unchecked(curve.g1_mul(curve.G1_GEN, 17)) * 3, 2)],
    },
    "ntt8_of_1_to_8": [hex(x) for x in poly.Domain(8).fft(list(range(1, 9)))],
}
json.dump(known, open(os.path.join(out, "known_answers.json"), "w"), indent=1)

proofs = {}
for name, run, nin, inputs in (("readme_pythagoras_3_4_5", builder.circuit_pythagoras, 3, [3, 4, 5]),
                               ("readme_pythagoras_bad_3_4_6", builder.circuit_pythagoras, 3, [3, 4, 6]),
                               ("additive_2_7_2_3_4", builder.circuit_additive, 5, [2, 7, 2, 3, 4]),
                               ("mulchain_13_gates", builder.make_mul_chain(13), 2, [3, 5]),
                               ("mulchain_61_gates", builder.make_mul_chain(61), 2, [3, 5])):
    c = builder.compile_circuit(run, nin, tau)
    p = plonk.prove(c, inputs, [0], blinders)
    proofs[name] = {"rows": c.rows, "perm": c.copy_constrains and None, "proof_hex": p.to_bytes().hex(),
                    "fixed_commitments": [curve.g1_serialize_unchecked(x).hex() for x in c.fixed_commitments],
                    "verifies": plonk.verify(c, p, use_trapdoor=True)}
    proofs[name].pop("perm")
json.dump(proofs, open(os.path.join(out, "proofs.json"), "w"), indent=1)
print("wrote", out)
