"""CPU: pins the Python spec oracle against external known answers and the committed golden
fixtures (the reference ships no golden vectors; SURVEY.md App. B lists the known answers)."""
import hashlib
import json
import os

from oracle.pyoracle import builder, curve, fields, kzg, permutation, plonk, poly, rng

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KNOWN = json.load(open(os.path.join(GOLD, "known_answers.json")))
PROOFS = json.load(open(os.path.join(GOLD, "proofs.json")))
R = fields.R_MOD


def test_field_constants():
    # SURVEY.md App. A.1 (ark-bls12-381 0.3.0 FrParameters / FqParameters)
    assert fields.FR_R == 0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe
    assert fields.FQ_R == 0x15f65ec3fa80e4935c071a97a256ec6d77ce5853705257455f48985753c758baebf4000bc40c0002760900000002fffd
    assert fields.FR_ROOT_OF_UNITY == 0x16a2a19edfe81f20d09b681922c813b4b63683508c2280b93829971f439f0d2b
    assert pow(fields.FR_ROOT_OF_UNITY, 1 << 32, R) == 1 and pow(fields.FR_ROOT_OF_UNITY, 1 << 31, R) != 1
    assert hex(fields.root_of_unity(8)) == KNOWN["omega_8"] == "0x345766f603fa66e78c0625cd70d77ce2b38b21c28713b7007228fd3397743f7a"


def test_curve_known_answers():
    assert curve.g1_is_on_curve(curve.G1_GEN) and curve.g2_is_on_curve(curve.G2_GEN)
    assert curve.g1_mul(curve.G1_GEN, R) is None
    p17 = curve.g1_mul(curve.G1_GEN, 17)  # kzg/src/lib.rs:96-105 expects commit == 17 G
    assert [hex(c) for c in p17] == KNOWN["g1_17"]
    assert p17[0] == 0x1098f178f84fc753a76bb63709e9be91eec3ff5f7f3a5f4836f34fe8a1a6d6c5578d8fd820573cef3a01e2bfef3eaf3a
    assert curve.g1_deserialize_unchecked(curve.g1_serialize_unchecked(p17)) == p17
    assert curve.g1_serialize_unchecked(None)[-1] == 0x40


def test_chacha_and_seed_expansion_vectors():
    blk = bytes(b for w in rng.chacha_block(bytes(32), 0, 0, 12) for b in w.to_bytes(4, "little"))
    assert blk.hex() == KNOWN["chacha12_zero_block0"]
    assert blk.hex().startswith("9bf49a6a0755f953811fce125f2683d5")
    c20 = bytes(b for w in rng.chacha_block(bytes(32), 0, 0, 20) for b in w.to_bytes(4, "little"))
    assert c20.hex().startswith("76b8e0ada0f13d90405d6ae55386bd28")  # RFC 7539 zero key/nonce block
    c8 = bytes(b for w in rng.chacha_block(bytes(32), 0, 0, 8) for b in w.to_bytes(4, "little"))
    assert c8.hex().startswith("3e00ef2f895f40d6")
    assert rng.seed_from_u64_key(0).hex() == KNOWN["seed_from_u64_0_key"]
    assert hashlib.blake2b(b"").hexdigest().startswith("786a02f742015903")


def test_rand_crates_own_value_stability_constants():
    """Pins taken from the pinned crates' OWN test suites (Cargo.lock: rand 0.8.4, rand_core 0.6.3, rand_chacha 0.3.1),
    i.e. known answers that do not come from this repository:
      * rand_core 0.6 `test_seed_from_u64`: the first 8 bytes `seed_from_u64(0)` expands to, as a little-endian u64,
        are 5029875928683246316  -> pins the PCG32 seed expansion of challenges.rs:38;
      * rand 0.8 `rngs::std::test::test_stdrng_construction`: StdRng::from_seed([1,0,0,0, 23,0,0,0, 200,1,0,0,
        210,30,0,0, 0...]) yields next_u64 = 10719222850664546238, and StdRng::from_rng of it then yields
        14064965282130556830  -> pins ChaCha12, rand_chacha's counter / stream word layout, the order output words are
        consumed in, next_u64 = lo | hi << 32 and fill_bytes;
      * ChaCha20 block 0 of the zero key (RFC 7539 / rand_chacha `test_chacha_true_values_a`): 0xade0b876, 0x903df1a0 ..."""
    import struct
    assert int.from_bytes(rng.seed_from_u64_key(0)[:8], "little") == 5029875928683246316
    r0 = rng.StdRng(bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16))
    assert r0.next_u64() == 10719222850664546238
    r1 = rng.StdRng(b"".join(struct.pack("<I", r0.next_u32()) for _ in range(8)))
    assert r1.next_u64() == 14064965282130556830
    assert rng.chacha_block(bytes(32), 0, 0, 20)[:4] == [0xade0b876, 0x903df1a0, 0xe56a5d40, 0x28bd8653]


def test_challenge_chain_golden():
    tr = curve.g1_serialize_unchecked(curve.g1_mul(curve.G1_GEN, 17)) * 3
    assert rng.challenge_seed(tr) == KNOWN["challenge_chain_17G_x3"]["seed"] == 13037422643194131432
    ch = rng.generate_challenges(tr, 2)
    assert [hex(fields.fr_to_mont(x)) for x in ch] == KNOWN["challenge_chain_17G_x3"]["challenges_mont"]


def test_pairing_bilinearity():
    a, b = 5, 7
    lhs = curve.pairing(curve.g1_mul(curve.G1_GEN, a), curve.g2_mul(curve.G2_GEN, b))
    rhs = curve.pairing(curve.G1_GEN, curve.G2_GEN).pow(a * b)
    assert lhs == rhs and not (lhs == curve.Fq12.one())


def test_reference_kzg_commit_test():
    """kzg/src/lib.rs:95-109."""
    srs = kzg.Srs.from_secret(2, 10)
    p = [1, 2, 3]
    com = kzg.commit(srs, p)
    assert com == curve.g1_mul(curve.G1_GEN, poly.evaluate(p, 2)) == curve.g1_mul(curve.G1_GEN, 17)
    assert poly.evaluate(p, 1) == 6
    opening = kzg.open_at(srs, p, 1)
    assert kzg.verify(srs, com, opening, 1)
    assert kzg.verify_trapdoor(srs, com, opening, 1)
    assert not kzg.verify_trapdoor(srs, com, (opening[0], 7), 1)


def test_reference_kzg_scalar_mul_test():
    """kzg/src/lib.rs:160-171."""
    srs = kzg.Srs.from_secret(rng.fr_rand_stream(1, 1)[0], 5)
    p = [1, 2, 3, 4, 5]
    assert curve.g1_mul(kzg.commit(srs, p), 9) == kzg.commit(srs, poly.scale(p, 9))


def test_reference_l0_test():
    """plonk/src/utils.rs:161-177 at a smaller domain: sum of L0 over the domain == 1."""
    n = 256
    l0 = plonk.l0_poly(n)
    assert l0 == [pow(n, -1, R)] * n
    assert sum(poly.Domain(n).fft(l0)) % R == 1


def test_reference_slicing_tests():
    """plonk/src/utils.rs:128-148,179-201: compact(point) evaluates like the whole polynomial and
    commit(compact) == compact_commitment(slice commitments)."""
    srs = kzg.Srs.from_secret(rng.fr_rand_stream(1, 1)[0], 8)
    p = list(range(1, 9))
    point = 4
    slices = [p[0:3], p[3:6], p[6:8]]
    comp = plonk.compact(slices, 3, point)
    assert poly.evaluate(comp, point) == poly.evaluate(p, point)
    lhs = kzg.commit(srs, comp)
    rhs = None
    for i, s in enumerate(slices):
        rhs = curve.g1_add(rhs, curve.g1_mul(kzg.commit(srs, s), pow(point, 3 * i, R)))
    assert lhs == rhs


def test_structural_goldens_appendix_c():
    gates, perm = builder.trace(builder.circuit_pythagoras, 3)
    assert gates == ["Mul", "Mul", "Mul", "Add"] + ["Dummy"] * 4
    rows = [[perm.perm[j + i * 8] for i in range(3)] for j in range(8)]
    assert rows == [[8, 0, 3], [9, 1, 11], [10, 2, 19], [16, 17, 18], [4, 12, 20], [5, 13, 21], [6, 14, 22], [7, 15, 23]]
    gates, perm = builder.trace(builder.make_mul_chain(5), 2)
    rows = [[perm.perm[j + i * 8] for i in range(3)] for j in range(8)]
    assert rows[:5] == [[0, 12, 1], [16, 8, 2], [17, 9, 3], [18, 10, 4], [19, 11, 20]]
    assert permutation.cosets(8) == [2, 3, 4] and permutation.cosets(1 << 20) == [2, 3, 4]


def test_golden_proofs_reproduce_and_verify():
    tau = rng.fr_rand_stream(1, 1)[0]
    blinders = rng.fr_rand_stream(2, 9)
    assert hex(fields.fr_to_mont(tau)) == KNOWN["tau_seed1_mont"]
    cases = {"readme_pythagoras_3_4_5": (builder.circuit_pythagoras, 3, [3, 4, 5]),
             "readme_pythagoras_bad_3_4_6": (builder.circuit_pythagoras, 3, [3, 4, 6]),
             "additive_2_7_2_3_4": (builder.circuit_additive, 5, [2, 7, 2, 3, 4]),
             "mulchain_13_gates": (builder.make_mul_chain(13), 2, [3, 5])}
    for name, (run, nin, inputs) in cases.items():
        c = builder.compile_circuit(run, nin, tau)
        p = plonk.prove(c, inputs, [0], blinders)
        assert p.to_bytes().hex() == PROOFS[name]["proof_hex"], name
        assert plonk.verify(c, p, use_trapdoor=True) == PROOFS[name]["verifies"], name


def test_readme_circuit_literal_prover_and_pairing_verify():
    """builder/test.rs:25-37: README circuit verifies (real pairings), bad inputs do not; the
    schoolbook (`naive_mul`) prover and the NTT prover produce identical bytes."""
    tau = rng.fr_rand_stream(1, 1)[0]
    blinders = rng.fr_rand_stream(2, 9)
    c = builder.compile_circuit(builder.circuit_pythagoras, 3, tau)
    lit = plonk.prove(c, [3, 4, 5], [0], blinders, literal=True)
    fast = plonk.prove(c, [3, 4, 5], [0], blinders, literal=False)
    assert lit.to_bytes() == fast.to_bytes() == bytes.fromhex(PROOFS["readme_pythagoras_3_4_5"]["proof_hex"])
    assert plonk.verify(c, lit)
    bad = plonk.prove(c, [3, 4, 6], [0], blinders)
    assert not plonk.verify(c, bad, use_trapdoor=True)


def test_fast_trapdoor_verifier_agrees_with_verify():
    from oracle.pyoracle import fastverify
    tau = rng.fr_rand_stream(1, 1)[0]
    blinders = rng.fr_rand_stream(2, 9)
    for run, nin, good, bad in ((builder.make_mul_chain(13), 2, [3, 5], None),
                                (builder.circuit_pythagoras, 3, [3, 4, 5], [3, 4, 6])):
        c = builder.compile_circuit(run, nin, tau)
        _, perm = builder.trace(run, nin)
        sig = c.copy_constrains.sigma_commitments(c.srs, c.domain)
        p = plonk.prove(c, good, [0], blinders)
        assert fastverify.verify_trapdoor(p.to_bytes()[:1472], c.rows, tau, perm.perm, c.fixed_commitments, sig)
        corrupt = bytearray(p.to_bytes()[:1472])
        corrupt[700] ^= 1
        assert not fastverify.verify_trapdoor(bytes(corrupt), c.rows, tau, perm.perm, c.fixed_commitments, sig)
        if bad:
            pb = plonk.prove(c, bad, [0], blinders)
            assert not fastverify.verify_trapdoor(pb.to_bytes()[:1472], c.rows, tau, perm.perm, c.fixed_commitments, sig)
