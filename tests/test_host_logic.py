"""CPU: host-side logic of the product package (tracing DSL mirror, permutation builder, synthetic
circuit generator, proof framing) against the oracle, and the C-ABI library's exported surface.
No compute calls -- there is no GPU here and the library has no CPU fallback."""
import os
import sys
import re

import pytest

from oracle.pyoracle import builder as obuilder, fields, permutation as operm, rng
from typlonk_b200 import field as F, ffi, synthetic
from typlonk_b200.permutation import PermutationBuilder
import py_tracer
from typlonk_b200.plonk import CircuitDescription, Proof

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Pythagoras(CircuitDescription):
    INPUTS = 3

    @staticmethod
    def run(inputs):
        a, b, c = inputs
        a = a.clone() * a
        b = b.clone() * b
        c = c.clone() * c
        d = a + b
        d.assert_eq(c)


def test_tracer_matches_oracle_tracer():
    gates, perm = py_tracer.trace(Pythagoras)
    ogates, operm_ = obuilder.trace(obuilder.circuit_pythagoras, 3)
    assert gates == ogates and perm.perm == operm_.perm
    for g in (1, 2, 5, 13, 29):
        gates, perm = py_tracer.trace(synthetic.mul_chain_description(g))
        ogates, operm_ = obuilder.trace(obuilder.make_mul_chain(g), 2)
        assert gates == ogates and perm.perm == operm_.perm


def test_direct_mul_chain_structure_equals_traced():
    for g in (1, 5, 13, 61, 125):
        gates, perm = synthetic.mul_chain_structure(g)
        tgates, tperm = py_tracer.trace(synthetic.mul_chain_description(g))
        assert gates == tgates and perm.perm == tperm.perm
    n = 64
    cols = synthetic.mul_chain_witness(61, n, blind=list(range(1, 10)))
    ocols = obuilder.witness_columns(obuilder.make_mul_chain(61), [3, 5], n, list(range(1, 10)))
    assert cols == ocols


def test_permutation_builder_invalid_tag():
    pb = PermutationBuilder.with_rows(4)
    assert pb.add_constrain((0, 0), (1, 3))
    assert not pb.add_constrain((0, 0), (1, 4))      # j out of range -> Err(()) in the reference
    assert pb.add_constrain((3, 0), (0, 0))          # sic: `i <= C` (permutation/src/lib.rs:46)
    with pytest.raises(ValueError):
        pb.add_constrains([((0, 0), (9, 9))])
    opb = operm.PermutationBuilder.with_rows(4)
    assert not opb.add_constrain((0, 0), (1, 4))


def test_field_helpers_round_trip():
    xs = rng.fr_rand_stream(8, 20) + [0, 1, fields.R_MOD - 1]
    assert F.fr_vec_from_bytes(F.fr_vec_to_bytes(xs)) == xs
    assert F.fr_vec_to_bytes(xs) == fields.fr_vec_to_mont_bytes(xs)
    assert F.g1_from_packed(F.g1_to_packed(None)) is None
    assert F.g1_serialize_unchecked(None)[-1] == 0x40


def test_proof_framing():
    p = Proof(bytes(range(256)) * 5 + bytes(192), [0, 1, 2])
    raw = p.to_bytes()
    assert len(raw) == 1472 + 8 + 3 * 32 and raw[1472:1480] == (3).to_bytes(8, "little")


def test_library_loads_and_exports_every_declared_symbol():
    lib = ffi.lib()
    header = open(os.path.join(ROOT, "include", "typlonk_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|const char\*)\s+(tp_[a-z0-9_]+)\(", header, flags=re.M)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(ffi.SYMBOLS) == declared


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of falling back to a CPU path."""
    import ctypes
    h = ctypes.c_void_p()
    rc = ffi.lib().tp_ctx_create(0, None, ctypes.byref(h))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if not has_gpu:
        assert rc == 7  # TP_ERR_NO_DEVICE
        with pytest.raises(ffi.TyplonkError):
            ffi.Context(0)


def test_multi_gpu_entry_points_validate_their_arguments():
    """tp_ctx_create_multi / tp_ctx_comm_init_rank / tp_comm_unique_id / tp_ctx_group_size without a device: bad
    arguments are refused, and a device group fails like a single context does (no CPU fallback)."""
    import ctypes as C
    L = ffi.lib()
    h = C.c_void_p()
    assert L.tp_ctx_create_multi(None, 2, C.byref(h)) == 1                      # TP_ERR_INVALID_ARG
    devs = (C.c_int * 2)(0, 0)
    assert L.tp_ctx_create_multi(devs, 0, C.byref(h)) == 1
    assert L.tp_ctx_create_multi(devs, 17, C.byref(h)) == 1                     # more than TP_MAX_GROUP
    assert L.tp_ctx_create_multi(devs, 2, None) == 1
    assert L.tp_ctx_group_size(None, None, None) == 1
    assert L.tp_ctx_comm_init_rank(None, 0, 1, None) == 1
    assert L.tp_comm_unique_id(None) == 1
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if not has_gpu:
        assert L.tp_ctx_create_multi(devs, 2, C.byref(h)) == 7                  # TP_ERR_NO_DEVICE
        with pytest.raises(ffi.TyplonkError):
            ffi.Context.multi([0, 0])


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "typlonk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace(
                    "independent of oracle/", "").replace("Product code -- independent of oracle/.", ""), f


def test_host_fr_rand_stream_matches_oracle():
    assert ffi.fr_rand_stream(5, 12) == fields.fr_vec_to_mont_bytes(rng.fr_rand_stream(5, 12))
    assert synthetic.tau() == rng.fr_rand_stream(1, 1)[0]
    assert synthetic.blinders() == rng.fr_rand_stream(2, 9)


def test_safegcd_fq_inversion_host_build(tmp_path):
    """typlonk_b200/csrc/fq_inv.cuh is plain C++: build it for the host and compare the division-step
    inversion with Python's modular inverse (constants, the 30-round bound, the 0 -> 0 convention)."""
    import random
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "fq_inv_test"
    subprocess.run(["g++", "-O2", "-o", str(exe), os.path.join(root, "tools", "host_tests", "fq_inv_test.cpp")], check=True)
    q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
    rnd = random.Random(11)
    xs = [0, 1, 2, q - 1, q - 2, (q - 1) // 2, (q + 1) // 2, 1 << 380, (1 << 192) + 1]
    xs += [rnd.randrange(q) for _ in range(2000)] + [rnd.randrange(1 << k) for k in range(1, 381, 2)]
    out = subprocess.run([str(exe)], input="\n".join("%x" % x for x in xs), capture_output=True, text=True,
                         check=True).stdout.split()
    assert len(out) == len(xs)
    for x, o in zip(xs, out):
        assert o != "FAIL"
        assert int(o, 16) == (pow(x, -1, q) if x else 0)


def test_quotient_by_cosets_identity():
    """The algebra behind the coset-sharded quotient (typlonk_b200/csrc/poly.cu k_quotient_combine, api.cu):
    for N of degree < 4n, the interpolants C_k of N on the cosets w4n^k H satisfy C_k = sum_j iota^(kj) N_j
    (N = sum_j X^(jn) N_j), so N_j = 1/4 sum_k iota^(-kj) C_k and floor(N / (X^n - 1)) has the chunks
    t_2 = N_3, t_1 = N_2 + t_2, t_0 = N_1 + t_1 -- checked against the oracle's long division
    (plonk/src/proof.rs:373, 504-508)."""
    import random
    from oracle.pyoracle import fields, poly
    M = fields.R_MOD
    rnd = random.Random(11)
    for n in (2, 8, 32):
        N = [rnd.randrange(M) for _ in range(4 * n - rnd.randrange(0, 3))]
        dom = poly.Domain(n)
        w4n = fields.root_of_unity(4 * n)
        iota = pow(w4n, n, M)
        assert iota * iota % M == M - 1
        C = []
        for k in range(4):
            g = pow(w4n, k, M)
            evals = [poly.evaluate(N, g * dom.element(i) % M) for i in range(n)]
            # inverse coset transform: interpolate on H the polynomial p(g X), then undo the scaling
            coeffs = dom.ifft(evals) + [0] * n
            ginv = pow(g, M - 2, M)
            C.append([coeffs[i] * pow(ginv, i, M) % M for i in range(n)])
        quarter = pow(4, M - 2, M)
        s = (M - iota) % M  # iota^-1
        Nj = [[quarter * sum(pow(s, k * j, M) * C[k][i] for k in range(4)) % M for i in range(n)] for j in range(4)]
        flat = [c for j in range(4) for c in Nj[j]]
        assert poly.strip(flat) == poly.strip(list(N))
        t2 = Nj[3]
        t1 = [(a + b) % M for a, b in zip(Nj[2], t2)]
        t0 = [(a + b) % M for a, b in zip(Nj[1], t1)]
        q, _r = poly.divide_by_vanishing_poly(N, n)
        assert poly.strip(t0 + t1 + t2) == poly.strip(list(q))


def test_fr_conversion_helpers_of_the_abi():
    """tp_fr_from_i64 / tp_fr_from_canonical / tp_fr_to_canonical (host only) against the Python field helpers."""
    import ctypes as C
    L = ffi.lib()
    L.tp_fr_from_i64.argtypes = [C.c_int64, C.c_void_p]
    for v in (0, 1, 5, -1, -7, 2**63 - 1, -2**63):
        out = (C.c_char * 32)()
        assert L.tp_fr_from_i64(v, out) == 0
        assert bytes(out) == F.fr_to_bytes(v % fields.R_MOD), v
    for x in rng.fr_rand_stream(12, 8) + [0, fields.R_MOD - 1]:
        mont = (C.c_char * 32)()
        assert L.tp_fr_from_canonical(x.to_bytes(32, "little"), mont) == 0
        assert bytes(mont) == F.fr_to_bytes(x)
        back = (C.c_char * 32)()
        assert L.tp_fr_to_canonical(bytes(mont), back) == 0
        assert int.from_bytes(bytes(back), "little") == x
    out = (C.c_char * 32)()
    assert L.tp_fr_from_canonical(fields.R_MOD.to_bytes(32, "little"), out) == 12      # TP_ERR_MALFORMED
    assert L.tp_fr_to_canonical(b"\xff" * 32, out) == 1                                  # not a field element


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md is the binding guide: every symbol the header declares must appear in it."""
    header = open(os.path.join(ROOT, "include", "typlonk_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|const char\*)\s+(tp_[a-z0-9_]+)\(", header, flags=re.M)))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for short in ("add_row", "add_constrain", "build", "destroy"):
        doc = doc.replace("…_" + short, "tp_permutation_builder_" + short)
    missing = [n for n in declared if n not in doc]
    assert not missing, missing


def test_library_stdrng_matches_the_rand_crates_own_constants():
    """The generator behind the product's Fiat-Shamir challenges (csrc/transcript.h), checked WITHOUT the oracle against
    constants from the pinned crates' own tests: rand_core 0.6 `test_seed_from_u64` (5029875928683246316) and rand 0.8
    `test_stdrng_construction` ([10719222850664546238, 14064965282130556830])."""
    import struct
    from typlonk_b200 import ffi
    # seed_from_u64(0): the ChaCha key is the PCG32 expansion; recover its first 8 bytes through a second generator
    # seeded with the oracle-free formula is not possible from outside, so compare streams instead: from_seed(key) and
    # seed_from_u64(0) must agree when key starts with the crate constant
    w = ffi.stdrng_words(4, seed32=bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16))
    assert w[0] | (w[1] << 32) == 10719222850664546238
    w = ffi.stdrng_words(10, seed32=bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16))
    seed1 = struct.pack("<8I", *w[2:10])
    w1 = ffi.stdrng_words(2, seed32=seed1)
    assert w1[0] | (w1[1] << 32) == 14064965282130556830
    # PCG32 expansion: the key seed_from_u64(0) builds starts with the bytes of 5029875928683246316 (little endian);
    # brute-force free check: a from_seed generator whose key is that expansion (known in full from the same recurrence)
    mul, inc, state, key = 6364136223846793005, 11634580027462260723, 0, b""
    for _ in range(8):
        state = (state * mul + inc) & (2**64 - 1)
        xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        key += struct.pack("<I", ((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF)
    assert int.from_bytes(key[:8], "little") == 5029875928683246316
    assert ffi.stdrng_words(40, seed_u64=0) == ffi.stdrng_words(40, seed32=key)


def test_generated_montgomery_header_is_current_and_its_carry_chains_check(tmp_path):
    """typlonk_b200/csrc/mont_gen.cuh is exactly what tools/gen_mont.py emits, and the generator's CPU simulator
    accepts every carry chain (random and edge operands against big-integer arithmetic) before emitting."""
    import subprocess
    out = tmp_path / "mont_gen.cuh"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_mont.py"), "--out", str(out)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "separated forms ok" in res.stdout and "Fq:" in res.stdout
    with open(os.path.join(ROOT, "typlonk_b200", "csrc", "mont_gen.cuh")) as f:
        assert out.read_text() == f.read(), "mont_gen.cuh is stale: run python tools/gen_mont.py"


def test_every_tunable_is_documented_in_the_header_and_the_integration_guide():
    """tp_ctx_set_option: the names the library accepts (api.cu) are the names include/typlonk_b200.h documents and
    INTEGRATION.md lists."""
    with open(os.path.join(ROOT, "typlonk_b200", "csrc", "api.cu")) as f:
        src = f.read()
    body = src[src.index("int tp_ctx_set_option("):src.index("int tp_ctx_get_stat(")]
    accepted = set(re.findall(r'strcmp\(name, "([a-z0-9_]+)"\)', body))
    assert len(accepted) >= 8
    with open(os.path.join(ROOT, "include", "typlonk_b200.h")) as f:
        hdr = f.read()
    doc = hdr[hdr.index("Tunables of the MSM"):hdr.index("int tp_ctx_set_option(")]
    documented = set(re.findall(r'"([a-z0-9_]+)"', doc))
    assert accepted == documented, (sorted(accepted - documented), sorted(documented - accepted))
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        guide = f.read()
    for name in accepted:
        assert '"%s"' % name in guide, name
