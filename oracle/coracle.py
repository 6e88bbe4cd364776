"""ORACLE (test infrastructure only): ctypes front end of oracle/c/oracle.cpp, the multi-threaded
C++ restatement of the reference prover (CPU).  Allowed callers: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline and --impl reference legs.  Built into oracle/_build/ (git-ignored)."""
import ctypes as C
import os
import struct
import subprocess
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "c" / "oracle.cpp"
OUT = HERE / "_build" / "liboracle.so"
_lib = None


def build(force=False) -> Path:
    OUT.parent.mkdir(exist_ok=True)
    if not force and OUT.exists() and OUT.stat().st_mtime >= SRC.stat().st_mtime:
        return OUT
    cmd = ["g++", "-O3", "-mbmi2", "-madx", "-fopenmp", "-shared", "-fPIC", "-std=c++17", str(SRC), "-o", str(OUT)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stderr)
    return OUT


def lib():
    global _lib
    if _lib is None:
        if not OUT.exists():
            build()
        _lib = C.CDLL(str(OUT))
        _lib.oracle_circuit_new.restype = C.c_void_p
        _lib.oracle_now_ms.restype = C.c_double
    return _lib


def threads() -> int:
    return lib().oracle_threads()


def fr_rand_stream(seed: int, count: int) -> bytes:
    out = (C.c_char * (32 * count))()
    lib().oracle_fr_rand_stream(C.c_uint64(seed), C.c_size_t(count), out)
    return bytes(out)


def ntt(data: bytes, log_n: int, inverse=False) -> bytes:
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    lib().oracle_ntt(buf, C.c_uint(log_n), C.c_int(1 if inverse else 0))
    return bytes(buf)


def msm(points_packed: bytes, scalars: bytes) -> bytes:
    n = len(scalars) // 32
    out = (C.c_char * 97)()
    lib().oracle_msm(points_packed, scalars, C.c_size_t(n), out)
    return bytes(out)


def srs(tau_mont: bytes, length: int) -> bytes:
    out = (C.c_char * (96 * length))()
    lib().oracle_srs(tau_mont, C.c_size_t(length), out)
    return bytes(out)


class Circuit:
    def __init__(self, tau_mont: bytes, selector_evals, perm, n):
        self._keep = [(C.c_char * len(b)).from_buffer_copy(b) for b in selector_evals]
        sel = (C.c_void_p * 5)(*[C.cast(x, C.c_void_p) for x in self._keep])
        perm_b = struct.pack("<%dQ" % len(perm), *perm)
        self.n = n
        self._h = C.c_void_p(lib().oracle_circuit_new(tau_mont, sel, perm_b, C.c_size_t(n)))

    def fixed_commitments(self):
        out = (C.c_char * (5 * 97))()
        lib().oracle_circuit_fixed_commitments(self._h, out)
        raw = bytes(out)
        return [raw[i * 97:(i + 1) * 97] for i in range(5)]

    def prove(self, advice_mont, public_inputs_mont):
        keep = [(C.c_char * len(b)).from_buffer_copy(b) for b in advice_mont]
        adv = (C.c_void_p * 3)(*[C.cast(x, C.c_void_p) for x in keep])
        out = (C.c_char * 1472)()
        rc = lib().oracle_prove(self._h, adv, public_inputs_mont, out)
        if rc != 0:
            raise AssertionError("oracle_prove failed with reference-panic code %d" % rc)
        return bytes(out)

    def commit_as_written(self, coeffs_mont: bytes) -> bytes:
        out = (C.c_char * 97)()
        lib().oracle_commit_as_written(self._h, coeffs_mont, C.c_size_t(len(coeffs_mont) // 32), out)
        return bytes(out)

    def close(self):
        if self._h:
            lib().oracle_circuit_free(self._h)
            self._h = None


def naive_mul(a: bytes, b: bytes) -> bytes:
    la, lb = len(a) // 32, len(b) // 32
    out = (C.c_char * (32 * (la + lb - 1)))()
    lib().oracle_naive_mul(a, C.c_size_t(la), b, C.c_size_t(lb), out)
    return bytes(out)


# ---- synthetic mul-chain workload built without any product code ------------------
_R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_RM = (1 << 256) % _R


def _mont(x):
    return (x % _R * _RM % _R).to_bytes(32, "little")


def mul_chain_inputs(log_n: int):
    """(tau, selector evals, perm, advice columns, public inputs) for the n = 2^log_n mul chain."""
    from .pyoracle import permutation as operm
    n = 1 << log_n
    gates = n - 3
    pb = operm.PermutationBuilder.with_rows(gates)
    for j in range(1, gates):
        pb.add_constrain((2, j - 1), (0, j))
        pb.add_constrain((1, 0), (1, j))
    perm = pb.build(n).perm
    one, zero = _mont(1), bytes(32)
    on = one * gates + zero * (n - gates)
    sel = [zero * n, zero * n, on, on, zero * n]
    blind = fr_rand_stream(2, 9)
    a, c = [], []
    x = 3
    for _ in range(gates):
        a.append(_mont(x))
        x = x * 5 % _R
        c.append(_mont(x))
    cols = [b"".join(a), _mont(5) * gates, b"".join(c)]
    cols = [col + zero * (n - 3 - gates) + blind[96 * k: 96 * k + 96] for k, col in enumerate(cols)]
    return fr_rand_stream(1, 1), sel, perm, cols, zero * n


def bench_prove(log_n: int, steps=1, warmup=0):
    """Time the CPU prover on the mul-chain circuit (witness -> proof bytes; setup excluded)."""
    tau, sel, perm, cols, pi = mul_chain_inputs(log_n)
    c = Circuit(tau, sel, perm, 1 << log_n)
    for _ in range(warmup):
        c.prove(cols, pi)
    t0 = time.perf_counter()
    for _ in range(max(steps, 1)):
        proof = c.prove(cols, pi)
    ms = (time.perf_counter() - t0) * 1e3 / max(steps, 1)
    c.close()
    return {"ms_per_step": ms, "threads": threads(), "log_n": log_n, "proof_digest": proof[:8].hex()}
