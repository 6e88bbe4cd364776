"""Synthetic circuits and inputs of BASELINE.json / SURVEY.md 8(d): the multiplication-chain
circuit `let [x,y]=inputs; for _ in 0..G { x = x * y.clone(); }` with G = n - 3 gates, inputs
[3, 5], public inputs [0]; tau / blinders / sweep inputs are `Fr::rand` streams of
`StdRng::seed_from_u64(1|2|3|4)`.

`mul_chain_direct` records that circuit into the library's native tracer with one bulk call
(tp_trace_gates) instead of running the Python closure gate by gate -- same recording, checked
against the closure in the CPU tests -- so that 2^20..2^22-gate setups take a fraction of a second
of host time.  `mul_chain_structure` / `mul_chain_witness` are the pure-Python statement of the same
circuit the tests compare with."""
import struct

from . import field as F
from .ffi import GATE_MUL, Context, Trace, fr_rand_stream
from .kzg import Srs
from .permutation import PermutationBuilder
from .permutation import Permutation
from .plonk import KIND_NAMES, CircuitDescription, CompiledCircuit

SEED_TAU, SEED_BLINDERS, SEED_MSM, SEED_NTT = 1, 2, 3, 4


def tau() -> int:
    return F.fr_from_bytes(fr_rand_stream(SEED_TAU, 1))


def blinders():
    return F.fr_vec_from_bytes(fr_rand_stream(SEED_BLINDERS, 9))


def mul_chain_description(gates: int):
    class MulChain(CircuitDescription):
        INPUTS = 2

        @staticmethod
        def run(inputs):
            x, y = inputs
            for _ in range(gates):
                x = x * y.clone()
    return MulChain


def mul_chain_structure(gates: int):
    """(gate list padded to n, permutation) exactly as the tracer + `fill` + `build` produce."""
    n = 2
    while n < gates + 3:
        n *= 2
    pb = PermutationBuilder.with_rows(gates)
    for j in range(1, gates):
        pb.add_constrain((2, j - 1), (0, j))
        pb.add_constrain((1, 0), (1, j))
    perm = pb.build(n)
    return ["Mul"] * gates + ["Dummy"] * (n - gates), perm


def mul_chain_witness(gates: int, n: int, x0=3, y=5, blind=None):
    """Columns a, b, c (canonical ints, n each; last three rows = blinders)."""
    blind = blinders() if blind is None else blind
    a, c = [0] * gates, [0] * gates
    x = x0 % F.R_MOD
    for j in range(gates):
        a[j] = x
        x = x * y % F.R_MOD
        c[j] = x
    cols = [a, [y % F.R_MOD] * gates, c]
    out = []
    for k, col in enumerate(cols):
        out.append(col + [0] * (n - 3 - gates) + list(blind[3 * k:3 * k + 3]))
    return out


def mul_chain_trace(gates: int) -> Trace:
    """The mul-chain closure recorded in bulk.  Variable ids follow builder.rs:339-370: inputs x = 0, y = 1;
    gate 0 places both and outputs id 2; every later gate allocates its output and then one copy id per
    (already placed) operand, so gate j >= 1 outputs id 3j."""
    import numpy as np
    t = Trace(2)
    if gates:
        j = np.arange(gates, dtype=np.uint64)
        out_ids = np.where(j == 0, 2, 3 * j).astype(np.uint64)
        lhs = np.concatenate([np.zeros(1, dtype=np.uint64), out_ids[:-1]])
        got = t.gates(np.full(gates, GATE_MUL, dtype=np.uint8), lhs, np.ones(gates, dtype=np.uint64))
        assert (got == out_ids).all()
    t.finish()
    return t


def mul_chain_direct(ctx: Context, log_n: int, tau_value=None) -> CompiledCircuit:
    n = 1 << log_n
    gates = n - 3
    t = mul_chain_trace(gates)
    assert t.rows == n
    srs = Srs.from_secret(ctx, tau() if tau_value is None else tau_value, n)
    sel_all = t.selectors()
    sel = [bytes(sel_all[k * n * 32:(k + 1) * n * 32]) for k in range(5)]
    perm_bytes = bytes(t.permutation())
    handle, fixed = ctx.circuit_compile(srs.handle, sel, perm_bytes, n)
    perm = Permutation(list(struct.unpack("<%dQ" % (3 * n), perm_bytes)))
    return CompiledCircuit(mul_chain_description(gates), ctx, srs, handle, n,
                           [F.g1_from_abi(c) for c in fixed], [KIND_NAMES[k] for k in t.gate_kinds()], perm, t)
