"""Synthetic circuits and inputs of BASELINE.json / SURVEY.md 8(d): the multiplication-chain
circuit `let [x,y]=inputs; for _ in 0..G { x = x * y.clone(); }` with G = n - 3 gates, inputs
[3, 5], public inputs [0]; tau / blinders / sweep inputs are `Fr::rand` streams of
`StdRng::seed_from_u64(1|2|3|4)`.

`mul_chain_direct` builds the gate list, copy constraints and witness columns of that circuit
without running the tracing DSL (same result, checked against the tracer in the CPU tests), so
that 2^20..2^22-gate setups take seconds of host time."""
import struct

from . import field as F
from .ffi import Context, fr_rand_stream
from .kzg import Srs
from .permutation import PermutationBuilder
from .plonk import GATE_ROWS, CircuitDescription, CompiledCircuit

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


def mul_chain_direct(ctx: Context, log_n: int, tau_value=None) -> CompiledCircuit:
    n = 1 << log_n
    gates = n - 3
    gate_list, perm = mul_chain_structure(gates)
    srs = Srs.from_secret(ctx, tau() if tau_value is None else tau_value, n)
    one = F.fr_to_bytes(1)
    zero = bytes(32)
    on = one * gates + zero * (n - gates)
    off = zero * n
    sel = [off, off, on, on, off]  # Mul = [0, 0, 1, 1, 0]
    assert GATE_ROWS["Mul"] == (0, 0, 1, 1, 0)
    perm_bytes = struct.pack("<%dQ" % len(perm.perm), *perm.perm)
    handle, fixed = ctx.circuit_compile(srs.handle, sel, perm_bytes, n)
    return CompiledCircuit(mul_chain_description(gates), ctx, srs, handle, n,
                           [F.g1_from_abi(c) for c in fixed], gate_list, perm)
