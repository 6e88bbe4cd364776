"""ORACLE (test infrastructure only).

Restatement of the reference's circuit DSL and compile step:
plonk/src/description.rs:4-16 (CircuitDescription / Var),
plonk/src/builder.rs:25-188 (CircuitBuilder, Context), :318-441 (Gate rows,
BuildVar, ComputeVar).  Debug printing (builder.rs:81,153,158,293) is not
reproduced.

A circuit is a Python callable `run(inputs: list[Var]) -> None` that uses `+`,
`*`, `.clone()` and `.assert_eq()` on its inputs, exactly like
`CircuitDescription::run<V: Var>`.
"""
from .fields import R_MOD
from .permutation import PermutationBuilder
from .poly import Domain, interpolate
from . import kzg

GATE_MUL, GATE_ADD, GATE_DUMMY = "Mul", "Add", "Dummy"

# builder.rs:318-324  [q_l, q_r, q_o, q_m, q_c]
GATE_ROWS = {
    GATE_MUL: (0, 0, 1, 1, 0),
    GATE_ADD: (1, 1, 1, 0, 0),
    GATE_DUMMY: (0, 0, 0, 0, 0),
}


class Context:
    """builder.rs:119-188 (InnerContext + Context)."""

    def __init__(self):
        self.gates = []
        self.permutation = PermutationBuilder()
        self.next_var_id = 0
        self.pending_eq = []
        self.var_map = {}

    def new_id(self) -> int:
        v = self.next_var_id
        self.next_var_id += 1
        return v

    def add_gate(self, gate) -> int:
        self.gates.append(gate)
        self.permutation.add_row()
        return len(self.gates) - 1

    def add_var(self, vid, tag):
        self.var_map[vid] = tag

    def get_var(self, vid):
        return self.var_map.get(vid)

    def add_eq(self, left, right):
        a, b = self.get_var(left), self.get_var(right)
        if a is not None and b is not None:
            if not self.permutation.add_constrain(a, b):
                raise ValueError("invalid tag")
        else:
            self.pending_eq.append((left, right))

    def finish(self):
        pending, self.pending_eq = self.pending_eq, []
        for left, right in pending:
            self.add_eq(left, right)
        assert not self.pending_eq
        # fill (builder.rs:47-58): pad with Dummy to the first 2^k >= gates + 3, k >= 1
        rows = len(self.gates)
        size = 2
        while size < rows + 3:
            size *= 2
        self.gates.extend([GATE_DUMMY] * (size - rows))
        return self.gates, self.permutation


class BuildVar:
    """builder.rs:327-378, 399-434."""

    def __init__(self, context: Context, vid: int):
        self.context = context
        self.id = vid

    @classmethod
    def input(cls, context: Context):
        return cls(context, context.new_id())

    def clone(self):
        return BuildVar(self.context, self.id)

    def _binary(self, rhs, gate):
        ctx = self.context
        j = ctx.add_gate(gate)
        out_id = ctx.new_id()
        ctx.add_var(out_id, (2, j))
        for vid, i in ((self.id, 0), (rhs.id, 1)):
            if ctx.get_var(vid) is not None:
                new_id = ctx.new_id()
                ctx.add_var(new_id, (i, j))
                ctx.add_eq(vid, new_id)
            else:
                ctx.add_var(vid, (i, j))
        return BuildVar(ctx, out_id)

    def __add__(self, rhs):
        return self._binary(rhs, GATE_ADD)

    def __mul__(self, rhs):
        return self._binary(rhs, GATE_MUL)

    def assert_eq(self, other):
        self.context.add_eq(self.id, other.id)


class ComputeVar:
    """builder.rs:332-336, 380-397, 435-441."""

    def __init__(self, value: int, advice):
        self.value = value % R_MOD
        self.advice = advice

    def clone(self):
        return ComputeVar(self.value, self.advice)

    def _binary(self, rhs, is_mul):
        left, right = self.value, rhs.value
        value = (left * right if is_mul else left + right) % R_MOD
        for col, v in zip(self.advice, (left, right, value)):
            col.append(v)
        return ComputeVar(value, self.advice)

    def __add__(self, rhs):
        return self._binary(rhs, False)

    def __mul__(self, rhs):
        return self._binary(rhs, True)

    def assert_eq(self, other):
        pass  # deliberately a no-op in the reference (builder.rs:437-440)


class CompiledCircuit:
    """plonk/src/lib.rs:18-35."""

    def __init__(self, run, n_inputs, selectors, fixed_commitments, copy_constrains, srs, domain, rows,
                 gates):
        self.run = run
        self.n_inputs = n_inputs
        self.q_l, self.q_r, self.q_o, self.q_m, self.q_c = selectors  # coefficient form
        self.fixed_commitments = fixed_commitments
        self.copy_constrains = copy_constrains
        self.srs = srs
        self.domain = domain
        self.rows = rows
        self.gates = gates


def trace(run, n_inputs: int):
    """Structural half of compile (builder.rs:61-66,80): gates + permutation."""
    ctx = Context()
    inputs = [BuildVar.input(ctx) for _ in range(n_inputs)]
    run(inputs)
    gates, permutation = ctx.finish()
    perm = permutation.build(len(gates))
    return gates, perm


def compile_circuit(run, n_inputs: int, tau: int) -> CompiledCircuit:
    """`CircuitBuilder::compile` (builder.rs:60-113) with the SRS secret passed in
    (the reference draws it from thread_rng, builder.rs:71 -> srs.rs:36-40)."""
    gates, perm = trace(run, n_inputs)
    rows = len(gates)
    domain = Domain(rows)
    srs = kzg.Srs.from_secret(tau, domain.size)
    cols = [[GATE_ROWS[g][k] for g in gates] for k in range(5)]
    compiled_perm = perm.compile()
    selectors = [interpolate(c, domain) for c in cols]
    commitments = [kzg.commit(srs, p) for p in selectors]
    return CompiledCircuit(run, n_inputs, selectors, commitments, compiled_perm, srs, domain, rows, gates)


def witness_columns(run, inputs, rows: int, blinders):
    """plonk/src/proof.rs:33-49: run the circuit on values, pad each column to
    rows-3 with zeros, append 3 blinders per column (order a0 a1 a2 b0 b1 b2 c0 c1 c2)."""
    advice = [[], [], []]
    run([ComputeVar(v, advice) for v in inputs])
    assert len(blinders) == 9
    out = []
    for k, col in enumerate(advice):
        col = list(col)
        if len(col) > rows - 3:
            col = col[: rows - 3]  # Vec::resize truncates
        col += [0] * (rows - 3 - len(col))
        col += [b % R_MOD for b in blinders[3 * k: 3 * k + 3]]
        out.append(col)
    return out


# ---- circuits used by the reference's tests and by BASELINE.json -------------

def circuit_pythagoras(inputs):
    """README.md:16-27 / plonk/src/builder/test.rs:12-23 (Circuit2), 3 inputs."""
    a, b, c = inputs
    a = a.clone() * a
    b = b.clone() * b
    c = c.clone() * c
    d = a + b
    d.assert_eq(c)


def circuit_additive(inputs):
    """plonk/src/builder/test.rs:3-11 (Circuit1), 5 inputs."""
    a, b, c, d, e = inputs
    x = (c + d) + e
    a = a + b
    a.assert_eq(x)


def make_mul_chain(gates: int):
    """`let [x,y]=inputs; for _ in 0..G { x = x * y.clone(); }`."""
    def run(inputs):
        x, y = inputs
        for _ in range(gates):
            x = x * y.clone()
    return run
