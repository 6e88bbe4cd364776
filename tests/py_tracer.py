"""Python restatement of the reference's two-pass tracing scheme -- test infrastructure only.

`_Context` / `BuildVar` follow plonk/src/builder.rs:119-188, 327-378 (the recording pass), `ComputeVar`
follows builder.rs:332-336, 380-397 (the witness pass, re-run per proof, proof.rs:33-49).  The CPU tests compare
the library's native tracer (csrc/trace.cpp) with these.
"""
from typlonk_b200 import field as F
from typlonk_b200.permutation import PermutationBuilder
from typlonk_b200.plonk import Var


class _Context:
    """builder.rs:119-188."""

    def __init__(self):
        self.gates = []
        self.permutation = PermutationBuilder()
        self.next_var_id = 0
        self.pending_eq = []
        self.var_map = {}

    def new_id(self):
        self.next_var_id += 1
        return self.next_var_id - 1

    def add_gate(self, gate):
        self.gates.append(gate)
        self.permutation.add_row()
        return len(self.gates) - 1

    def add_eq(self, left, right):
        a, b = self.var_map.get(left), self.var_map.get(right)
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
        size = 2
        while size < len(self.gates) + 3:  # fill(), builder.rs:47-58
            size *= 2
        self.gates += ["Dummy"] * (size - len(self.gates))
        return self.gates, self.permutation


class BuildVar(Var):
    """builder.rs:327-378."""

    def __init__(self, context, vid):
        self.context, self.id = context, vid

    def clone(self):
        return BuildVar(self.context, self.id)

    def _binary(self, rhs, gate):
        ctx = self.context
        j = ctx.add_gate(gate)
        out = ctx.new_id()
        ctx.var_map[out] = (2, j)
        for vid, i in ((self.id, 0), (rhs.id, 1)):
            if vid in ctx.var_map:
                new_id = ctx.new_id()
                ctx.var_map[new_id] = (i, j)
                ctx.add_eq(vid, new_id)
            else:
                ctx.var_map[vid] = (i, j)
        return BuildVar(ctx, out)

    def __add__(self, rhs):
        return self._binary(rhs, "Add")

    def __mul__(self, rhs):
        return self._binary(rhs, "Mul")

    def assert_eq(self, other):
        self.context.add_eq(self.id, other.id)


class ComputeVar(Var):
    """builder.rs:332-336, 380-397; assert_eq is a no-op as in the reference (:435-441)."""

    def __init__(self, value, advice):
        self.value, self.advice = value % F.R_MOD, advice

    def clone(self):
        return ComputeVar(self.value, self.advice)

    def _binary(self, rhs, mul):
        l, r = self.value, rhs.value
        v = (l * r if mul else l + r) % F.R_MOD
        self.advice[0].append(l)
        self.advice[1].append(r)
        self.advice[2].append(v)
        return ComputeVar(v, self.advice)

    def __add__(self, rhs):
        return self._binary(rhs, False)

    def __mul__(self, rhs):
        return self._binary(rhs, True)

    def assert_eq(self, other):
        pass



def trace(desc):
    """(gate kinds incl. Dummy padding, flat permutation) of a CircuitDescription, by the recording pass."""
    ctx = _Context()
    desc.run([BuildVar(ctx, ctx.new_id()) for _ in range(desc.INPUTS)])
    gates, permutation = ctx.finish()
    return gates, permutation.build(len(gates))


def witness(desc, rows, inputs, blinders):
    """proof.rs:33-49 by re-running the closure over ComputeVar, as the reference does (canonical ints)."""
    advice = [[], [], []]
    desc.run([ComputeVar(v, advice) for v in inputs])
    assert len(blinders) == 9
    cols = []
    for k, col in enumerate(advice):
        col = col[: rows - 3] + [0] * max(0, rows - 3 - len(col))
        cols.append(col + [b % F.R_MOD for b in blinders[3 * k: 3 * k + 3]])
    return cols
