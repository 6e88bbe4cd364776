"""CPU: the library's native tracer / copy-constraint compiler / witness generator (csrc/trace.cpp, host only, no
device calls) against the oracle's restatement of plonk/src/builder.rs + permutation/src/lib.rs:62-93, against the
structural goldens of SURVEY.md App. C, and against the Python two-pass mirror."""
import random
import struct

import numpy as np
import pytest

from oracle.pyoracle import builder as obuilder, fields, permutation as operm
from typlonk_b200 import field as F, ffi, synthetic
import py_tracer
from typlonk_b200.plonk import GATE_ROWS, KIND_NAMES, CircuitDescription, TraceVar


def _native(run, n_inputs):
    t = ffi.Trace(n_inputs)
    run([TraceVar(t, k) for k in range(n_inputs)])
    t.finish()
    return t


def _perm(t):
    return list(struct.unpack("<%dQ" % (3 * t.rows), bytes(t.permutation())))


def _kinds(t):
    return [KIND_NAMES[k] for k in t.gate_kinds()]


def _witness_ints(t, inputs, blinders):
    cols = t.witness(F.fr_vec_to_bytes(inputs), F.fr_vec_to_bytes(blinders))
    return [F.fr_vec_from_bytes(bytes(c)) for c in cols]


def test_readme_circuit_structural_goldens():
    """gates, rows, permutation printed by row, witness rows of the README circuit."""
    t = _native(obuilder.circuit_pythagoras, 3)
    assert (t.rows, t.gate_count) == (8, 4)
    assert _kinds(t) == ["Mul", "Mul", "Mul", "Add"] + ["Dummy"] * 4
    perm = _perm(t)
    by_row = [[perm[j + i * 8] for i in range(3)] for j in range(8)]
    assert by_row == [[8, 0, 3], [9, 1, 11], [10, 2, 19], [16, 17, 18], [4, 12, 20], [5, 13, 21], [6, 14, 22],
                      [7, 15, 23]]
    cols = _witness_ints(t, [3, 4, 5], list(range(1, 10)))
    rows = list(zip(*cols))
    assert rows[:5] == [(3, 3, 9), (4, 4, 16), (5, 5, 25), (9, 16, 25), (0, 0, 0)]
    assert rows[5:] == [(1, 4, 7), (2, 5, 8), (3, 6, 9)]


def test_mul_chain_structural_golden():
    t = _native(obuilder.make_mul_chain(5), 2)
    perm = _perm(t)
    by_row = [[perm[j + i * 8] for i in range(3)] for j in range(8)]
    assert by_row[:5] == [[0, 12, 1], [16, 8, 2], [17, 9, 3], [18, 10, 4], [19, 11, 20]]
    assert by_row[5:] == [[5, 13, 21], [6, 14, 22], [7, 15, 23]]


@pytest.mark.parametrize("run,n_inputs,inputs", [
    (obuilder.circuit_pythagoras, 3, [3, 4, 5]),
    (obuilder.circuit_additive, 5, [1, 2, 3, 4, 5]),
    (obuilder.make_mul_chain(1), 2, [3, 5]),
    (obuilder.make_mul_chain(13), 2, [3, 5]),
    (obuilder.make_mul_chain(125), 2, [fields.R_MOD - 1, 7]),
])
def test_native_tracer_equals_oracle(run, n_inputs, inputs):
    t = _native(run, n_inputs)
    ogates, operm_ = obuilder.trace(run, n_inputs)
    assert _kinds(t) == ogates and t.rows == len(ogates)
    sel = t.selectors()
    for k in range(5):
        want = fields.fr_vec_to_mont_bytes([obuilder.GATE_ROWS[g][k] for g in ogates])
        assert bytes(sel[k * t.rows * 32:(k + 1) * t.rows * 32]) == want
    assert _perm(t) == operm_.perm
    blind = list(range(11, 20))
    assert _witness_ints(t, inputs, blind) == obuilder.witness_columns(run, inputs, t.rows, blind)
    assert GATE_ROWS == obuilder.GATE_ROWS


def _random_circuit(seed, n_inputs, n_ops):
    """A closure with clones, repeated operands (x * x), chained outputs and assert_eq between earlier results."""
    def run(inputs):
        rnd = random.Random(seed)
        pool = list(inputs)
        for _ in range(n_ops):
            a, b = rnd.choice(pool), rnd.choice(pool)
            a, b = a.clone(), b.clone()
            pool.append(a * b if rnd.random() < 0.5 else a + b)
            if rnd.random() < 0.2 and len(pool) > n_inputs + 1:
                x, y = rnd.sample(pool[n_inputs:], 2)
                x.assert_eq(y)
        # every input takes part in at least one gate, so no equality stays parked
        acc = pool[0].clone()
        for v in pool[1:n_inputs]:
            acc = acc + v.clone()
        pool[0].assert_eq(pool[-1])
    return run


@pytest.mark.parametrize("seed", range(12))
def test_random_circuits_native_vs_oracle_and_mirror(seed):
    n_inputs, n_ops = 1 + seed % 4, 5 + 9 * seed
    run = _random_circuit(seed, n_inputs, n_ops)
    t = _native(run, n_inputs)
    ogates, operm_ = obuilder.trace(run, n_inputs)
    assert _kinds(t) == ogates
    assert _perm(t) == operm_.perm

    class D(CircuitDescription):
        INPUTS = n_inputs
    D.run = staticmethod(run)
    mgates, mperm = py_tracer.trace(D)
    assert mgates == ogates and mperm.perm == operm_.perm
    inputs = [random.Random(seed + 100).randrange(fields.R_MOD) for _ in range(n_inputs)]
    blind = [random.Random(seed + 200 + k).randrange(fields.R_MOD) for k in range(9)]
    assert _witness_ints(t, inputs, blind) == obuilder.witness_columns(run, inputs, t.rows, blind)
    # a second permutation() call sees no constraints (mem::take, permutation/src/lib.rs:67)
    assert _perm(t) == list(range(3 * t.rows))


def test_pending_equalities_resolve_at_finish():
    """assert_eq on a variable that is only placed LATER is parked and resolved by finish (builder.rs:160-186)."""
    def run(inputs):
        x, y = inputs
        x.assert_eq(y)          # neither is placed yet
        z = x.clone() * y.clone()
        z.assert_eq(x)
    t = _native(run, 2)
    ogates, operm_ = obuilder.trace(run, 2)
    assert _kinds(t) == ogates and _perm(t) == operm_.perm


def test_unplaced_variable_and_argument_errors():
    t = ffi.Trace(3)
    out = t.gate(ffi.GATE_MUL, 0, 1)
    t.assert_eq(2, out)  # input 2 never enters a gate
    with pytest.raises(ffi.TyplonkError) as e:
        t.finish()
    assert e.value.code == 11  # assert!(inner.pending_eq.is_empty()), builder.rs:177
    t = ffi.Trace(1)
    with pytest.raises(ffi.TyplonkError):
        t.gate(ffi.GATE_MUL, 0, 7)      # unknown variable
    with pytest.raises(ffi.TyplonkError):
        t.gate(ffi.GATE_DUMMY, 0, 0)    # not an operator
    t.gate(ffi.GATE_ADD, 0, 0)
    t.finish()
    assert (t.rows, t.gate_count) == (4, 1)
    with pytest.raises(ffi.TyplonkError):
        t.gate(ffi.GATE_ADD, 0, 0)      # finished
    with pytest.raises(AssertionError):
        t.witness(bytes(64), bytes(288))
    with pytest.raises(ffi.TyplonkError):
        t.witness(b"\xff" * 32, bytes(288))   # input limbs >= r are not a field element
    with pytest.raises(ffi.TyplonkError):
        t.witness(bytes(32), bytes(256) + b"\xff" * 32)   # nor are such blinder limbs


def test_fill_sizes():
    """fill(): the first power of two >= gates + 3, at least 2 (builder.rs:47-58)."""
    for gates, rows in [(0, 4), (1, 4), (2, 8), (5, 8), (6, 16), (13, 16), (14, 32), (61, 64), (62, 128)]:
        t = synthetic.mul_chain_trace(gates)
        assert (t.rows, t.gate_count) == (rows, gates)


def test_bulk_mul_chain_equals_closure_and_python_witness():
    for g in (1, 2, 5, 61, 1021):
        t = synthetic.mul_chain_trace(g)
        c = _native(obuilder.make_mul_chain(g), 2)
        assert t.rows == c.rows and t.gate_kinds() == c.gate_kinds()
        assert bytes(t.selectors()) == bytes(c.selectors())
        assert _perm(t) == _perm(c)
    log_n = 16
    n = 1 << log_n
    t = synthetic.mul_chain_trace(n - 3)
    _, sperm = synthetic.mul_chain_structure(n - 3)
    assert _perm(t) == sperm.perm
    blind = synthetic.blinders()
    cols = t.witness(F.fr_vec_to_bytes([3, 5]), F.fr_vec_to_bytes(blind))
    want = synthetic.mul_chain_witness(n - 3, n, blind=blind)
    for k in range(3):
        assert bytes(cols[k]) == F.fr_vec_to_bytes(want[k])


def test_native_permutation_builder_equals_oracle():
    rnd = random.Random(5)
    for trial in range(20):
        rows = rnd.choice([1, 2, 4, 8, 16])
        nb, ob = ffi.NativePermutationBuilder(rows), operm.PermutationBuilder.with_rows(rows)
        for _ in range(rnd.randrange(0, 40)):
            l = (rnd.randrange(3), rnd.randrange(rows))
            r = (rnd.randrange(3), rnd.randrange(rows + (trial % 3 == 0)))
            assert nb.add_constrain(l, r) == ob.add_constrain(l, r)
        size = rows * rnd.choice([1, 2])
        assert nb.build(size) == ob.build(size).perm
    nb = ffi.NativePermutationBuilder()
    nb.add_row()
    assert nb.add_constrain((3, 0), (0, 0))          # sic: `i <= C` passes check_tag (permutation/src/lib.rs:46)
    assert not nb.add_constrain((4, 0), (0, 0))
    assert not nb.add_constrain((0, 0), (0, 1))
    with pytest.raises(ffi.TyplonkError) as e:       # ...and the reference then indexes out of bounds in build
        nb.build(1)
    assert e.value.code == 10


def test_cycles_are_a_permutation_with_the_right_classes():
    """Property at a size the oracle does not need: perm is a bijection whose cycles are exactly the copy classes."""
    g = (1 << 14) - 3
    t = synthetic.mul_chain_trace(g)
    n = t.rows
    perm = np.frombuffer(bytes(t.permutation()), dtype=np.uint64).astype(np.int64)
    assert np.array_equal(np.sort(perm), np.arange(3 * n))
    # class of y: column 1, rows 0..g-1 -- one cycle of length g
    seen, k = 0, n
    while True:
        seen += 1
        k = int(perm[k])
        if k == n:
            break
    assert seen == g
    # c[j] = a[j+1]
    for j in (0, 1, g - 2):
        assert int(perm[2 * n + j]) == j + 1 and int(perm[j + 1]) == 2 * n + j
