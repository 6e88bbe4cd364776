"""Host-side mirror of the reference's `permutation` crate API (permutation/src/lib.rs:12-195,
permutation/src/proving.rs:7-31).  Cycle building is host bookkeeping;
the sigma/id tables and the grand product run on the GPU."""
from . import field as F
from .ffi import Context

C = 3


class PermutationBuilder:
    """permutation/src/lib.rs:28-93 (constraints iterate in first-insertion order; see DESIGN.md
    on the reference's HashMap order)."""

    def __init__(self, rows=0):
        self.constrains = {}
        self.rows = rows

    @classmethod
    def with_rows(cls, rows):
        return cls(rows)

    def add_row(self):
        self.rows += 1

    def _check_tag(self, tag):
        return tag[0] <= C and tag[1] < self.rows

    def add_constrain(self, left, right):
        if not (self._check_tag(left) and self._check_tag(right)):
            return False
        self.constrains.setdefault(left, []).append(right)
        return True

    def add_constrains(self, pairs):
        for left, right in pairs:
            if not self.add_constrain(left, right):
                raise ValueError("invalid tag")

    def build(self, size):
        n = size * C
        mapping = list(range(n))
        aux = list(range(n))
        sizes = [1] * n
        cons, self.constrains = self.constrains, {}
        for (li, lj), rights in cons.items():
            left = lj + li * size
            for (ri, rj) in rights:
                right = rj + ri * size
                if aux[left] == aux[right]:
                    continue
                if sizes[aux[left]] < sizes[aux[right]]:
                    left, right = right, left
                sizes[aux[left]] += sizes[aux[right]]
                nxt, tgt = right, aux[left]
                while True:
                    aux[nxt] = tgt
                    nxt = mapping[nxt]
                    if aux[nxt] == tgt:
                        break
                mapping[left], mapping[right] = mapping[right], mapping[left]
        return Permutation(mapping)


class Permutation:
    def __init__(self, perm):
        self.perm = perm

    def compile(self, ctx: Context) -> "CompiledPermutation":
        """Permutation::compile (permutation/src/lib.rs:101-128) on the device (tp_permutation_compile)."""
        import struct
        n = len(self.perm) // C
        ids, sgs, ks = ctx.permutation_compile(struct.pack("<%dQ" % len(self.perm), *self.perm), n)
        return CompiledPermutation(ctx, [F.fr_vec_from_bytes(bytes(b)) for b in ids],
                                   [F.fr_vec_from_bytes(bytes(b)) for b in sgs], [F.fr_from_bytes(k) for k in ks])


class CompiledPermutation:
    """permutation/src/lib.rs:156-195: `cols` (id, sigma) per column + cosets."""

    def __init__(self, ctx: Context, ids, sigmas, cosets):
        self.ctx = ctx
        self.ids = ids          # 3 lists of n canonical ints
        self.sigmas = sigmas
        self.cosets = cosets
        self.rows = len(ids[0])

    def prove(self, values, beta: int, gamma: int):
        """CompiledPermutation::prove (proving.rs:7-31): n + 1 running products."""
        out = self.ctx.perm_prove([F.fr_vec_to_bytes(v) for v in values],
                                  [F.fr_vec_to_bytes(v) for v in self.ids],
                                  [F.fr_vec_to_bytes(v) for v in self.sigmas],
                                  F.fr_to_bytes(beta), F.fr_to_bytes(gamma))
        return F.fr_vec_from_bytes(out)
