"""ORACLE (test infrastructure only).

Restatement of the reference's `permutation` crate:
permutation/src/lib.rs:12-195 (Tag, PermutationBuilder, Permutation,
CompiledPermutation) and permutation/src/proving.rs:7-31 (grand product).

Determinism note: the reference stores constraints in a
randomly-seeded HashMap and iterates it in `build` (lib.rs:30,67-68).  This
oracle iterates keys in first-insertion order; for every circuit used by the
tests each equivalence class is fed from a single key (or has size 2), for which
the resulting sigma is iteration-order independent.
"""
from .fields import R_MOD, fr_inv
from .poly import Domain, interpolate, evaluate
from . import kzg as _kzg

C = 3  # columns; the plonk crate instantiates PermutationBuilder<3> (builder.rs:27)


class PermutationBuilder:
    def __init__(self, rows: int = 0):
        self.constrains = {}  # Tag (i, j) -> [Tag]; dict keeps insertion order
        self.rows = rows

    @classmethod
    def with_rows(cls, rows: int):
        return cls(rows)

    def add_row(self):
        self.rows += 1

    def check_tag(self, tag) -> bool:
        i, j = tag
        return i <= C and j < self.rows  # sic: `i <= &C` (lib.rs:46)

    def add_constrain(self, left, right):
        """lib.rs:48-56; returns False where the reference returns Err(())."""
        if not (self.check_tag(left) and self.check_tag(right)):
            return False
        self.constrains.setdefault(left, []).append(right)
        return True

    def add_constrains(self, pairs):
        for left, right in pairs:
            if not self.add_constrain(left, right):
                raise ValueError("invalid tag")  # `.unwrap()` on Err, lib.rs:59

    def build(self, size: int):
        """lib.rs:62-93: merge cycles by swapping mapping[left], mapping[right]."""
        length = size * C
        mapping = list(range(length))
        aux = list(range(length))
        sizes = [1] * length
        constrains, self.constrains = self.constrains, {}
        for (li, lj), rights in constrains.items():
            left = lj + li * size
            for (ri, rj) in rights:
                right = rj + ri * size
                if aux[left] == aux[right]:
                    continue
                if sizes[aux[left]] < sizes[aux[right]]:
                    left, right = right, left
                sizes[aux[left]] += sizes[aux[right]]
                nxt = right
                aux_left = aux[left]
                while True:
                    aux[nxt] = aux_left
                    nxt = mapping[nxt]
                    if aux[nxt] == aux_left:
                        break
                mapping[left], mapping[right] = mapping[right], mapping[left]
        return Permutation(mapping)


def cosets(gates: int):
    """lib.rs:141-154: first C field elements k >= 1 with k^n != 1."""
    domain = Domain(gates)
    out = []
    k = 1
    for _ in range(C):
        while domain.evaluate_vanishing_polynomial(k) == 0:
            k += 1
        out.append(k)
        k += 1
    return out


class Permutation:
    def __init__(self, perm):
        self.perm = perm

    def compile(self):
        """lib.rs:101-128: cols[i][j] = (id = k_i w^j, sigma = k_i' w^j')."""
        assert len(self.perm) % C == 0
        rows = len(self.perm) // C
        ks = cosets(rows)
        roots = Domain(rows).elements()
        cols = []
        for i in range(C):
            col = []
            for j in range(rows):
                index = self.perm[i * rows + j]
                ti, tj = index // rows, index % rows
                value = ks[ti] * roots[tj] % R_MOD
                tag = ks[i] * roots[j] % R_MOD
                col.append((tag, value))
            cols.append(col)
        return CompiledPermutation(cols, ks, rows)


class CompiledPermutation:
    def __init__(self, cols, cosets_, rows):
        self.cols = cols
        self.cosets = cosets_
        self.rows = rows

    def prove(self, values, beta: int, gamma: int):
        """proving.rs:7-31: n+1 running products, out[0] = 1; one inversion per cell
        (`numerator / denominator`), raising where the reference panics on a zero
        denominator."""
        out = [1]
        state = 1
        for j in range(self.rows):
            row_val = 1
            for i in range(C):
                cell_val = values[i][j]
                tag, value = self.cols[i][j]
                numerator = (cell_val + beta * tag + gamma) % R_MOD
                denominator = (cell_val + beta * value + gamma) % R_MOD
                row_val = row_val * numerator % R_MOD * fr_inv(denominator) % R_MOD
            state = state * row_val % R_MOD
            out.append(state)
        return out

    def sigma_polys(self, domain: Domain):
        return [interpolate([cell[1] for cell in col], domain) for col in self.cols]

    def sigma_evals(self, point: int, domain: Domain):
        """lib.rs:165-177."""
        return [evaluate(p, point) for p in self.sigma_polys(domain)]

    def sigma_commitments(self, srs, domain: Domain):
        """lib.rs:178-194."""
        return [_kzg.commit(srs, p) for p in self.sigma_polys(domain)]
