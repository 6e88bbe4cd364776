"""ORACLE (test infrastructure only): Python big-integer specification oracle of
fabrizio-m/TyPLONK (kzg, permutation, plonk crates) for small sizes."""
