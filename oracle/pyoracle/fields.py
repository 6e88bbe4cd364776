"""ORACLE (test infrastructure only -- never imported by the product path).

BLS12-381 scalar field Fr and base field Fq as plain Python integers.

The reference (fabrizio-m/TyPLONK) takes these from the un-vendored crates
ark-ff 0.3.0 / ark-bls12-381 0.3.0 (Cargo.lock); the types enter at
kzg/src/lib.rs:2,12-14 (`Fr`, `G1Point`).  Values here are *canonical* integers;
`to_mont`/`from_mont` convert to the Montgomery representation arkworks keeps in
memory (R = 2^256 for Fr, 2^384 for Fq, little-endian u64 limbs), which is the
layout that crosses the C ABI.

Parity: UNPINNED by the reference (it ships no golden vectors); pinned here by
external known answers (tests/test_oracle_known_answers.py).
"""

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001  # Fr order r
Q_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB

FR_R = (1 << 256) % R_MOD      # Montgomery R mod r
FQ_R = (1 << 384) % Q_MOD
FR_RINV = pow(FR_R, -1, R_MOD)
FQ_RINV = pow(FQ_R, -1, Q_MOD)

FR_TWO_ADICITY = 32
FR_GENERATOR = 7
# ark-bls12-381 FrParameters::TWO_ADIC_ROOT_OF_UNITY = 7^((r-1)/2^32)
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_TWO_ADICITY, R_MOD)


def fr(x: int) -> int:
    """`Fr::from(i32/u64)`: negatives wrap mod r (plonk/src/utils.rs:152-153)."""
    return x % R_MOD


def fr_inv(x: int) -> int:
    if x % R_MOD == 0:
        # ark-ff `Div` does `inverse().unwrap()` -> panic (permutation/src/proving.rs:21)
        raise ZeroDivisionError("Fr inverse of zero")
    return pow(x, -1, R_MOD)


def fr_to_mont(x: int) -> int:
    return (x * FR_R) % R_MOD


def fr_from_mont(x: int) -> int:
    return (x * FR_RINV) % R_MOD


def fq_to_mont(x: int) -> int:
    return (x * FQ_R) % Q_MOD


def fq_from_mont(x: int) -> int:
    return (x * FQ_RINV) % Q_MOD


def root_of_unity(n: int) -> int:
    """ark-ff 0.3 `FftField::get_root_of_unity(n)`: size = n.next_power_of_two(),
    omega = TWO_ADIC_ROOT_OF_UNITY squared (32 - log2 size) times."""
    size = 1
    log = 0
    while size < n:
        size <<= 1
        log += 1
    assert log <= FR_TWO_ADICITY
    w = FR_ROOT_OF_UNITY
    for _ in range(FR_TWO_ADICITY - log):
        w = w * w % R_MOD
    return w


# ---- byte / limb helpers (the C-ABI layout) ---------------------------------

def fr_mont_bytes(x: int) -> bytes:
    """32 bytes: 4 little-endian u64 limbs of the Montgomery form (ark_ff::Fp256.0)."""
    return fr_to_mont(x).to_bytes(32, "little")


def fr_from_mont_bytes(b: bytes) -> int:
    return fr_from_mont(int.from_bytes(b, "little"))


def fq_mont_bytes(x: int) -> bytes:
    return fq_to_mont(x).to_bytes(48, "little")


def fq_from_mont_bytes(b: bytes) -> int:
    return fq_from_mont(int.from_bytes(b, "little"))


def fr_vec_to_mont_bytes(v) -> bytes:
    return b"".join(fr_mont_bytes(x) for x in v)


def fr_vec_from_mont_bytes(b: bytes):
    return [fr_from_mont_bytes(b[i:i + 32]) for i in range(0, len(b), 32)]
