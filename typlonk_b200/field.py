"""Host-side BLS12-381 field constants and Montgomery (de)serialisation helpers for the Python
mirror of the reference API.  Values cross the C ABI as ark_ff does in memory: little-endian
limbs of x * R mod p (R = 2^256 for Fr, 2^384 for Fq)."""

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
Q_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_FR_R = (1 << 256) % R_MOD
_FR_RINV = pow(_FR_R, -1, R_MOD)
_FQ_R = (1 << 384) % Q_MOD
_FQ_RINV = pow(_FQ_R, -1, Q_MOD)


def fr_to_bytes(x: int) -> bytes:
    return (x % R_MOD * _FR_R % R_MOD).to_bytes(32, "little")


def fr_from_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little") * _FR_RINV % R_MOD


def fr_vec_to_bytes(xs) -> bytes:
    return b"".join((x % R_MOD * _FR_R % R_MOD).to_bytes(32, "little") for x in xs)


def fr_vec_from_bytes(b: bytes):
    return [int.from_bytes(b[i:i + 32], "little") * _FR_RINV % R_MOD for i in range(0, len(b), 32)]


def fq_to_bytes(x: int) -> bytes:
    return (x % Q_MOD * _FQ_R % Q_MOD).to_bytes(48, "little")


def fq_from_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little") * _FQ_RINV % Q_MOD


def g1_from_abi(b: bytes):
    """97-byte ABI point -> (x, y) canonical ints, or None for infinity."""
    assert len(b) == 97
    if b[96]:
        return None
    return (fq_from_bytes(b[:48]), fq_from_bytes(b[48:96]))


def g1_to_packed(pt) -> bytes:
    """(x, y) -> 96-byte packed SRS record; None -> all zero."""
    if pt is None:
        return bytes(96)
    return fq_to_bytes(pt[0]) + fq_to_bytes(pt[1])


def g1_from_packed(b: bytes):
    if b == bytes(96):
        return None
    return (fq_from_bytes(b[:48]), fq_from_bytes(b[48:96]))


def g1_serialize_unchecked(pt) -> bytes:
    """ark-serialize 0.3 uncompressed G1 (plonk/src/proof/challenges.rs:17-22)."""
    if pt is None:
        out = bytearray(96)
        out[48] = 1
        out[95] |= 1 << 6
        return bytes(out)
    return pt[0].to_bytes(48, "little") + pt[1].to_bytes(48, "little")


def g1_to_abi(pt) -> bytes:
    """(x, y) / None -> 97-byte ABI point (infinity = (0, 1), flag 1, like GroupAffine::zero())."""
    if pt is None:
        return fq_to_bytes(0) + fq_to_bytes(1) + b"\x01"
    return fq_to_bytes(pt[0]) + fq_to_bytes(pt[1]) + b"\x00"


def g2_to_abi(pt) -> bytes:
    """((x0, x1), (y0, y1)) / None -> 193-byte ABI G2 record."""
    if pt is None:
        return fq_to_bytes(0) * 2 + fq_to_bytes(1) + fq_to_bytes(0) + b"\x01"
    (x0, x1), (y0, y1) = pt
    return fq_to_bytes(x0) + fq_to_bytes(x1) + fq_to_bytes(y0) + fq_to_bytes(y1) + b"\x00"


def g2_from_abi(b: bytes):
    assert len(b) == 193
    if b[192]:
        return None
    v = [fq_from_bytes(b[i:i + 48]) for i in range(0, 192, 48)]
    return ((v[0], v[1]), (v[2], v[3]))
