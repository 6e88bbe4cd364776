"""Wire formats (csrc/wire.cu).  CPU part: the proof codec (host only) against the oracle's
proof bytes and the committed golden proofs, and its rejection of every kind of malformed input.  GPU part: the SRS
codec, whose Montgomery <-> canonical conversion and validation run on the device."""
import json
import os
import struct

import pytest

from oracle.pyoracle import curve, fields, kzg as okzg, rng
from typlonk_b200 import field as F, ffi
from typlonk_b200.plonk import Proof

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PROOFS = json.load(open(os.path.join(GOLD, "proofs.json")))


@pytest.mark.parametrize("name", sorted(PROOFS))
def test_proof_codec_round_trips_golden_proofs(name):
    raw = bytes.fromhex(PROOFS[name]["proof_hex"])
    p = Proof.from_bytes(raw)
    assert p.fixed == raw[:1472]
    n = struct.unpack("<Q", raw[1472:1480])[0]
    assert len(p.public_inputs) == n
    assert p.to_bytes() == raw
    fixed, pis = ffi.proof_decode(raw)
    assert ffi.proof_encode(fixed, pis) == raw


def _golden():
    return bytearray(bytes.fromhex(PROOFS["readme_pythagoras_3_4_5"]["proof_hex"]))


def test_proof_decode_rejects_malformed():
    raw = _golden()
    for cut in (0, 100, 1472, 1479, len(raw) - 1):
        with pytest.raises(ffi.Malformed):
            ffi.proof_decode(bytes(raw[:cut]))
    with pytest.raises(ffi.Malformed):
        ffi.proof_decode(bytes(raw) + b"\0")                      # trailing byte
    bad = _golden()
    bad[1472:1480] = struct.pack("<Q", 1 << 61)                    # absurd count must not overflow the length check
    with pytest.raises(ffi.Malformed):
        ffi.proof_decode(bytes(bad))
    bad = _golden()
    bad[0] ^= 1                                                    # a.commitment leaves the curve
    with pytest.raises(ffi.Malformed):
        ffi.proof_decode(bytes(bad))
    bad = _golden()
    bad[95] |= 0x80                                                # compressed-form sign flag
    with pytest.raises(ffi.Malformed):
        ffi.proof_decode(bytes(bad))
    bad = _golden()
    bad[0:48] = fields.Q_MOD.to_bytes(48, "little")                # x = q: not canonical
    with pytest.raises(ffi.Malformed):
        ffi.proof_decode(bytes(bad))
    bad = _golden()
    bad[192:224] = fields.R_MOD.to_bytes(32, "little")             # a(zeta) = r: not canonical
    with pytest.raises(ffi.Malformed):
        ffi.proof_decode(bytes(bad))
    bad = _golden()
    bad[len(bad) - 32:] = (fields.R_MOD + 5).to_bytes(32, "little")  # last public input
    with pytest.raises(ffi.Malformed):
        ffi.proof_decode(bytes(bad))
    # infinity is a valid point: x = 0, y = 1, bit 6 of the last byte (GroupAffine::zero())
    ok = _golden()
    ok[0:96] = F.g1_serialize_unchecked(None)
    fixed, _ = ffi.proof_decode(bytes(ok))
    assert fixed[:96] == F.g1_serialize_unchecked(None)
    # ... and ONLY in that encoding: the flag over any other (x, y) is a second spelling of the same point
    for x, y in ((5, 7), (0, 0), (0, 2), (1, 1)):
        bad = _golden()
        rec = bytearray(x.to_bytes(48, "little") + y.to_bytes(48, "little"))
        rec[95] |= 0x40
        bad[0:96] = rec
        with pytest.raises(ffi.Malformed):
            ffi.proof_decode(bytes(bad))
        # the verifier's own parser treats it as malformed too (no device needed: rejected while parsing)
        with pytest.raises(ffi.TyplonkError):
            ffi.proof_challenges(bytes(bad[:1472]))


def test_proof_points_subgroup_check():
    """tp_proof_points_in_subgroup: golden proofs pass; a point that is on the curve but outside the prime-order
    subgroup (cofactor component) passes tp_proof_decode's curve check and is caught here."""
    raw = _golden()
    assert ffi.proof_points_in_subgroup(bytes(raw[:1472]))
    # a curve point of the full group: x = small, y = sqrt(x^3 + 4); almost surely not in the r-torsion
    q = fields.Q_MOD
    x = 1
    while True:
        rhs = (x * x * x + 4) % q
        y = pow(rhs, (q + 1) // 4, q)
        if y * y % q == rhs:      # on the curve; inside the subgroup only with probability 1 / cofactor ~ 2^-126
            break
        x += 1
    bad = bytearray(raw[:1472])
    bad[0:96] = x.to_bytes(48, "little") + y.to_bytes(48, "little")
    ffi.proof_decode(bytes(bad) + bytes(raw[1472:]))          # on the curve: decode accepts
    assert not ffi.proof_points_in_subgroup(bytes(bad))


def test_proof_encode_checks_its_arguments():
    raw = _golden()
    with pytest.raises(ffi.TyplonkError):
        ffi.proof_encode(bytes(raw[:1472]), b"\xff" * 32)         # Montgomery limbs >= r
    assert ffi.proof_encode(bytes(raw[:1472]), b"") == bytes(raw[:1472]) + bytes(8)


# ---- SRS codec (device) ------------------------------------------------------------------------------------------
def _g2_wire(pt):
    x, y = pt
    return b"".join(v.to_bytes(48, "little") for v in (x.a, x.b, y.a, y.b))


@pytest.mark.gpu
def test_srs_wire_round_trip_and_oracle_bytes():
    ctx = ffi.Context(0)
    tau = rng.fr_rand_stream(1, 1)[0]
    gates = 29
    from typlonk_b200.kzg import KzgScheme, Srs
    srs = Srs.from_secret(ctx, tau, gates)
    raw = srs.to_bytes()
    osrs = okzg.Srs.from_secret(tau, gates)
    want = struct.pack("<Q", gates + 3) + b"".join(F.g1_serialize_unchecked(p) for p in osrs.g1)
    want += _g2_wire(osrs.g2) + _g2_wire(osrs.g2s)
    assert raw == want
    for check in (0, 1, 2):
        back = Srs.from_bytes(ctx, raw, check)
        assert back.handle.download() == srs.handle.download()
        assert back.handle.g2() == srs.handle.g2()
        poly = rng.fr_rand_stream(9, gates + 3)
        assert KzgScheme(back).commit(poly) == KzgScheme(srs).commit(poly)
    ctx.close()


@pytest.mark.gpu
def test_srs_wire_rejects_bad_points():
    ctx = ffi.Context(0)
    from typlonk_b200.kzg import Srs
    tau = rng.fr_rand_stream(1, 1)[0]
    raw = bytearray(Srs.from_secret(ctx, tau, 61).to_bytes())
    n = 64

    def expect(bad, check, index=None):
        with pytest.raises(ffi.Malformed) as e:
            Srs.from_bytes(ctx, bytes(bad), check)
        if index is not None:
            assert "index %d" % index in str(e.value)

    with pytest.raises(ffi.Malformed):
        Srs.from_bytes(ctx, bytes(raw[:-1]), 0)
    bad = bytearray(raw)
    bad[0:8] = struct.pack("<Q", n + 1)
    expect(bad, 0)
    bad = bytearray(raw)
    off = 8 + 96 * 17
    bad[off] ^= 1                                  # off the curve: passes unchecked, fails check >= 1
    Srs.from_bytes(ctx, bytes(bad), 0)
    expect(bad, 1, 17)
    bad = bytearray(raw)
    bad[off:off + 48] = fields.Q_MOD.to_bytes(48, "little")
    expect(bad, 0, 17)
    bad = bytearray(raw)
    bad[off + 95] |= 0x80
    expect(bad, 0, 17)
    # a point on the curve but outside the prime-order subgroup: (x, y) with x^3 + 4 a square, cofactor not cleared
    x = 0
    while True:
        x += 1
        rhs = (x ** 3 + 4) % fields.Q_MOD
        y = pow(rhs, (fields.Q_MOD + 1) // 4, fields.Q_MOD)   # q = 3 mod 4
        if y * y % fields.Q_MOD == rhs and curve.g1_mul((x, y), fields.R_MOD - 1) != curve.g1_neg((x, y)):
            break
    bad = bytearray(raw)
    off = 8 + 96 * 40
    bad[off:off + 96] = x.to_bytes(48, "little") + y.to_bytes(48, "little")
    Srs.from_bytes(ctx, bytes(bad), 1)
    expect(bad, 2, 40)
    # two bad records: the lowest index is reported
    bad[8 + 96 * 5] ^= 1
    expect(bad, 2, 5)
    # infinity records survive the round trip as the all-zero device record
    ok = bytearray(raw)
    ok[off:off + 96] = F.g1_serialize_unchecked(None)
    s = Srs.from_bytes(ctx, bytes(ok), 2)
    assert s.handle.download(40, 1) == bytes(96)
    assert bytes(s.to_bytes()) == bytes(ok)
    # a G2 point off the curve
    bad = bytearray(raw)
    bad[8 + 96 * n] ^= 1
    expect(bad, 1)
    ctx.close()


@pytest.mark.gpu
def test_srs_wire_at_scale_subgroup_checked():
    """2^16 + 3 points: serialise, deserialise with the full subgroup check, identical device SRS."""
    ctx = ffi.Context(0)
    from typlonk_b200.kzg import Srs
    srs = Srs.from_secret(ctx, rng.fr_rand_stream(1, 1)[0], 1 << 16)
    raw = srs.to_bytes()
    assert len(raw) == 8 + 96 * ((1 << 16) + 3) + 384
    back = Srs.from_bytes(ctx, raw, 2)
    assert back.handle.download() == srs.handle.download()
    ctx.close()


def test_proof_codec_is_canonical_under_random_corruption():
    """Property: whatever bytes come in, tp_proof_decode either rejects them or accepts an encoding that re-encodes to
    exactly the same bytes (one canonical encoding per proof)."""
    import random
    rnd = random.Random(2024)
    accepted = rejected = 0
    for name in sorted(PROOFS):
        raw = bytes.fromhex(PROOFS[name]["proof_hex"])
        for _ in range(60):
            bad = bytearray(raw)
            for _ in range(rnd.choice([1, 1, 2, 5])):
                pos = rnd.randrange(len(bad))
                bad[pos] ^= 1 << rnd.randrange(8)
            try:
                fixed, pis = ffi.proof_decode(bytes(bad))
            except ffi.Malformed:
                rejected += 1
                continue
            accepted += 1
            assert ffi.proof_encode(fixed, pis) == bytes(bad)
    # flips in a scalar's low bits are still canonical scalars (accepted); flips in a point leave the curve (rejected)
    assert accepted > 0 and rejected > 0
