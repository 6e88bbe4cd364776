"""ORACLE (test infrastructure only).

Fiat-Shamir challenge derivation of the reference, plonk/src/proof/challenges.rs:9-46:
  transcript bytes = serialize_unchecked(commitments...)      (:17-22)
  Blake2b-512 -> first 8 bytes LE -> u64 seed                 (:31-38)
  StdRng::seed_from_u64(seed)  (rand 0.8.4: PCG32 seed expansion -> ChaCha12)
  Fr::rand x N                 (ark-ff 0.3.0 `Standard` sampling for Fp256)

None of rand / rand_chacha / ark-ff is vendored in the reference tree; the
algorithms are restated from their published definitions.  The ChaCha core is
pinned against the RFC 7539-style zero-key vectors; the rest of the chain is
"spec-from-memory"  and documented as UNPINNED.
"""
import hashlib
import struct

from .fields import R_MOD, fr_from_mont

_MASK32 = 0xFFFFFFFF
_MASK64 = 0xFFFFFFFFFFFFFFFF


def _rotl32(x, n):
    return ((x << n) | (x >> (32 - n))) & _MASK32


def _quarter(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & _MASK32; s[d] = _rotl32(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & _MASK32; s[b] = _rotl32(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & _MASK32; s[d] = _rotl32(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & _MASK32; s[b] = _rotl32(s[b] ^ s[c], 7)


def chacha_block(key: bytes, counter: int, stream: int = 0, rounds: int = 12):
    """One 64-byte ChaCha block as 16 u32 words.  rand_chacha layout: words 12-13 =
    64-bit block counter, words 14-15 = 64-bit stream id."""
    const = (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)
    k = struct.unpack("<8I", key)
    init = list(const) + list(k) + [counter & _MASK32, (counter >> 32) & _MASK32,
                                    stream & _MASK32, (stream >> 32) & _MASK32]
    s = list(init)
    for _ in range(rounds // 2):
        _quarter(s, 0, 4, 8, 12); _quarter(s, 1, 5, 9, 13)
        _quarter(s, 2, 6, 10, 14); _quarter(s, 3, 7, 11, 15)
        _quarter(s, 0, 5, 10, 15); _quarter(s, 1, 6, 11, 12)
        _quarter(s, 2, 7, 8, 13); _quarter(s, 3, 4, 9, 14)
    return [(x + y) & _MASK32 for x, y in zip(s, init)]


def seed_from_u64_key(state: int) -> bytes:
    """rand_core 0.6 `SeedableRng::seed_from_u64`: PCG32 expands the u64 to 32 bytes."""
    MUL = 6364136223846793005
    INC = 11634580027462260723
    out = b""
    for _ in range(8):
        state = (state * MUL + INC) & _MASK64
        xorshifted = (((state >> 18) ^ state) >> 27) & _MASK32
        rot = state >> 59
        x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & _MASK32
        out += struct.pack("<I", x)
    return out


class StdRng:
    """rand 0.8 `StdRng` = ChaCha12Rng; words are consumed sequentially, next_u64 =
    lo | hi << 32 (challenges.rs:38 `StdRng::seed_from_u64`)."""

    def __init__(self, key: bytes):
        self.key = key
        self.counter = 0
        self.buf = []

    @classmethod
    def seed_from_u64(cls, seed: int) -> "StdRng":
        return cls(seed_from_u64_key(seed))

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = chacha_block(self.key, self.counter, 0, 12)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)


def fr_rand(rng: StdRng) -> int:
    """ark-ff 0.3 `Fp256::rand`: draw 4 u64 limbs (limb 0 first), clear the top
    REPR_SHAVE_BITS = 1 bit, accept if < r; the accepted integer IS the Montgomery
    representation.  Returns the canonical value.  (challenges.rs:44, proof.rs:46,
    srs.rs:38)"""
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= _MASK64 >> 1
        v = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
        if v < R_MOD:
            return fr_from_mont(v)


def fr_rand_stream(seed: int, count: int):
    """`count` x Fr::rand from StdRng::seed_from_u64(seed) (SURVEY.md 8(d) inputs)."""
    rng = StdRng.seed_from_u64(seed)
    return [fr_rand(rng) for _ in range(count)]


def challenge_seed(transcript: bytes) -> int:
    """Blake2b-512 of the transcript, first 8 bytes little-endian (challenges.rs:31-38)."""
    return int.from_bytes(hashlib.blake2b(transcript).digest()[:8], "little")


def generate_challenges(transcript: bytes, n: int):
    """`ChallengeGenerator::generate_challenges::<N>` (challenges.rs:40-45)."""
    rng = StdRng.seed_from_u64(challenge_seed(transcript))
    return [fr_rand(rng) for _ in range(n)]
