"""ORACLE (test infrastructure only).

Fiat-Shamir transcript of the shuffle argument: ark-marlin 0.3 `FiatShamirRng<Blake2s>`
as used by the reference at barnett-smart-card-protocol/src/discrete_log_cards/mod.rs:408,436
(`FiatShamirRng::<Blake2s>::from_seed(&to_bytes![SHUFFLE_RNG_SEED]?)`, seed string mod.rs:84).

Restated from SURVEY.md Appendix A4 (FiatShamirRng), A5 (rand_chacha ChaCha20) and A1
(ark-ff 0.3 `UniformRand for Fp256`): the upstream crates are not present in this
environment -- PARITY UNPINNED at byte level.  ChaCha20 and Blake2s themselves are pinned to
RFC 7539 / RFC 7693 vectors in tests/test_oracle_math.py.
"""
import hashlib
import struct

from . import stark

SHUFFLE_RNG_SEED = b"Shuffle Proof"  # mod.rs:84


def _rotl(v, c):
    return ((v << c) & 0xFFFFFFFF) | (v >> (32 - c))


def chacha20_block(key_words, counter):
    """One 64-byte ChaCha20 block; 64-bit block counter in words 12..13, stream id 0
    (rand_chacha layout)."""
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [
        counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, 0, 0]
    x = st[:]

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF
        x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF
        x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF
        x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF
        x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & 0xFFFFFFFF for i in range(16)]


class ChaCha20Rng:
    """rand_chacha 0.3 `ChaCha20Rng::from_seed(seed)`: keystream words in order."""

    def __init__(self, seed32):
        assert len(seed32) == 32
        self.key = struct.unpack("<8I", seed32)
        self.counter = 0
        self.buf = []

    def next_u32(self):
        if not self.buf:
            self.buf = chacha20_block(self.key, self.counter)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self):
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)


# scalar field the challenges are drawn from: the Stark curve's by default; `bayer_groth.curve(...)`
# switches it together with the group (second instantiation: ark_bls12_377::Fr, 253 bits)
SCALAR_MODULUS = stark.N


def fr_rand(rng):
    """ark-ff 0.3 `Fp256::rand`: draw 4 x u64 (limb 0 first) as the raw
    Montgomery representation, clear the top REPR_SHAVE_BITS = 256 - bitlen(modulus) bits
    (4 for the Stark scalar field, 3 for BLS12-377's), accept iff < modulus.
    Returns the canonical value raw * R^-1 mod n."""
    n = SCALAR_MODULUS
    rinv = pow(1 << 256, -1, n)
    shave = 256 - n.bit_length()
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= 0xFFFFFFFFFFFFFFFF >> shave
        raw = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
        if raw < n:
            return raw * rinv % n


class FiatShamirRng:
    """ark-marlin 0.3 `FiatShamirRng<Blake2s>`."""

    def __init__(self, seed_bytes=SHUFFLE_RNG_SEED):
        self.seed = hashlib.blake2s(seed_bytes).digest()
        self.rng = ChaCha20Rng(self.seed)

    def absorb(self, data):
        self.seed = hashlib.blake2s(bytes(data) + self.seed).digest()
        self.rng = ChaCha20Rng(self.seed)

    def challenge(self):
        return fr_rand(self.rng)


class SeededStream:
    """Synthetic-input PRNG of SURVEY.md section 8(d): ChaCha20 keyed by a u64 seed
    (LE, zero padded to 32 bytes).  `scalar()` = uniform in [0, n) by 252-bit
    mask-and-reject of 32 keystream bytes (little-endian)."""

    def __init__(self, seed):
        self.rng = ChaCha20Rng(int(seed).to_bytes(8, "little") + bytes(24))

    def scalar(self):
        while True:
            limbs = [self.rng.next_u64() for _ in range(4)]
            v = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
            v &= (1 << 252) - 1
            if v < stark.N:
                return v

    def below(self, bound):
        """uniform integer in [0, bound) by rejection on next_u64."""
        lim = (1 << 64) - ((1 << 64) % bound)
        while True:
            v = self.rng.next_u64()
            if v < lim:
                return v % bound

    def permutation(self, size):
        """Fisher-Yates: for i = size-1 .. 1 swap(i, below(i+1))."""
        perm = list(range(size))
        for i in range(size - 1, 0, -1):
            j = self.below(i + 1)
            perm[i], perm[j] = perm[j], perm[i]
        return perm
