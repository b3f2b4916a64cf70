"""Pins the oracle's primitives against mathematics and published vectors (the reference ships
no golden vectors: SURVEY.md section 8(c))."""
import hashlib
import random

from oracle.py import stark
from oracle.py.transcript import ChaCha20Rng, FiatShamirRng, SeededStream, chacha20_block, fr_rand


def _is_prime(n):
    # deterministic Miller-Rabin for the sizes used here (first 13 primes as bases + random)
    if n < 2:
        return False
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    rnd = random.Random(1)
    for a in [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41] + [rnd.randrange(2, n - 1) for _ in range(8)]:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def test_stark_constants():
    assert stark.P == 2**251 + 17 * 2**192 + 1
    assert _is_prime(stark.P) and _is_prime(stark.N)
    assert stark.P.bit_length() == 252 and stark.N.bit_length() == 252
    assert stark.is_on_curve(stark.G)
    assert stark.mul(stark.G, stark.N - 1) == stark.neg(stark.G)
    # order: (N-1)G + G = O
    assert stark.add(stark.mul(stark.G, stark.N - 1), stark.G) is stark.INF
    # Montgomery constants used by the CUDA field code
    assert (-pow(stark.P, -1, 2**32)) % 2**32 == 0xFFFFFFFF
    assert (1 << 256) % stark.P == 0x07fffffffffffdf0ffffffffffffffffffffffffffffffffffffffffffffffe1
    assert (-pow(stark.N, -1, 2**64)) % 2**64 == 0xbb6b3c4ce8bde631


def test_group_laws():
    rnd = random.Random(2)
    pts = [stark.mul(stark.G, rnd.randrange(1, stark.N)) for _ in range(6)]
    for p in pts:
        assert stark.is_on_curve(p)
        assert stark.add(p, stark.neg(p)) is stark.INF
        assert stark.add(p, stark.INF) == p and stark.add(stark.INF, p) == p
        assert stark.add(p, p) == stark.mul(p, 2)
    a, b, c = pts[:3]
    assert stark.add(stark.add(a, b), c) == stark.add(a, stark.add(b, c))
    assert stark.add(a, b) == stark.add(b, a)
    k1, k2 = rnd.randrange(stark.N), rnd.randrange(stark.N)
    assert stark.add(stark.mul(a, k1), stark.mul(a, k2)) == stark.mul(a, (k1 + k2) % stark.N)
    assert stark.mul(stark.mul(a, k1), k2) == stark.mul(a, k1 * k2 % stark.N)
    assert stark.mul(a, 0) is stark.INF and stark.mul(a, stark.N) is stark.INF
    assert stark.msm([a, b, c], [k1, k2, 1]) == stark.add(stark.add(stark.mul(a, k1), stark.mul(b, k2)), c)


def test_point_bytes_roundtrip():
    p = stark.mul(stark.G, 12345)
    assert stark.point_from_bytes65(stark.point_to_bytes65(p)) == p
    assert stark.point_from_bytes64(stark.point_to_bytes64(p)) == p
    assert stark.point_to_bytes65(stark.INF) == bytes(32) + (1).to_bytes(32, "little") + b"\x01"
    assert stark.point_from_bytes64(bytes(64)) is stark.INF
    assert stark.fe_to_bytes(1) == b"\x01" + bytes(31)


def test_chacha20_rfc7539_block():
    # RFC 7539 section 2.3.2 uses a 32-bit counter + 96-bit nonce; with an all-zero key,
    # counter and nonce every layout coincides: first keystream block of ChaCha20.
    words = chacha20_block([0] * 8, 0)
    ks = b"".join(w.to_bytes(4, "little") for w in words)
    assert ks.hex().startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7")
    # block 1 of the same stream (counter increments in word 12)
    words = chacha20_block([0] * 8, 1)
    ks = b"".join(w.to_bytes(4, "little") for w in words)
    assert ks.hex().startswith("9f07e7be5551387a98ba977c732d080dcb0f29a048e3656912c6533e32ee7aed")


def test_chacha_matches_cryptography_package():
    try:
        from cryptography.hazmat.primitives.ciphers import Cipher, algorithms
    except Exception:  # pragma: no cover
        import pytest
        pytest.skip("cryptography not importable")
    key = bytes(range(32))
    enc = Cipher(algorithms.ChaCha20(key, bytes(16)), mode=None).encryptor()
    want = enc.update(bytes(256))
    rng = ChaCha20Rng(key)
    got = b"".join(rng.next_u32().to_bytes(4, "little") for _ in range(64))
    assert got == want


def test_blake2s_seed_of_shuffle_transcript():
    # Blake2s("Shuffle Proof")
    assert hashlib.blake2s(b"Shuffle Proof").hexdigest() == \
        "99df86eeefd21867b5ea2a0194c5e8dd819aa01221dcdcbc5ff6b16f9303b656"
    # RFC 7693 appendix B: BLAKE2s-256("abc")
    assert hashlib.blake2s(b"abc").hexdigest() == \
        "508c5e8c327c14e2e1a72ba34eeb452f37458b209ed63a294d999b4c86675982"


def test_fiat_shamir_rng_semantics():
    fs = FiatShamirRng()
    assert fs.seed == hashlib.blake2s(b"Shuffle Proof").digest()
    c1 = fs.challenge()
    c2 = fs.challenge()
    assert c1 != c2 and 0 <= c1 < stark.N
    fs2 = FiatShamirRng()
    assert fs2.challenge() == c1          # deterministic
    fs2.absorb(b"hello")
    assert fs2.seed == hashlib.blake2s(b"hello" + hashlib.blake2s(b"Shuffle Proof").digest()).digest()
    assert fs2.challenge() != c2          # stream restarted from the new seed


def test_fr_rand_is_montgomery_interpretation():
    rng = ChaCha20Rng(bytes(32))
    v = fr_rand(rng)
    rng2 = ChaCha20Rng(bytes(32))
    while True:
        limbs = [rng2.next_u64() for _ in range(4)]
        limbs[3] &= (1 << 60) - 1
        raw = sum(l << (64 * i) for i, l in enumerate(limbs))
        if raw < stark.N:
            break
    assert v * (1 << 256) % stark.N == raw


def test_seeded_stream_permutation():
    st = SeededStream(1)
    p = st.permutation(52)
    assert sorted(p) == list(range(52)) and p != list(range(52))
    assert SeededStream(1).permutation(52) == p
    s = SeededStream(1).scalar()
    assert 0 <= s < stark.N
