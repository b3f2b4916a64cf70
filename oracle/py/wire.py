"""ORACLE (test infrastructure only).

ark-serialize 0.3 wire format for the objects that cross the network in a round (SURVEY.md section
8(f) rank 2, Appendix A3; the reference bounds every public type by `CanonicalSerialize +
CanonicalDeserialize`, src/lib.rs:45-71, and measures proof sizes with `serialized_size`,
examples/parameter_selection.rs:95).  The crate is not in the container: the layout below is restated
from recall [UPSTREAM-RECALL] -- PARITY UNPINNED at byte level against upstream.

  field element            32 bytes little-endian canonical
  SW affine, compressed    x (32 bytes LE) with flags in the top bits of the last byte:
                           bit 7 = y is the lexicographically larger of (y, -y), bit 6 = infinity
                           (the Stark prime has 252 bits, so bits 252..255 of x are free)
  Vec<T>                   u64 LE length, then the items
  ciphertext               c1, c2 (two compressed points)

Decompression solves y^2 = x^3 + x + b.  p - 1 = 2^192 * (2^59 + 17): Tonelli-Shanks with a 192-bit
two-adic part, the expensive step the GPU path batches.
"""
from . import stark

P = stark.P
FLAG_LARGER, FLAG_INF = 0x80, 0x40


def compress(pt):
    if pt is stark.INF:
        return bytes(31) + bytes([FLAG_INF])
    x, y = pt
    b = bytearray(stark.fe_to_bytes(x))
    if y > P - y:
        b[31] |= FLAG_LARGER
    return bytes(b)


def decompress(b):
    """-> point; raises ValueError for non-canonical x, stray flag bits or x not on the curve."""
    assert len(b) == 32
    flags = b[31] & 0xC0
    x = int.from_bytes(b[:31] + bytes([b[31] & 0x3F]), "little")
    if flags & FLAG_INF:
        if x != 0 or flags & FLAG_LARGER:
            raise ValueError("bad infinity encoding")
        return stark.INF
    if x >= P:
        raise ValueError("x not canonical")
    y = stark.fq_sqrt((x * x * x + stark.A * x + stark.B) % P)
    if y is None:
        raise ValueError("x is not the abscissa of a curve point")
    if (y > P - y) != bool(flags & FLAG_LARGER):
        y = (P - y) % P
    return (x, y)


def deck_serialize(deck):
    """Vec<MaskedCard>: u64 LE length, then c1, c2 compressed per card."""
    out = len(deck).to_bytes(8, "little")
    for c1, c2 in deck:
        out += compress(c1) + compress(c2)
    return out


def deck_deserialize(b):
    n = int.from_bytes(b[:8], "little")
    if len(b) != 8 + 64 * n:
        raise ValueError("length prefix does not match the buffer")
    return [(decompress(b[8 + 64 * i:40 + 64 * i]), decompress(b[40 + 64 * i:72 + 64 * i])) for i in range(n)]


# ----------------------------------------------------------------------------- second curve (BLS12-377 G1)
def compress_generic(pt, curve):
    """Same encoding over any `weierstrass.Curve`: x in `curve.fe_bytes` little-endian bytes, flags in the top bits of
    the last byte (BLS12-377: 377-bit prime in 48 bytes, bits 377..383 free)."""
    nb = curve.fe_bytes
    if pt is None:
        return bytes(nb - 1) + bytes([FLAG_INF])
    x, y = pt
    b = bytearray(int(x).to_bytes(nb, "little"))
    if y > curve.P - y:
        b[nb - 1] |= FLAG_LARGER
    return bytes(b)


def deck_serialize_generic(deck, curve):
    out = len(deck).to_bytes(8, "little")
    for c1, c2 in deck:
        out += compress_generic(c1, curve) + compress_generic(c2, curve)
    return out


def proof_serialize_generic(flat_proof, m, n, curve):
    """The flat C-ABI proof (points of 2 * fe_bytes, 32-byte scalars) with every point compressed."""
    runs = [(True, 5 * m + 4), (False, 2 * n + 3), (True, 3), (False, 2 * n + 2), (True, 6 * m + 1), (False, n + 4)]
    pb, out, pos = 2 * curve.fe_bytes, b"", 0
    for is_pts, count in runs:
        if is_pts:
            for _ in range(count):
                out += compress_generic(curve.point_from_bytes(flat_proof[pos:pos + pb]), curve)
                pos += pb
        else:
            out += flat_proof[pos:pos + 32 * count]
            pos += 32 * count
    assert pos == len(flat_proof)
    return out


def sqrt_mod(a, p):
    """A square root of a mod the odd prime p (Tonelli-Shanks), or None for a non-residue."""
    a %= p
    if a == 0:
        return 0
    if pow(a, (p - 1) // 2, p) != 1:
        return None
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, r = t * c % p, r * b % p
    return r


def decompress_generic(b, curve, subgroup_check=True):
    """Inverse of compress_generic, validating as ark-serialize 0.3 does: canonical x, no stray flag bits, x on the curve,
    and (for a curve with a cofactor) the point in the order-N subgroup.  Raises ValueError otherwise."""
    nb = curve.fe_bytes
    assert len(b) == nb
    flags = b[nb - 1] & 0xC0
    x = int.from_bytes(b[:nb - 1] + bytes([b[nb - 1] & 0x3F]), "little")
    if flags & FLAG_INF:
        if x != 0 or flags & FLAG_LARGER:
            raise ValueError("bad infinity encoding")
        return None
    if x >= curve.P:
        raise ValueError("x not canonical")
    y = sqrt_mod(x * x * x + curve.A * x + curve.B, curve.P)
    if y is None:
        raise ValueError("x is not the abscissa of a curve point")
    if (y > curve.P - y) != bool(flags & FLAG_LARGER):
        y = (curve.P - y) % curve.P
    if subgroup_check and not in_subgroup((x, y), curve):
        raise ValueError("point outside the prime-order subgroup")
    return (x, y)


def in_subgroup(pt, curve):
    """N * pt == O by plain double-and-add (`curve.mul` reduces its scalar mod N, so it cannot be used for this)."""
    acc = None
    for bit in bin(curve.N)[2:]:
        acc = curve.add(acc, acc)
        if bit == "1":
            acc = curve.add(acc, pt)
    return acc is None


def deck_deserialize_generic(b, curve):
    nb = curve.fe_bytes
    n = int.from_bytes(b[:8], "little")
    if len(b) != 8 + 2 * nb * n:
        raise ValueError("length prefix does not match the buffer")
    at = lambda k: decompress_generic(b[8 + nb * k:8 + nb * (k + 1)], curve)
    return [(at(2 * i), at(2 * i + 1)) for i in range(n)]
