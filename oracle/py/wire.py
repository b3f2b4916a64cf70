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
