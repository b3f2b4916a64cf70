"""ORACLE (test infrastructure only -- never imported by the product path).

BLS12-377 G1: the curve the reference's own benchmark harness instantiates the protocol over
(`type Curve = ark_bls12_377::G1Projective`, reference
barnett-smart-card-protocol/examples/parameter_selection.rs:25-26).  Constants restated from
the published curve definition (ark-bls12-377 0.3, not in this container) and pinned against
mathematics here and in tests/test_oracle_bls12_377.py: with the BLS parameter
x = 0x8508c00000000001,  r = x^4 - x^2 + 1  and  q = (x - 1)^2 * r / 3 + x  (both prime),
G on y^2 = x^3 + 1, r * G = O, cofactor (x - 1)^2 / 3.

Field element = 48 bytes little-endian canonical; scalar = 32 bytes; C-ABI point = x || y
(96 bytes, all-zero = identity).
"""
from .weierstrass import Curve

X_PARAM = 0x8508C00000000001
Q = 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
R = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
GX = 0x008848DEFE740A67C8FC6225BF87FF5485951E2CAA9D41BB188282C8BD37CB5CD5481512FFCD394EEAB9B16EB21BE9EF
GY = 0x01914A69C5102EFF1F674F5D30AFEEC4BD7FB348CA3E52D96D182AD44FB82305C2FE3D3634A9591AFD82DE55559C8EA6
COFACTOR = 0x170B5D44300000000000000000000000

CURVE = Curve("bls12_377_g1", Q, R, 0, 1, (GX, GY), fe_bytes=48)

# module-level aliases with the same names stark.py exports
P, N, A, B, G, INF = Q, R, 0, 1, CURVE.G, None
is_on_curve, neg, add, sub, mul, msm = CURVE.is_on_curve, CURVE.neg, CURVE.add, CURVE.sub, CURVE.mul, CURVE.msm
fe_to_bytes, scalar_to_bytes = CURVE.fe_to_bytes, CURVE.scalar_to_bytes
point_to_bytes, point_from_bytes = CURVE.point_to_bytes, CURVE.point_from_bytes
