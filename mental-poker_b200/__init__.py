"""mental-poker_b200: B200-native Bayer-Groth shuffle-proof engine (hot path only).

Host-side mirror of the reference's shuffle surface
(`BarnettSmartProtocol::{setup, shuffle_and_remask, verify_shuffle}`, reference
barnett-smart-card-protocol/src/lib.rs:74-78,181-197) over the C ABI of
`lib/libmpshuffle.so` (include/mpshuffle.h).  The directory name contains a hyphen, so import
it through `__graft_entry__.load_package()` (registers it as `mental_poker_b200`).

There is no CPU fallback: importing works without a GPU (so symbols can be inspected), but
every compute entry point needs a CUDA device and raises otherwise.
"""
from ._lib import lib, lib_path, MpError, check, Context  # noqa: F401
from . import dist  # noqa: F401
from . import bls12_377  # noqa: F401  (second curve: BLS12-377 G1 group layer)
