"""ORACLE (test infrastructure only).

The sigma protocols either side of the shuffle in a Barnett-Smart round (SURVEY.md section 8(f),
rank 1), behind the reference's glue in barnett-smart-card-protocol/src/discrete_log_cards/mod.rs:

  prove_key_ownership / verify_key_ownership   mod.rs:132-165   Schnorr identification
  mask / verify_mask                           mod.rs:182-240   Chaum-Pedersen DL equality
  remask / verify_remask                       mod.rs:242-299   Chaum-Pedersen DL equality
  compute_reveal_token / verify_reveal         mod.rs:301-354   Chaum-Pedersen DL equality
  Fiat-Shamir seeds                            mod.rs:80-83

The bodies of `schnorr_identification::SchnorrIdentification::{prove,verify}` and
`chaum_pedersen_dl_equality::DLEquality::{prove,verify}` live in the un-vendored, unpinned git
dependency `proof-essentials` (Cargo.toml:18), absent here.  They are restated as the textbook
protocols in the shape the reference's call sites fix (parameters, statement, witness, one
transcript absorb, one challenge), marked [UPSTREAM-RECALL]; the transcript byte order is THIS
repository's definition -- PARITY UNPINNED at byte level against upstream.  What the reference's
own tests pin (masking.rs:64-107, remasking.rs:65-114, reveal.rs:43-84, tests.rs:48-78) is
behavioural: prove -> verify == Ok, and a wrong statement fails with "Chaum-Pedersen" /
"Schnorr Identification"; tests/test_oracle_sigma.py reproduces exactly that.

Point encoding inside an absorb: ark-ec 0.3 `ToBytes` for affine points, 65 bytes (as in
oracle/py/transcript.py); labels are raw ASCII.
"""
import contextlib

from . import stark
from .stark import N as Q
from .transcript import FiatShamirRng

KEY_OWN_RNG_SEED = b"Key Ownership Proof"   # mod.rs:80
MASKING_RNG_SEED = b"Masking Proof"         # mod.rs:81
REMASKING_RNG_SEED = b"Remasking Proof"     # mod.rs:82
REVEAL_RNG_SEED = b"Reveal Proof"           # mod.rs:83

OK = 0
ERR_CHAUM_PEDERSEN = 5        # "Chaum-Pedersen"           (masking.rs:103-105)
ERR_SCHNORR = 6               # "Schnorr Identification"   (tests.rs:72-77)
ERR_STRINGS = {ERR_CHAUM_PEDERSEN: "Chaum-Pedersen", ERR_SCHNORR: "Schnorr Identification"}

P65 = stark.point_to_bytes65


@contextlib.contextmanager
def curve(name):
    """`with curve("bls12_377"): ...` runs these protocols over BLS12-377 G1 (group, challenge field, byte widths), the
    reference's second instantiation (examples/parameter_selection.rs:25-29).  Test infrastructure: not re-entrant."""
    global stark, Q, P65
    from . import bayer_groth as bg
    saved = (stark, Q, P65)
    with bg.curve(name) as grp:
        stark, Q, P65 = grp, grp.N, grp.point_to_bytes65
        try:
            yield grp
        finally:
            stark, Q, P65 = saved


# ----------------------------------------------------------------------------- the two protocols
def schnorr_prove(g, pk, sk, omega, seed):
    """[UPSTREAM-RECALL] commit = omega*g; c = H(label, g, pk, commit); opening = omega - c*sk."""
    commit = stark.mul(g, omega)
    fs = FiatShamirRng(seed)
    fs.absorb(b"schnorr_identity" + P65(g) + P65(pk) + P65(commit))
    c = fs.challenge()
    return commit, (omega - c * sk) % Q


def schnorr_verify(g, pk, proof, seed):
    """opening*g + c*pk == commit, else "Schnorr Identification" (tests.rs:72-77)."""
    commit, opening = proof
    fs = FiatShamirRng(seed)
    fs.absorb(b"schnorr_identity" + P65(g) + P65(pk) + P65(commit))
    c = fs.challenge()
    lhs = stark.add(stark.mul(g, opening), stark.mul(pk, c))
    return OK if lhs == commit else ERR_SCHNORR


def cp_challenge(g, h, s0, s1, a, b, seed):
    fs = FiatShamirRng(seed)
    fs.absorb(b"chaum_pedersen" + P65(g) + P65(h) + P65(s0) + P65(s1) + P65(a) + P65(b))
    return fs.challenge()


def cp_prove(g, h, s0, s1, x, omega, seed):
    """[UPSTREAM-RECALL] statement (s0, s1) = (x*g, x*h): a = omega*g, b = omega*h,
    c = H(label, g, h, s0, s1, a, b), r = omega + c*x."""
    a, b = stark.mul(g, omega), stark.mul(h, omega)
    c = cp_challenge(g, h, s0, s1, a, b, seed)
    return a, b, (omega + c * x) % Q


def cp_verify(g, h, s0, s1, proof, seed):
    """r*g == a + c*s0 and r*h == b + c*s1, else "Chaum-Pedersen"."""
    a, b, r = proof
    c = cp_challenge(g, h, s0, s1, a, b, seed)
    ok = stark.mul(g, r) == stark.add(a, stark.mul(s0, c)) and stark.mul(h, r) == stark.add(b, stark.mul(s1, c))
    return OK if ok else ERR_CHAUM_PEDERSEN


# ----------------------------------------------------------------------------- reference glue
def prove_key_ownership(g, pk, sk, info, omega):
    """mod.rs:132-148: seed = to_bytes![KEY_OWN_RNG_SEED, player_public_info]."""
    return schnorr_prove(g, pk, sk, omega, KEY_OWN_RNG_SEED + bytes(info))


def verify_key_ownership(g, pk, info, proof):
    """mod.rs:150-165."""
    return schnorr_verify(g, pk, proof, KEY_OWN_RNG_SEED + bytes(info))


def mask(g, shared_key, card, r, omega):
    """mod.rs:182-214 with masking.rs:10-20: masked = (r*g, card + r*pk); Chaum-Pedersen over
    parameters (g, pk), statement (masked.0, masked.1 - card), witness r."""
    masked = (stark.mul(g, r), stark.add(card, stark.mul(shared_key, r)))
    s1 = stark.sub(masked[1], card)
    return masked, cp_prove(g, shared_key, masked[0], s1, r, omega, MASKING_RNG_SEED)


def verify_mask(g, shared_key, card, masked, proof):
    """mod.rs:216-240."""
    return cp_verify(g, shared_key, masked[0], stark.sub(masked[1], card), proof, MASKING_RNG_SEED)


def remask(g, shared_key, original, alpha, omega):
    """mod.rs:242-272 with remasking.rs:9-22: remasked = original + (alpha*g, alpha*pk); statement =
    remasked - original."""
    remasked = (stark.add(original[0], stark.mul(g, alpha)), stark.add(original[1], stark.mul(shared_key, alpha)))
    s0, s1 = stark.sub(remasked[0], original[0]), stark.sub(remasked[1], original[1])
    return remasked, cp_prove(g, shared_key, s0, s1, alpha, omega, REMASKING_RNG_SEED)


def verify_remask(g, shared_key, original, remasked, proof):
    """mod.rs:274-299."""
    s0, s1 = stark.sub(remasked[0], original[0]), stark.sub(remasked[1], original[1])
    return cp_verify(g, shared_key, s0, s1, proof, REMASKING_RNG_SEED)


def compute_reveal_token(g, sk, pk, masked, omega):
    """mod.rs:301-328: token = sk*masked.0; parameters (masked.0, g), statement (token, pk), witness sk."""
    token = stark.mul(masked[0], sk)
    return token, cp_prove(masked[0], g, token, pk, sk, omega, REVEAL_RNG_SEED)


def verify_reveal(g, pk, token, masked, proof):
    """mod.rs:330-354."""
    return cp_verify(masked[0], g, token, pk, proof, REVEAL_RNG_SEED)


# ----------------------------------------------------------------------------- byte formats (C ABI)
def cp_proof_bytes(proof):
    """160 bytes: a (64) | b (64) | r (32), canonical little-endian."""
    return stark.point_to_bytes64(proof[0]) + stark.point_to_bytes64(proof[1]) + stark.fe_to_bytes(proof[2])


def cp_proof_from_bytes(b):
    return stark.point_from_bytes64(b[:64]), stark.point_from_bytes64(b[64:128]), stark.fe_from_bytes(b[128:160])


def schnorr_proof_bytes(proof):
    """96 bytes: commit (64) | opening (32)."""
    return stark.point_to_bytes64(proof[0]) + stark.fe_to_bytes(proof[1])


def schnorr_proof_from_bytes(b):
    return stark.point_from_bytes64(b[:64]), stark.fe_from_bytes(b[64:96])
