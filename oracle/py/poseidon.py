"""ORACLE (test infrastructure only — never imported by the product path).

Poseidon x^5 permutation/hash over BLS12-381 Fr, widths t = 2..17.

Follows /root/reference/src/zk/poseidon/mod.rs:
  :15-22  state = [0] ++ inputs (zero "capacity" lane 0)
  :24-42  hash(): R_F/2 full rounds, R_P partial rounds (S-box on lane 0), R_F/2 full rounds;
          digest = lane 1
  :56-61  round constants consumed strictly sequentially, t per round
  :63-71  dense MDS product  new[j] = sum_k mds[j][k] * old[k]
  :74-79  S-box x -> x^5
and params/mod.rs:39-57 for which file lines hold R_F, R_P, constants and the MDS matrix.
Pinned by the 16 known answers of `test_hash_samples` (mod.rs:115-149) — see tests/test_oracle_py.py.
"""
import os
import struct
from .field import R_MOD

_PARAMS = None
PARAMS_PATH = os.path.join(os.path.dirname(__file__), "..", "..", "bazuka_b200", "data", "poseidon_params.bin")

MAX_ARITY = 16


def load_params(path=PARAMS_PATH):
    """-> {t: (R_F, R_P, round_constants[int], mds[t][t])} from the table made by
    tools/extract_poseidon_params.py."""
    global _PARAMS
    if _PARAMS is not None:
        return _PARAMS
    blob = open(path, "rb").read()
    assert blob[:8] == b"BZKPOSv1"
    (n,) = struct.unpack_from("<I", blob, 8)
    off = 12
    out = {}
    for _ in range(n):
        t, rf, rp, nrc = struct.unpack_from("<IIII", blob, off)
        off += 16
        rc = [int.from_bytes(blob[off + 32 * i : off + 32 * i + 32], "little") for i in range(nrc)]
        off += 32 * nrc
        m = [int.from_bytes(blob[off + 32 * i : off + 32 * i + 32], "little") for i in range(t * t)]
        off += 32 * t * t
        out[t] = (rf, rp, rc, [m[j * t : (j + 1) * t] for j in range(t)])
    assert off == len(blob)
    _PARAMS = out
    return out


def permute(state):
    """the full permutation on a width-t state (returns the new state)."""
    t = len(state)
    rf, rp, rc, mds = load_params()[t]
    s = [x % R_MOD for x in state]
    off = 0
    for rnd in range(rf + rp):
        s = [(x + rc[off + i]) % R_MOD for i, x in enumerate(s)]
        off += t
        if rnd < rf // 2 or rnd >= rf // 2 + rp:
            s = [pow(x, 5, R_MOD) for x in s]
        else:
            s[0] = pow(s[0], 5, R_MOD)
        s = [sum(mds[j][k] * s[k] for k in range(t)) % R_MOD for j in range(t)]
    return s


def poseidon(vals):
    """`poseidon::poseidon(vals)` — arity len(vals) in 1..16."""
    assert 1 <= len(vals) <= MAX_ARITY
    return permute([0] + list(vals))[1]
