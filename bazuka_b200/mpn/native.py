"""Native (non-circuit) arithmetic the witness builder needs, on Python integers.

Follows /root/reference/src/zk/poseidon/mod.rs:24-84 (Poseidon), /root/reference/src/crypto/jubjub/
curve.rs:19-164 + mod.rs:108-168 (twisted Edwards curve, EdDSA-Poseidon), /root/reference/src/zk/mod.rs:
262-271 (`ZkScalar::new`), /root/reference/src/zk/state/mod.rs:218-264,310-420 (4-ary Merkle state:
node = Poseidon-4 of its children, missing = level default, proofs leaf-first with self skipped)."""
import functools
import hashlib
import os
import struct

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_PARAMS = None
_PARAMS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "poseidon_params.bin")


def poseidon_params():
    global _PARAMS
    if _PARAMS is None:
        blob = open(_PARAMS_PATH, "rb").read()
        assert blob[:8] == b"BZKPOSv1"
        (n,) = struct.unpack_from("<I", blob, 8)
        off, out = 12, {}
        for _ in range(n):
            t, rf, rp, nrc = struct.unpack_from("<IIII", blob, off)
            off += 16
            rc = [int.from_bytes(blob[off + 32 * i: off + 32 * i + 32], "little") for i in range(nrc)]
            off += 32 * nrc
            m = [int.from_bytes(blob[off + 32 * i: off + 32 * i + 32], "little") for i in range(t * t)]
            off += 32 * t * t
            out[t] = (rf, rp, rc, [m[j * t:(j + 1) * t] for j in range(t)])
        _PARAMS = out
    return _PARAMS


def poseidon(vals):
    t = len(vals) + 1
    rf, rp, rc, mds = poseidon_params()[t]
    s = [0] + [v % R for v in vals]
    off = 0
    for rnd in range(rf + rp):
        s = [(x + rc[off + i]) % R for i, x in enumerate(s)]
        off += t
        if rnd < rf // 2 or rnd >= rf // 2 + rp:
            s = [pow(x, 5, R) for x in s]
        else:
            s[0] = pow(s[0], 5, R)
        s = [sum(mds[j][k] * s[k] for k in range(t)) % R for j in range(t)]
    return s[1]


def hash_to_scalar(data: bytes) -> int:
    """`hash_to_scalar` = ZkScalar::new(sha3_256(data)) (src/zk/mod.rs:218-220,262-271)."""
    return int.from_bytes(hashlib.sha3_256(data).digest(), "little") % R


def fr_sqrt(a):
    """a square root of a in Fr (Tonelli-Shanks; r - 1 = 2^32 * odd), or None."""
    a %= R
    if a == 0:
        return 0
    if pow(a, (R - 1) // 2, R) != 1:
        return None
    q, s = R - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 7  # generator of Fr^* (src/zk/mod.rs:204) is a non-residue
    m, c, t, r_ = s, pow(z, q, R), pow(a, q, R), pow(a, (q + 1) // 2, R)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % R
            i += 1
        b = pow(c, 1 << (m - i - 1), R)
        m, c = i, b * b % R
        t, r_ = t * c % R, r_ * b % R
    return r_


# ---------------------------------------------------------------- JubJub (a = -1, d below)
JJ_A = R - 1
JJ_D = 19257038036680949359750312669786877991949435402254120286184196891950884077233
JJ_BASE = (28867639725710769449342053336011988556061781325688749245863888315629457631946, 18)
JJ_ORDER = 6554484396890773809930967563523245729705921265872317281365359162392183254199


def jj_on_curve(p):
    x, y = p
    return (y * y - x * x) % R == (1 + JJ_D * x * x % R * y * y) % R


def jj_double(p):
    x, y = p
    xx = pow((JJ_A * x * x + y * y) % R, -1, R)
    yy = pow((2 - JJ_A * x * x - y * y) % R, -1, R)
    return (2 * x * y % R * xx % R, (y * y - JJ_A * x * x) % R * yy % R)


def jj_add(p, q):
    """PointAffine::add_assign (curve.rs:19-34)."""
    if p == q:
        return jj_double(p)
    x1, y1 = p
    x2, y2 = q
    k = JJ_D * x1 % R * x2 % R * y1 % R * y2 % R
    xx, yy = pow((1 + k) % R, -1, R), pow((1 - k) % R, -1, R)
    return ((x1 * y2 + y1 * x2) % R * xx % R, (y1 * y2 - JJ_A * x1 * x2) % R * yy % R)


def jj_mul(p, k):
    """PointAffine::multiply: MSB-first double-and-add over the 256 LE bits (curve.rs:58-68);
    the identity is (0, 1)."""
    acc = (0, 1)
    for i in range(255, -1, -1):
        acc = jj_add(acc, acc) if acc != (0, 1) else acc
        if (k >> i) & 1:
            acc = jj_add(acc, p) if acc != (0, 1) else p
    return acc


JJ_BASE_COFACTOR = jj_mul(JJ_BASE, 8)


def jj_compress(p):
    return (p[0], p[1] & 1 == 1)


@functools.lru_cache(maxsize=65536)
def jj_decompress(c):
    """PointCompressed::decompress (curve.rs:78-88).  Memoised: the Fr square root is ~20 modular powers and the
    same few thousand keys recur in every batch."""
    x, odd = c
    y = fr_sqrt((1 - JJ_A * x * x) % R * pow((1 - JJ_D * x * x) % R, -1, R) % R)
    assert y is not None
    if (y & 1 == 1) != odd:
        y = (-y) % R
    return (x, y)


def jj_decompress_checked(c):
    """`PublicKey::is_on_curve` (src/crypto/jubjub/mod.rs:71-73) as the builders use it: the decompressed point, or None
    when x is not the abscissa of a curve point (the reference's `.sqrt().unwrap()` would panic there)."""
    x, odd = c
    if not 0 <= x < R:
        return None
    den = (1 - JJ_D * x * x) % R
    if den == 0:
        return None
    y = fr_sqrt((1 - JJ_A * x * x) % R * pow(den, -1, R) % R)
    if y is None:
        return None
    if (y & 1 == 1) != odd:
        y = (-y) % R
    return (x, y)


def eddsa_keys(seed: bytes):
    """JubJub::generate_keys (mod.rs:112-125) -> (pk_affine, sk dict)."""
    randomness = hash_to_scalar(seed)
    scalar = hash_to_scalar(randomness.to_bytes(32, "little"))
    point = jj_mul(JJ_BASE, scalar)
    return point, {"public_key": point, "randomness": randomness, "scalar": scalar}


def eddsa_sign(sk, message):
    """JubJub::sign (mod.rs:126-150)."""
    r = poseidon([sk["randomness"], message])
    rr = jj_mul(JJ_BASE, r)
    h = poseidon([rr[0], rr[1], sk["public_key"][0], sk["public_key"][1], message])
    s = (r + h * sk["scalar"]) % JJ_ORDER
    return {"r": rr, "s": s}


def eddsa_verify(pk, message, sig):
    """JubJub::verify (mod.rs:151-167), pk affine."""
    if not jj_on_curve(pk) or not jj_on_curve(sig["r"]):
        return False
    h = poseidon([sig["r"][0], sig["r"][1], pk[0], pk[1], message])
    return jj_add(jj_mul(pk, h), sig["r"]) == jj_mul(JJ_BASE, sig["s"])


# ---------------------------------------------------------------- sparse 4-ary Merkle tree
class SparseTree4:
    """`List {log4_size, item}` of a ZkStateModel as the state manager keeps it: node = Poseidon-4 of
    its 4 children, absent = the level's default (compress_default, src/zk/mod.rs:401-423)."""

    def __init__(self, log4_size, default_leaf):
        self.depth = log4_size
        self.defaults = [default_leaf]
        for _ in range(log4_size):
            d = self.defaults[-1]
            self.defaults.append(poseidon([d, d, d, d]))
        self.levels = [dict() for _ in range(log4_size + 1)]  # level 0 = leaves

    def get(self, level, idx):
        return self.levels[level].get(idx, self.defaults[level])

    @property
    def root(self):
        return self.get(self.depth, 0)

    def set_leaf(self, index, value):
        self._put(0, index, value)
        idx = index
        for lvl in range(self.depth):
            parent = idx >> 2
            kids = [self.get(lvl, 4 * parent + k) for k in range(4)]
            self._put(lvl + 1, parent, poseidon(kids))
            idx = parent

    def _put(self, level, idx, value):
        if value == self.defaults[level]:
            self.levels[level].pop(idx, None)  # defaults are deleted, not stored (state/mod.rs:385-389)
        else:
            self.levels[level][idx] = value

    def prove(self, index):
        """leaf-first list of [3 siblings] in ascending child order, self skipped (state/mod.rs:233-258)."""
        out, idx = [], index
        for lvl in range(self.depth):
            base = (idx >> 2) << 2
            out.append([self.get(lvl, base + k) for k in range(4) if base + k != idx])
            idx >>= 2
        return out


def calc_root(index, value, proof):
    """what the merkle gadget recomputes (gadgets/merkle/mod.rs:53-65): child position = 2 index bits."""
    cur, idx = value, index
    for sib in proof:
        pos = idx & 3
        kids = list(sib)
        kids.insert(pos, cur)
        cur = poseidon(kids)
        idx >>= 2
    return cur
