"""ORACLE (test infrastructure only — never imported by the product path).

BLS12-381 G1 / G2 group law, the (x,y,inf) Montgomery wire images and a plain optimal-ate pairing,
all on Python big integers.  Slow, obviously-correct arbiter for the C oracle and the CUDA path.

Follows:
  * wire structs `(Fp,Fp,bool)` / `((Fp,Fp),(Fp,Fp),bool)`:
      /root/reference/src/zk/groth16/mod.rs:19-38 (transmuted images of bls12_381::G1Affine/G2Affine)
  * verifier equation used by `groth16_verify`: /root/reference/src/zk/groth16/mod.rs:67-121
    (bellman 0.14.0 `verify_proof`: e(A,B) = e(alpha,beta) * e(sum x_i ic_i, gamma) * e(C, delta)).
bls12_381 0.8.0 and bellman 0.14.0 are un-vendored crates.io dependencies (Cargo.toml:27-28); the
curve (y^2 = x^3 + 4 over Fp, twist y^2 = x^3 + 4(u+1) over Fp2, BLS parameter
x = -0xd201000000010000) is the published BLS12-381 definition restated here.
"""
from .field import P_MOD, R_MOD, fp_to_mont_bytes, fp_from_mont_bytes

P = P_MOD
BLS_X = 0xD201000000010000  # |x|; the curve parameter is -BLS_X

G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G2_GEN = (
    (
        0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
        0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
    ),
    (
        0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
        0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE,
    ),
)


# ------------------------------------------------------------------ Fp2 = Fp[u]/(u^2+1)
def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_sqr(a):
    return f2_mul(a, a)


def f2_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
    return (a[0] * n % P, (-a[1]) * n % P)


def f2_muls(a, s):
    return (a[0] * s % P, a[1] * s % P)


F2_ZERO = (0, 0)
F2_ONE = (1, 0)
B1 = 4  # G1: y^2 = x^3 + 4
B2 = (4, 4)  # G2: y^2 = x^3 + 4(u+1)


class _Fld:
    """Minimal field vtable so one group-law implementation serves G1 (Fp) and G2 (Fp2)."""

    def __init__(self, add, sub, neg, mul, inv, zero, one, b):
        self.add, self.sub, self.neg, self.mul, self.inv = add, sub, neg, mul, inv
        self.zero, self.one, self.b = zero, one, b


FP = _Fld(
    lambda a, b: (a + b) % P,
    lambda a, b: (a - b) % P,
    lambda a: (-a) % P,
    lambda a, b: a * b % P,
    lambda a: pow(a, -1, P),
    0,
    1,
    B1,
)
FP2 = _Fld(f2_add, f2_sub, f2_neg, f2_mul, f2_inv, F2_ZERO, F2_ONE, B2)


# ------------------------------------------------------------------ affine group law (None = identity)
def on_curve(F, pt):
    if pt is None:
        return True
    x, y = pt
    return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), F.b)


def neg(F, pt):
    return None if pt is None else (pt[0], F.neg(pt[1]))


def add(F, p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if y1 == y2:
            if y1 == F.zero:
                return None
            three = F.add(F.add(F.one, F.one), F.one)
            lam = F.mul(F.mul(three, F.mul(x1, x1)), F.inv(F.add(y1, y1)))
        else:
            return None
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


# Jacobian arithmetic (used for anything longer than a handful of additions)
def _jdbl(F, p):
    X, Y, Z = p
    if Z == F.zero:
        return p
    A = F.mul(X, X)
    B = F.mul(Y, Y)
    C = F.mul(B, B)
    t = F.add(X, B)
    D = F.sub(F.sub(F.mul(t, t), A), C)
    D = F.add(D, D)
    E = F.add(F.add(A, A), A)
    Fq = F.mul(E, E)
    X3 = F.sub(Fq, F.add(D, D))
    C8 = F.add(C, C)
    C8 = F.add(C8, C8)
    C8 = F.add(C8, C8)
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
    Z3 = F.mul(F.add(Y, Y), Z)
    return (X3, Y3, Z3)


def _jadd(F, p, q):
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    if Z1 == F.zero:
        return q
    if Z2 == F.zero:
        return p
    Z1Z1 = F.mul(Z1, Z1)
    Z2Z2 = F.mul(Z2, Z2)
    U1 = F.mul(X1, Z2Z2)
    U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
    S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
    if U1 == U2:
        if S1 == S2:
            return _jdbl(F, p)
        return (F.one, F.one, F.zero)
    H = F.sub(U2, U1)
    Rr = F.sub(S2, S1)
    HH = F.mul(H, H)
    HHH = F.mul(H, HH)
    V = F.mul(U1, HH)
    X3 = F.sub(F.sub(F.mul(Rr, Rr), HHH), F.add(V, V))
    Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
    Z3 = F.mul(F.mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def to_jac(F, pt):
    return (F.one, F.one, F.zero) if pt is None else (pt[0], pt[1], F.one)


def from_jac(F, j):
    X, Y, Z = j
    if Z == F.zero:
        return None
    zi = F.inv(Z)
    zi2 = F.mul(zi, zi)
    return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))


def mul(F, pt, k):
    """[k]pt by left-to-right double-and-add (k is a plain non-negative integer)."""
    if pt is None or k == 0:
        return None
    acc = to_jac(F, None)
    base = to_jac(F, pt)
    for bit in bin(k)[2:]:
        acc = _jdbl(F, acc)
        if bit == "1":
            acc = _jadd(F, acc, base)
    return from_jac(F, acc)


def msm_naive(F, bases, scalars):
    """sum_i [s_i] P_i, the definition — arbiter for every Pippenger variant."""
    acc = to_jac(F, None)
    for b, s in zip(bases, scalars):
        if b is None or s % R_MOD == 0:
            continue
        acc = _jadd(F, acc, to_jac(F, mul(F, b, s % R_MOD)))
    return from_jac(F, acc)


# ------------------------------------------------------------------ wire images
def g1_to_bytes(pt) -> bytes:
    """104-byte in-memory image of bls12_381::G1Affine {x, y, infinity} (Montgomery limbs).
    identity = (x=0, y=R(1), inf=1) as bls12_381 0.8.0 `G1Affine::identity()` (ext)."""
    if pt is None:
        return fp_to_mont_bytes(0) + fp_to_mont_bytes(1) + bytes([1]) + bytes(7)
    return fp_to_mont_bytes(pt[0]) + fp_to_mont_bytes(pt[1]) + bytes(8)


def g1_from_bytes(b: bytes):
    if b[96] != 0:
        return None
    return (fp_from_mont_bytes(b[0:48]), fp_from_mont_bytes(b[48:96]))


def g2_to_bytes(pt) -> bytes:
    """200-byte image of bls12_381::G2Affine {x:(c0,c1), y:(c0,c1), infinity}."""
    if pt is None:
        return fp_to_mont_bytes(0) * 2 + fp_to_mont_bytes(1) + fp_to_mont_bytes(0) + bytes([1]) + bytes(7)
    (x0, x1), (y0, y1) = pt
    return fp_to_mont_bytes(x0) + fp_to_mont_bytes(x1) + fp_to_mont_bytes(y0) + fp_to_mont_bytes(y1) + bytes(8)


def g2_from_bytes(b: bytes):
    if b[192] != 0:
        return None
    return (
        (fp_from_mont_bytes(b[0:48]), fp_from_mont_bytes(b[48:96])),
        (fp_from_mont_bytes(b[96:144]), fp_from_mont_bytes(b[144:192])),
    )


# ------------------------------------------------------------------ Fp12 as Fp[w]/(w^12 - 2 w^6 + 2)
# (u = w^6 - 1 satisfies u^2 = -1, and the sextic twist is untwisted by dividing by w^2 / w^3.)
_F12_MOD_LOW = [2, 0, 0, 0, 0, 0, -2, 0, 0, 0, 0, 0]  # w^12 = -2 + 2 w^6  (i.e. minus these coeffs)


def f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for k in range(22, 11, -1):  # reduce w^k, k>=12, using w^12 = 2 w^6 - 2
        c = t[k]
        if c:
            t[k - 6] += 2 * c
            t[k - 12] -= 2 * c
    return [x % P for x in t[:12]]


F12_ONE = [1] + [0] * 11


def f12_pow(a, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_mul(r, r)
        if bit == "1":
            r = f12_mul(r, a)
    return r


def _poly_deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def f12_inv(a):
    """extended Euclid over Fp[w]."""
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [2, 0, 0, 0, 0, 0, (-2) % P, 0, 0, 0, 0, 0, 1]
    while _poly_deg(low):
        # r = high / low (polynomial rounded division)
        dh, dl = _poly_deg(high), _poly_deg(low)
        temp = list(high)
        q = [0] * 13
        inv_lead = pow(low[dl], -1, P)
        for i in range(dh - dl, -1, -1):
            q[i] = temp[dl + i] * inv_lead % P
            for c in range(dl + 1):
                temp[c + i] = (temp[c + i] - low[c] * q[i]) % P
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * q[j]) % P
                new[i + j] = (new[i + j] - low[i] * q[j]) % P
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, P)
    return [x * inv0 % P for x in lm[:12]]


def _f12_from_fp(x):
    return [x % P] + [0] * 11


def _f12_from_fp2(x):
    # a + b u with u = w^6 - 1  ->  (a - b) + b w^6
    v = [0] * 12
    v[0] = (x[0] - x[1]) % P
    v[6] = x[1] % P
    return v


_W = [0, 1] + [0] * 10
_W2_INV = f12_inv(f12_mul(_W, _W))
_W3_INV = f12_inv(f12_mul(f12_mul(_W, _W), _W))


def _untwist(q):
    """E'(Fp2) -> E(Fp12): (x, y) -> (x / w^2, y / w^3)."""
    return (f12_mul(_f12_from_fp2(q[0]), _W2_INV), f12_mul(_f12_from_fp2(q[1]), _W3_INV))


def _f12_sub(a, b):
    return [(x - y) % P for x, y in zip(a, b)]


def _f12_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def _line(p1, p2, t):
    """value at t of the line through p1,p2 (points of E(Fp12), affine)."""
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if x1 != x2:
        m = f12_mul(_f12_sub(y2, y1), f12_inv(_f12_sub(x2, x1)))
        return _f12_sub(f12_mul(m, _f12_sub(xt, x1)), _f12_sub(yt, y1))
    if y1 == y2:
        three_x2 = f12_mul(_f12_from_fp(3), f12_mul(x1, x1))
        m = f12_mul(three_x2, f12_inv(_f12_add(y1, y1)))
        return _f12_sub(f12_mul(m, _f12_sub(xt, x1)), _f12_sub(yt, y1))
    return _f12_sub(xt, x1)


def _e12_add(p1, p2):
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2 and y1 == y2:
        m = f12_mul(f12_mul(_f12_from_fp(3), f12_mul(x1, x1)), f12_inv(_f12_add(y1, y1)))
    else:
        m = f12_mul(_f12_sub(y2, y1), f12_inv(_f12_sub(x2, x1)))
    x3 = _f12_sub(_f12_sub(f12_mul(m, m), x1), x2)
    y3 = _f12_sub(f12_mul(m, _f12_sub(x1, x3)), y1)
    return (x3, y3)


def miller_loop(q_g2, p_g1):
    """f_{|x|,Q}(P) on E(Fp12) (ate pairing, loop over |x|; conjugation for the sign of x is
    applied by the caller through the final inversion-free trick f -> f^-1 being absorbed:
    for *product-equals-one* checks the common sign is irrelevant, for pairing values we invert)."""
    if q_g2 is None or p_g1 is None:
        return F12_ONE
    Q = _untwist(q_g2)
    Pt = (_f12_from_fp(p_g1[0]), _f12_from_fp(p_g1[1]))
    R = Q
    f = F12_ONE
    for bit in bin(BLS_X)[3:]:
        f = f12_mul(f12_mul(f, f), _line(R, R, Pt))
        R = _e12_add(R, R)
        if bit == "1":
            f = f12_mul(f, _line(R, Q, Pt))
            R = _e12_add(R, Q)
    return f


_FINAL_EXP = (P**12 - 1) // R_MOD


def final_exponentiation(f):
    return f12_pow(f, _FINAL_EXP)


def pairing(q_g2, p_g1):
    """e(P, Q) up to the fixed automorphism induced by using |x| (consistent across calls, bilinear)."""
    return final_exponentiation(miller_loop(q_g2, p_g1))


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 for pairs (P_i in G1, Q_i in G2) — one shared final exponentiation."""
    f = F12_ONE
    for p_g1, q_g2 in pairs:
        f = f12_mul(f, miller_loop(q_g2, p_g1))
    return final_exponentiation(f) == F12_ONE
