"""The reference's R1CS gadgets, restated on bazuka_b200.mpn.cs.

  Number, UnsignedInteger, mux, extract_bool/assert_true/boolean_or
        /root/reference/src/zk/groth16/gadgets/common/{number,uint,mux,boolean}.rs
  poseidon                /root/reference/src/zk/groth16/gadgets/poseidon/mod.rs:8-95
  merkle (4-ary)          /root/reference/src/zk/groth16/gadgets/merkle/mod.rs:21-79
  AllocatedPoint / EdDSA  /root/reference/src/zk/groth16/gadgets/eddsa/mod.rs:14-280
Constraint emission order = the source order of those files."""
from . import native as N
from .cs import LC, ONE, R, AllocatedBit, AllocatedNum, Boolean


class Number:
    """`Number(LinearCombination, Option<value>)` — number.rs:10-11"""
    __slots__ = ("lc", "value")

    def __init__(self, lc, value):
        self.lc, self.value = lc, value % R

    @staticmethod
    def zero():
        return Number(LC(), 0)

    @staticmethod
    def one():
        return Number(LC({ONE: 1}), 1)

    @staticmethod
    def constant(v):
        return Number(LC({ONE: v % R}), v)

    @staticmethod
    def of(x, coeff=1):
        """From<AllocatedNum> / From<AllocatedBit> / From<(coeff, AllocatedNum)> / From<UnsignedInteger>"""
        if isinstance(x, Number):
            return x
        if isinstance(x, UnsignedInteger):
            return x.num
        return Number(LC({x.var: coeff % R}), x.value * coeff)

    def add_constant(self, c):
        return Number(self.lc.add_term(c % R, ONE), self.value + c)

    def add_num(self, coeff, num):
        return Number(self.lc.add_term(coeff % R, num.var), self.value + coeff * num.value)

    def __add__(self, other):
        return Number(self.lc + other.lc, self.value + other.value)

    def __sub__(self, other):
        return Number(self.lc - other.lc, self.value - other.value)

    def add_scaled(self, coeff, other):  # Add<(Fr, Number)>
        return Number(self.lc + other.lc.scaled(coeff % R), self.value + coeff * other.value)

    def mul(self, cs, other):
        out = AllocatedNum.alloc(cs, self.value * other.value, ("mul", self.lc, other.lc))
        cs.enforce(self.lc, other.lc, LC({out.var: 1}))
        return out

    def compress(self, cs):
        return self.mul(cs, Number.one())

    def is_zero(self, cs):
        """number.rs:75-111 (2 constraints + the bit's booleanity)"""
        z = 1 if self.value == 0 else 0
        is_zero = AllocatedBit.alloc(cs, z, ("iszero", self.lc))
        inv = AllocatedNum.alloc(cs, 0 if z else pow(self.value, -1, R), ("invz", self.lc))
        cs.enforce(LC() - self.lc, LC({inv.var: 1}), LC({is_zero.var: 1, ONE: R - 1}))
        cs.enforce(LC({is_zero.var: 1}), self.lc, LC())
        return Boolean.is_(is_zero)

    def is_equal(self, cs, other):
        return (self - other).is_zero(cs)

    def assert_equal(self, cs, other):
        cs.enforce(self.lc, LC({ONE: 1}), other.lc)

    def assert_equal_if_enabled(self, cs, enabled, other):
        """number.rs:132-178"""
        if enabled.kind == "is":
            e = enabled.bit
            eis = cs.alloc(self.value if e.value else 0, ("mul", LC({e.var: 1}), self.lc))
            cs.enforce(LC({e.var: 1}), self.lc, LC({eis: 1}))
            cs.enforce(LC({e.var: 1}), other.lc, LC({eis: 1}))
        elif enabled.kind == "const":
            if enabled.const:
                self.assert_equal(cs, other)
        else:
            raise NotImplementedError


class UnsignedInteger:
    """uint.rs:10-134"""
    __slots__ = ("bits", "num")

    def __init__(self, bits, num):
        self.bits, self.num = bits, num

    @property
    def value(self):
        return self.num.value

    @staticmethod
    def alloc(cs, val, nbits):
        return UnsignedInteger.constrain(cs, Number.of(AllocatedNum.alloc(cs, val)), nbits)

    @staticmethod
    def alloc_64(cs, val):
        return UnsignedInteger.alloc(cs, val, 64)

    @staticmethod
    def constrain(cs, num, nbits):
        bits, allc, coeff = [], LC(), 1
        for i in range(nbits):
            bit = AllocatedBit.alloc(cs, (num.value >> i) & 1, ("bit", num.lc, i))
            allc = allc.add_term(coeff, bit.var)
            bits.append(bit)
            coeff = coeff * 2 % R
        cs.enforce(allc, LC({ONE: 1}), num.lc)
        return UnsignedInteger(bits, num)

    def lt(self, cs, other):
        assert len(self.bits) == len(other.bits)
        n = len(self.bits)
        sub = (self.num - other.num).add_constant(pow(2, n + 1, R))
        sb = UnsignedInteger.constrain(cs, sub, n + 2)
        return Boolean.is_(sb.bits[n])

    def gt(self, cs, other):
        return other.lt(cs, self)

    def lte(self, cs, other):
        return self.gt(cs, other).not_()

    def gte(self, cs, other):
        return self.lt(cs, other).not_()


def extract_bool(b):
    if b.kind == "is":
        return Number.of(b.bit)
    if b.kind == "not":
        return Number.one() - Number.of(b.bit)
    return Number.one() if b.const else Number.zero()


def assert_true(cs, b):
    extract_bool(b).assert_equal(cs, Number.one())


def assert_true_if_enabled(cs, enabled, cond):
    extract_bool(cond).assert_equal_if_enabled(cs, enabled, Number.one())


def boolean_or(cs, a, b):
    return Boolean.and_(cs, a.not_(), b.not_()).not_()


def mux(cs, select, a, b):
    """select ? b : a — mux.rs:7-47"""
    if select.kind == "is":
        s = select.bit
        ret = AllocatedNum.alloc(cs, b.value if s.value else a.value, ("select", LC({s.var: 1}), a.lc, b.lc))
        cs.enforce(a.lc - b.lc, LC({s.var: 1}), a.lc.add_term(R - 1, ret.var))
        return ret
    if select.kind == "not":
        ns = select.bit
        ret = AllocatedNum.alloc(cs, a.value if ns.value else b.value, ("select", LC({ns.var: 1}), b.lc, a.lc))
        cs.enforce(b.lc - a.lc, LC({ns.var: 1}), b.lc.add_term(R - 1, ret.var))
        return ret
    raise NotImplementedError


# ---------------------------------------------------------------------------------------------
# Poseidon gadget
# ---------------------------------------------------------------------------------------------
def _sbox(cs, a):
    a2 = a.mul(cs, a)
    a4 = a2.mul(cs, a2)
    return a.mul(cs, Number.of(a4))


def poseidon(cs, vals):
    elems = [Number.zero()] + list(vals)
    t = len(elems)
    rf, rp, rc, mds = N.poseidon_params()[t]
    off = 0

    def product_mds(v):
        out = []
        for j in range(t):
            acc = Number.zero()
            for k in range(t):
                acc = acc.add_scaled(mds[j][k], v[k])
            out.append(acc)
        return out

    for rnd in range(rf + rp):
        elems = [e.add_constant(rc[off + i]) for i, e in enumerate(elems)]
        off += t
        if rnd < rf // 2 or rnd >= rf // 2 + rp:
            elems = [Number.of(_sbox(cs, e)) for e in elems]
        else:
            first = Number.of(_sbox(cs, elems[0]))
            elems = [first] + [Number.of(e.compress(cs)) for e in elems[1:]]
        elems = product_mds(elems)
    return elems[1]


# ---------------------------------------------------------------------------------------------
# 4-ary Merkle gadget
# ---------------------------------------------------------------------------------------------
def _merge_hash_poseidon4(cs, s0, s1, v, p):
    b0, b1 = Boolean.is_(s0), Boolean.is_(s1)
    and_ = Boolean.and_(cs, b0, b1)
    or_ = boolean_or(cs, b0, b1)
    p0, p1, p2 = (Number.of(x) for x in p)
    v0 = mux(cs, or_, v, p0)
    v1p = mux(cs, b0, p0, v)
    v1 = mux(cs, b1, Number.of(v1p), p1)
    v2p = mux(cs, b0, v, p2)
    v2 = mux(cs, b1, p1, Number.of(v2p))
    v3 = mux(cs, and_, p2, v)
    return poseidon(cs, [Number.of(v0), Number.of(v1), Number.of(v2), Number.of(v3)])


def calc_root_poseidon4(cs, index, val, proof):
    assert len(index.bits) == 2 * len(proof)
    cur = val
    for lvl, p in enumerate(proof):
        cur = _merge_hash_poseidon4(cs, index.bits[2 * lvl], index.bits[2 * lvl + 1], cur, p)
    return cur


def check_proof_poseidon4(cs, enabled, index, val, proof, root):
    new_root = calc_root_poseidon4(cs, index, val, proof)
    root.assert_equal_if_enabled(cs, enabled, new_root)


def alloc_proof(cs, proof):
    return [[AllocatedNum.alloc(cs, s) for s in level] for level in proof]


# ---------------------------------------------------------------------------------------------
# EdDSA on JubJub
# ---------------------------------------------------------------------------------------------
class AllocatedPoint:
    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x, self.y = x, y

    @staticmethod
    def alloc(cs, pt, recipe_args=None):
        """recipe_args = (lcX1, lcY1, lcX2, lcY2) when the point is the sum of two points (witness hint),
        None when it is an external input."""
        if recipe_args is None:
            return AllocatedPoint(AllocatedNum.alloc(cs, pt[0]), AllocatedNum.alloc(cs, pt[1]))
        return AllocatedPoint(AllocatedNum.alloc(cs, pt[0], ("jjx",) + tuple(recipe_args)),
                              AllocatedNum.alloc(cs, pt[1], ("jjy",) + tuple(recipe_args)))

    @property
    def value(self):
        return (self.x.value, self.y.value)

    def is_null(self, cs):
        xz = Number.of(self.x).is_zero(cs)
        yz = Number.of(self.y).is_zero(cs)
        return Boolean.and_(cs, xz, yz)

    def is_equal(self, cs, other):
        xe = Number.of(self.x).is_equal(cs, Number.of(other.x))
        ye = Number.of(self.y).is_equal(cs, Number.of(other.y))
        return Boolean.and_(cs, xe, ye)

    def assert_on_curve(self, cs, enabled):
        x2 = self.x.mul(cs, self.x)
        y2 = self.y.mul(cs, self.y)
        x2y2 = x2.mul(cs, y2)
        lhs = Number.of(y2) - Number.of(x2)
        rhs = Number.of(x2y2, N.JJ_D) + Number.one()
        lhs.assert_equal_if_enabled(cs, enabled, rhs)

    @staticmethod
    def _sum_value(a, b):
        if not N.jj_on_curve(a) or not N.jj_on_curve(b):
            return (0, 0)  # "If invalid, do not need to calculate" — eddsa/mod.rs:85-88
        return N.jj_add(a, b)

    def add_const(self, cs, b):
        s = AllocatedPoint.alloc(cs, AllocatedPoint._sum_value(self.value, b),
                                 (LC({self.x.var: 1}), LC({self.y.var: 1}), LC({ONE: b[0] % R}), LC({ONE: b[1] % R})))
        bx, by = b
        k = N.JJ_D * bx % R * by % R
        common = self.x.mul(cs, self.y)
        cs.enforce(LC({ONE: 1, common.var: k}), LC({s.x.var: 1}), LC({self.x.var: by % R}).add_term(bx % R, self.y.var))
        cs.enforce(LC({ONE: 1, common.var: (-k) % R}), LC({s.y.var: 1}),
                   LC({self.y.var: by % R}).add_term((-(N.JJ_A * bx)) % R, self.x.var))
        return s

    def add(self, cs, other):
        s = AllocatedPoint.alloc(cs, AllocatedPoint._sum_value(self.value, other.value),
                                 (LC({self.x.var: 1}), LC({self.y.var: 1}), LC({other.x.var: 1}), LC({other.y.var: 1})))
        common = self.x.mul(cs, other.x).mul(cs, self.y).mul(cs, other.y)
        x1 = self.x.mul(cs, other.y)
        x2 = self.y.mul(cs, other.x)
        cs.enforce(LC({ONE: 1, common.var: N.JJ_D}), LC({s.x.var: 1}), LC({x1.var: 1}).add_term(1, x2.var))
        y1 = self.y.mul(cs, other.y)
        y2 = self.x.mul(cs, other.x)
        cs.enforce(LC({ONE: 1, common.var: (-N.JJ_D) % R}), LC({s.y.var: 1}),
                   LC({y1.var: 1}).add_term((-N.JJ_A) % R, y2.var))
        return s

    def mul(self, cs, b):
        bits = b.to_bits_le_strict(cs)[::-1]
        res = AllocatedPoint(mux(cs, bits[0], Number.zero(), Number.of(self.x)),
                             mux(cs, bits[0], Number.constant(1), Number.of(self.y)))
        for bit in bits[1:]:
            res = res.add(cs, res)
            rpb = res.add(cs, self)
            res = AllocatedPoint(mux(cs, bit, Number.of(res.x), Number.of(rpb.x)),
                                 mux(cs, bit, Number.of(res.y), Number.of(rpb.y)))
        return res


def base_mul(cs, base, b):
    bits = b.to_bits_le_strict(cs)[::-1]
    res = AllocatedPoint(mux(cs, bits[0], Number.zero(), Number.constant(base[0])),
                         mux(cs, bits[0], Number.constant(1), Number.constant(base[1])))
    for bit in bits[1:]:
        res = res.add(cs, res)
        rpb = res.add_const(cs, base)
        res = AllocatedPoint(mux(cs, bit, Number.of(res.x), Number.of(rpb.x)),
                             mux(cs, bit, Number.of(res.y), Number.of(rpb.y)))
    return res


def mul_cofactor(cs, p):
    q = p.add(cs, p)
    q = q.add(cs, q)
    return q.add(cs, q)


def verify_eddsa(cs, enabled, pk, msg, sig_r, sig_s):
    h = poseidon(cs, [Number.of(sig_r.x), Number.of(sig_r.y), Number.of(pk.x), Number.of(pk.y), msg]).compress(cs)
    sb = base_mul(cs, N.JJ_BASE_COFACTOR, sig_s)
    rpha = pk.mul(cs, h)
    rpha = rpha.add(cs, sig_r)
    rpha = mul_cofactor(cs, rpha)
    Number.of(rpha.x).assert_equal_if_enabled(cs, enabled, Number.of(sb.x))
    Number.of(rpha.y).assert_equal_if_enabled(cs, enabled, Number.of(sb.y))
