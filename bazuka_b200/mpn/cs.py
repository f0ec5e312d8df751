"""bellman's constraint-system vocabulary, restated (bellman 0.14.0 is an un-vendored crates.io
dependency of the reference, /root/reference/Cargo.toml:27; the reference's gadgets and circuits are
written against exactly these types — `ConstraintSystem::{alloc, alloc_input, enforce}`,
`LinearCombination`, `gadgets::num::AllocatedNum`, `gadgets::boolean::{AllocatedBit, Boolean}`).

`ConstraintSystem` here plays the role of bellman's `ProvingAssignment` and `KeypairAssembly` at once:
it records every R1CS row (for the CSR the GPU prover and the setup consume) and every variable's value
(the witness).  Variables: Input(0) = ONE, Input(i), Aux(j); after synthesis z = inputs ++ aux.
Constraint and variable ORDER follow the gadget source order; since no proving key of the reference
exists only self-consistency between setup and prover is required, and both consume
the same CSR."""
import numpy as np

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_RMONT = (1 << 256) % R


class LC:
    """linear combination: {var: coeff}; var = 2*i for Input(i), 2*j+1 for Aux(j)."""
    __slots__ = ("t",)

    def __init__(self, terms=None):
        self.t = terms if terms is not None else {}

    def copy(self):
        return LC(dict(self.t))

    def add_term(self, coeff, var):
        out = dict(self.t)
        c = (out.get(var, 0) + coeff) % R
        out[var] = c
        return LC(out)

    def __add__(self, other):
        out = dict(self.t)
        for v, c in other.t.items():
            out[v] = (out.get(v, 0) + c) % R
        return LC(out)

    def __sub__(self, other):
        out = dict(self.t)
        for v, c in other.t.items():
            out[v] = (out.get(v, 0) - c) % R
        return LC(out)

    def scaled(self, k):
        return LC({v: c * k % R for v, c in self.t.items()})


ONE = 0  # Input(0)


class ConstraintSystem:
    def __init__(self, record=False):
        self.inputs = [1]
        self.aux = []
        self.rows = []  # (LC, LC, LC)
        # optional "witness program": for every aux variable, HOW its value follows from earlier ones
        # (see witness_program.py).  None = not recording.
        self.recipes = [] if record else None

    # bellman ConstraintSystem
    def alloc(self, value, recipe=None):
        """recipe: None / ('raw',) = external input of the circuit (a field of the transition), or one of
        ('mul', lcA, lcB), ('bit', lc, i), ('iszero', lc), ('invz', lc), ('select', lcS, lcA, lcB),
        ('jjx' | 'jjy', lcX1, lcY1, lcX2, lcY2)."""
        self.aux.append(value % R)
        if self.recipes is not None:
            self.recipes.append(recipe if recipe is not None else ("raw",))
        return 2 * (len(self.aux) - 1) + 1

    def alloc_input(self, value):
        self.inputs.append(value % R)
        return 2 * (len(self.inputs) - 1)

    def enforce(self, a, b, c):
        self.rows.append((a, b, c))

    def value(self, var):
        return self.inputs[var >> 1] if var % 2 == 0 else self.aux[var >> 1]

    def eval(self, lc):
        return sum(c * self.value(v) for v, c in lc.t.items()) % R

    # ------------------------------------------------------------------ results
    @property
    def num_constraints(self):
        return len(self.rows)

    def is_satisfied(self):
        for i, (a, b, c) in enumerate(self.rows):
            if self.eval(a) * self.eval(b) % R != self.eval(c):
                return False, i
        return True, -1

    def to_csr(self):
        """-> (num_inputs, num_aux, [(rowptr, col, val)] x3, inputs [ni,4], aux [na,4]) with Montgomery
        uint64 images, zero coefficients dropped (bellman's `eval` skips them as well)."""
        ni = len(self.inputs)

        def zidx(var):
            return (var >> 1) if var % 2 == 0 else ni + (var >> 1)

        mats = []
        for k in range(3):
            rp = np.zeros(len(self.rows) + 1, dtype=np.uint64)
            cols, vals = [], []
            for j, row in enumerate(self.rows):
                for v, c in row[k].t.items():
                    if c:
                        cols.append(zidx(v))
                        vals.append(c)
                rp[j + 1] = len(cols)
            mats.append((rp, np.array(cols, dtype=np.uint32), to_mont(vals)))
        return ni, len(self.aux), mats, to_mont(self.inputs), to_mont(self.aux)


def to_mont(values):
    """canonical Python ints -> [n,4] uint64 Montgomery images (`ZkScalar` memory layout)."""
    buf = b"".join(((v % R) * _RMONT % R).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy() if values else np.zeros((0, 4), dtype=np.uint64)


# ---------------------------------------------------------------------------------------------
# bellman::gadgets::num::AllocatedNum
# ---------------------------------------------------------------------------------------------
class AllocatedNum:
    __slots__ = ("var", "value")

    def __init__(self, var, value):
        self.var, self.value = var, value

    @staticmethod
    def alloc(cs, value, recipe=None):
        value %= R
        return AllocatedNum(cs.alloc(value, recipe), value)

    def inputize(self, cs):
        inp = cs.alloc_input(self.value)
        cs.enforce(LC({inp: 1}), LC({ONE: 1}), LC({self.var: 1}))

    def mul(self, cs, other):
        la, lb = LC({self.var: 1}), LC({other.var: 1})
        out = AllocatedNum.alloc(cs, self.value * other.value, ("mul", la, lb))
        cs.enforce(la, lb, LC({out.var: 1}))
        return out

    def to_bits_le_strict(self, cs):
        """bellman `AllocatedNum::to_bits_le_strict`: 255 bits, big-endian walk over r-1 with runs of
        ones AND-ed together so that the bit pattern cannot exceed r-1, then one unpacking row."""
        a_bits = [(self.value >> i) & 1 for i in range(256)][::-1]
        b_bits = [((R - 1) >> i) & 1 for i in range(256)][::-1]
        result, last_run, current_run, found_one = [], None, [], False
        me = LC({self.var: 1})
        for pos, (a_bit, b) in enumerate(zip(a_bits, b_bits)):
            bit_index = 255 - pos  # little-endian index of this bit
            found_one |= bool(b)
            if not found_one:
                assert a_bit == 0
                continue
            if b:
                bit = AllocatedBit.alloc(cs, a_bit, ("bit", me, bit_index))
                current_run.append(bit)
                result.append(bit)
            else:
                if current_run:
                    if last_run is not None:
                        current_run.append(last_run)
                    cur = None
                    for v in current_run:  # kary_and
                        cur = v if cur is None else AllocatedBit.and_(cs, cur, v)
                    last_run = cur
                    current_run = []
                bit = AllocatedBit.alloc_conditionally(cs, a_bit, last_run, ("bit", me, bit_index))
                result.append(bit)
        assert not current_run
        lc, coeff = LC(), 1
        for bit in reversed(result):
            lc = lc.add_term(coeff, bit.var)
            coeff = coeff * 2 % R
        lc = lc.add_term(R - 1, self.var)
        cs.enforce(LC(), LC(), lc)
        return [Boolean.is_(b) for b in reversed(result)]


# ---------------------------------------------------------------------------------------------
# bellman::gadgets::boolean
# ---------------------------------------------------------------------------------------------
class AllocatedBit:
    __slots__ = ("var", "value")

    def __init__(self, var, value):
        self.var, self.value = var, value

    @staticmethod
    def alloc(cs, value, recipe=None):
        value = 1 if value else 0
        var = cs.alloc(value, recipe)
        cs.enforce(LC({ONE: 1, var: R - 1}), LC({var: 1}), LC())  # (1 - a) * a = 0
        return AllocatedBit(var, value)

    @staticmethod
    def alloc_conditionally(cs, value, must_be_false, recipe=None):
        value = 1 if value else 0
        var = cs.alloc(value, recipe)
        # (1 - must_be_false - a) * a = 0
        cs.enforce(LC({ONE: 1, must_be_false.var: R - 1}).add_term(R - 1, var), LC({var: 1}), LC())
        return AllocatedBit(var, value)

    @staticmethod
    def and_(cs, a, b):
        la, lb = LC({a.var: 1}), LC({b.var: 1})
        out = AllocatedBit(cs.alloc(a.value & b.value, ("mul", la, lb)), a.value & b.value)
        cs.enforce(la, lb, LC({out.var: 1}))
        return out

    @staticmethod
    def and_not(cs, a, b):
        v = a.value & (1 - b.value)
        la, lb = LC({a.var: 1}), LC({ONE: 1, b.var: R - 1})
        out = AllocatedBit(cs.alloc(v, ("mul", la, lb)), v)
        cs.enforce(la, lb, LC({out.var: 1}))
        return out

    @staticmethod
    def nor(cs, a, b):
        v = (1 - a.value) & (1 - b.value)
        la, lb = LC({ONE: 1, a.var: R - 1}), LC({ONE: 1, b.var: R - 1})
        out = AllocatedBit(cs.alloc(v, ("mul", la, lb)), v)
        cs.enforce(la, lb, LC({out.var: 1}))
        return out


class Boolean:
    """Is(bit) | Not(bit) | Constant(bool)"""
    __slots__ = ("kind", "bit", "const")

    def __init__(self, kind, bit=None, const=None):
        self.kind, self.bit, self.const = kind, bit, const

    @staticmethod
    def is_(bit):
        return Boolean("is", bit)

    @staticmethod
    def constant(b):
        return Boolean("const", None, bool(b))

    def not_(self):
        if self.kind == "const":
            return Boolean.constant(not self.const)
        return Boolean("not" if self.kind == "is" else "is", self.bit)

    @property
    def value(self):
        if self.kind == "const":
            return 1 if self.const else 0
        return self.bit.value if self.kind == "is" else 1 - self.bit.value

    @staticmethod
    def and_(cs, a, b):
        """bellman `Boolean::and`"""
        if a.kind == "const" or b.kind == "const":
            c, x = (a, b) if a.kind == "const" else (b, a)
            return x if c.const else Boolean.constant(False)
        if a.kind == "is" and b.kind == "is":
            return Boolean.is_(AllocatedBit.and_(cs, a.bit, b.bit))
        if a.kind == "is" and b.kind == "not":
            return Boolean.is_(AllocatedBit.and_not(cs, a.bit, b.bit))
        if a.kind == "not" and b.kind == "is":
            return Boolean.is_(AllocatedBit.and_not(cs, b.bit, a.bit))
        return Boolean.is_(AllocatedBit.nor(cs, a.bit, b.bit))
