"""Witness generation for UpdateCircuit as a straight-line PROGRAM, so that it can run on the GPU.

bellman computes a witness by running the circuit's `synthesize` with value closures
(/root/reference/src/mpn/circuits/update_circuit.rs:81-469 and the gadgets under
/root/reference/src/zk/groth16/gadgets/): every `alloc` carries the rule that derives its value from
earlier variables.  The gadgets in this package record those rules ("recipes", cs.py `alloc`), and a
transition slot of the update circuit is the same instruction sequence for every slot, so one
compiled program + one thread per slot replaces the per-slot Python/Rust synthesis:

    op j   writes block variable j (z index  ni + P_aux + slot*A_tx + j)
    RAW    external value number `imm` of the slot (raw_values(tr) order)
    MUL    lc0 * lc1
    BIT    bit `imm` of the canonical value of lc0
    ISZERO lc0 == 0 ;  INVZ  lc0^-1 (0 for 0)
    SELECT lc0 ? lc2 : lc1
    JJ     twisted-Edwards sum of (lc0,lc1)+(lc2,lc3) -> variables j, j+1  ((0,0) when an input is off-curve)
    NOP    (the y half of a JJ)

Linear combinations live in a pool (deduplicated); variables are addressed by SLOT: 0 = ONE,
1..n_ext = "externals" (variables the block reads but does not define: for the update circuit the accepted
fee token and the state root entering the slot), 1 + n_ext + j = block variable j.
`run_reference` interprets the program with Python integers (CPU tests); csrc/witness.cu is the device
interpreter."""
from dataclasses import dataclass

import numpy as np

from . import native as N
from . import update as U
from .cs import LC, ONE, R, AllocatedNum, ConstraintSystem, to_mont
from .fastsynth import FAKE_STATE_VAR
from .gadgets import Number

OP_RAW, OP_MUL, OP_BIT, OP_ISZERO, OP_INVZ, OP_SELECT, OP_JJ, OP_NOP = range(8)
SLOT_ONE = 0


def raw_values(tr, A, T):
    """the external values of one slot, in `_tx_block`'s allocation order (update.py)."""
    flat = lambda proof: [s for level in proof for s in level]
    dst_pk = N.jj_decompress(tr.tx.dst_pub_key)
    out = [1 if tr.enabled else 0, tr.src_token_index, tr.src_fee_token_index, tr.dst_token_index,
           tr.src_before.tx_nonce, tr.src_before.withdraw_nonce, tr.src_before.address[0], tr.src_before.address[1],
           tr.src_before_balances_hash, tr.dst_before_balances_hash,
           tr.src_before_balance.token_id, tr.src_before_balance.amount,
           tr.src_before_fee_balance.token_id, tr.src_before_fee_balance.amount]
    out += flat(tr.src_balance_proof)
    out += [tr.tx.amount.amount, tr.tx.fee.amount]
    out += flat(tr.src_fee_balance_proof)
    out += [tr.tx.nonce, tr.src_index, tr.tx.amount.token_id, tr.tx.fee.token_id,
            tr.dst_before_balance.token_id, tr.dst_before_balance.amount]
    out += flat(tr.dst_balance_proof)
    out += flat(tr.src_proof)
    out += [dst_pk[0], dst_pk[1], tr.dst_index, tr.dst_before.tx_nonce, tr.dst_before.withdraw_nonce,
            tr.dst_before.address[0], tr.dst_before.address[1]]
    out += flat(tr.dst_proof)
    out += [tr.tx.sig["r"][0], tr.tx.sig["r"][1], tr.tx.sig["s"]]
    return [v % R for v in out]


@dataclass
class WitnessProgram:
    A: int
    T: int
    ops: np.ndarray        # int32 [n_ops, 6]: opcode, lc0, lc1, lc2, lc3, imm
    lc_ptr: np.ndarray     # int32 [n_lc + 1]
    lc_slot: np.ndarray    # int32 [n_terms]
    lc_coef: np.ndarray    # int32 [n_terms]   index into coefs (0 = the constant 1)
    coefs: list            # canonical ints
    n_raw: int
    n_ext: int             # external slots (values supplied per slot by the caller)
    p_aux: int = 0         # update circuit: prologue aux count
    state_out: int = 0     # block-local index of the state root leaving the slot
    final_fee: int = 0     # update circuit: block-local index of the slot's accepted fee

    @property
    def n_ops(self):
        return len(self.ops)

    @property
    def block0(self):
        return 1 + self.n_ext

    @property
    def n_slots(self):
        return self.block0 + len(self.ops)

    def coefs_mont(self):
        return to_mont(self.coefs)


def compile_block(recipes, first_aux, externals) -> WitnessProgram:
    """recipes: the recorded rules of the block's aux variables (aux index first_aux + j for recipes[j]);
    externals: variable ids the block may read without defining them, in external-slot order."""
    n_ops, n_ext = len(recipes), len(externals)
    ext_slot = {v: 1 + k for k, v in enumerate(externals)}
    block0 = 1 + n_ext

    def slot_of(var):
        if var == ONE:
            return SLOT_ONE
        if var in ext_slot:
            return ext_slot[var]
        assert var % 2 == 1 and first_aux <= (var >> 1) < first_aux + n_ops, f"block recipe reads foreign variable {var}"
        return block0 + (var >> 1) - first_aux

    coef_index = {1: 0}
    coefs = [1]
    pool, lc_ptr, lc_slot, lc_coef = {}, [0], [], []

    def lc_id(lc, upto):
        terms = tuple(sorted((slot_of(v), c) for v, c in lc.t.items() if c))
        for s, _ in terms:
            assert s < block0 + upto, "recipe reads a variable allocated later"
        got = pool.get(terms)
        if got is None:
            got = pool[terms] = len(lc_ptr) - 1
            for s, c in terms:
                lc_slot.append(s)
                ci = coef_index.get(c)
                if ci is None:
                    ci = coef_index[c] = len(coefs)
                    coefs.append(c)
                lc_coef.append(ci)
            lc_ptr.append(len(lc_slot))
        return got

    ops = np.zeros((n_ops, 6), dtype=np.int32)
    n_raw = 0
    for j, rec in enumerate(recipes):
        kind = rec[0]
        if kind == "raw":
            ops[j] = (OP_RAW, 0, 0, 0, 0, n_raw)
            n_raw += 1
        elif kind == "mul":
            ops[j] = (OP_MUL, lc_id(rec[1], j), lc_id(rec[2], j), 0, 0, 0)
        elif kind == "bit":
            ops[j] = (OP_BIT, lc_id(rec[1], j), 0, 0, 0, rec[2])
        elif kind == "iszero":
            ops[j] = (OP_ISZERO, lc_id(rec[1], j), 0, 0, 0, 0)
        elif kind == "invz":
            ops[j] = (OP_INVZ, lc_id(rec[1], j), 0, 0, 0, 0)
        elif kind == "select":
            ops[j] = (OP_SELECT, lc_id(rec[1], j), lc_id(rec[2], j), lc_id(rec[3], j), 0, 0)
        elif kind == "jjx":
            nxt = recipes[j + 1]
            assert nxt[0] == "jjy" and all(a is b for a, b in zip(nxt[1:], rec[1:]))
            ops[j] = (OP_JJ, lc_id(rec[1], j), lc_id(rec[2], j), lc_id(rec[3], j), lc_id(rec[4], j), 0)
        elif kind == "jjy":
            assert recipes[j - 1][0] == "jjx"
            ops[j] = (OP_NOP, 0, 0, 0, 0, 0)
        else:
            raise ValueError(kind)
    return WitnessProgram(0, 0, ops, np.array(lc_ptr, dtype=np.int32), np.array(lc_slot, dtype=np.int32),
                          np.array(lc_coef, dtype=np.int32), coefs, n_raw, n_ext)


def compile_update_block(A, T) -> WitnessProgram:
    """record one slot of UpdateCircuit (a null transition: the instruction sequence does not depend on values)
    and compile it; externals = [accepted_fee_token, state root entering the slot]."""
    tr = U.UpdateTransition.null(A, T)
    circ = U.UpdateCircuit(A, T, 0, transitions=[tr])
    cs = ConstraintSystem(record=True)
    _, fee_tok, _, _ = circ._prologue(cs)
    p_aux = len(cs.aux)
    state_in = AllocatedNum(FAKE_STATE_VAR, 0)
    state_out, _ = circ._tx_block(cs, tr, state_in, fee_tok, Number.zero())
    prog = compile_block(cs.recipes[p_aux:], p_aux, [fee_tok.var, FAKE_STATE_VAR])
    assert prog.n_raw == len(raw_values(tr, A, T))
    prog.A, prog.T, prog.p_aux = A, T, p_aux
    prog.state_out, prog.final_fee = (state_out.var >> 1) - p_aux, (circ._last_final_fee.var >> 1) - p_aux
    return prog


def compile_update_epilogue(slot_prog: WitnessProgram, B) -> WitnessProgram:
    """the part of `synthesize` after the slot loop (update_circuit.rs:470-493): the Poseidon(fee_token, fee_sum)
    gadget.  Externals = [accepted_fee_token, slot 0's accepted fee, ..., slot n-1's accepted fee]."""
    A, T, n, a_tx, p_aux = slot_prog.A, slot_prog.T, 1 << (2 * B), slot_prog.n_ops, slot_prog.p_aux
    circ = U.UpdateCircuit(A, T, B)
    cs = ConstraintSystem(record=True)
    _, fee_tok, aux_wit, claimed = circ._prologue(cs)
    cs.aux.extend([0] * (n * a_tx))
    cs.recipes.extend([("raw",)] * (n * a_tx))
    fee_vars = [2 * (p_aux + k * a_tx + slot_prog.final_fee) + 1 for k in range(n)]
    start = len(cs.aux)
    last_state = AllocatedNum(2 * (p_aux + (n - 1) * a_tx + slot_prog.state_out) + 1, 0)
    circ._epilogue(cs, last_state, fee_tok, aux_wit, claimed, Number(LC({v: 1 for v in fee_vars}), 0))
    prog = compile_block(cs.recipes[start:], start, [fee_tok.var] + fee_vars)
    prog.A, prog.T = A, T
    return prog


def run_reference(prog: WitnessProgram, raws, ext):
    """interpret the program for one slot with Python integers -> the block's aux values (canonical).
    ext: the slot's external values (update circuit: [fee_token, state_in])."""
    V = [0] * prog.n_slots
    V[SLOT_ONE] = 1
    assert len(ext) == prog.n_ext
    for k, v in enumerate(ext):
        V[1 + k] = v % R
    SLOT_BLOCK0 = prog.block0
    ptr, slots, cidx, coefs = prog.lc_ptr, prog.lc_slot, prog.lc_coef, prog.coefs

    def ev(l):
        return sum(coefs[cidx[k]] * V[slots[k]] for k in range(ptr[l], ptr[l + 1])) % R

    for j, (op, a0, a1, a2, a3, imm) in enumerate(prog.ops.tolist()):
        d = SLOT_BLOCK0 + j
        if op == OP_RAW:
            V[d] = raws[imm] % R
        elif op == OP_MUL:
            V[d] = ev(a0) * ev(a1) % R
        elif op == OP_BIT:
            V[d] = (ev(a0) >> imm) & 1
        elif op == OP_ISZERO:
            V[d] = 1 if ev(a0) == 0 else 0
        elif op == OP_INVZ:
            x = ev(a0)
            V[d] = pow(x, -1, R) if x else 0
        elif op == OP_SELECT:
            V[d] = ev(a2) if ev(a0) else ev(a1)
        elif op == OP_JJ:
            p, q = (ev(a0), ev(a1)), (ev(a2), ev(a3))
            s = N.jj_add(p, q) if N.jj_on_curve(p) and N.jj_on_curve(q) else (0, 0)
            V[d], V[d + 1] = s
        elif op == OP_NOP:
            pass
        else:
            raise ValueError(op)
    return V[SLOT_BLOCK0:]


def slot_roots(circ):
    """state root entering every slot (update() records pre_root for real transitions; padding slots keep the
    final state)."""
    roots, cur = [], circ.state
    n = len(circ.transitions)
    last_enabled = max([k for k, t in enumerate(circ.transitions) if t.enabled], default=-1)
    for k, tr in enumerate(circ.transitions):
        if tr.enabled:
            roots.append(tr.pre_root)
        elif k > last_enabled:
            roots.append(circ.next_state if last_enabled >= 0 else circ.state)
        else:  # a disabled slot between enabled ones keeps the state of the next enabled slot's pre_root
            nxt = next(t for t in circ.transitions[k:] if t.enabled)
            roots.append(nxt.pre_root)
    return roots
