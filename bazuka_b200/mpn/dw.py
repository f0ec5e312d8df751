"""MPN deposit and withdraw paths: transition builders and circuits.

  reveal gadget        /root/reference/src/zk/groth16/gadgets/reveal/mod.rs:13-64
  DepositCircuit       /root/reference/src/mpn/circuits/deposit_circuit.rs:47-293
  WithdrawCircuit      /root/reference/src/mpn/circuits/withdraw_circuit.rs:49-413
  deposit()/withdraw() /root/reference/src/mpn/deposit.rs:11-233, withdraw.rs:10-259 (state updates, proof
                       order, `aux_data` = root of the List<Struct> of the batch; calldata formats
                       deposit = Poseidon-2(pk.x, pk.y), withdraw = Poseidon-6(pk.x, pk.y, nonce, R.x, R.y, s),
                       withdraw message = Poseidon-2(fingerprint, nonce), src/core/transaction.rs:177-189)
The L1 side of a deposit/withdraw (`ContractDeposit/ContractWithdraw`, signatures, fingerprints) is
out of scope: `fingerprint` is an opaque scalar here, as it is for the circuit."""
from dataclasses import dataclass, field

from . import gadgets as G
from . import native as N
from .cs import LC, ONE, AllocatedBit, AllocatedNum, Boolean, ConstraintSystem
from .gadgets import Number, UnsignedInteger
from .update import Money, MpnAccount, MpnState, NULL_POINT


# ---------------------------------------------------------------- reveal: state model -> root, in circuit
def reveal_list_of_structs(cs, log4_size, rows):
    """`reveal` for List{log4_size, Struct{k scalars}}: rows = list of lists of Number."""
    assert len(rows) == 1 << (2 * log4_size)
    leaves = [G.poseidon(cs, r) for r in rows]
    while len(leaves) != 1:
        leaves = [G.poseidon(cs, leaves[i:i + 4]) for i in range(0, len(leaves), 4)]
    return leaves[0]


def native_list_root(log4_size, rows):
    leaves = [N.poseidon(r) for r in rows]
    assert len(leaves) == 1 << (2 * log4_size)
    while len(leaves) != 1:
        leaves = [N.poseidon(leaves[i:i + 4]) for i in range(0, len(leaves), 4)]
    return leaves[0]


def _public_inputs(cs, c):
    commitment = AllocatedNum.alloc(cs, c.commitment); commitment.inputize(cs)
    height = AllocatedNum.alloc(cs, c.height); height.inputize(cs)
    state = AllocatedNum.alloc(cs, c.state); state.inputize(cs)
    aux = AllocatedNum.alloc(cs, c.aux_data); aux.inputize(cs)
    nxt = AllocatedNum.alloc(cs, c.next_state); nxt.inputize(cs)
    return state, aux, nxt


# ---------------------------------------------------------------- deposit
@dataclass
class MpnDeposit:
    mpn_address: tuple = (0, False)   # compressed public key
    token_id: int = 0
    amount: int = 0
    src: object = None                # `payment.src`, the L1 account paying (any hashable id; None = not tracked): once one of its
                                      # deposits is rejected, its later ones in the batch are too (deposit.rs:33,68-83)


@dataclass
class DepositTransition:
    enabled: bool
    tx: MpnDeposit
    before: MpnAccount
    before_balances_hash: int
    before_balance: Money
    proof: list
    account_index: int
    token_index: int
    balance_proof: list
    pre_root: int = 0   # state root before this transition (bookkeeping for the GPU witness path)

    @staticmethod
    def null(A, T):
        z3 = lambda n: [[0, 0, 0] for _ in range(n)]
        return DepositTransition(False, MpnDeposit(), MpnAccount(), 0, Money(), z3(A), 0, 0, z3(T))


def deposit(state: MpnState, deposits, log4_batch):
    """-> (public {state, aux_data, next_state}, transitions)."""
    prev, trans = state.root, []
    n = 1 << (2 * log4_batch)
    rejected_srcs = set()                 # deposit.rs:33 `rejected_pub_keys`

    def reject(d):
        if d.src is not None:
            rejected_srcs.add(d.src)

    for d in deposits:
        if len(trans) == n:
            break
        addr = N.jj_decompress_checked(d.mpn_address)
        if addr is None:
            reject(d)
            continue
        # deposit.rs:40-54: chain index table, then this fork's new accounts, else mpn_account_count + |new_account_indices|
        idx = state.index_of(addr)
        is_new = idx is None
        if is_new:
            idx = state.new_index()
        if idx >> (2 * state.A):
            reject(d)
            continue
        before = state.get(idx)
        ti = before.find_token_index(state.T, d.token_id, True)
        if (ti is None or (d.src is not None and d.src in rejected_srcs)
                or (N.jj_on_curve(before.address) and before.address != addr)):
            reject(d)
            continue
        proof, bproof = state.prove(idx), state.prove_token(idx, ti)
        pre_root = state.root
        bal = before.tokens.get(ti)
        after = before.copy()
        after.address = addr
        after.tokens.setdefault(ti, Money(d.token_id, 0)).amount += d.amount
        state.write(idx, after)
        if is_new:
            state.new_account_indices[addr] = idx
        trans.append(DepositTransition(True, d, before, before.tokens_tree(state.T).root,
                                       Money(bal.token_id, bal.amount) if bal else Money(), proof, idx, ti, bproof, pre_root))
    rows = []
    for k in range(n):
        if k < len(trans):
            d = trans[k].tx
            a = N.jj_decompress(d.mpn_address)
            rows.append([1, d.token_id, d.amount, N.poseidon([a[0], a[1]])])
        else:
            rows.append([0, 0, 0, 0])
    return {"state": prev, "aux_data": native_list_root(log4_batch, rows), "next_state": state.root}, trans


class DepositCircuit:
    def __init__(self, A, T, B, commitment=0, height=0, state=0, aux_data=0, next_state=0, transitions=None):
        self.A, self.T, self.B = A, T, B
        self.commitment, self.height, self.state, self.aux_data, self.next_state = commitment, height, state, aux_data, next_state
        n = 1 << (2 * B)
        tr = list(transitions or [])
        self.transitions = tr + [DepositTransition.null(A, T) for _ in range(n - len(tr))]

    def synthesize(self, cs: ConstraintSystem):
        A, T = self.A, self.T
        num = Number.of
        state_wit, aux_wit, claimed = _public_inputs(cs, self)
        wits, rows = [], []
        for tr in self.transitions:
            w, row = self._phase1(cs, tr)
            wits.append(w)
            rows.append(row)
        tx_root = reveal_list_of_structs(cs, self.B, rows)
        cs.enforce(LC({aux_wit.var: 1}), LC({ONE: 1}), tx_root.lc)
        for tr, w in zip(self.transitions, wits):
            state_wit = self._phase2(cs, tr, w, state_wit)
        cs.enforce(LC({state_wit.var: 1}), LC({ONE: 1}), LC({claimed.var: 1}))
        return cs

    def _phase1(self, cs, tr):
        """the slot's transaction fields and its row of the revealed batch (deposit_circuit.rs, first loop)"""
        num = Number.of
        enabled = AllocatedBit.alloc(cs, tr.enabled)
        token_id = AllocatedNum.alloc(cs, tr.tx.token_id)
        amount = UnsignedInteger.alloc_64(cs, tr.tx.amount)
        pub_key = G.AllocatedPoint.alloc(cs, N.jj_decompress(tr.tx.mpn_address))
        pk_hash = G.poseidon(cs, [num(pub_key.x), num(pub_key.y)])
        calldata = G.mux(cs, Boolean.is_(enabled), Number.zero(), pk_hash)
        return (Boolean.is_(enabled), token_id, amount, pub_key), [num(enabled), num(token_id), num(amount), num(calldata)]

    def _phase2(self, cs, tr, wits, state_wit):
        """the slot's state transition (deposit_circuit.rs, second loop); returns the new state variable"""
        A, T = self.A, self.T
        num = Number.of
        enabled, tx_token_id, tx_amount, tx_pub_key = wits
        tx_index = UnsignedInteger.alloc(cs, tr.account_index, 2 * A)
        tx_token_index = UnsignedInteger.alloc(cs, tr.token_index, 2 * T)
        tx_pub_key.assert_on_curve(cs, enabled)
        src_tx_nonce = AllocatedNum.alloc(cs, tr.before.tx_nonce)
        src_withdraw_nonce = AllocatedNum.alloc(cs, tr.before.withdraw_nonce)
        src_addr = G.AllocatedPoint.alloc(cs, tr.before.address)
        src_balances_hash = AllocatedNum.alloc(cs, tr.before_balances_hash)
        src_token_id = AllocatedNum.alloc(cs, tr.before_balance.token_id)
        src_balance = AllocatedNum.alloc(cs, tr.before_balance.amount)
        src_token_balance_hash = G.poseidon(cs, [num(src_token_id), num(src_balance)])
        bproof = G.alloc_proof(cs, tr.balance_proof)
        G.check_proof_poseidon4(cs, enabled, tx_token_index, src_token_balance_hash, bproof, num(src_balances_hash))
        src_hash = G.poseidon(cs, [num(src_tx_nonce), num(src_withdraw_nonce), num(src_addr.x), num(src_addr.y), num(src_balances_hash)])
        proof = G.alloc_proof(cs, tr.proof)
        is_null_tok = num(src_token_id).is_zero(cs)
        is_eq_tok = num(src_token_id).is_equal(cs, num(tx_token_id))
        G.assert_true(cs, G.boolean_or(cs, is_null_tok, is_eq_tok))
        is_null_addr = src_addr.is_null(cs)
        is_eq_addr = src_addr.is_equal(cs, tx_pub_key)
        G.assert_true(cs, G.boolean_or(cs, is_null_addr, is_eq_addr))
        G.check_proof_poseidon4(cs, enabled, tx_index, src_hash, proof, num(state_wit))
        new_bal_hash = G.poseidon(cs, [num(tx_token_id), num(src_balance) + num(tx_amount)])
        new_balances_hash = G.calc_root_poseidon4(cs, tx_token_index, new_bal_hash, bproof)
        new_hash = G.poseidon(cs, [num(src_tx_nonce), num(src_withdraw_nonce), num(tx_pub_key.x), num(tx_pub_key.y), new_balances_hash])
        next_state = G.calc_root_poseidon4(cs, tx_index, new_hash, proof)
        return G.mux(cs, enabled, num(state_wit), next_state)


# ---------------------------------------------------------------- withdraw
@dataclass
class MpnWithdraw:
    mpn_address: tuple = (0, False)
    mpn_withdraw_nonce: int = 0
    mpn_sig: dict = field(default_factory=lambda: {"r": NULL_POINT, "s": 0})
    amount: Money = field(default_factory=Money)
    fee: Money = field(default_factory=Money)
    fingerprint: int = 0           # `ContractWithdraw::fingerprint()` of the L1 payment (opaque here)
    calldata: object = None        # `payment.calldata` when the caller has the L1 payment: must equal expected_calldata()
                                   # (`verify_calldata`, src/core/transaction.rs:177-182; withdraw.rs:77); None = not checked

    def expected_calldata(self):
        a = N.jj_decompress(self.mpn_address)
        return N.poseidon([a[0], a[1], self.mpn_withdraw_nonce, self.mpn_sig["r"][0], self.mpn_sig["r"][1], self.mpn_sig["s"]])

    def message(self):
        return N.poseidon([self.fingerprint, self.mpn_withdraw_nonce])

    def sign(self, sk):
        self.mpn_sig = N.eddsa_sign(sk, self.message())


@dataclass
class WithdrawTransition:
    enabled: bool
    tx: MpnWithdraw
    before: MpnAccount
    before_token_balance: Money
    before_fee_balance: Money
    proof: list
    account_index: int
    token_index: int
    token_balance_proof: list
    before_token_hash: int
    fee_token_index: int
    fee_balance_proof: list
    pre_root: int = 0   # state root before this transition (bookkeeping for the GPU witness path)

    @staticmethod
    def null(A, T):
        z3 = lambda n: [[0, 0, 0] for _ in range(n)]
        return WithdrawTransition(False, MpnWithdraw(), MpnAccount(), Money(), Money(), z3(A), 0, 0, z3(T), 0, 0, z3(T))


def withdraw(state: MpnState, withdraws, log4_batch):
    prev, trans = state.root, []
    n = 1 << (2 * log4_batch)
    for w in withdraws:
        if len(trans) == n:
            break
        addr = N.jj_decompress_checked(w.mpn_address)
        idx = state.index_of(addr) if addr is not None else None
        if idx is None:
            continue
        before = state.get(idx)
        ti = before.find_token_index(state.T, w.amount.token_id, False)
        fi = before.find_token_index(state.T, w.fee.token_id, False)
        if ti is None or fi is None or w.mpn_withdraw_nonce != before.withdraw_nonce + 1:
            continue
        if w.calldata is not None and w.calldata != w.expected_calldata():
            continue
        if before.tokens[ti].amount < w.amount.amount or not N.eddsa_verify(addr, w.message(), w.mpn_sig):
            continue
        proof, tproof = state.prove(idx), state.prove_token(idx, ti)
        pre_root = state.root
        tok = before.tokens[ti]
        after = before.copy()
        after.tokens[ti].amount -= w.amount.amount
        state.write(idx, after)
        feeb = after.tokens[fi]
        if feeb.amount < w.fee.amount:
            state.write(idx, before)
            continue
        fee_before = Money(feeb.token_id, feeb.amount)
        fproof = state.prove_token(idx, fi)
        after.tokens[fi].amount -= w.fee.amount
        after.withdraw_nonce += 1
        state.write(idx, after)
        trans.append(WithdrawTransition(True, w, before, Money(tok.token_id, tok.amount), fee_before, proof, idx, ti, tproof,
                                        before.tokens_tree(state.T).root, fi, fproof, pre_root))
    rows = []
    for k in range(n):
        if k < len(trans):
            w = trans[k].tx
            a = N.jj_decompress(w.mpn_address)
            cd = N.poseidon([a[0], a[1], w.mpn_withdraw_nonce, w.mpn_sig["r"][0], w.mpn_sig["r"][1], w.mpn_sig["s"]])
            rows.append([1, w.amount.token_id, w.amount.amount, w.fee.token_id, w.fee.amount, w.fingerprint, cd])
        else:
            rows.append([0] * 7)
    return {"state": prev, "aux_data": native_list_root(log4_batch, rows), "next_state": state.root}, trans


class WithdrawCircuit:
    def __init__(self, A, T, B, commitment=0, height=0, state=0, aux_data=0, next_state=0, transitions=None):
        self.A, self.T, self.B = A, T, B
        self.commitment, self.height, self.state, self.aux_data, self.next_state = commitment, height, state, aux_data, next_state
        n = 1 << (2 * B)
        tr = list(transitions or [])
        self.transitions = tr + [WithdrawTransition.null(A, T) for _ in range(n - len(tr))]

    def synthesize(self, cs: ConstraintSystem):
        A, T = self.A, self.T
        num = Number.of
        state_wit, aux_wit, claimed = _public_inputs(cs, self)
        wits, rows = [], []
        for tr in self.transitions:
            w, row = self._phase1(cs, tr)
            wits.append(w)
            rows.append(row)
        tx_root = reveal_list_of_structs(cs, self.B, rows)
        cs.enforce(LC({aux_wit.var: 1}), LC({ONE: 1}), tx_root.lc)
        for tr, w in zip(self.transitions, wits):
            state_wit = self._phase2(cs, tr, w, state_wit)
        cs.enforce(LC({state_wit.var: 1}), LC({ONE: 1}), LC({claimed.var: 1}))
        return cs

    def _phase1(self, cs, tr):
        num = Number.of
        enabled = AllocatedBit.alloc(cs, tr.enabled)
        amount_token_id = AllocatedNum.alloc(cs, tr.tx.amount.token_id)
        amount = UnsignedInteger.alloc_64(cs, tr.tx.amount.amount)
        fee_token_id = AllocatedNum.alloc(cs, tr.tx.fee.token_id)
        fee = UnsignedInteger.alloc_64(cs, tr.tx.fee.amount)
        fingerprint = AllocatedNum.alloc(cs, tr.tx.fingerprint if tr.enabled else 0)
        pub_key = G.AllocatedPoint.alloc(cs, N.jj_decompress(tr.tx.mpn_address))
        nonce = AllocatedNum.alloc(cs, tr.tx.mpn_withdraw_nonce)
        sig_r = G.AllocatedPoint.alloc(cs, tr.tx.mpn_sig["r"])
        sig_s = AllocatedNum.alloc(cs, tr.tx.mpn_sig["s"])
        cd_hash = G.poseidon(cs, [num(pub_key.x), num(pub_key.y), num(nonce), num(sig_r.x), num(sig_r.y), num(sig_s)])
        calldata = G.mux(cs, Boolean.is_(enabled), Number.zero(), cd_hash)
        return ((Boolean.is_(enabled), amount_token_id, amount, fee_token_id, fee, fingerprint, pub_key, nonce, sig_r, sig_s),
                [num(enabled), num(amount_token_id), num(amount), num(fee_token_id), num(fee), num(fingerprint), num(calldata)])

    def _phase2(self, cs, tr, wits, state_wit):
        A, T = self.A, self.T
        num = Number.of
        enabled, tx_amount_token_id, tx_amount, tx_fee_token_id, tx_fee, fingerprint, tx_pub_key, tx_nonce, tx_sig_r, tx_sig_s = wits
        tx_index = UnsignedInteger.alloc(cs, tr.account_index, 2 * A)
        tx_token_index = UnsignedInteger.alloc(cs, tr.token_index, 2 * T)
        tx_fee_token_index = UnsignedInteger.alloc(cs, tr.fee_token_index, 2 * T)
        tx_pub_key.assert_on_curve(cs, enabled)
        tx_hash = G.poseidon(cs, [num(fingerprint), num(tx_nonce)])
        tx_sig_r.assert_on_curve(cs, enabled)
        G.verify_eddsa(cs, enabled, tx_pub_key, tx_hash, tx_sig_r, tx_sig_s)
        src_tx_nonce = AllocatedNum.alloc(cs, tr.before.tx_nonce)
        src_withdraw_nonce = AllocatedNum.alloc(cs, tr.before.withdraw_nonce)
        src_addr = G.AllocatedPoint.alloc(cs, tr.before.address)
        src_addr.assert_on_curve(cs, enabled)
        before_token_hash = AllocatedNum.alloc(cs, tr.before_token_hash)
        src_token_id = AllocatedNum.alloc(cs, tr.before_token_balance.token_id)
        num(src_token_id).assert_equal(cs, num(tx_amount_token_id))
        src_balance = AllocatedNum.alloc(cs, tr.before_token_balance.amount)
        src_token_balance_hash = G.poseidon(cs, [num(src_token_id), num(src_balance)])
        tproof = G.alloc_proof(cs, tr.token_balance_proof)
        G.check_proof_poseidon4(cs, enabled, tx_token_index, src_token_balance_hash, tproof, num(before_token_hash))
        new_token_balance_hash = G.poseidon(cs, [num(src_token_id), num(src_balance) - num(tx_amount)])
        balance_middle_root = G.calc_root_poseidon4(cs, tx_token_index, new_token_balance_hash, tproof)
        src_fee_token_id = AllocatedNum.alloc(cs, tr.before_fee_balance.token_id)
        num(src_fee_token_id).assert_equal(cs, num(tx_fee_token_id))
        src_fee_balance = AllocatedNum.alloc(cs, tr.before_fee_balance.amount)
        src_fee_token_balance_hash = G.poseidon(cs, [num(src_fee_token_id), num(src_fee_balance)])
        fproof = G.alloc_proof(cs, tr.fee_balance_proof)
        G.check_proof_poseidon4(cs, enabled, tx_fee_token_index, src_fee_token_balance_hash, fproof, balance_middle_root)
        new_fee_token_balance_hash = G.poseidon(cs, [num(src_fee_token_id), num(src_fee_balance) - num(tx_fee)])
        src_hash = G.poseidon(cs, [num(src_tx_nonce), num(src_withdraw_nonce), num(src_addr.x), num(src_addr.y), num(before_token_hash)])
        proof = G.alloc_proof(cs, tr.proof)
        G.check_proof_poseidon4(cs, enabled, tx_index, src_hash, proof, num(state_wit))
        num(tx_nonce).assert_equal_if_enabled(cs, enabled, num(src_withdraw_nonce) + Number.constant(1))
        balance_final_root = G.calc_root_poseidon4(cs, tx_fee_token_index, new_fee_token_balance_hash, fproof)
        new_hash = G.poseidon(cs, [num(src_tx_nonce), num(src_withdraw_nonce) + Number.constant(1), num(tx_pub_key.x), num(tx_pub_key.y), balance_final_root])
        next_state = G.calc_root_poseidon4(cs, tx_index, new_hash, proof)
        return G.mux(cs, enabled, num(state_wit), next_state)
