"""MPN update path: ledger state, `update()` transition builder and `UpdateCircuit`.

  MpnState / MpnAccount   /root/reference/src/mpn/mod.rs:219-240 (state model), src/zk/mod.rs:60-94,
                          src/zk/state/mod.rs:93-208 (get/set_mpn_account over the 4-ary Poseidon state)
  update()                /root/reference/src/mpn/update.rs:8-299 (acceptance rules, proof order,
                          intermediate-root handling when amount and fee share a token)
  UpdateTransition.null   /root/reference/src/mpn/mod.rs:513-537
  UpdateCircuit           /root/reference/src/mpn/circuits/update_circuit.rs:49-494 — synthesize() emits the
                          same constraints in the same order: 5 public inputs, then per transition the
                          balance/fee/account Merkle checks, range checks, tx hash and EdDSA, then the
                          fee commitment and the next-state equality."""
from dataclasses import dataclass, field

from . import native as N
from .cs import LC, ONE, AllocatedBit, AllocatedNum, Boolean, ConstraintSystem
from . import gadgets as G
from .gadgets import Number, UnsignedInteger

NULL_POINT = (0, 0)                       # PointAffine::default()
NULL_DST = (0, N.R - 1)                   # PublicKey::default().0.decompress()  
ZIESHA = 1                                # ContractId::Ziesha -> ZkScalar::ONE (src/zk/mod.rs:280-288)


@dataclass
class Money:
    token_id: int = 0
    amount: int = 0


@dataclass
class MpnAccount:
    tx_nonce: int = 0
    withdraw_nonce: int = 0
    address: tuple = NULL_POINT
    tokens: dict = field(default_factory=dict)  # token index -> Money

    def copy(self):
        return MpnAccount(self.tx_nonce, self.withdraw_nonce, self.address,
                          {k: Money(v.token_id, v.amount) for k, v in self.tokens.items()})

    def tokens_tree(self, log4_token):
        t = N.SparseTree4(log4_token, N.poseidon([0, 0]))
        for i, m in self.tokens.items():
            t.set_leaf(i, N.poseidon([m.token_id, m.amount]))
        return t

    def find_token_index(self, log4_token, token_id, empty_allowed):
        """src/zk/mod.rs:95-117: the slot holding token_id, else (if allowed) the first free slot."""
        for i, m in sorted(self.tokens.items()):
            if m.token_id == token_id:
                return i
        if empty_allowed:
            for i in range(1 << (2 * log4_token)):
                if i not in self.tokens:
                    return i
        return None

    def leaf_hash(self, log4_token):
        return N.poseidon([self.tx_nonce, self.withdraw_nonce, self.address[0], self.address[1],
                           self.tokens_tree(log4_token).root])


class MpnState:
    """the MPN contract state: List<log4 A>(Struct[tx_nonce, withdraw_nonce, pk.x, pk.y, List<log4 T>(...)])."""

    def __init__(self, log4_tree, log4_token):
        self.A, self.T = log4_tree, log4_token
        self.accounts = {}
        self.tree = N.SparseTree4(log4_tree, MpnAccount().leaf_hash(log4_token))
        # The CHAIN's tables, which the builders only read (`get_mpn_account_indices`, `get_mpn_account_count`,
        # src/mpn/update.rs:29,47-70): they change when a block is applied (commit_accounts), not when a batch is built.
        self.address_index = {}
        self.account_count = 0
        # `new_account_indices` (src/mpn/mod.rs:330): accounts created by the batches built so far on this fork, threaded
        # through deposit -> withdraw -> update by prepare_works
        self.new_account_indices = {}
        self.state_size = 0            # ZkCompressedState.state_size: non-zero scalar leaves (src/zk/state/mod.rs:327-341)

    @property
    def root(self):
        return self.tree.root

    @property
    def compressed(self):
        """`ZkCompressedState {state_hash, state_size}` — what goes into `MpnWork.new_root` (src/mpn/mod.rs:264-270)."""
        return (self.tree.root, self.state_size)

    def get(self, idx):
        return self.accounts.get(idx, MpnAccount()).copy()

    @staticmethod
    def leaf_count(acc):
        if acc is None:
            return 0
        return ((acc.tx_nonce != 0) + (acc.withdraw_nonce != 0) + (acc.address[0] != 0) + (acc.address[1] != 0)
                + sum((m.token_id != 0) + (m.amount != 0) for m in acc.tokens.values()))

    def write(self, idx, acc):
        """`set_mpn_account` (src/zk/state/mod.rs:140-208): what the transition builders do to the state."""
        self.state_size += self.leaf_count(acc) - self.leaf_count(self.accounts.get(idx))
        self.accounts[idx] = acc.copy()
        self.tree.set_leaf(idx, acc.leaf_hash(self.T))

    def set(self, idx, acc):
        """load an account as the chain knows it: state + the chain's address index and account count."""
        self.write(idx, acc)
        if acc.address != NULL_POINT:
            self.address_index.setdefault(acc.address, idx)
        self.account_count = max(self.account_count, idx + 1)

    def prove(self, idx):
        return self.tree.prove(idx)

    def prove_token(self, idx, token_index):
        return self.get(idx).tokens_tree(self.T).prove(token_index)

    def index_of(self, address):
        """update.rs:47-70: the chain's index table first, then the accounts created earlier on this fork."""
        i = self.address_index.get(address)
        return i if i is not None else self.new_account_indices.get(address)

    def new_index(self):
        """update.rs:66: `mpn_account_count + new_account_indices.len()`"""
        return self.account_count + len(self.new_account_indices)

    def commit_accounts(self):
        """the block built on this fork was applied: its new accounts enter the chain's address index."""
        for addr, i in self.new_account_indices.items():
            self.address_index.setdefault(addr, i)
            self.account_count = max(self.account_count, i + 1)
        self.new_account_indices = {}

    def snapshot(self):
        return (dict(self.accounts), _clone_tree(self.tree), self.state_size)

    def restore(self, snap):
        self.accounts, self.tree, self.state_size = snap

    def fork(self):
        """`db.fork_on_ram()` (src/mpn/mod.rs:313)"""
        import copy
        return copy.deepcopy(self)


@dataclass
class MpnTransaction:
    nonce: int = 0
    src_pub_key: tuple = (0, False)       # compressed
    dst_pub_key: tuple = (0, False)
    amount: Money = field(default_factory=Money)
    fee: Money = field(default_factory=Money)
    sig: dict = field(default_factory=lambda: {"r": NULL_POINT, "s": 0})

    def hash(self):
        d = N.jj_decompress(self.dst_pub_key)
        return N.poseidon([self.nonce, d[0], d[1], self.amount.token_id, self.amount.amount,
                           self.fee.token_id, self.fee.amount])

    def sign(self, sk):
        self.sig = N.eddsa_sign(sk, self.hash())


@dataclass
class UpdateTransition:
    enabled: bool
    tx: MpnTransaction
    src_before: MpnAccount
    src_before_balances_hash: int
    src_before_balance: Money
    src_before_fee_balance: Money
    src_proof: list
    src_index: int
    src_token_index: int
    src_balance_proof: list
    src_fee_token_index: int
    src_fee_balance_proof: list
    dst_before: MpnAccount
    dst_before_balances_hash: int
    dst_before_balance: Money
    dst_proof: list
    dst_index: int
    dst_token_index: int
    dst_balance_proof: list
    pre_root: int = 0   # state root before this transition (bookkeeping for the parallel synthesiser)

    @staticmethod
    def null(A, T):
        z3 = lambda n: [[0, 0, 0] for _ in range(n)]
        return UpdateTransition(False, MpnTransaction(), MpnAccount(), 0, Money(), Money(), z3(A), 0, 0, z3(T), 0, z3(T),
                                MpnAccount(), 0, Money(), z3(A), 0, 0, z3(T))


def update(state: MpnState, txs, log4_batch, fee_token=ZIESHA):
    """-> (public dict {state, aux_data, next_state}, transitions (accepted only), rejected)."""
    A, T = state.A, state.T
    prev_root = state.root
    transitions, rejected, fee_sum = [], [], 0
    for tx in txs:
        if len(transitions) == 1 << (2 * log4_batch):
            break
        # the reference's pre-filter (update.rs:31-38): fee token, and both keys must decompress to curve points
        # (reported here with the rejected transactions; the reference drops them silently)
        src_addr, dst_addr = N.jj_decompress_checked(tx.src_pub_key), N.jj_decompress_checked(tx.dst_pub_key)
        if tx.fee.token_id != fee_token or src_addr is None or dst_addr is None:
            rejected.append(tx)
            continue
        # update.rs:47-70: chain index table, then this fork's new accounts; unknown sender -> rejected,
        # unknown receiver -> index mpn_account_count + |new_account_indices|
        src_index = state.index_of(src_addr)
        if src_index is None:
            rejected.append(tx)
            continue
        dst_index = state.index_of(dst_addr)
        dst_new = dst_index is None
        if dst_new:
            dst_index = state.new_index()
        if dst_index >> (2 * A):
            rejected.append(tx)
            continue
        src_before, dst_before0 = state.get(src_index), state.get(dst_index)
        sti = src_before.find_token_index(T, tx.amount.token_id, False)
        dti = dst_before0.find_token_index(T, tx.amount.token_id, True)
        sfi = src_before.find_token_index(T, tx.fee.token_id, False)
        if sti is None or dti is None or sfi is None:
            rejected.append(tx)
            continue
        src_token = src_before.tokens[sti]
        dst_token0 = dst_before0.tokens.get(dti)
        if (tx.nonce != src_before.tx_nonce + 1 or src_before.address != src_addr
                or (N.jj_on_curve(dst_before0.address) and dst_before0.address != dst_addr)
                or (dst_token0 is not None and src_token.token_id != dst_token0.token_id)
                or src_token.token_id != tx.amount.token_id or src_token.amount < tx.amount.amount):
            rejected.append(tx)
            continue
        snap = state.snapshot()
        pre_root = state.root
        src_proof = state.prove(src_index)
        src_balance_proof = state.prove_token(src_index, sti)
        src_after = src_before.copy()
        src_after.tx_nonce += 1
        src_after.tokens[sti].amount -= tx.amount.amount
        state.write(src_index, src_after)
        src_fee_token = src_after.tokens.get(sfi)
        if src_fee_token is None or src_fee_token.token_id != tx.fee.token_id or src_fee_token.amount < tx.fee.amount:
            state.restore(snap)
            rejected.append(tx)
            continue
        src_fee_token = Money(src_fee_token.token_id, src_fee_token.amount)
        src_fee_balance_proof = state.prove_token(src_index, sfi)
        src_after.tokens[sfi].amount -= tx.fee.amount
        state.write(src_index, src_after)
        dst_proof = state.prove(dst_index)
        dst_balance_proof = state.prove_token(dst_index, dti)
        dst_before = state.get(dst_index)
        dst_token = dst_before.tokens.get(dti)
        dst_after = dst_before.copy()
        dst_after.address = dst_addr
        dst_after.tokens.setdefault(dti, Money(tx.amount.token_id, 0)).amount += tx.amount.amount
        state.write(dst_index, dst_after)
        if dst_new:
            state.new_account_indices[dst_addr] = dst_index
        transitions.append(UpdateTransition(
            True, tx, src_before, src_before.tokens_tree(T).root, Money(src_token.token_id, src_token.amount), src_fee_token,
            src_proof, src_index, sti, src_balance_proof, sfi, src_fee_balance_proof,
            dst_before, dst_before.tokens_tree(T).root, Money(dst_token.token_id, dst_token.amount) if dst_token else Money(),
            dst_proof, dst_index, dti, dst_balance_proof, pre_root))
        fee_sum += tx.fee.amount
    public = {"state": prev_root, "aux_data": N.poseidon([fee_token, fee_sum]), "next_state": state.root}
    return public, transitions, rejected


def _clone_tree(t):
    c = N.SparseTree4.__new__(N.SparseTree4)
    c.depth, c.defaults, c.levels = t.depth, t.defaults, [dict(l) for l in t.levels]
    return c


# ---------------------------------------------------------------------------------------------
class UpdateCircuit:
    """`impl Circuit<Fr> for UpdateCircuit` — update_circuit.rs:49-494"""

    def __init__(self, log4_tree, log4_token, log4_batch, commitment=0, height=0, state=0, aux_data=0, next_state=0,
                 fee_token=ZIESHA, transitions=None):
        self.A, self.T, self.B = log4_tree, log4_token, log4_batch
        self.commitment, self.height, self.state, self.aux_data, self.next_state = commitment, height, state, aux_data, next_state
        self.fee_token = fee_token
        n = 1 << (2 * log4_batch)
        trans = list(transitions or [])
        assert len(trans) <= n
        self.transitions = trans + [UpdateTransition.null(log4_tree, log4_token) for _ in range(n - len(trans))]

    @staticmethod
    def empty(log4_tree, log4_token, log4_batch):
        """`MpnCircuit::empty` (update_circuit.rs:29-46): the shape used for parameter generation."""
        return UpdateCircuit(log4_tree, log4_token, log4_batch, fee_token=0)

    def _prologue(self, cs):
        commitment_wit = AllocatedNum.alloc(cs, self.commitment); commitment_wit.inputize(cs)
        height_wit = AllocatedNum.alloc(cs, self.height); height_wit.inputize(cs)
        state_wit = AllocatedNum.alloc(cs, self.state); state_wit.inputize(cs)
        accepted_fee_token = AllocatedNum.alloc(cs, self.fee_token)
        aux_wit = AllocatedNum.alloc(cs, self.aux_data); aux_wit.inputize(cs)
        claimed_next_state_wit = AllocatedNum.alloc(cs, self.next_state); claimed_next_state_wit.inputize(cs)
        return state_wit, accepted_fee_token, aux_wit, claimed_next_state_wit

    def _epilogue(self, cs, state_wit, accepted_fee_token, aux_wit, claimed_next_state_wit, fee_sum):
        fee_sum_and_token_hash = G.poseidon(cs, [Number.of(accepted_fee_token), fee_sum])
        cs.enforce(LC({aux_wit.var: 1}), LC({ONE: 1}), fee_sum_and_token_hash.lc)
        cs.enforce(LC({state_wit.var: 1}), LC({ONE: 1}), LC({claimed_next_state_wit.var: 1}))

    def _tx_block(self, cs, tr, state_wit, accepted_fee_token, fee_sum):
        """one transition slot (update_circuit.rs:81-469); returns the new (state_wit, fee_sum)."""
        A, T = self.A, self.T
        num = Number.of
        enabled = Boolean.is_(AllocatedBit.alloc(cs, tr.enabled))
        tx_src_token_index = UnsignedInteger.alloc(cs, tr.src_token_index, 2 * T)
        tx_src_fee_token_index = UnsignedInteger.alloc(cs, tr.src_fee_token_index, 2 * T)
        tx_dst_token_index = UnsignedInteger.alloc(cs, tr.dst_token_index, 2 * T)
        src_tx_nonce = AllocatedNum.alloc(cs, tr.src_before.tx_nonce)
        src_withdraw_nonce = AllocatedNum.alloc(cs, tr.src_before.withdraw_nonce)
        src_addr = G.AllocatedPoint.alloc(cs, tr.src_before.address)
        src_addr.assert_on_curve(cs, enabled)
        src_before_balances_hash = AllocatedNum.alloc(cs, tr.src_before_balances_hash)
        dst_before_balances_hash = AllocatedNum.alloc(cs, tr.dst_before_balances_hash)
        src_token_id = AllocatedNum.alloc(cs, tr.src_before_balance.token_id)
        src_balance = UnsignedInteger.alloc_64(cs, tr.src_before_balance.amount)
        src_token_balance_hash = G.poseidon(cs, [num(src_token_id), num(src_balance)])
        src_fee_token_id = AllocatedNum.alloc(cs, tr.src_before_fee_balance.token_id)
        src_fee_balance = UnsignedInteger.alloc_64(cs, tr.src_before_fee_balance.amount)
        src_fee_token_balance_hash = G.poseidon(cs, [num(src_fee_token_id), num(src_fee_balance)])
        src_balance_proof = G.alloc_proof(cs, tr.src_balance_proof)
        G.check_proof_poseidon4(cs, enabled, tx_src_token_index, src_token_balance_hash, src_balance_proof, num(src_before_balances_hash))
        tx_amount = UnsignedInteger.alloc_64(cs, tr.tx.amount.amount)
        tx_fee = UnsignedInteger.alloc_64(cs, tr.tx.fee.amount)
        new_token_balance_hash = G.poseidon(cs, [num(src_token_id), num(src_balance) - num(tx_amount)])
        balance_middle_root = G.calc_root_poseidon4(cs, tx_src_token_index, new_token_balance_hash, src_balance_proof)
        src_fee_balance_proof = G.alloc_proof(cs, tr.src_fee_balance_proof)
        G.check_proof_poseidon4(cs, enabled, tx_src_fee_token_index, src_fee_token_balance_hash, src_fee_balance_proof, balance_middle_root)
        new_fee_token_balance_hash = G.poseidon(cs, [num(src_fee_token_id), num(src_fee_balance) - num(tx_fee)])
        src_balance_final_root = G.calc_root_poseidon4(cs, tx_src_fee_token_index, new_fee_token_balance_hash, src_fee_balance_proof)
        tx_nonce = AllocatedNum.alloc(cs, tr.tx.nonce)
        tx_src_index = UnsignedInteger.alloc(cs, tr.src_index, 2 * A)
        tx_amount_token_id = AllocatedNum.alloc(cs, tr.tx.amount.token_id)
        tx_fee_token_id = AllocatedNum.alloc(cs, tr.tx.fee.token_id)
        num(accepted_fee_token).assert_equal_if_enabled(cs, enabled, num(tx_fee_token_id))
        num(src_token_id).assert_equal(cs, num(tx_amount_token_id))
        num(src_fee_token_id).assert_equal(cs, num(tx_fee_token_id))
        src_hash = G.poseidon(cs, [num(src_tx_nonce), num(src_withdraw_nonce), num(src_addr.x), num(src_addr.y), num(src_before_balances_hash)])
        dst_token_id = AllocatedNum.alloc(cs, tr.dst_before_balance.token_id)
        dst_balance = AllocatedNum.alloc(cs, tr.dst_before_balance.amount)
        dst_token_balance_hash = G.poseidon(cs, [num(dst_token_id), num(dst_balance)])
        new_dst_token_balance_hash = G.poseidon(cs, [num(tx_amount_token_id), num(dst_balance) + num(tx_amount)])
        dst_balance_proof = G.alloc_proof(cs, tr.dst_balance_proof)
        G.check_proof_poseidon4(cs, enabled, tx_dst_token_index, dst_token_balance_hash, dst_balance_proof, num(dst_before_balances_hash))
        dst_balance_final_root = G.calc_root_poseidon4(cs, tx_dst_token_index, new_dst_token_balance_hash, dst_balance_proof)
        src_proof = G.alloc_proof(cs, tr.src_proof)
        G.check_proof_poseidon4(cs, enabled, tx_src_index, src_hash, src_proof, num(state_wit))
        new_src_tx_nonce = num(src_tx_nonce) + Number.constant(1)
        new_src_hash = G.poseidon(cs, [new_src_tx_nonce, num(src_withdraw_nonce), num(src_addr.x), num(src_addr.y), src_balance_final_root])
        middle_root = G.calc_root_poseidon4(cs, tx_src_index, new_src_hash, src_proof)
        tx_dst_addr = G.AllocatedPoint.alloc(cs, N.jj_decompress(tr.tx.dst_pub_key))
        tx_dst_addr.assert_on_curve(cs, enabled)
        tx_dst_index = UnsignedInteger.alloc(cs, tr.dst_index, 2 * A)
        dst_tx_nonce = AllocatedNum.alloc(cs, tr.dst_before.tx_nonce)
        dst_withdraw_nonce = AllocatedNum.alloc(cs, tr.dst_before.withdraw_nonce)
        dst_addr = G.AllocatedPoint.alloc(cs, tr.dst_before.address)
        dst_hash = G.poseidon(cs, [num(dst_tx_nonce), num(dst_withdraw_nonce), num(dst_addr.x), num(dst_addr.y), num(dst_before_balances_hash)])
        dst_proof = G.alloc_proof(cs, tr.dst_proof)
        is_dst_null = dst_addr.is_null(cs)
        is_dst_and_tx_dst_equal = dst_addr.is_equal(cs, tx_dst_addr)
        addr_valid = G.boolean_or(cs, is_dst_null, is_dst_and_tx_dst_equal)
        G.assert_true(cs, addr_valid)
        G.check_proof_poseidon4(cs, enabled, tx_dst_index, dst_hash, dst_proof, middle_root)
        new_dst_hash = G.poseidon(cs, [num(dst_tx_nonce), num(dst_withdraw_nonce), num(tx_dst_addr.x), num(tx_dst_addr.y), dst_balance_final_root])
        next_state = G.calc_root_poseidon4(cs, tx_dst_index, new_dst_hash, dst_proof)
        state_wit = G.mux(cs, enabled, num(state_wit), next_state)
        tx_balance_plus_fee_64 = UnsignedInteger.constrain(cs, num(tx_amount) + num(tx_fee), 64)
        is_lte = tx_balance_plus_fee_64.lte(cs, src_balance)
        G.assert_true(cs, is_lte)
        num(tx_nonce).assert_equal_if_enabled(cs, enabled, num(src_tx_nonce) + Number.constant(1))
        final_fee = G.mux(cs, enabled, Number.zero(), num(tx_fee))
        fee_sum = fee_sum.add_num(1, final_fee)
        self._last_final_fee = final_fee
        tx_hash = G.poseidon(cs, [num(tx_nonce), num(tx_dst_addr.x), num(tx_dst_addr.y), num(tx_amount_token_id), num(tx_amount),
                                  num(tx_fee_token_id), num(tx_fee)])
        tx_sig_r = G.AllocatedPoint.alloc(cs, tr.tx.sig["r"])
        tx_sig_r.assert_on_curve(cs, enabled)
        tx_sig_s = AllocatedNum.alloc(cs, tr.tx.sig["s"])
        G.verify_eddsa(cs, enabled, src_addr, tx_hash, tx_sig_r, tx_sig_s)
        return state_wit, fee_sum

    def synthesize(self, cs: ConstraintSystem):
        state_wit, accepted_fee_token, aux_wit, claimed = self._prologue(cs)
        fee_sum = Number.zero()
        for tr in self.transitions:
            state_wit, fee_sum = self._tx_block(cs, tr, state_wit, accepted_fee_token, fee_sum)
        self._epilogue(cs, state_wit, accepted_fee_token, aux_wit, claimed, fee_sum)
        return cs
