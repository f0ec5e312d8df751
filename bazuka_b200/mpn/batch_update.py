"""`mpn::update::update` with the hashing done in batches on the GPU.

The reference builds the transitions of an update batch one transaction at a time: per transaction 5
`prove` calls and 3 `set_mpn_account` calls on the Poseidon state, O(700) sequential hashes
(/root/reference/src/mpn/update.rs:40-258, /root/reference/src/zk/state/mod.rs:218-264,310-420).  `update()` in
update.py restates that loop.  Here the same transitions come out of two phases:

  1. ledger logic, sequential, NO hashing: acceptance rules, balances, nonces, slot choice — on an
     in-memory mirror of the touched accounts (identical decisions to update());
  2. hashing, batched: token-leaf Poseidon-2 and account-leaf Poseidon-5 as two batch calls, and the
     token forest and the state tree as two calls of the versioned level-synchronous tree update
     (csrc/poseidon.cu `k_tree4_versioned_level`: one launch per level for the whole batch), which returns
     the proof every write saw and the root every write produced.

The result is field-for-field the output of update() (tests compare them), the state object ends in the
same tree, and a 256-transaction batch costs ~20 kernel launches instead of ~180 000 dependent hashes."""
import numpy as np

from . import native as N
from .cs import R, to_mont
from .update import ZIESHA, Money, MpnAccount, MpnState, UpdateTransition

_RINV = pow(1 << 256, -1, R)


def _from_mont_rows(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int.from_bytes(row.tobytes(), "little") * _RINV % R for row in a]


class GpuTreeHasher:
    """the two batched primitives on a `bazuka_b200.Context` (values cross as Python ints)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def poseidon_batch(self, rows):
        if not rows:
            return []
        arity = len(rows[0])
        flat = to_mont([v for r in rows for v in r]).reshape(len(rows), arity, 4)
        return _from_mont_rows(self.ctx.poseidon(flat))

    def tree_update(self, depth, tree_ids, indices, leaf_values, init_proofs):
        n = len(indices)
        if n == 0:
            return [[] for _ in range(depth + 1)], []
        init = to_mont([v for p in init_proofs for lvl in p for v in lvl]).reshape(n, depth, 3, 4)
        vals, proofs = self.ctx.tree4_versioned_update(depth, np.asarray(tree_ids, dtype=np.uint32), np.asarray(indices, dtype=np.uint64),
                                                        to_mont(leaf_values), init)
        v = _from_mont_rows(vals)
        p = _from_mont_rows(proofs)
        vals_i = [v[l * n:(l + 1) * n] for l in range(depth + 1)]
        proofs_i = [[p[(e * depth + l) * 3:(e * depth + l) * 3 + 3] for l in range(depth)] for e in range(n)]
        return vals_i, proofs_i


_TOKEN_DEFAULTS = {}


def _token_defaults(T):
    if T not in _TOKEN_DEFAULTS:
        _TOKEN_DEFAULTS[T] = N.SparseTree4(T, N.poseidon([0, 0])).defaults
    return _TOKEN_DEFAULTS[T]


def update_batched(hasher, state: MpnState, txs, log4_batch, fee_token=ZIESHA):
    """same contract as update.update(): -> (public dict, transitions, rejected); `state` is advanced."""
    A, T = state.A, state.T
    cap = 1 << (2 * log4_batch)
    prev_root = state.root
    # ------------------------------------------------------------------ phase 1: ledger logic on a mirror
    mirror = {}                                  # account index -> MpnAccount (current value)
    pending = dict(state.new_account_indices)    # this fork's new accounts (update.rs:47-70), committed with the batch

    def index_of(addr):
        i = state.address_index.get(addr)
        return i if i is not None else pending.get(addr)

    def get(i):
        if i not in mirror:
            mirror[i] = state.accounts.get(i, MpnAccount()).copy()
        return mirror[i].copy()

    plan, rejected, fee_sum = [], [], 0
    for tx in txs:
        if len(plan) == cap:
            break
        src_addr, dst_addr = N.jj_decompress_checked(tx.src_pub_key), N.jj_decompress_checked(tx.dst_pub_key)
        if tx.fee.token_id != fee_token or src_addr is None or dst_addr is None:
            rejected.append(tx)
            continue
        src_index = index_of(src_addr)
        if src_index is None:
            rejected.append(tx)
            continue
        dst_index = index_of(dst_addr)
        dst_new = dst_index is None
        if dst_new:
            dst_index = state.account_count + len(pending)
        if dst_index >> (2 * A):
            rejected.append(tx)
            continue
        src_before, dst_before0 = get(src_index), get(dst_index)
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
        src_mid = src_before.copy()
        src_mid.tx_nonce += 1
        src_mid.tokens[sti].amount -= tx.amount.amount
        fee_tok = src_mid.tokens.get(sfi)
        if fee_tok is None or fee_tok.token_id != tx.fee.token_id or fee_tok.amount < tx.fee.amount:
            rejected.append(tx)
            continue
        src_fee_token = Money(fee_tok.token_id, fee_tok.amount)
        src_after = src_mid.copy()
        src_after.tokens[sfi].amount -= tx.fee.amount
        mirror[src_index] = src_after.copy()
        dst_before = get(dst_index)
        dst_token = dst_before.tokens.get(dti)
        dst_after = dst_before.copy()
        dst_after.address = dst_addr
        dst_after.tokens.setdefault(dti, Money(tx.amount.token_id, 0)).amount += tx.amount.amount
        mirror[dst_index] = dst_after.copy()
        if dst_new:
            pending[dst_addr] = dst_index
        plan.append(dict(tx=tx, src_index=src_index, dst_index=dst_index, sti=sti, sfi=sfi, dti=dti, src_before=src_before,
                         src_mid=src_mid, src_after=src_after, dst_before=dst_before, dst_after=dst_after,
                         src_token=Money(src_token.token_id, src_token.amount), src_fee_token=src_fee_token,
                         dst_token=Money(dst_token.token_id, dst_token.amount) if dst_token else Money()))
        fee_sum += tx.fee.amount
    # ------------------------------------------------------------------ phase 2a: the token forest
    touched = list(dict.fromkeys(i for p in plan for i in (p["src_index"], p["dst_index"])))
    forest = _TokenForest(state, touched)        # pre-batch token trees enter as writes into empty trees
    for p in plan:
        p["e1"] = forest.write(p["src_index"], p["sti"], p["src_mid"].tokens[p["sti"]])
        p["e2"] = forest.write(p["src_index"], p["sfi"], p["src_after"].tokens[p["sfi"]])
        p["e3"] = forest.write(p["dst_index"], p["dti"], p["dst_after"].tokens[p["dti"]])
    forest.run(hasher)
    acct_rows = []
    for p in plan:
        p["src_before_balances_hash"] = forest.root(p["src_index"])
        p["src_balance_proof"], p["src_fee_balance_proof"] = forest.proofs[p["e1"]], forest.proofs[p["e2"]]
        r1 = forest.applied(p["src_index"], p["e1"])
        r2 = forest.applied(p["src_index"], p["e2"])
        p["dst_before_balances_hash"] = forest.root(p["dst_index"])
        p["dst_balance_proof"] = forest.proofs[p["e3"]]
        r3 = forest.applied(p["dst_index"], p["e3"])
        sm, sa, da = p["src_mid"], p["src_after"], p["dst_after"]
        acct_rows.append([sm.tx_nonce, sm.withdraw_nonce, sm.address[0], sm.address[1], r1])
        acct_rows.append([sa.tx_nonce, sa.withdraw_nonce, sa.address[0], sa.address[1], r2])
        acct_rows.append([da.tx_nonce, da.withdraw_nonce, da.address[0], da.address[1], r3])
    # ------------------------------------------------------------------ phase 2b: the state tree
    s_leaves = hasher.poseidon_batch(acct_rows)
    s_idx = [i for p in plan for i in (p["src_index"], p["src_index"], p["dst_index"])]
    init = [state.tree.prove(i) for i in s_idx]
    s_vals, s_proofs = hasher.tree_update(A, [0] * len(s_idx), s_idx, s_leaves, init)
    transitions, root = [], prev_root
    for k, p in enumerate(plan):
        transitions.append(UpdateTransition(
            True, p["tx"], p["src_before"], p["src_before_balances_hash"], p["src_token"], p["src_fee_token"],
            s_proofs[3 * k], p["src_index"], p["sti"], p["src_balance_proof"], p["sfi"], p["src_fee_balance_proof"],
            p["dst_before"], p["dst_before_balances_hash"], p["dst_token"], s_proofs[3 * k + 2], p["dst_index"], p["dti"],
            p["dst_balance_proof"], root))
        root = s_vals[A][3 * k + 2]
    # ------------------------------------------------------------------ commit: accounts + the nodes every write left behind
    _commit(state, mirror, touched, s_idx, s_vals)
    state.new_account_indices = pending
    assert state.root == root
    public = {"state": prev_root, "aux_data": hasher.poseidon_batch([[fee_token, fee_sum]])[0], "next_state": root}
    return public, transitions, rejected


# ---------------------------------------------------------------------------------------------
# deposit / withdraw builders on the same primitives
# ---------------------------------------------------------------------------------------------
class _Ledger:
    """phase-1 mirror shared by the builders: touched accounts, address index, free-slot counter"""

    def __init__(self, state):
        self.state, self.mirror = state, {}
        self.pending = dict(state.new_account_indices)

    def index_of(self, addr):
        i = self.state.address_index.get(addr)
        return i if i is not None else self.pending.get(addr)

    def new_index(self):
        return self.state.account_count + len(self.pending)

    def get(self, i):
        if i not in self.mirror:
            self.mirror[i] = self.state.accounts.get(i, MpnAccount()).copy()
        return self.mirror[i].copy()

    def put(self, i, acc):
        self.mirror[i] = acc.copy()


class _TokenForest:
    """collects the ordered token-leaf writes of a batch; pre-batch tokens enter as writes into empty trees"""

    def __init__(self, state, accounts):
        self.T = state.T
        self.tree_of = {a: k for k, a in enumerate(accounts)}
        self.rows, self.tree, self.idx = [], [], []
        for acc in accounts:
            for i, m in state.accounts.get(acc, MpnAccount()).tokens.items():
                self.write(acc, i, m)
        self.n_init = len(self.rows)

    def write(self, acc, index, money):
        self.rows.append([money.token_id, money.amount]); self.tree.append(self.tree_of[acc]); self.idx.append(index)
        return len(self.rows) - 1

    def run(self, hasher):
        tdef = _token_defaults(self.T)
        leaves = hasher.poseidon_batch(self.rows)
        self.vals, self.proofs = hasher.tree_update(self.T, self.tree, self.idx, leaves, [[[tdef[l]] * 3 for l in range(self.T)]] * len(self.rows))
        self.cur = {k: tdef[self.T] for k in self.tree_of.values()}
        for e in range(self.n_init):
            self.cur[self.tree[e]] = self.vals[self.T][e]

    def root(self, acc):
        return self.cur[self.tree_of[acc]]

    def applied(self, acc, e):
        """advance the replay past write e; returns the root after it"""
        self.cur[self.tree_of[acc]] = self.vals[self.T][e]
        return self.vals[self.T][e]


def _commit(state, mirror, touched, s_idx, s_vals):
    A = state.A
    for e, i in enumerate(s_idx):
        node = i
        for lvl in range(A + 1):
            state.tree._put(lvl, node, s_vals[lvl][e])
            node >>= 2
    for i in touched:
        state.state_size += state.leaf_count(mirror[i]) - state.leaf_count(state.accounts.get(i))
        state.accounts[i] = mirror[i].copy()


def _list_root(hasher, rows):
    leaves = hasher.poseidon_batch(rows)
    while len(leaves) != 1:
        leaves = hasher.poseidon_batch([leaves[i:i + 4] for i in range(0, len(leaves), 4)])
    return leaves[0]


def deposit_batched(hasher, state: MpnState, deposits, log4_batch):
    """dw.deposit() with batched hashing (/root/reference/src/mpn/deposit.rs:11-233): same transitions, public inputs, state."""
    from .dw import DepositTransition
    n, prev = 1 << (2 * log4_batch), state.root
    led, plan = _Ledger(state), []
    rejected_srcs = set()                 # deposit.rs:33 `rejected_pub_keys`

    def reject(d):
        if d.src is not None:
            rejected_srcs.add(d.src)

    for d in deposits:
        if len(plan) == n:
            break
        addr = N.jj_decompress_checked(d.mpn_address)
        if addr is None:
            reject(d)
            continue
        idx = led.index_of(addr)
        is_new = idx is None
        if is_new:
            idx = led.new_index()
        if idx >> (2 * state.A):
            reject(d)
            continue
        before = led.get(idx)
        ti = before.find_token_index(state.T, d.token_id, True)
        if (ti is None or (d.src is not None and d.src in rejected_srcs)
                or (N.jj_on_curve(before.address) and before.address != addr)):
            reject(d)
            continue
        bal = before.tokens.get(ti)
        after = before.copy()
        after.address = addr
        after.tokens.setdefault(ti, Money(d.token_id, 0)).amount += d.amount
        led.put(idx, after)
        if is_new:
            led.pending[addr] = idx
        plan.append(dict(d=d, idx=idx, ti=ti, before=before, after=after, bal=Money(bal.token_id, bal.amount) if bal else Money(), addr=addr))
    touched = list(dict.fromkeys(p["idx"] for p in plan))
    forest = _TokenForest(state, touched)
    for p in plan:
        p["e"] = forest.write(p["idx"], p["ti"], p["after"].tokens[p["ti"]])
    forest.run(hasher)
    acct_rows = []
    for p in plan:
        p["before_balances_hash"] = forest.root(p["idx"])
        p["bproof"] = forest.proofs[p["e"]]
        a = p["after"]
        acct_rows.append([a.tx_nonce, a.withdraw_nonce, a.address[0], a.address[1], forest.applied(p["idx"], p["e"])])
    s_idx = [p["idx"] for p in plan]
    s_vals, s_proofs = hasher.tree_update(state.A, [0] * len(plan), s_idx, hasher.poseidon_batch(acct_rows), [state.tree.prove(i) for i in s_idx])
    trans, root = [], prev
    for k, p in enumerate(plan):
        trans.append(DepositTransition(True, p["d"], p["before"], p["before_balances_hash"], p["bal"], s_proofs[k], p["idx"], p["ti"], p["bproof"], root))
        root = s_vals[state.A][k]
    _commit(state, led.mirror, touched, s_idx, s_vals)
    state.new_account_indices = led.pending
    pk_hashes = hasher.poseidon_batch([[p["addr"][0], p["addr"][1]] for p in plan])
    rows = [[1, p["d"].token_id, p["d"].amount, pk_hashes[k]] for k, p in enumerate(plan)] + [[0, 0, 0, 0]] * (n - len(plan))
    return {"state": prev, "aux_data": _list_root(hasher, rows), "next_state": root}, trans


def withdraw_batched(hasher, state: MpnState, withdraws, log4_batch):
    """dw.withdraw() with batched hashing (/root/reference/src/mpn/withdraw.rs:10-259)."""
    from .dw import WithdrawTransition
    n, prev = 1 << (2 * log4_batch), state.root
    led, plan = _Ledger(state), []
    for w in withdraws:
        if len(plan) == n:
            break
        addr = N.jj_decompress_checked(w.mpn_address)
        idx = led.index_of(addr) if addr is not None else None
        if idx is None:
            continue
        before = led.get(idx)
        ti = before.find_token_index(state.T, w.amount.token_id, False)
        fi = before.find_token_index(state.T, w.fee.token_id, False)
        if ti is None or fi is None or w.mpn_withdraw_nonce != before.withdraw_nonce + 1:
            continue
        if w.calldata is not None and w.calldata != w.expected_calldata():
            continue
        if before.tokens[ti].amount < w.amount.amount or not N.eddsa_verify(addr, w.message(), w.mpn_sig):
            continue
        tok = before.tokens[ti]
        mid = before.copy()
        mid.tokens[ti].amount -= w.amount.amount
        feeb = mid.tokens[fi]
        if feeb.amount < w.fee.amount:
            continue
        fee_before = Money(feeb.token_id, feeb.amount)
        after = mid.copy()
        after.tokens[fi].amount -= w.fee.amount
        after.withdraw_nonce += 1
        led.put(idx, after)
        plan.append(dict(w=w, idx=idx, ti=ti, fi=fi, before=before, mid=mid, after=after, tok=Money(tok.token_id, tok.amount), fee_before=fee_before, addr=addr))
    touched = list(dict.fromkeys(p["idx"] for p in plan))
    forest = _TokenForest(state, touched)
    for p in plan:
        p["e1"] = forest.write(p["idx"], p["ti"], p["mid"].tokens[p["ti"]])
        p["e2"] = forest.write(p["idx"], p["fi"], p["after"].tokens[p["fi"]])
    forest.run(hasher)
    acct_rows = []
    for p in plan:
        p["before_token_hash"] = forest.root(p["idx"])
        p["tproof"], p["fproof"] = forest.proofs[p["e1"]], forest.proofs[p["e2"]]
        m, a = p["mid"], p["after"]
        acct_rows.append([m.tx_nonce, m.withdraw_nonce, m.address[0], m.address[1], forest.applied(p["idx"], p["e1"])])
        acct_rows.append([a.tx_nonce, a.withdraw_nonce, a.address[0], a.address[1], forest.applied(p["idx"], p["e2"])])
    s_idx = [p["idx"] for p in plan for _ in (0, 1)]
    s_vals, s_proofs = hasher.tree_update(state.A, [0] * len(s_idx), s_idx, hasher.poseidon_batch(acct_rows), [state.tree.prove(i) for i in s_idx])
    trans, root = [], prev
    for k, p in enumerate(plan):
        trans.append(WithdrawTransition(True, p["w"], p["before"], p["tok"], p["fee_before"], s_proofs[2 * k], p["idx"], p["ti"], p["tproof"],
                                        p["before_token_hash"], p["fi"], p["fproof"], root))
        root = s_vals[state.A][2 * k + 1]
    _commit(state, led.mirror, touched, s_idx, s_vals)
    cds = hasher.poseidon_batch([[p["addr"][0], p["addr"][1], p["w"].mpn_withdraw_nonce, p["w"].mpn_sig["r"][0], p["w"].mpn_sig["r"][1], p["w"].mpn_sig["s"]]
                                 for p in plan])
    rows = [[1, p["w"].amount.token_id, p["w"].amount.amount, p["w"].fee.token_id, p["w"].fee.amount, p["w"].fingerprint, cds[k]]
            for k, p in enumerate(plan)] + [[0] * 7] * (n - len(plan))
    return {"state": prev, "aux_data": _list_root(hasher, rows), "next_state": root}, trans
