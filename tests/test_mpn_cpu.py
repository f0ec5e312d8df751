"""CPU tier: the MPN host layer (bazuka_b200/mpn) against the reference's own fixtures and test shapes:
gadget truth tables (gadgets/common/test.rs:51-56,113-141,189-198), EdDSA accept / reject / disabled
(gadgets/eddsa/test.rs:64-95), Poseidon / Merkle gadgets vs the native functions, the empty-MPN-root
constant, JubJub constants, and UpdateCircuit satisfiability for the reference's tested shape
(A=3,T=3,B=1, mpn/circuits/test.rs:117-149) with null and real transitions."""
import copy
import json
import os

import pytest

from bazuka_b200.mpn import cs as C, gadgets as G, native as N, update as U

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def make_state(A, T, nacc, bal=10 ** 12):
    st, keys = U.MpnState(A, T), []
    for i in range(nacc):
        pk, sk = N.eddsa_keys(b"acct%d" % i)
        keys.append((pk, sk))
        st.set(i, U.MpnAccount(0, 0, pk, {0: U.Money(U.ZIESHA, bal)}))
    return st, keys


def transfer(keys, s, d, nonce, amount=1000, fee=10):
    tx = U.MpnTransaction(nonce, N.jj_compress(keys[s][0]), N.jj_compress(keys[d][0]), U.Money(U.ZIESHA, amount), U.Money(U.ZIESHA, fee))
    tx.sign(keys[s][1])
    return tx


def test_native_poseidon_kats_and_empty_root():
    kats = json.load(open(f"{GOLD}/poseidon_kats.json"))["expected_decimal"]
    assert all(N.poseidon(list(range(n))) == int(k) for n, k in enumerate(kats, 1))
    st = U.MpnState(30, 1)
    want = [h for h in json.load(open(f"{GOLD}/empty_root.json"))["hex_scalars_in_vector"] if int(h, 16)][0]
    assert st.root == int(want, 16)


def test_jubjub_constants_and_eddsa():
    assert N.jj_on_curve(N.JJ_BASE) and N.jj_mul(N.JJ_BASE, N.JJ_ORDER) == (0, 1)
    assert N.JJ_BASE_COFACTOR == N.jj_mul(N.JJ_BASE, 8)
    p = N.jj_mul(N.JJ_BASE, 12345)
    assert N.jj_add(p, N.jj_mul(N.JJ_BASE, 55)) == N.jj_mul(N.JJ_BASE, 12400)
    assert N.jj_decompress(N.jj_compress(p)) == p
    assert N.jj_decompress((0, False)) == U.NULL_DST
    pk, sk = N.eddsa_keys(b"ABC")
    sig = N.eddsa_sign(sk, 777)
    assert N.eddsa_verify(pk, 777, sig) and not N.eddsa_verify(pk, 778, sig)


@pytest.mark.parametrize("a,b", [(0, 0), (1, 0), (0, 1), (5, 5), (2 ** 64 - 1, 2 ** 64 - 1), (7, 2 ** 63), (2 ** 64 - 1, 0)])
def test_comparison_truth_tables(a, b):
    cs = C.ConstraintSystem()
    ua, ub = G.UnsignedInteger.alloc_64(cs, a), G.UnsignedInteger.alloc_64(cs, b)
    got = (ua.lt(cs, ub).value, ua.lte(cs, ub).value, ua.gt(cs, ub).value, ua.gte(cs, ub).value,
           G.Number.of(ua).is_equal(cs, G.Number.of(ub)).value)
    assert got == (int(a < b), int(a <= b), int(a > b), int(a >= b), int(a == b))
    assert cs.is_satisfied()[0]


def test_boolean_gadgets_and_mux():
    for x in (0, 1):
        for y in (0, 1):
            cs = C.ConstraintSystem()
            bx, by = C.Boolean.is_(C.AllocatedBit.alloc(cs, x)), C.Boolean.is_(C.AllocatedBit.alloc(cs, y))
            assert G.boolean_or(cs, bx, by).value == (x | y)
            assert C.Boolean.and_(cs, bx, by.not_()).value == (x & (1 - y))
            assert C.Boolean.and_(cs, bx.not_(), by.not_()).value == ((1 - x) & (1 - y))
            m = G.mux(cs, bx, G.Number.constant(10), G.Number.constant(20))
            assert m.value == (20 if x else 10)
            assert cs.is_satisfied()[0]
    # a non-boolean witness for a bit is rejected by the booleanity row
    cs = C.ConstraintSystem()
    bit = C.AllocatedBit.alloc(cs, 1)
    cs.aux[bit.var >> 1] = 2
    assert not cs.is_satisfied()[0]


def test_strict_bit_decomposition():
    for v in (0, 1, N.R - 1, N.R - 2, 2 ** 254, 123456789):
        cs = C.ConstraintSystem()
        bits = C.AllocatedNum.alloc(cs, v).to_bits_le_strict(cs)
        assert len(bits) == 255 and sum(b.value << i for i, b in enumerate(bits)) == v
        assert cs.is_satisfied()[0]
    assert cs.num_constraints in range(380, 400)  # ~388


def test_poseidon_and_merkle_gadgets_match_native():
    for vals in ([1, 2], [3, 4, 5, 6], [1, 2, 3, 4, 5], [9, 8, 7, 6, 5, 4, 3]):
        cs = C.ConstraintSystem()
        out = G.poseidon(cs, [G.Number.of(C.AllocatedNum.alloc(cs, v)) for v in vals])
        assert out.value == N.poseidon(vals) and cs.is_satisfied()[0]
    assert cs.num_constraints == 762  # arity 7: 24t + R_P(t+2)
    tree = N.SparseTree4(3, 0)
    tree.set_leaf(37, 1234)
    cs = C.ConstraintSystem()
    idx = G.UnsignedInteger.alloc(cs, 37, 6)
    root = G.calc_root_poseidon4(cs, idx, G.Number.of(C.AllocatedNum.alloc(cs, 1234)), G.alloc_proof(cs, tree.prove(37)))
    assert root.value == tree.root and cs.is_satisfied()[0]


def test_eddsa_gadget_accept_reject_disabled():
    pk, sk = N.eddsa_keys(b"ABC")
    sig = N.eddsa_sign(sk, 123456)

    def run(enabled, m):
        cs = C.ConstraintSystem()
        en = C.Boolean.is_(C.AllocatedBit.alloc(cs, enabled))
        G.verify_eddsa(cs, en, G.AllocatedPoint.alloc(cs, pk), G.Number.of(C.AllocatedNum.alloc(cs, m)),
                       G.AllocatedPoint.alloc(cs, sig["r"]), C.AllocatedNum.alloc(cs, sig["s"]))
        return cs
    good = run(1, 123456)
    assert good.is_satisfied()[0] and abs(good.num_constraints - 10057) < 5
    assert not run(1, 123457).is_satisfied()[0]
    assert run(0, 123457).is_satisfied()[0]


def test_update_circuit_reference_shape_null_and_real():
    circ = U.UpdateCircuit(3, 3, 1, state=5, next_state=5, aux_data=N.poseidon([U.ZIESHA, 0]))
    cs = circ.synthesize(C.ConstraintSystem())
    assert cs.is_satisfied()[0] and len(cs.inputs) == 6
    n_null = cs.num_constraints
    st, keys = make_state(3, 3, 3)
    keys.append(N.eddsa_keys(b"newcomer"))
    txs = [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1, amount=5), transfer(keys, 0, 3, 2, amount=77, fee=3),
           transfer(keys, 2, 0, 5)]  # the last one has a wrong nonce
    pub, trans, rej = U.update(st, txs, 1)
    assert len(trans) == 3 and len(rej) == 1
    cs = U.UpdateCircuit(3, 3, 1, commitment=42, height=7, transitions=trans, **pub).synthesize(C.ConstraintSystem())
    assert cs.is_satisfied()[0] and cs.num_constraints == n_null  # shape is witness-independent
    assert cs.inputs[1:] == [42, 7, pub["state"], pub["aux_data"], pub["next_state"]]
    bad = dict(pub, next_state=pub["next_state"] + 1)
    assert not U.UpdateCircuit(3, 3, 1, transitions=trans, **bad).synthesize(C.ConstraintSystem()).is_satisfied()[0]
    forged = copy.deepcopy(trans)
    forged[0].tx.sig["s"] = (forged[0].tx.sig["s"] + 1) % N.JJ_ORDER
    assert not U.UpdateCircuit(3, 3, 1, transitions=forged, **pub).synthesize(C.ConstraintSystem()).is_satisfied()[0]
    over = copy.deepcopy(trans)
    over[1].tx.amount.amount += 1  # amount no longer matches the signed hash / balances
    assert not U.UpdateCircuit(3, 3, 1, transitions=over, **pub).synthesize(C.ConstraintSystem()).is_satisfied()[0]


def test_deposit_and_withdraw_circuits():
    """same shape as the reference's circuit tests (A=3,T=3,B=1; mpn/circuits/test.rs:157-229) with
    real transitions; aux_data is the revealed root of the batch (deposit.rs:178-218, withdraw.rs:190-245)."""
    from bazuka_b200.mpn import dw as D
    st, keys = make_state(3, 3, 2)
    newpk, _ = N.eddsa_keys(b"dep-new")
    deps = [D.MpnDeposit(N.jj_compress(keys[0][0]), U.ZIESHA, 500), D.MpnDeposit(N.jj_compress(newpk), 77, 9),
            D.MpnDeposit(N.jj_compress(keys[1][0]), 77, 1)]
    pub, tr = D.deposit(st, deps, 1)
    assert len(tr) == 3
    cs = D.DepositCircuit(3, 3, 1, commitment=3, height=1, transitions=tr, **pub).synthesize(C.ConstraintSystem())
    assert cs.is_satisfied()[0]
    assert D.DepositCircuit(3, 3, 1).synthesize(C.ConstraintSystem()).num_constraints == cs.num_constraints
    assert not D.DepositCircuit(3, 3, 1, transitions=tr, **dict(pub, aux_data=pub["aux_data"] + 1)).synthesize(C.ConstraintSystem()).is_satisfied()[0]
    null_pub = {"state": 5, "next_state": 5, "aux_data": D.native_list_root(1, [[0] * 4] * 4)}
    assert D.DepositCircuit(3, 3, 1, **null_pub).synthesize(C.ConstraintSystem()).is_satisfied()[0]
    ws = []
    for i, amt in enumerate([100, 5]):
        w = D.MpnWithdraw(N.jj_compress(keys[i][0]), 1, amount=U.Money(U.ZIESHA, amt), fee=U.Money(U.ZIESHA, 2), fingerprint=1000 + i)
        w.sign(keys[i][1])
        ws.append(w)
    pub, tr = D.withdraw(st, ws, 1)
    assert len(tr) == 2
    assert D.WithdrawCircuit(3, 3, 1, transitions=tr, **pub).synthesize(C.ConstraintSystem()).is_satisfied()[0]
    t2 = copy.deepcopy(tr)
    t2[0].tx.fingerprint += 1  # signature no longer covers the payment
    assert not D.WithdrawCircuit(3, 3, 1, transitions=t2, **pub).synthesize(C.ConstraintSystem()).is_satisfied()[0]
    null_pub = {"state": 5, "next_state": 5, "aux_data": D.native_list_root(1, [[0] * 7] * 4)}
    assert D.WithdrawCircuit(3, 3, 1, **null_pub).synthesize(C.ConstraintSystem()).is_satisfied()[0]


def test_parallel_synthesis_equals_sequential():
    """the template/replicate synthesiser used for production-size batches yields exactly the R1CS and
    witness of `UpdateCircuit.synthesize` (4 slots: 3 signed transfers + 1 null)."""
    import numpy as np
    from bazuka_b200.mpn import fastsynth as F
    st, keys = make_state(3, 3, 3)
    keys.append(N.eddsa_keys(b"newcomer"))
    txs = [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1, amount=5), transfer(keys, 0, 3, 2, amount=77, fee=3)]
    pub, trans, _ = U.update(st, txs, 1)
    circ = U.UpdateCircuit(3, 3, 1, commitment=42, height=7, transitions=trans, **pub)
    ni, na, mats, inputs, aux = circ.synthesize(C.ConstraintSystem()).to_csr()
    ni2, na2, mats2, inputs2, aux_canon = F.synthesize_update(circ, workers=2)
    assert (ni, na) == (ni2, na2) and (inputs == inputs2).all()
    assert (F.canon_to_mont_host(aux_canon) == aux).all()
    for (rp1, c1, v1), (rp2, c2, v2) in zip(mats, mats2):
        assert (rp1 == rp2).all() and (c1 == c2).all() and (v1 == v2).all()


def test_witness_program_reproduces_synthesised_witness():
    """the compiled straight-line witness program (what csrc/witness.cu interprets, one thread per slot)
    yields, slot by slot, exactly the aux values `UpdateCircuit.synthesize` assigns — for signed transfers,
    a transfer to a new account and a null slot."""
    from bazuka_b200.mpn import witness_program as W
    st, keys = make_state(3, 3, 3)
    keys.append(N.eddsa_keys(b"newcomer"))
    txs = [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1, amount=5), transfer(keys, 0, 3, 2, amount=77, fee=3)]
    pub, trans, _ = U.update(st, txs, 1)
    circ = U.UpdateCircuit(3, 3, 1, commitment=42, height=7, transitions=trans, **pub)
    cs = circ.synthesize(C.ConstraintSystem(record=True))
    prog = W.compile_update_block(3, 3)
    n = len(circ.transitions)
    a_tx = prog.n_ops
    assert len(cs.aux) > prog.p_aux + n * a_tx
    roots = W.slot_roots(circ)
    kinds = {r[0] for r in cs.recipes}
    assert kinds <= {"raw", "mul", "bit", "iszero", "invz", "select", "jjx", "jjy"}
    for k, tr in enumerate(circ.transitions):
        raws = W.raw_values(tr, 3, 3)
        assert len(raws) == prog.n_raw
        got = W.run_reference(prog, raws, [circ.fee_token, roots[k]])
        want = cs.aux[prog.p_aux + k * a_tx: prog.p_aux + (k + 1) * a_tx]
        assert got == want, (k, next(i for i, (g, w) in enumerate(zip(got, want)) if g != w))
    # chaining: the state root leaving slot k enters slot k+1
    for k in range(n - 1):
        assert cs.aux[prog.p_aux + k * a_tx + prog.state_out] == roots[k + 1]


def _assert_same_transitions(t1, t2):
    import dataclasses
    assert len(t1) == len(t2)
    for a, b in zip(t1, t2):
        da, db = dataclasses.asdict(a), dataclasses.asdict(b)
        for k in da:
            assert da[k] == db[k], k


def _batch_scenario():
    """transfers incl. a new account, a self-referencing chain, same-token fee, and rejected ones
    (bad nonce, unknown sender, overdraft, wrong fee token)."""
    st, keys = make_state(3, 3, 3)
    keys.append(N.eddsa_keys(b"newcomer"))
    keys.append(N.eddsa_keys(b"stranger"))
    txs = [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1, amount=5), transfer(keys, 0, 3, 2, amount=77, fee=3),
           transfer(keys, 0, 1, 9),                     # bad nonce
           transfer(keys, 4, 1, 1),                     # unknown sender
           transfer(keys, 2, 0, 1, amount=10 ** 13),    # overdraft
           transfer(keys, 3, 0, 1, amount=50, fee=2),   # the newcomer spends what it just received
           transfer(keys, 1, 1, 2, amount=7),           # to itself
           transfer(keys, 2, 3, 1, amount=1, fee=1)]
    bad_fee = transfer(keys, 0, 2, 3)
    bad_fee.fee = U.Money(7, 1)
    bad_fee.sign(keys[0][1])
    txs.insert(4, bad_fee)
    return st, txs


def test_batched_transition_builder_equals_sequential_update():
    """update_batched (ledger logic first, then hashing in batches through the versioned tree update) returns
    exactly update()'s transitions, public inputs, rejections and final state — here with the host stand-in
    for the GPU primitives, which applies the writes one by one like the reference's set_data/prove loop."""
    import copy
    from bazuka_b200.mpn import batch_update as BU
    from oracle.py.state import HostTreeHasher
    st1, txs = _batch_scenario()
    st2 = copy.deepcopy(st1)
    pub1, tr1, rej1 = U.update(st1, txs, 2)
    pub2, tr2, rej2 = BU.update_batched(HostTreeHasher(N.poseidon), st2, txs, 2)
    assert len(tr1) == 6 and len(rej1) == 4
    assert pub1 == pub2 and rej1 == rej2
    _assert_same_transitions(tr1, tr2)
    assert st1.root == st2.root and st1.tree.levels == st2.tree.levels
    assert {i: dataclasses_asdict(a) for i, a in st1.accounts.items()} == {i: dataclasses_asdict(a) for i, a in st2.accounts.items()}
    # the batch cap: only 4^B transactions are taken
    st3, txs3 = _batch_scenario()
    st4 = copy.deepcopy(st3)
    p3, t3, r3 = U.update(st3, txs3, 0)
    p4, t4, r4 = BU.update_batched(HostTreeHasher(N.poseidon), st4, txs3, 0)
    assert len(t3) == 1 and p3 == p4 and r3 == r4 and st3.root == st4.root
    _assert_same_transitions(t3, t4)


def dataclasses_asdict(a):
    import dataclasses
    return dataclasses.asdict(a)


def _dw_scenario(kind):
    from bazuka_b200.mpn import dw as D
    st, keys = make_state(3, 3, 2)
    if kind == "deposit":
        newpk, _ = N.eddsa_keys(b"dep-new")
        deps = [D.MpnDeposit(N.jj_compress(keys[0][0]), U.ZIESHA, 500), D.MpnDeposit(N.jj_compress(newpk), 77, 9),
                D.MpnDeposit(N.jj_compress(keys[1][0]), 77, 1)]
        pub, tr = D.deposit(st, deps, 1)
        return D.DepositCircuit(3, 3, 1, commitment=3, height=1, transitions=tr, **pub)
    ws = []
    for i, amt in enumerate([100, 5]):
        w = D.MpnWithdraw(N.jj_compress(keys[i][0]), 1, amount=U.Money(U.ZIESHA, amt), fee=U.Money(U.ZIESHA, 2), fingerprint=1000 + i)
        w.sign(keys[i][1])
        ws.append(w)
    pub, tr = D.withdraw(st, ws, 1)
    return D.WithdrawCircuit(3, 3, 1, commitment=9, height=2, transitions=tr, **pub)


@pytest.mark.parametrize("kind", ["deposit", "withdraw"])
def test_two_phase_witness_programs_reproduce_synthesis(kind):
    """deposit / withdraw: phase-1 program, host reveal, phase-2 program (reading phase-1 raws and the entering
    state as externals) assemble to exactly `synthesize`'s aux vector."""
    from bazuka_b200.mpn import dw_witness as DW, witness_program as W
    circ = _dw_scenario(kind)
    cs = circ.synthesize(C.ConstraintSystem())
    assert cs.is_satisfied()[0]
    pg = DW.TwoPhasePrograms(kind, 3, 3)
    inputs, pro, rev = DW.host_parts(pg, circ)
    roots = DW.slot_roots(circ)
    b1, b2 = [], []
    for k, tr in enumerate(circ.transitions):
        r1, r2 = pg.raws_of(tr, 3, 3)
        b1 += W.run_reference(pg.prog1, r1, [])
        out2 = W.run_reference(pg.prog2, r2, pg.ext_values(r1, roots[k]))
        b2 += out2
        if k + 1 < len(roots):
            assert out2[pg.state_out] == roots[k + 1]
    assert inputs == cs.inputs
    assert pro + b1 + rev + b2 == cs.aux


def test_batched_deposit_and_withdraw_builders_equal_sequential():
    """deposit_batched / withdraw_batched (batched hashing through the versioned tree update; here the host
    stand-in) == dw.deposit / dw.withdraw: transitions, public inputs, final state — including a deposit to a new
    address, a second token, two deposits to one account, a rejected withdraw (bad nonce) and a chain of two
    withdraws from one account."""
    import copy
    from bazuka_b200.mpn import batch_update as BU, dw as D
    from oracle.py.state import HostTreeHasher
    h = HostTreeHasher(N.poseidon)
    st1, keys = make_state(3, 3, 2)
    newpk, _ = N.eddsa_keys(b"dep-new")
    deps = [D.MpnDeposit(N.jj_compress(keys[0][0]), U.ZIESHA, 500), D.MpnDeposit(N.jj_compress(newpk), 77, 9),
            D.MpnDeposit(N.jj_compress(keys[1][0]), 77, 1), D.MpnDeposit(N.jj_compress(keys[0][0]), 77, 4),
            D.MpnDeposit(N.jj_compress(newpk), 77, 1),
            # a key that does not decompress lists its L1 source; the source's next deposit goes with it (deposit.rs:33,68-83)
            D.MpnDeposit((6, False), 77, 1, "carol"), D.MpnDeposit(N.jj_compress(keys[1][0]), 77, 2, "carol")]
    st2 = copy.deepcopy(st1)
    pub1, tr1 = D.deposit(st1, deps, 2)
    pub2, tr2 = BU.deposit_batched(h, st2, deps, 2)
    assert len(tr1) == 5 and pub1 == pub2
    _assert_same_transitions(tr1, tr2)
    assert st1.root == st2.root and st1.tree.levels == st2.tree.levels
    ws = []
    for i, amt, nonce in ((0, 100, 1), (1, 5, 1), (0, 7, 5), (0, 30, 2)):
        w = D.MpnWithdraw(N.jj_compress(keys[i][0]), nonce, amount=U.Money(U.ZIESHA, amt), fee=U.Money(U.ZIESHA, 2), fingerprint=1000 + amt)
        w.sign(keys[i][1])
        ws.append(w)
    ws[0].calldata = ws[0].expected_calldata()          # `verify_calldata` (withdraw.rs:77) when the payment's calldata is known
    bad = D.MpnWithdraw(N.jj_compress(keys[1][0]), 2, amount=U.Money(U.ZIESHA, 1), fee=U.Money(U.ZIESHA, 0), fingerprint=7)
    bad.sign(keys[1][1])
    bad.calldata = bad.expected_calldata() + 1
    ws.insert(2, bad)
    pub1, tr1 = D.withdraw(st1, ws, 1)
    pub2, tr2 = BU.withdraw_batched(h, st2, ws, 1)
    assert len(tr1) == 3 and pub1 == pub2 and all(t.tx is not bad for t in tr1 + tr2)
    _assert_same_transitions(tr1, tr2)
    assert st1.root == st2.root and st1.tree.levels == st2.tree.levels
    assert {i: dataclasses_asdict(a) for i, a in st1.accounts.items()} == {i: dataclasses_asdict(a) for i, a in st2.accounts.items()}


def test_update_epilogue_program_reproduces_synthesis():
    """the part of the witness after the slot loop (fee-commitment Poseidon gadget) as a program whose externals are
    the fee token and every slot's accepted fee — what bzk_mpn_update_witness runs after the slots."""
    from bazuka_b200.mpn import witness_program as W
    st, keys = make_state(3, 3, 3)
    pub, trans, _ = U.update(st, [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1, amount=5, fee=7)], 1)
    circ = U.UpdateCircuit(3, 3, 1, commitment=4, height=2, transitions=trans, **pub)
    cs = circ.synthesize(C.ConstraintSystem())
    prog = W.compile_update_block(3, 3)
    ep = W.compile_update_epilogue(prog, 1)
    fees = [tr.tx.fee.amount if tr.enabled else 0 for tr in circ.transitions]
    assert ep.n_raw == 0 and ep.n_ext == 5
    assert W.run_reference(ep, [], [circ.fee_token] + fees) == cs.aux[prog.p_aux + 4 * prog.n_ops:]


def test_genesis_mpn_addresses_decompress_on_curve():
    """reference fixture: the 211 `jub…` keys of the genesis allocation
    (/root/reference/src/config/initials.rs:13027+) all parse (x < r) and decompress — Fr square root plus the parity
    rule of curve.rs:78-88 — to points on the curve; the Python restatement and libbzk's host C++ agree on each."""
    import ctypes as ct
    import json
    import numpy as np
    from bazuka_b200 import _lib
    addrs = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "genesis_mpn_addresses.json")))["addresses"]
    assert len(addrs) == 211
    lib = _lib.load()
    canon = lambda v: np.frombuffer((v % N.R).to_bytes(32, "little"), dtype=np.uint64).copy()
    d = canon(N.JJ_D)
    for s in addrs:
        odd, x = s[3] == "3", int(s[4:], 16)
        assert x < N.R
        p = N.jj_decompress((x, odd))
        assert p[0] == x and (p[1] & 1 == 1) == odd and N.jj_on_curve(p)
        xc, out = canon(x), np.zeros((2, 4), dtype=np.uint64)
        assert lib.bzk_jubjub_decompress(ct.c_void_p(d.ctypes.data), ct.c_void_p(xc.ctypes.data), int(odd), ct.c_void_p(out.ctypes.data)) == 0
        assert int.from_bytes(out[0].tobytes(), "little") == p[0] and int.from_bytes(out[1].tobytes(), "little") == p[1]
    # x = 2 is not the abscissa of a point? whichever it is, both sides must agree
    for x in (2, 3, 5, 7, 11):
        rhs = (1 + x * x) * pow((1 - N.JJ_D * x * x) % N.R, -1, N.R) % N.R
        xc, out = canon(x), np.zeros((2, 4), dtype=np.uint64)
        st = lib.bzk_jubjub_decompress(ct.c_void_p(d.ctypes.data), ct.c_void_p(xc.ctypes.data), 0, ct.c_void_p(out.ctypes.data))
        assert (st == 0) == (N.fr_sqrt(rhs) is not None)


@pytest.mark.parametrize("shape", [(3, 3, 1), (5, 2, 0)])
def test_native_circuit_compiler_equals_python(shape):
    """csrc/mpn_circuit.cu — UpdateCircuit, the gadgets and bellman's vocabulary in C++, structure-only — emits
    exactly the R1CS (three CSR matrices: row pointers, columns, Montgomery coefficients) and exactly the slot and
    epilogue witness programs (ops, LC pool, coefficient table) of the Python definition."""
    from bazuka_b200.mpn import witness_program as W
    from bazuka_b200.mpn.native_circuit import NativeUpdateCircuit
    A, T, B = shape
    nc = NativeUpdateCircuit(A, T, B)
    ni, na, mats = nc.r1cs()
    cs = U.UpdateCircuit(A, T, B).synthesize(C.ConstraintSystem())
    pni, pna, pmats, _, _ = cs.to_csr()
    assert (ni, na, nc.num_constraints) == (pni, pna, cs.num_constraints)
    for (rp, col, val), (prp, pcol, pval) in zip(mats, pmats):
        assert (rp == prp).all() and (col == pcol).all() and (val == pval).all()
    want = W.compile_update_block(A, T)
    for got, ref in ((nc.program(0), want), (nc.program(1), W.compile_update_epilogue(want, B))):
        assert (got.n_raw, got.n_ext) == (ref.n_raw, ref.n_ext)
        assert (got.ops == ref.ops).all() and (got.lc_ptr == ref.lc_ptr).all()
        assert (got.lc_slot == ref.lc_slot).all() and (got.lc_coef == ref.lc_coef).all() and got.coefs == ref.coefs
    assert (nc.p_aux, nc.slot_vars, nc.state_out, nc.final_fee) == (want.p_aux, want.n_ops, want.state_out, want.final_fee)
    nc.free()


@pytest.mark.parametrize("kind", ["deposit", "withdraw"])
def test_native_two_phase_circuit_compiler_equals_python(kind):
    """DepositCircuit / WithdrawCircuit in C++ (csrc/mpn_circuit.cu): R1CS of a 4-slot batch, both slot programs and
    the placement data (row positions, phase-2 externals) equal the Python definition's."""
    from bazuka_b200.mpn import dw as D, dw_witness as DW
    from bazuka_b200.mpn.native_circuit import NativeTwoPhaseCircuit
    nc = NativeTwoPhaseCircuit(kind, 3, 3, 1)
    ni, na, mats = nc.r1cs()
    circ_cls = D.DepositCircuit if kind == "deposit" else D.WithdrawCircuit
    cs = circ_cls(3, 3, 1).synthesize(C.ConstraintSystem())
    pni, pna, pmats, _, _ = cs.to_csr()
    assert (ni, na, nc.num_constraints) == (pni, pna, cs.num_constraints)
    for (rp, col, val), (prp, pcol, pval) in zip(mats, pmats):
        assert (rp == prp).all() and (col == pcol).all() and (val == pval).all()
    pg = DW.TwoPhasePrograms(kind, 3, 3)
    assert (nc.p_aux, nc.n1, nc.n2, nc.state_out) == (pg.p_aux, pg.n1, pg.n2, pg.state_out)
    assert list(nc.row_local) == pg.row_local and nc.ext_src == pg.ext_src
    assert nc.reveal_vars == na - nc.p_aux - 4 * (nc.n1 + nc.n2)
    for got, ref in ((nc.program(0), pg.prog1), (nc.program(1), pg.prog2)):
        assert (got.n_raw, got.n_ext) == (ref.n_raw, ref.n_ext)
        assert (got.ops == ref.ops).all() and (got.lc_ptr == ref.lc_ptr).all()
        assert (got.lc_slot == ref.lc_slot).all() and (got.lc_coef == ref.lc_coef).all() and got.coefs == ref.coefs
    nc.free()


@pytest.mark.parametrize("kind", ["deposit", "withdraw"])
def test_two_phase_witness_entirely_through_programs(kind):
    """phase 1, the reveal and phase 2 as three witness programs (the reveal's externals are the rows produced by phase 1)
    reproduce `synthesize`'s aux vector with no gadget evaluation at all; the natively compiled reveal program is
    identical to the Python one."""
    from bazuka_b200.mpn import dw_witness as DW, witness_program as W
    from bazuka_b200.mpn.native_circuit import NativeTwoPhaseCircuit
    circ = _dw_scenario(kind)
    cs = circ.synthesize(C.ConstraintSystem())
    pg = DW.TwoPhasePrograms(kind, 3, 3)
    rv = DW.compile_reveal_program(pg, 1)
    roots = DW.slot_roots(circ)
    b1, b2, rows = [], [], []
    for k, tr in enumerate(circ.transitions):
        r1, r2 = pg.raws_of(tr, 3, 3)
        out1 = W.run_reference(pg.prog1, r1, [])
        b1 += out1
        rows += [out1[j] for j in pg.row_local]
        b2 += W.run_reference(pg.prog2, r2, pg.ext_values(r1, roots[k]))
    rev = W.run_reference(rv, [], rows)
    assert cs.aux[:pg.p_aux] + b1 + rev + b2 == cs.aux
    nc = NativeTwoPhaseCircuit(kind, 3, 3, 1)
    got = nc.program(2)
    assert (got.n_raw, got.n_ext) == (rv.n_raw, rv.n_ext) == (0, 4 * len(pg.row_local))
    assert (got.ops == rv.ops).all() and (got.lc_ptr == rv.lc_ptr).all() and (got.lc_slot == rv.lc_slot).all()
    assert (got.lc_coef == rv.lc_coef).all() and got.coefs == rv.coefs
    nc.free()


def test_reference_sum_hasher_membership_kat(monkeypatch):
    """the reference's own state-tree test (`test_zk_list_membership_proof`, /root/reference/src/zk/test/mod.rs:44-71):
    with the additive SumHasher, a list of 4^4 scalars i -> i, every proof's siblings plus the leaf sum to 32640.
    Checked on the sparse tree restatement (`set_leaf` / `prove`) and on the write-by-write reference semantics of
    the batched tree update (oracle.py.state.sequential_tree_updates), whose proofs are taken just before each write."""
    from oracle.py.state import sequential_tree_updates
    summing = lambda vals: sum(vals) % N.R
    monkeypatch.setattr(N, "poseidon", summing)
    t = N.SparseTree4(4, 0)
    for i in range(256):
        t.set_leaf(i, i)
    assert t.root == 32640
    for i in range(256):
        assert (i + sum(v for lvl in t.prove(i) for v in lvl)) % N.R == 32640
    # batched semantics: first pass writes i -> i into the empty tree, second pass rewrites the same values, so the
    # second pass's "proof before the write" is the proof in the final tree
    empty = N.SparseTree4(4, 0)
    idx = list(range(256)) * 2
    init = [empty.prove(i) for i in idx]
    vals, proofs = sequential_tree_updates(4, [0] * 512, idx, idx, init, summing)
    assert vals[4][255] == 32640 and vals[4][511] == 32640
    for e in range(256, 512):
        assert (idx[e] + sum(v for lvl in proofs[e] for v in lvl)) % N.R == 32640
    for e in range(256):  # during the first pass the proof sees exactly the leaves written so far
        assert sum(v for lvl in proofs[e] for v in lvl) % N.R == sum(range(e))


# ------------------------------------------------------------------ the reference's ledger rules, pinned (update.rs:29-70,256-266)
def _recount(st):
    return sum(U.MpnState.leaf_count(a) for a in st.accounts.values())


def _off_curve_key():
    x = 2
    while N.jj_decompress_checked((x, False)) is not None:
        x += 1
    return (x, False)


def test_update_prefilter_drops_keys_that_do_not_decompress():
    """update.rs:31-38: a transaction whose source or destination key is not on the curve is not eligible — the batch goes
    on without it (round 1 aborted the whole batch)."""
    from bazuka_b200.mpn import batch_update as BU
    from oracle.py.state import HostTreeHasher
    import copy
    st, keys = make_state(3, 3, 3)
    bad_dst = transfer(keys, 0, 1, 1)
    bad_dst.dst_pub_key = _off_curve_key()
    bad_src = transfer(keys, 1, 2, 1)
    bad_src.src_pub_key = _off_curve_key()
    txs = [bad_dst, transfer(keys, 0, 1, 1), bad_src, transfer(keys, 1, 2, 1, amount=5)]
    st2 = copy.deepcopy(st)
    pub, trans, rej = U.update(st, txs, 1)
    assert [t.tx for t in trans] == [txs[1], txs[3]] and rej == [txs[0], txs[2]]
    pub2, trans2, rej2 = BU.update_batched(HostTreeHasher(N.poseidon), st2, txs, 1)
    assert pub2 == pub and rej2 == rej and st2.root == st.root


def test_new_account_index_is_chain_count_plus_new_accounts_and_threads_across_batches():
    """update.rs:47-70: the receiver of an unknown address gets index `mpn_account_count + new_account_indices.len()`;
    the map is threaded across the batches built on one fork (mod.rs:330), so a later batch finds the newcomer as a
    SENDER; a fresh fork that did not see the map rejects it; once the block is applied the chain table knows it."""
    st, keys = make_state(3, 3, 3)
    st.account_count = 10                       # the chain has indexed ten accounts (seven of them empty here)
    new1, new2 = N.eddsa_keys(b"newcomer"), N.eddsa_keys(b"second")
    keys += [new1, new2]
    pub, trans, rej = U.update(st, [transfer(keys, 0, 3, 1, amount=500), transfer(keys, 1, 4, 1, amount=7)], 1)
    assert [t.dst_index for t in trans] == [10, 11] and not rej
    assert st.new_account_indices == {new1[0]: 10, new2[0]: 11} and st.account_count == 10
    # next batch on the same fork: the newcomer spends; and a third new address continues the numbering
    third = N.eddsa_keys(b"third")
    keys.append(third)
    fresh = st.fork()
    fresh.new_account_indices = {}              # a fork that never saw the first batch's map
    pub, trans, rej = U.update(st, [transfer(keys, 3, 0, 1, amount=50, fee=1), transfer(keys, 0, 5, 2, amount=1)], 1)
    assert [(t.src_index, t.dst_index) for t in trans] == [(10, 0), (0, 12)] and not rej
    _, trans_f, rej_f = U.update(fresh, [transfer(keys, 3, 0, 1, amount=50, fee=1)], 1)
    assert not trans_f and len(rej_f) == 1
    st.commit_accounts()
    assert st.account_count == 13 and st.address_index[new1[0]] == 10 and not st.new_account_indices
    pub, trans, rej = U.update(st, [transfer(keys, 4, 3, 1, amount=2, fee=1)], 1)
    assert [(t.src_index, t.dst_index) for t in trans] == [(11, 10)]


def test_state_size_counts_non_zero_leaves_through_every_builder():
    """`ZkCompressedState.state_size` (update.rs:29,256-266; state/mod.rs:327-341): +1 when a scalar leaf becomes non-zero,
    -1 when it returns to zero — through update, deposit, withdraw and their batched twins."""
    import copy
    from bazuka_b200.mpn import batch_update as BU, dw as D
    from oracle.py.state import HostTreeHasher
    st, keys = make_state(3, 3, 2, bal=1000)
    assert st.state_size == _recount(st) == 2 * 4          # (pk.x, pk.y, token id, balance) per account; nonces are 0
    keys.append(N.eddsa_keys(b"newcomer"))
    st_b = copy.deepcopy(st)
    txs = [transfer(keys, 0, 2, 1, amount=990, fee=10),     # empties account 0's balance: -1; creates 4 leaves at the newcomer; nonce +1
           transfer(keys, 1, 0, 1, amount=5, fee=1)]
    U.update(st, txs, 1)
    assert st.state_size == _recount(st) == 8 + 4 + 1 - 1 + 1 + 1   # newcomer 4, nonce(0), balance(0) gone, nonce(1), balance(0) back
    BU.update_batched(HostTreeHasher(N.poseidon), st_b, txs, 1)
    assert st_b.state_size == st.state_size and st_b.compressed == st.compressed
    dep = [D.MpnDeposit(N.jj_compress(N.eddsa_keys(b"dep-new")[0]), 77, 9), D.MpnDeposit(N.jj_compress(keys[0][0]), 77, 4)]
    st_b = copy.deepcopy(st)
    D.deposit(st, dep, 1)
    BU.deposit_batched(HostTreeHasher(N.poseidon), st_b, dep, 1)
    assert st.state_size == _recount(st) == st_b.state_size and st.root == st_b.root
    w = D.MpnWithdraw(N.jj_compress(keys[1][0]), 1, amount=U.Money(U.ZIESHA, 100), fee=U.Money(U.ZIESHA, 2), fingerprint=4242)
    w.sign(keys[1][1])
    st_b = copy.deepcopy(st)
    D.withdraw(st, [w], 1)
    BU.withdraw_batched(HostTreeHasher(N.poseidon), st_b, [w], 1)
    assert st.state_size == _recount(st) == st_b.state_size and st.root == st_b.root


def test_deposit_rejection_follows_the_l1_source_and_withdraw_checks_calldata():
    """deposit.rs:33,68-83: a rejected deposit puts its L1 source (`payment.src`) on a list and the source's later deposits in
    the call are rejected as well (their L1 nonces would no longer line up); withdraw.rs:77 / transaction.rs:177-182: a withdrawal
    whose payment carries a calldata other than Poseidon(address, nonce, signature) is rejected.  Untracked entries (src None,
    calldata None) behave as before."""
    import copy
    from bazuka_b200.mpn import dw as D
    st, keys = make_state(3, 3, 2)
    stranger = N.eddsa_keys(b"stranger")[0]
    # account 0's token tree gets filled so that a deposit of a fifth token finds no slot: 2^(2*3) = 64 slots
    full = st.get(0).copy()
    for slot in range(1, 64):
        full.tokens[slot] = U.Money(1000 + slot, 1)
    st.set(0, full)
    mk = lambda pk, tok, amt, src: D.MpnDeposit(N.jj_compress(pk), tok, amt, src)
    deps = [mk(keys[1][0], 77, 5, "alice"),            # accepted
            mk(keys[0][0], 9999, 1, "bob"),            # no free slot -> rejected, bob listed
            mk(keys[1][0], 77, 6, "bob"),              # would be fine, but bob is listed
            mk(stranger, 77, 7, "alice"),              # alice is still fine (new account)
            mk(keys[0][0], 9998, 1, None),             # rejected, nobody listed
            mk(keys[1][0], 77, 8, None)]               # accepted
    pub, trans = D.deposit(copy.deepcopy(st), deps, 2)
    assert [t.tx.amount for t in trans] == [5, 7, 8]
    # withdrawals
    st2, keys2 = make_state(3, 3, 2)
    ws = []
    for i, cd in ((0, "good"), (1, "bad"), (1, None)):
        w = D.MpnWithdraw(N.jj_compress(keys2[i][0]), 1, amount=U.Money(U.ZIESHA, 10 + i), fee=U.Money(U.ZIESHA, 1), fingerprint=500 + i)
        w.sign(keys2[i][1])
        w.calldata = {"good": w.expected_calldata(), "bad": w.expected_calldata() + 1, None: None}[cd]
        ws.append(w)
    pub, trans = D.withdraw(st2, ws, 1)
    assert [t.tx is w for t, w in zip(trans, (ws[0], ws[2]))] == [True, True] and len(trans) == 2
