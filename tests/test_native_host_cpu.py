"""CPU tier: libbzk's HOST sources (csrc/mpn_host.cu: ledger + the three transition builders + the witness drivers), compiled
unmodified with g++ into tests/hostshim/_mpn_shim.so, against the Python restatement of the reference
(/root/reference/src/mpn/{update,deposit,withdraw}.rs).  The GPU side they reach — batched Poseidon, the versioned tree update,
the witness interpreter launch — is replaced by host stand-ins (tests/hostshim/mpn_shim.cpp), so what is checked here is the C++
ledger logic, the row assembly and the witness layout; the `-m gpu` tier (tests/test_gpu_mpn.py) repeats the same scenarios over
the real kernels."""
import ctypes as ct

import numpy as np

from test_mpn_cpu import _batch_scenario, _off_curve_key, make_state, transfer


def _ptr(a):
    return ct.c_void_p(a.ctypes.data)


def _canon_rows(values):
    from bazuka_b200.mpn.gpu_witness import _canon_rows as f
    return f(values)


def _load(hostmpn, st, A, T):
    from bazuka_b200.mpn.ledger import NativeLedger
    led = NativeLedger(hostmpn, A, T)
    for i, a in st.accounts.items():
        led.set_account(i, a)
    assert led.root == st.root
    return led


def test_host_build_reproduces_the_reference_empty_root(hostmpn):
    """`compress_default` of the MPN state model, A=30, T=1 (/root/reference/src/node/api/get_explorer_blocks.rs:29)"""
    from bazuka_b200.mpn.ledger import NativeLedger
    led = NativeLedger(hostmpn, 30, 1)
    assert led.root == int("501a18871f186db1437e77e2c33acfa81405608cc60806399347215dbe98f714", 16)
    led.free()


def test_native_update_builder_rows_and_witness(hostmpn):
    """bzk_mpn_update_build over two consecutive batches (new account, self-transfer, same-token fee, four kinds of rejection):
    accepted set, every row of circuit inputs, the entering roots, the public values and the final state equal `update()`'s;
    bzk_mpn_update_witness over those rows writes exactly `UpdateCircuit::synthesize`'s assignment."""
    from bazuka_b200.mpn import cs as C, update as U, witness_program as W
    from bazuka_b200.mpn.gpu_witness import upload_program
    st, txs = _batch_scenario()
    led = _load(hostmpn, st, 3, 3)
    assert led.n_raw == len(W.raw_values(U.UpdateTransition.null(3, 3), 3, 3))
    prog = W.compile_update_block(3, 3)
    h_slot = upload_program(hostmpn, prog)
    for batch, B in ((txs, 2), (txs[6:], 1)):
        pub, trans, rej = U.update(st, batch, B)
        circ = U.UpdateCircuit(3, 3, B, commitment=5, height=1, transitions=trans, **pub)
        raws, ext, acc, public, n_acc = led.update_build(batch, B)
        assert n_acc == len(trans) and public == pub and led.root == st.root
        assert [t for t, a in zip(batch, acc) if a] == [t.tx for t in trans if t.enabled]
        want_raws = np.stack([_canon_rows(W.raw_values(tr, 3, 3)) for tr in circ.transitions])
        want_ext = np.stack([_canon_rows([circ.fee_token, r]) for r in W.slot_roots(circ)])
        assert (raws == want_raws).all(), np.nonzero((raws != want_raws).any(axis=2))
        assert (ext == want_ext).all()
        assert led.info()["state_size"] == st.state_size
        # the whole-batch witness driver on these rows
        epi = W.compile_update_epilogue(prog, B)
        h_epi = upload_program(hostmpn, epi)
        n = 1 << (2 * B)
        ni, na, mats, inputs, aux = circ.synthesize(C.ConstraintSystem()).to_csr()
        z_in, z_aux = np.zeros((6, 4), np.uint64), np.zeros((prog.p_aux + n * prog.n_ops + epi.n_ops, 4), np.uint64)
        assert z_aux.shape == aux.shape
        pro = _canon_rows([5, 1, pub["state"], circ.fee_token, pub["aux_data"], pub["next_state"]])
        hostmpn._check(hostmpn._l.bzk_mpn_update_witness(hostmpn._h, h_slot, h_epi, n, 3, prog.n_ops, epi.n_ops, _ptr(raws), _ptr(ext), prog.n_raw,
                                                         _ptr(pro), _ptr(z_in), _ptr(z_aux)))
        assert (z_in == inputs).all()
        bad = np.nonzero((z_aux != aux).any(axis=1))[0]
        assert len(bad) == 0, (len(bad), bad[:8])
        # a program of another shape is refused before anything is read or written
        assert hostmpn._l.bzk_mpn_update_witness(hostmpn._h, h_slot, h_epi, n, 3, prog.n_ops + 1, epi.n_ops, _ptr(raws), _ptr(ext), prog.n_raw,
                                                 _ptr(pro), _ptr(z_in), _ptr(z_aux)) == -1
        hostmpn._l.bzk_witness_program_free(hostmpn._h, h_epi)
    hostmpn._l.bzk_witness_program_free(hostmpn._h, h_slot)
    led.free()


def test_native_ledger_rules(hostmpn):
    """the rules of /root/reference/src/mpn/update.rs the C++ ledger must share with the Python restatement: keys that do not
    decompress are filtered (:31-38); a new receiver gets `mpn_account_count + |new_account_indices|` and the map threads across
    the batches of one fork (:47-70); `state_size` (:29,256-266); forks are independent (`fork_on_ram`, mod.rs:313);
    `set_account` drops zero-id token slots and releases an overwritten address."""
    from bazuka_b200.mpn import native as N, update as U
    from bazuka_b200.mpn.ledger import NativeLedger
    st, keys = make_state(3, 3, 3)
    led = _load(hostmpn, st, 3, 3)
    st.account_count = 10
    led.set_account(9, U.MpnAccount())
    assert led.info() == {"state_hash": st.root, "state_size": st.state_size, "account_count": 10, "pending_accounts": 0}
    keys += [N.eddsa_keys(b"newcomer"), N.eddsa_keys(b"second")]
    bad = transfer(keys, 0, 1, 1)
    bad.dst_pub_key = _off_curve_key()
    batch1 = [bad, transfer(keys, 0, 3, 1, amount=500), transfer(keys, 1, 4, 1, amount=7)]
    fork = led.fork()
    pub_py, trans, rej = U.update(st, batch1, 1)
    raws, ext, acc, pub, n_acc = fork.update_build(batch1, 1)
    assert acc.tolist() == [False, True, True] and n_acc == 2 and pub == pub_py and rej == [bad]
    assert [t.dst_index for t in trans] == [10, 11]
    assert fork.info() == {"state_hash": st.root, "state_size": st.state_size, "account_count": 10, "pending_accounts": 2}
    assert led.info()["state_hash"] != st.root and led.info()["pending_accounts"] == 0
    keys.append(N.eddsa_keys(b"third"))
    batch2 = [transfer(keys, 3, 0, 1, amount=50, fee=1), transfer(keys, 0, 5, 2, amount=1)]
    pub_py, trans, rej = U.update(st, batch2, 1)
    raws, ext, acc, pub, n_acc = fork.update_build(batch2, 1)
    assert acc.all() and pub == pub_py and [(t.src_index, t.dst_index) for t in trans] == [(10, 0), (0, 12)]
    other = led.fork()
    _, _, acc_o, _, n_o = other.update_build(batch2[:1], 1)
    assert n_o == 0 and not acc_o.any()
    st.commit_accounts(); fork.commit_accounts()
    assert fork.info() == {"state_hash": st.root, "state_size": st.state_size, "account_count": 13, "pending_accounts": 0}
    batch3 = [transfer(keys, 4, 3, 1, amount=2, fee=1)]
    pub_py, trans, _ = U.update(st, batch3, 1)
    _, _, acc, pub, _ = fork.update_build(batch3, 1)
    assert acc.all() and pub == pub_py and fork.info()["state_size"] == st.state_size
    z = NativeLedger(hostmpn, 3, 3)
    z.set_account(0, U.MpnAccount(0, 0, keys[0][0], {0: U.Money(U.ZIESHA, 5), 1: U.Money(0, 9)}))
    ref = U.MpnState(3, 3)
    ref.set(0, U.MpnAccount(0, 0, keys[0][0], {0: U.Money(U.ZIESHA, 5)}))
    assert z.root == ref.root and z.info()["state_size"] == ref.state_size
    z.set_account(0, U.MpnAccount(0, 0, keys[1][0], {0: U.Money(U.ZIESHA, 5)}))
    _, _, acc, _, _ = z.update_build([transfer(keys, 0, 1, 1, amount=1)], 0)
    assert not acc.any()
    for l in (led, fork, other, z):
        l.free()


def _dw_batches(kind, keys):
    from bazuka_b200.mpn import dw as D, native as N, update as U
    new1, new2 = N.eddsa_keys(b"dep-new")[0], N.eddsa_keys(b"dep-new-2")[0]
    if kind == "deposit":
        mk = lambda pk, tok, amt, src=None: D.MpnDeposit(N.jj_compress(pk), tok, amt, src)
        return [[mk(keys[0][0], U.ZIESHA, 500, "a"), mk(new1, 77, 9), mk(keys[1][0], 77, 1, "b"), mk(keys[0][0], 77, 4, "a"), mk(new2, 5, 5)],
                [mk(new1, 78, 3), D.MpnDeposit((6, False), 77, 1, "carol"), mk(new2, 5, 1, "carol"), mk(new2, 5, 2, "dave")]]

    def mk(i, amt, nonce, fee=2, sk=None, tok=U.ZIESHA):
        w = D.MpnWithdraw(N.jj_compress(keys[i][0]), nonce, amount=U.Money(tok, amt), fee=U.Money(U.ZIESHA, fee), fingerprint=1000 + amt)
        w.sign(sk or keys[i][1])
        return w
    stranger = D.MpnWithdraw(N.jj_compress(new1), 1, amount=U.Money(U.ZIESHA, 1), fee=U.Money(U.ZIESHA, 0), fingerprint=1)
    good_cd, bad_cd = mk(0, 100, 1), mk(2, 3, 1)
    good_cd.calldata, bad_cd.calldata = good_cd.expected_calldata(), bad_cd.expected_calldata() + 1
    return [[good_cd, mk(1, 5, 1, sk=keys[0][1]), mk(1, 5, 1), mk(0, 30, 2), mk(2, 7, 2), stranger, bad_cd],
            [mk(2, 10**15, 1), mk(2, 7, 1, tok=12345), mk(0, 1, 3), mk(1, 1, 2, fee=10**15), mk(1, 1, 2)]]


import pytest


@pytest.mark.parametrize("kind", ["deposit", "withdraw"])
def test_native_deposit_withdraw_builders_rows_and_witness(hostmpn, kind):
    """bzk_mpn_{deposit,withdraw}_build + bzk_mpn_dw_witness: accepted set over two consecutive batches (new account, repeated
    account, a key that does not decompress and its L1 source's later deposit, bad signature / nonce / balance / calldata /
    unknown key), phase rows, entering roots, reveal rows, public values, `state_size` and the witness equal the Python
    builder's transitions and `synthesize`'s assignment."""
    from bazuka_b200.mpn import cs as C, dw as D, dw_witness as DW
    from bazuka_b200.mpn.gpu_witness import upload_program
    A = T = 3
    B = 1
    st, keys = make_state(A, T, 3)
    led = _load(hostmpn, st, A, T)
    batches = _dw_batches(kind, keys)
    seq, build, raws_of = (D.deposit, led.deposit_build, DW.deposit_raws) if kind == "deposit" else (D.withdraw, led.withdraw_build, DW.withdraw_raws)
    progs = DW.TwoPhasePrograms(kind, A, T)
    rev = DW.compile_reveal_program(progs, B)
    hs = [upload_program(hostmpn, p) for p in (progs.prog1, progs.prog2, rev)]
    ext_src = np.array([-1 if s[0] == "state" else s[1] for s in progs.ext_src], dtype=np.int32)
    circ_cls = D.DepositCircuit if kind == "deposit" else D.WithdrawCircuit
    n = 1 << (2 * B)
    for items in batches:
        pub, trans = seq(st, items, B)
        rows = build(items, B)
        assert rows["public"] == pub and rows["n_accepted"] == len(trans) and led.root == st.root
        assert [it for it, a in zip(items, rows["accepted"]) if a] == [t.tx for t in trans]
        assert 0 < len(trans) < len(items)
        circ = circ_cls(A, T, B, commitment=3, height=1, transitions=trans, **pub)
        want = [raws_of(t, A, T) for t in circ.transitions]
        assert (rows["raws1"].reshape(-1, 4) == _canon_rows([v for a, _ in want for v in a])).all()
        assert (rows["raws2"].reshape(-1, 4) == _canon_rows([v for _, b in want for v in b])).all()
        assert (rows["roots"] == _canon_rows(DW.slot_roots(circ))).all()
        assert (rows["reveal"].reshape(-1, 4) == _canon_rows([v for r in DW.reveal_rows_native(kind, circ) for v in r])).all()
        assert led.info()["state_size"] == st.state_size
        ni, na, mats, inputs, aux = circ.synthesize(C.ConstraintSystem()).to_csr()
        z_in, z_aux = np.zeros((6, 4), np.uint64), np.zeros((5 + n * (progs.n1 + progs.n2) + rev.n_ops, 4), np.uint64)
        assert z_aux.shape == aux.shape
        head = _canon_rows([3, 1, pub["state"], pub["aux_data"], pub["next_state"]])
        hostmpn._check(hostmpn._l.bzk_mpn_dw_witness(hostmpn._h, hs[0], hs[1], hs[2], n, _ptr(rows["raws1"]), _ptr(rows["raws2"]), _ptr(rows["roots"]),
                                                     _ptr(ext_src), len(ext_src), _ptr(rows["reveal"]), _ptr(head), _ptr(z_in), _ptr(z_aux)))
        assert (z_in == inputs).all()
        bad = np.nonzero((z_aux != aux).any(axis=1))[0]
        assert len(bad) == 0, (len(bad), bad[:8])
    for h in hs:
        hostmpn._l.bzk_witness_program_free(hostmpn._h, h)
    led.free()


def test_native_update_builder_at_the_production_tree_shape(hostmpn):
    """A = 15, T = 3 (/root/reference/src/config/blockchain.rs:22-26), accounts spread over the 2^30 leaves (indices above 2^16 and
    2^29 - 1, so the newcomer lands in the upper half of the tree), one batch of four slots: rows, entering roots and public values equal `update()`'s, and the rows
    recovered from the WIRE image of the resulting work (csrc/mpn_wire.cu) equal the builder's own."""
    from bazuka_b200.mpn import native as N, update as U, wire as Wr, witness_program as W, works as Wk
    from bazuka_b200._lib import PARAMS_PATH
    from test_wire_cpu import _config
    A, T, B = 15, 3, 1
    st, keys = U.MpnState(A, T), []
    spots = [0, 70_000, (1 << 29) - 1]
    for i, idx in enumerate(spots):
        pk, sk = N.eddsa_keys(b"acct%d" % i)
        keys.append((pk, sk))
        st.set(idx, U.MpnAccount(0, 0, pk, {0: U.Money(U.ZIESHA, 10 ** 9), 63: U.Money(77, 5)}))
    st.account_count = 1 << 29
    keys.append(N.eddsa_keys(b"newcomer"))
    txs = [transfer(keys, 0, 1, 1), transfer(keys, 2, 0, 1, amount=9, fee=1), transfer(keys, 1, 3, 1, amount=3), transfer(keys, 0, 2, 7)]
    led = _load(hostmpn, st, A, T)
    assert led.info()["account_count"] == 1 << 29
    pub, trans, rej = U.update(st, txs, B)
    assert len(trans) == 3 and rej == [txs[3]] and trans[2].dst_index == 1 << 29
    raws, ext, acc, public, n_acc = led.update_build(txs, B)
    assert acc.tolist() == [True, True, True, False] and public == pub and led.root == st.root
    circ = U.UpdateCircuit(A, T, B, commitment=5, height=1, transitions=trans, **pub)
    want_raws = np.stack([_canon_rows(W.raw_values(tr, A, T)) for tr in circ.transitions])
    want_ext = np.stack([_canon_rows([circ.fee_token, r]) for r in W.slot_roots(circ)])
    assert (raws == want_raws).all() and (ext == want_ext).all()
    # the same rows from the wire image of the work
    cfg = dict(_config(), log4_tree_size=A, log4_token_tree_size=T, log4_update_batch_size=B)
    work = {"config": cfg, "public_inputs": dict(pub, height=1), "data": ("update", Wk.transitions_to_wire("update", trans)),
            "new_root": {"state_hash": st.root, "state_size": st.state_size}, "reward": 1}
    blob = Wr.work_to_bytes(work)
    lib = hostmpn._l
    h, hasher = ct.c_void_p(), ct.c_void_p()
    assert lib.bzk_mpn_work_decode(blob, len(blob), ct.byref(h), None) == 0
    pb = open(PARAMS_PATH, "rb").read()
    assert lib.bzk_poseidon_host_create(pb, len(pb), ct.byref(hasher)) == 0
    raws2, ext2 = np.zeros_like(raws), np.zeros_like(ext)
    canon = lambda v: np.frombuffer((v % N.R).to_bytes(32, "little"), dtype=np.uint64).copy()
    jj_d, fee = canon(N.JJ_D), canon(U.ZIESHA)
    assert lib.bzk_mpn_work_update_rows(h, hasher, _ptr(jj_d), _ptr(fee), _ptr(raws2), _ptr(ext2)) == 0
    assert (raws2 == raws).all() and (ext2 == ext).all()
    lib.bzk_mpn_work_free(h); lib.bzk_poseidon_host_free(hasher)
    led.free()


def test_reference_withdraw_scenario_at_the_production_config(hostmpn):
    """the reference's own transition-builder test (/root/reference/src/mpn/withdraw.rs:270-353) on the native ledger: production
    config (A = 15, T = 3, deposit / withdraw batches of 64), a fresh state, `TxBuilder::new("ABC")` deposits 10056 of token 123 into
    its own MPN account (a new account: index 0), then withdraws 30 with a fee of 26 in the same token — one accepted transition
    each, as the reference asserts; an empty withdraw batch builds too (`test_withdraw_empty`).  Rows, roots and public values
    equal the Python restatement's."""
    from bazuka_b200.mpn import dw as D, dw_witness as DW, native as N, update as U
    from bazuka_b200.mpn.ledger import NativeLedger
    A, T, B = 15, 3, 3
    st = U.MpnState(A, T)
    led = NativeLedger(hostmpn, A, T)
    assert led.root == st.root
    # test_withdraw_empty
    pub0, tr0 = D.withdraw(st, [], B)
    rows0 = led.withdraw_build([], B)
    assert rows0["n_accepted"] == 0 == len(tr0) and rows0["public"] == pub0 and pub0["state"] == pub0["next_state"] == st.root
    pk, sk = N.eddsa_keys(b"ABC")
    dep = D.MpnDeposit(N.jj_compress(pk), 123, 10056, "abc-l1")
    pub, trans = D.deposit(st, [dep], B)
    rows = led.deposit_build([dep], B)
    assert len(trans) == 1 == rows["n_accepted"] and rows["public"] == pub and led.root == st.root and trans[0].account_index == 0
    circ = D.DepositCircuit(A, T, B, commitment=0, height=0, transitions=trans, **pub)
    want = [DW.deposit_raws(t, A, T) for t in circ.transitions]
    assert len(want) == 64
    assert (rows["raws1"].reshape(-1, 4) == _canon_rows([v for a, _ in want for v in a])).all()
    assert (rows["raws2"].reshape(-1, 4) == _canon_rows([v for _, b in want for v in b])).all()
    assert (rows["roots"] == _canon_rows(DW.slot_roots(circ))).all()
    w = D.MpnWithdraw(N.jj_compress(pk), 1, amount=U.Money(123, 30), fee=U.Money(123, 26), fingerprint=4242)
    w.sign(sk)
    pub, trans = D.withdraw(st, [w], B)
    rows = led.withdraw_build([w], B)
    assert len(trans) == 1 == rows["n_accepted"] and rows["public"] == pub and led.root == st.root
    assert st.accounts[0].tokens[0].amount == 10056 - 30 - 26 and st.accounts[0].withdraw_nonce == 1
    circ = D.WithdrawCircuit(A, T, B, commitment=0, height=0, transitions=trans, **pub)
    want = [DW.withdraw_raws(t, A, T) for t in circ.transitions]
    assert (rows["raws1"].reshape(-1, 4) == _canon_rows([v for a, _ in want for v in a])).all()
    assert (rows["raws2"].reshape(-1, 4) == _canon_rows([v for _, b in want for v in b])).all()
    assert (rows["reveal"].reshape(-1, 4) == _canon_rows([v for r in DW.reveal_rows_native("withdraw", circ) for v in r])).all()
    assert led.info()["state_size"] == st.state_size
    led.free()
