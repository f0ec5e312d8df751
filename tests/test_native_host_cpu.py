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


def test_randomised_differential_of_the_update_ledger_rules(hostmpn):
    """random transfer traffic on a SMALL state (A = 2: sixteen accounts, T = 1: four token slots per account) so that every rule
    fires many times — unknown and undecompressible keys, wrong nonces, overdrafts, fee tokens that are not the accepted one or not
    held, transfers to oneself, new receivers until the account tree is full, token trees that fill up, tokens whose slot was
    chosen by the first free index: over consecutive batches on one ledger the C++ builder and the Python restatement of
    `update()` must accept the same transactions, produce the same rows / roots / public values and end in the same state."""
    import random
    from bazuka_b200.mpn import native as N, update as U, witness_program as W
    from bazuka_b200.mpn.ledger import NativeLedger
    A, T, B = 2, 1, 1
    rng = random.Random(77)
    keys = [N.eddsa_keys(b"rk%d" % i) for i in range(24)]
    tokens = [U.ZIESHA, 5, 6, 7, 8, 9]
    st = U.MpnState(A, T)
    for i in range(5):
        st.set(i, U.MpnAccount(0, 0, keys[i][0], {0: U.Money(U.ZIESHA, 10 ** 6), 1: U.Money(tokens[1 + i % 3], 500)}))
    led = NativeLedger(hostmpn, A, T)
    for i, a in st.accounts.items():
        led.set_account(i, a)
    assert led.root == st.root
    known = list(range(5))                          # key numbers that own an account
    nonces = {i: 0 for i in known}
    total_acc = total = 0
    for batch_no in range(40):
        txs = []
        for _ in range(rng.randint(1, 7)):
            s = rng.choice(known) if rng.random() < 0.85 else rng.randrange(len(keys))
            d = rng.randrange(len(keys)) if rng.random() < 0.4 else rng.choice(known)
            tok = rng.choice(tokens) if rng.random() < 0.5 else U.ZIESHA
            amount = rng.choice([0, 1, 3, 50, 499, 10 ** 5, 10 ** 7])
            fee_tok = U.ZIESHA if rng.random() < 0.9 else rng.choice(tokens)
            nonce = nonces.get(s, 0) + 1 + (1 if rng.random() < 0.1 else 0)
            tx = U.MpnTransaction(nonce, N.jj_compress(keys[s][0]), N.jj_compress(keys[d][0]), U.Money(tok, amount), U.Money(fee_tok, rng.choice([0, 1, 7])))
            tx.sign(keys[s][1])
            if rng.random() < 0.05:
                tx.dst_pub_key = _off_curve_key()        # (after signing: the message hash decompresses the key)
            txs.append(tx)
        pub, trans, rej = U.update(st, txs, B)
        raws, ext, acc, public, n_acc = led.update_build(txs, B)
        got = [t for t, a in zip(txs, acc) if a]
        assert got == [t.tx for t in trans], (batch_no, [txs.index(t) for t in got], [txs.index(t.tx) for t in trans])
        assert public == pub and led.root == st.root and led.info()["state_size"] == st.state_size, batch_no
        circ = U.UpdateCircuit(A, T, B, commitment=1, height=1, transitions=trans, **pub)
        want = np.stack([_canon_rows(W.raw_values(tr, A, T)) for tr in circ.transitions])
        assert (raws == want).all(), batch_no
        assert (ext == np.stack([_canon_rows([circ.fee_token, r]) for r in W.slot_roots(circ)])).all(), batch_no
        for t in trans:                               # bookkeeping for the generator
            s = next(i for i, k in enumerate(keys) if N.jj_compress(k[0]) == t.tx.src_pub_key)
            nonces[s] = t.tx.nonce
            d = next((i for i, k in enumerate(keys) if N.jj_compress(k[0]) == t.tx.dst_pub_key), None)
            if d is not None and d not in known:
                known.append(d)
                nonces.setdefault(d, 0)
        total += len(txs)
        total_acc += len(trans)
        if batch_no % 10 == 9:                        # a block boundary: the new accounts enter the chain's index table
            st.commit_accounts(); led.commit_accounts()
    assert total_acc > 30 and total - total_acc > 30, (total, total_acc)
    assert len(st.accounts) >= 12                     # the tree filled up with newcomers
    led.free()


def test_randomised_differential_of_the_deposit_and_withdraw_rules(hostmpn):
    """the same for `deposit()` / `withdraw()`: random deposits (new accounts until the tree is full, tokens until an account's four
    slots are full, keys that do not decompress, L1 sources whose earlier deposit was rejected) alternating with random withdrawals
    (wrong nonce, wrong signer, unknown key, overdraft of the token or of the fee, fee in a token the account does not hold, wrong
    calldata) on one small ledger: same accepted sets, rows, roots, revealed rows, public values and `state_size` throughout."""
    import random
    from bazuka_b200.mpn import dw as D, dw_witness as DW, native as N, update as U
    from bazuka_b200.mpn.ledger import NativeLedger
    A, T, B = 2, 1, 1
    rng = random.Random(4242)
    keys = [N.eddsa_keys(b"dk%d" % i) for i in range(22)]
    tokens = [U.ZIESHA, 5, 6, 7, 8, 9]
    st = U.MpnState(A, T)
    led = NativeLedger(hostmpn, A, T)
    wn = {}                                          # key number -> withdraw nonce
    n_dep = n_wd = n_dep_acc = n_wd_acc = 0
    for round_no in range(30):
        deps = []
        for _ in range(rng.randint(1, 6)):
            k = rng.randrange(len(keys))
            addr = (6, False) if rng.random() < 0.08 else N.jj_compress(keys[k][0])
            deps.append(D.MpnDeposit(addr, rng.choice(tokens), rng.choice([1, 40, 1000]), rng.choice(["a", "b", "c", None])))
        pub, trans = D.deposit(st, deps, B)
        rows = led.deposit_build(deps, B)
        assert [d for d, a in zip(deps, rows["accepted"]) if a] == [t.tx for t in trans], round_no
        assert rows["public"] == pub and led.root == st.root and led.info()["state_size"] == st.state_size, round_no
        circ = D.DepositCircuit(A, T, B, commitment=0, height=0, transitions=trans, **pub)
        want = [DW.deposit_raws(t, A, T) for t in circ.transitions]
        assert (rows["raws1"].reshape(-1, 4) == _canon_rows([v for a, _ in want for v in a])).all(), round_no
        assert (rows["raws2"].reshape(-1, 4) == _canon_rows([v for _, b in want for v in b])).all(), round_no
        assert (rows["roots"] == _canon_rows(DW.slot_roots(circ))).all(), round_no
        assert (rows["reveal"].reshape(-1, 4) == _canon_rows([v for r in DW.reveal_rows_native("deposit", circ) for v in r])).all(), round_no
        n_dep += len(deps); n_dep_acc += len(trans)
        owners = {tuple(a.address): i for i, a in st.accounts.items()}
        wds = []
        for _ in range(rng.randint(1, 6)):
            k = rng.randrange(len(keys))
            acc = st.accounts.get(owners.get(tuple(keys[k][0])))
            held = [m.token_id for m in acc.tokens.values()] if acc else tokens
            tok = rng.choice(held) if rng.random() < 0.8 else rng.choice(tokens)
            ftok = rng.choice(held) if rng.random() < 0.8 else rng.choice(tokens)
            nonce = wn.get(k, 0) + 1 + (1 if rng.random() < 0.1 else 0)
            w = D.MpnWithdraw(N.jj_compress(keys[k][0]), nonce, amount=U.Money(tok, rng.choice([0, 1, 30, 5000])), fee=U.Money(ftok, rng.choice([0, 1, 2000])),
                              fingerprint=rng.randrange(1, 10 ** 9))
            w.sign(keys[rng.randrange(len(keys))][1] if rng.random() < 0.1 else keys[k][1])
            if rng.random() < 0.3:
                w.calldata = w.expected_calldata() + (1 if rng.random() < 0.3 else 0)
            wds.append(w)
        pub, trans = D.withdraw(st, wds, B)
        rows = led.withdraw_build(wds, B)
        assert [x for x, a in zip(wds, rows["accepted"]) if a] == [t.tx for t in trans], round_no
        assert rows["public"] == pub and led.root == st.root and led.info()["state_size"] == st.state_size, round_no
        circ = D.WithdrawCircuit(A, T, B, commitment=0, height=0, transitions=trans, **pub)
        want = [DW.withdraw_raws(t, A, T) for t in circ.transitions]
        assert (rows["raws1"].reshape(-1, 4) == _canon_rows([v for a, _ in want for v in a])).all(), round_no
        assert (rows["raws2"].reshape(-1, 4) == _canon_rows([v for _, b in want for v in b])).all(), round_no
        assert (rows["roots"] == _canon_rows(DW.slot_roots(circ))).all(), round_no
        assert (rows["reveal"].reshape(-1, 4) == _canon_rows([v for r in DW.reveal_rows_native("withdraw", circ) for v in r])).all(), round_no
        for t in trans:
            k = next(i for i, kk in enumerate(keys) if N.jj_compress(kk[0]) == t.tx.mpn_address)
            wn[k] = t.tx.mpn_withdraw_nonce
        n_wd += len(wds); n_wd_acc += len(trans)
        if round_no % 7 == 6:
            st.commit_accounts(); led.commit_accounts()
    assert n_dep_acc > 25 and n_dep - n_dep_acc > 10 and n_wd_acc > 15 and n_wd - n_wd_acc > 25, (n_dep, n_dep_acc, n_wd, n_wd_acc)
    led.free()
