"""CPU tier: the worker protocol's bincode images (mpn/wire.py), `prepare_works` and the `MpnWork::verify` commitment
(mpn/works.py) — /root/reference/src/mpn/mod.rs:264-424, /root/reference/src/client/messages.rs:368-397."""
import hashlib
import struct

import numpy as np
import pytest

from bazuka_b200.mpn import dw as D, native as N, update as U, wire as Wr, works as Wk
from test_mpn_cpu import make_state, transfer

R = N.R


def _mont(x):
    return ((x << 256) % R).to_bytes(32, "little")


def _vk_blob(n_ic=6, fill=7):
    """a syntactically valid 878+97n-byte `Groth16VerifyingKey` image (points are opaque to the codec)"""
    return bytes([fill]) * 870 + struct.pack("<Q", n_ic) + bytes([fill + 1]) * (97 * n_ic)


def _config(num=(1, 1, 1)):
    return {"log4_tree_size": 3, "log4_token_tree_size": 3, "log4_deposit_batch_size": 1, "log4_withdraw_batch_size": 1, "log4_update_batch_size": 1,
            "mpn_contract_id": 0x1234, "mpn_num_update_batches": num[2], "mpn_num_deposit_batches": num[0], "mpn_num_withdraw_batches": num[1],
            "deposit_vk": _vk_blob(fill=1), "withdraw_vk": _vk_blob(fill=3), "update_vk": _vk_blob(fill=5)}


def test_leaf_encodings_are_bincode():
    w = Wr.Writer()
    Wr.enc_money(w, {"token_id": "ziesha", "amount": 5})
    assert bytes(w.b) == bytes.fromhex("01000000" "0500000000000000")
    w = Wr.Writer()
    Wr.enc_money(w, {"token_id": 9, "amount": 1})
    assert bytes(w.b) == bytes.fromhex("02000000") + _mont(9) + struct.pack("<Q", 1)      # Custom(scalar): raw Montgomery limbs
    w = Wr.Writer()
    Wr.enc_contract_id(w, None)
    assert bytes(w.b) == bytes(4)                                                           # Null(PhantomData): the tag only
    w = Wr.Writer()
    Wr.enc_pubkey(w, (3, True))
    assert bytes(w.b) == _mont(3) + b"\x01"                                                 # PointCompressed(ZkScalar, bool)
    w = Wr.Writer()
    Wr.enc_account(w, {"tx_nonce": 2, "withdraw_nonce": 1, "address": (4, 5), "tokens": {0: {"token_id": "ziesha", "amount": 7}}})
    assert bytes(w.b) == struct.pack("<II", 2, 1) + _mont(4) + _mont(5) + struct.pack("<QQ", 1, 0) + bytes.fromhex("01000000") + struct.pack("<Q", 7)
    w = Wr.Writer()
    Wr.enc_zkproof(w, bytes(range(256)) + bytes(131))
    assert len(w.b) == 391 and bytes(w.b[:4]) == bytes(4)                                   # /root/reference/src/zk/mod.rs:646-651
    for bad in (b"\x02", bytes.fromhex("03000000"), _mont(0)[:31]):
        with pytest.raises(ValueError):
            r = Wr.Reader(bad)
            (r.bool() if len(bad) == 1 else Wr.dec_contract_id(r) if len(bad) == 4 else r.fr())
    with pytest.raises(ValueError):
        Wr.Reader((R).to_bytes(32, "little")).fr()                                          # limbs must be reduced


def test_commitment_formula():
    """mod.rs:283-285: sha3-256 of bincode((prover, reward)) as a little-endian integer mod r"""
    addr, reward = bytes(range(32)), 123_456_789
    image = struct.pack("<Q", 32) + addr + struct.pack("<Q", reward)
    want = int.from_bytes(hashlib.sha3_256(image).digest(), "little") % R
    assert Wk.Wr.commitment(addr, reward) == want
    assert Wr.commitment(addr, reward + 1) != want and Wr.commitment(bytes(32), reward) != want


def _scenario():
    st, keys = make_state(3, 3, 3)
    newcomer = N.eddsa_keys(b"dep-new")
    keys.append(newcomer)
    deposits = [D.MpnDeposit(N.jj_compress(newcomer[0]), U.ZIESHA, 5000), D.MpnDeposit(N.jj_compress(keys[0][0]), 77, 9)]
    w = D.MpnWithdraw(N.jj_compress(keys[1][0]), 1, amount=U.Money(U.ZIESHA, 100), fee=U.Money(U.ZIESHA, 2))
    payment = {"memo": "rent", "contract_id": 0x1234, "withdraw_circuit_id": 0, "calldata": 0, "dst": bytes(range(32)),
               "amount": {"token_id": "ziesha", "amount": 100}, "fee": {"token_id": "ziesha", "amount": 2}}
    w.fingerprint = Wk.withdraw_fingerprint(payment)        # of the payment with calldata zeroed, so it can be signed first
    w.sign(keys[1][1])
    payment["calldata"] = w.expected_calldata()             # `verify_calldata`: Poseidon(address, nonce, signature)
    # the depositor's brand-new account (index account_count + 0) spends in the update batch of the same block
    updates = [transfer(keys, 3, 0, 1, amount=40, fee=1), transfer(keys, 0, 2, 1)]
    return st, keys, deposits, [w], {0: payment}, updates


def test_prepare_works_orders_batches_on_one_fork_and_round_trips():
    st, keys, deposits, withdraws, wpay, updates = _scenario()
    root0, size0 = st.root, st.state_size
    rewards = {"deposit": 11, "withdraw": 22, "update": 33}
    works, fork = Wk.prepare_works(_config(), st, deposits, withdraws, updates, rewards, height=9, withdraw_payments=wpay)
    assert st.root == root0 and st.state_size == size0 and not st.new_account_indices        # built on a fork
    assert [works[i]["data"][0] for i in range(3)] == ["deposit", "withdraw", "update"]       # mod.rs:353-414
    assert [works[i]["reward"] for i in range(3)] == [11, 22, 33]
    assert works[0]["public_inputs"]["state"] == root0
    for i in range(2):                                                                        # each batch starts where the last ended
        assert works[i + 1]["public_inputs"]["state"] == works[i]["new_root"]["state_hash"] == works[i]["public_inputs"]["next_state"]
    assert works[2]["new_root"] == {"state_hash": fork.root, "state_size": fork.state_size}
    dep, upd = works[0]["data"][1], works[2]["data"][1]
    assert [t["account_index"] for t in dep] == [3, 0]                                        # newcomer: mpn_account_count + 0
    assert [(t["src_index"], t["dst_index"]) for t in upd] == [(3, 0), (0, 2)]                # found through new_account_indices
    assert fork.new_account_indices == {keys[3][0]: 3}
    # bytes: encode -> decode -> same value -> same bytes; and the message envelopes
    for work in works.values():
        blob = Wr.work_to_bytes(work)
        back = Wr.work_from_bytes(blob)
        assert back == work and Wr.work_to_bytes(back) == blob
        with pytest.raises(ValueError):
            Wr.work_from_bytes(blob[:-1])
        with pytest.raises(ValueError):
            Wr.work_from_bytes(blob + b"\x00")
    resp = Wr.get_mpn_work_response_to_bytes(works)
    assert Wr.get_mpn_work_response_from_bytes(resp) == works
    assert struct.unpack("<Q", resp[:8])[0] == 3 and struct.unpack("<Q", resp[8:16])[0] == 0  # HashMap<usize, MpnWork>
    # the head of a work is its MpnConfig: five u8, the contract id, three usize, then the three tagged keys
    blob = Wr.work_to_bytes(works[0])
    assert blob[:5] == bytes([3, 3, 1, 1, 1]) and blob[5:9] == bytes.fromhex("02000000") and blob[9:41] == _mont(0x1234)
    assert struct.unpack("<QQQ", blob[41:65]) == (1, 1, 1) and blob[65:69] == bytes(4) and blob[69:69 + 870] == bytes([1]) * 870
    # the withdraw work carries the L1 payment and the builder-side fingerprint is derived from it
    wt = Wk.wire_to_transitions("withdraw", works[1]["data"][1])[0]
    assert wt.tx.fingerprint == withdraws[0].fingerprint and works[1]["data"][1][0]["tx"]["payment"]["memo"] == "rent"
    # wire -> builder dataclasses -> wire is the identity
    for i, kind in enumerate(["deposit", "withdraw", "update"]):
        items = works[i]["data"][1]
        pays = [t["tx"]["payment"] for t in items] if kind != "update" else None
        assert Wk.transitions_to_wire(kind, Wk.wire_to_transitions(kind, items), pays) == items
    prover, proofs = Wr.post_mpn_solution_request_from_bytes(Wr.post_mpn_solution_request(bytes(range(32)), {0: bytes(387), 2: bytes([1]) * 387}))
    assert prover == bytes(range(32)) and proofs == {0: bytes(387), 2: bytes([1]) * 387}
    assert Wr.post_mpn_solution_response_from_bytes(struct.pack("<Q", 2)) == 2 and Wr.post_mpn_worker_response_from_bytes(b"\x01") is True


def test_several_batches_of_one_kind_continue_each_other():
    """`mpn_num_update_batches` > 1: the same transaction list is offered to every batch; what a batch accepted is stale
    (nonce) for the next one, which takes the rest (mod.rs:396-414)."""
    st, keys = make_state(3, 3, 3)
    updates = [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1), transfer(keys, 2, 0, 1), transfer(keys, 0, 2, 2),
               transfer(keys, 1, 0, 2), transfer(keys, 2, 1, 2)]
    works, fork = Wk.prepare_works(_config(num=(0, 0, 2)), st, [], [], updates, {"deposit": 0, "withdraw": 0, "update": 5})
    got = [[t["tx"]["nonce"] for t in works[i]["data"][1]] for i in range(2)]
    assert len(works) == 2 and got[0] == [1, 1, 1, 2] and got[1] == [2, 2]
    seq = make_state(3, 3, 3)[0]
    U.update(seq, updates, 2)
    assert fork.root == seq.root and fork.state_size == seq.state_size


def test_entering_roots_are_recovered_from_the_wire():
    """the builders' `pre_root` bookkeeping (what the GPU witness path feeds each slot) does not travel in an `MpnWork`; it
    is recomputed from each transition's own account, Merkle proof and index — equal to the builders' values."""
    st, keys, deposits, withdraws, wpay, updates = _scenario()
    works, _ = Wk.prepare_works(_config(), st, deposits, withdraws, updates, {"deposit": 1, "withdraw": 2, "update": 3}, withdraw_payments=wpay)
    f = st.fork()
    _, dep = D.deposit(f, deposits, 1)
    _, wd = D.withdraw(f, withdraws, 1)
    _, up, _ = U.update(f, updates, 1)
    for i, (kind, want) in enumerate((("deposit", dep), ("withdraw", wd), ("update", up))):
        got = Wk.wire_to_transitions(kind, works[i]["data"][1])
        assert [t.pre_root for t in got] == [t.pre_root for t in want] and all(t.pre_root for t in got)


def test_final_delta_lists_the_changed_leaves():
    """mod.rs:17-45,416-417: what `prepare_works` hands the chain besides the works — every scalar leaf the batches changed."""
    st, keys, deposits, withdraws, wpay, updates = _scenario()
    _, fork = Wk.prepare_works(_config(), st, deposits, withdraws, updates, {"deposit": 1, "withdraw": 2, "update": 3}, withdraw_payments=wpay)
    d = Wk.final_delta(st, fork)
    # replaying the delta on the old leaves gives the new ledger's leaves
    old = _leaves(st)
    for loc, v in d.items():
        if v is None:
            old.pop(loc, None)
        else:
            old[loc] = v
    assert {k: v for k, v in old.items() if v != 0} == {k: v for k, v in _leaves(fork).items() if v != 0}
    assert fork.state_size == sum(1 for v in old.values() if v != 0)
    assert (3, 2) in d and (3, 4, 0, 1) in d          # the depositor's new account: its key and its balance
    assert all(len(loc) in (2, 4) for loc in d)
    w = Wr.Writer()
    Wk.enc_delta(w, d)
    assert struct.unpack("<Q", bytes(w.b[:8]))[0] == len(d)


def _leaves(state):
    out = {}
    for idx, acc in state.accounts.items():
        for f, v in enumerate((acc.tx_nonce, acc.withdraw_nonce, acc.address[0], acc.address[1])):
            out[(idx, f)] = v
        for slot, m in acc.tokens.items():
            out[(idx, 4, slot, 0)] = m.token_id
            out[(idx, 4, slot, 1)] = m.amount
    return out


def test_prepare_works_rejects_a_withdrawal_whose_payment_calldata_is_wrong():
    """withdraw.rs:77 through prepare_works: the calldata of the L1 payment that comes with a withdrawal is checked"""
    st, keys, deposits, withdraws, wpay, updates = _scenario()
    wpay[0]["calldata"] += 1
    works, _ = Wk.prepare_works(_config(), st, deposits, withdraws, updates, {"deposit": 1, "withdraw": 2, "update": 3}, withdraw_payments=wpay)
    kinds = {w["data"][0]: w for w in works.values()}
    assert not any(t["enabled"] for t in kinds["withdraw"]["data"][1])
    st, keys, deposits, withdraws, wpay, updates = _scenario()
    works, _ = Wk.prepare_works(_config(), st, deposits, withdraws, updates, {"deposit": 1, "withdraw": 2, "update": 3}, withdraw_payments=wpay)
    kinds = {w["data"][0]: w for w in works.values()}
    assert sum(t["enabled"] for t in kinds["withdraw"]["data"][1]) == 1
