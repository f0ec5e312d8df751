"""CPU tier: libbzk's native worker protocol (csrc/mpn_wire.cu, host code compiled unmodified into tests/hostshim/_mpn_shim.so)
against the Python restatement (bazuka_b200/mpn/wire.py, works.py): sha3-256, the commitment, bincode of MpnWork and the
messages byte for byte, and the rows a work's transitions feed the witness drivers."""
import ctypes as ct
import hashlib
import struct

import numpy as np
import pytest

from bazuka_b200.mpn import dw as D, dw_witness as DW, native as N, update as U, wire as Wr, witness_program as W, works as Wk
from test_wire_cpu import _config, _scenario

R = N.R


def _ptr(a):
    return ct.c_void_p(a.ctypes.data)


def _canon(v):
    return np.frombuffer((v % R).to_bytes(32, "little"), dtype=np.uint64).copy()


def _canon_rows(values):
    from bazuka_b200.mpn.gpu_witness import _canon_rows as f
    return f(values)


def _int(a):
    return int.from_bytes(np.ascontiguousarray(a).tobytes(), "little")


class _Work:
    def __init__(self, lib, blob):
        self.lib, self.h = lib, ct.c_void_p()
        st = lib.bzk_mpn_work_decode(blob, len(blob), ct.byref(self.h), None)
        if st != 0:
            raise ValueError(st)

    def encode(self):
        n = ct.c_size_t()
        assert self.lib.bzk_mpn_work_encode(self.h, None, 0, ct.byref(n)) == 0
        buf = ct.create_string_buffer(n.value)
        assert self.lib.bzk_mpn_work_encode(self.h, buf, n.value, ct.byref(n)) == 0
        return buf.raw

    def info(self):
        dt = np.dtype([("kind", "<u4"), ("A", "<u4"), ("T", "<u4"), ("B", "<u4"), ("n", "<u8"), ("height", "<u8"), ("state", "<u8", 4), ("aux", "<u8", 4),
                       ("next", "<u8", 4), ("root", "<u8", 4), ("size", "<u8"), ("reward", "<u8")])
        out = np.zeros(1, dtype=dt)
        assert self.lib.bzk_mpn_work_get_info(self.h, _ptr(out)) == 0
        return out[0]

    def free(self):
        self.lib.bzk_mpn_work_free(self.h)


@pytest.fixture(scope="module")
def works():
    st, keys, deposits, withdraws, wpay, updates = _scenario()
    w, _ = Wk.prepare_works(_config(), st, deposits, withdraws, updates, {"deposit": 11, "withdraw": 22, "update": 33}, height=9, withdraw_payments=wpay)
    return w


@pytest.fixture(scope="module")
def host_hasher(hostmpn):
    from bazuka_b200 import _lib
    blob = open(_lib.PARAMS_PATH, "rb").read()
    h = ct.c_void_p()
    assert hostmpn._l.bzk_poseidon_host_create(blob, len(blob), ct.byref(h)) == 0
    return h


def test_sha3_and_commitment(hostmpn):
    lib = hostmpn._l
    for n in (0, 1, 31, 135, 136, 137, 271, 272, 273, 1000):
        data = bytes((7 * i + n) & 0xFF for i in range(n))
        out = ct.create_string_buffer(32)
        assert lib.bzk_sha3_256(data, n, out) == 0
        assert out.raw == hashlib.sha3_256(data).digest(), n
    for addr, reward in ((bytes(range(32)), 123_456_789), (bytes(32), 0), (bytes([255]) * 32, 2**64 - 1)):
        out = np.zeros(4, np.uint64)
        assert lib.bzk_mpn_commitment(addr, reward, _ptr(out)) == 0
        assert _int(out) == Wr.commitment(addr, reward)


def test_work_bincode_round_trips_byte_for_byte(hostmpn, works):
    lib = hostmpn._l
    for i, work in works.items():
        blob = Wr.work_to_bytes(work)
        w = _Work(lib, blob)
        assert w.encode() == blob
        info = w.info()
        kind = ["deposit", "withdraw", "update"][info["kind"]]
        assert kind == work["data"][0] and (info["A"], info["T"], info["B"]) == (3, 3, 1) and info["n"] == len(work["data"][1])
        p = work["public_inputs"]
        assert (info["height"], _int(info["state"]), _int(info["aux"]), _int(info["next"])) == (p["height"], p["state"], p["aux_data"], p["next_state"])
        assert (_int(info["root"]), info["size"], info["reward"]) == (work["new_root"]["state_hash"], work["new_root"]["state_size"], work["reward"])
        vk, n = ct.c_void_p(), ct.c_size_t()
        assert lib.bzk_mpn_work_vk(w.h, ct.byref(vk), ct.byref(n)) == 0
        assert ct.string_at(vk, n.value) == work["config"][kind + "_vk"]
        pub = np.zeros((5, 4), np.uint64)
        addr = bytes(range(32))
        assert lib.bzk_mpn_work_public_inputs(w.h, addr, _ptr(pub)) == 0
        from bazuka_b200.mpn.cs import to_mont
        assert (pub == to_mont(Wk.work_public_inputs(work, addr))).all()
        w.free()
        # malformed images are refused, not mis-read
        for bad in (blob[:-1], blob + b"\x00", blob[:5] + b"\x07" + blob[6:]):
            with pytest.raises(ValueError):
                _Work(lib, bad)
    blob = Wr.work_to_bytes(works[2])
    off = blob.index(((works[2]["public_inputs"]["state"] << 256) % R).to_bytes(32, "little"))
    with pytest.raises(ValueError):                                    # scalar limbs must be reduced
        _Work(lib, blob[:off] + R.to_bytes(32, "little") + blob[off + 32:])


def test_message_envelopes(hostmpn, works):
    lib = hostmpn._l
    resp = Wr.get_mpn_work_response_to_bytes(works)
    ids, hs, n = np.zeros(8, np.uint64), (ct.c_void_p * 8)(), ct.c_uint64()
    assert lib.bzk_mpn_get_work_response_decode(resp, len(resp), _ptr(ids), hs, 8, ct.byref(n)) == 0
    assert n.value == 3 and ids[:3].tolist() == [0, 1, 2]
    for k in range(3):
        ln = ct.c_size_t()
        assert lib.bzk_mpn_work_encode(hs[k], None, 0, ct.byref(ln)) == 0 and ln.value == len(Wr.work_to_bytes(works[k]))
        lib.bzk_mpn_work_free(hs[k])
    assert lib.bzk_mpn_get_work_response_decode(resp[:-3], len(resp) - 3, _ptr(ids), hs, 8, ct.byref(n)) == -1
    addr = bytes(range(32))
    out = ct.create_string_buffer(40)
    assert lib.bzk_mpn_get_work_request_encode(addr, out) == 0 and out.raw == Wr.get_mpn_work_request(addr)
    proofs = {0: bytes(387), 2: bytes([1]) * 387}
    want = Wr.post_mpn_solution_request(addr, proofs)
    ln = ct.c_size_t()
    pid, blob = np.array([0, 2], np.uint64), b"".join(proofs.values())
    assert lib.bzk_mpn_post_solution_request_encode(addr, _ptr(pid), blob, 2, None, 0, ct.byref(ln)) == 0 and ln.value == len(want)
    buf = ct.create_string_buffer(ln.value)
    assert lib.bzk_mpn_post_solution_request_encode(addr, _ptr(pid), blob, 2, buf, ln.value, ct.byref(ln)) == 0 and buf.raw == want
    acc = ct.c_uint64()
    assert lib.bzk_mpn_post_solution_response_decode(struct.pack("<Q", 2), 8, ct.byref(acc)) == 0 and acc.value == 2


def test_rows_of_a_work_equal_the_python_prover(hostmpn, works, host_hasher):
    """what the external prover feeds the witness programs, derived natively from the WIRE image of each work: equal to
    works.MpnProver's path (wire -> builder dataclasses -> raw_values / deposit_raws / withdraw_raws, entering roots recomputed
    from each transition's own proof, revealed rows with their calldata hash, the withdraw fingerprint from the payment)."""
    lib = hostmpn._l
    A = T = 3
    jj_d, fee = _canon(N.JJ_D), _canon(U.ZIESHA)
    for i, work in works.items():
        kind, items = work["data"]
        trans = Wk.wire_to_transitions(kind, items)
        p = work["public_inputs"]
        w = _Work(lib, Wr.work_to_bytes(work))
        n = 4                                            # 4^B slots: the work carries 2 transitions, the prover pads
        if kind == "update":
            circ = U.UpdateCircuit(A, T, 1, fee_token=U.ZIESHA, commitment=1, height=p["height"], state=p["state"], aux_data=p["aux_data"],
                                   next_state=p["next_state"], transitions=trans)
            n_raw = 32 + 9 * T + 6 * A
            raws, ext = np.zeros((n, n_raw, 4), np.uint64), np.zeros((n, 2, 4), np.uint64)
            assert lib.bzk_mpn_work_update_rows(w.h, host_hasher, _ptr(jj_d), _ptr(fee), _ptr(raws), _ptr(ext)) == 0
            want_raws = np.stack([_canon_rows(W.raw_values(tr, A, T)) for tr in circ.transitions])
            want_ext = np.stack([_canon_rows([circ.fee_token, r]) for r in W.slot_roots(circ)])
            assert (raws == want_raws).all(), np.nonzero((raws != want_raws).any(axis=2))
            assert (ext == want_ext).all()
            # a deposit-shaped call on an update work is refused
            assert lib.bzk_mpn_work_dw_rows(w.h, host_hasher, _ptr(jj_d), _ptr(raws), _ptr(raws), _ptr(raws), _ptr(raws)) == -1
        else:
            cls, raws_of, w1, w2, wr = ((D.DepositCircuit, DW.deposit_raws, 5, 9 + 3 * T + 3 * A, 4) if kind == "deposit" else
                                        (D.WithdrawCircuit, DW.withdraw_raws, 12, 12 + 6 * T + 3 * A, 7))
            circ = cls(A, T, 1, commitment=1, height=p["height"], state=p["state"], aux_data=p["aux_data"], next_state=p["next_state"], transitions=trans)
            r1, r2 = np.zeros((n, w1, 4), np.uint64), np.zeros((n, w2, 4), np.uint64)
            roots, rev = np.zeros((n, 4), np.uint64), np.zeros((n, wr, 4), np.uint64)
            assert lib.bzk_mpn_work_dw_rows(w.h, host_hasher, _ptr(jj_d), _ptr(r1), _ptr(r2), _ptr(roots), _ptr(rev)) == 0
            want = [raws_of(t, A, T) for t in circ.transitions]
            assert (r1.reshape(-1, 4) == _canon_rows([v for a, _ in want for v in a])).all()
            assert (r2.reshape(-1, 4) == _canon_rows([v for _, b in want for v in b])).all()
            assert (roots == _canon_rows(DW.slot_roots(circ))).all()
            assert (rev.reshape(-1, 4) == _canon_rows([v for r in DW.reveal_rows_native(kind, circ) for v in r])).all()
            assert 0 < len(items) < n and all(t["enabled"] for t in items)                 # real slots and padded ones
        w.free()


def _compile(lib, kind, A, T, B):
    from bazuka_b200 import _lib
    blob = open(_lib.PARAMS_PATH, "rb").read()
    jj = np.ascontiguousarray(np.stack([_canon(N.JJ_D), _canon(N.JJ_BASE_COFACTOR[0]), _canon(N.JJ_BASE_COFACTOR[1])]))
    h = ct.c_void_p()
    if kind == "update":
        st = lib.bzk_mpn_update_circuit_compile(A, T, B, blob, len(blob), _ptr(jj), ct.byref(h))
    else:
        st = lib.bzk_mpn_dw_circuit_compile({"deposit": 1, "withdraw": 2}[kind], A, T, B, blob, len(blob), _ptr(jj), ct.byref(h))
    assert st == 0
    return h, blob


def test_work_bytes_to_satisfying_witness_through_the_native_prover_object(hostmpn, works):
    """bzk_mpn_prover_create / bzk_mpn_prover_prove_work (csrc/mpn_prover.cu) on the host build: the circuit compiled by C++, its
    programs and R1CS "uploaded", then for the wire image of each work: rows -> witness drivers -> resident z -> the prove call,
    which in this tier (tests/hostshim/mpn_shim.cpp) checks a(z) * b(z) = c(z) on every constraint of the compiled circuit.  So a
    work that came off the wire yields a satisfying assignment; a tampered public input does not; a work of another kind is
    refused.  (The MSM / NTT half of the prove call is the GPU tier's.)"""
    lib = hostmpn._l
    lib.shim_last_unsat_row.restype = ct.c_uint64
    addr = bytes(range(32))
    rs = np.zeros((2, 4), np.uint64)
    provers = {}
    for kind in ("deposit", "withdraw", "update"):
        c, blob = _compile(lib, kind, 3, 3, 1)
        k4 = np.zeros(4, np.uint32)
        assert lib.bzk_mpn_circuit_kind(c, _ptr(k4)) == 0 and k4.tolist() == [{"update": 0, "deposit": 1, "withdraw": 2}[kind], 3, 3, 1]
        p = ct.c_void_p()
        jj_d, fee = _canon(N.JJ_D), _canon(U.ZIESHA)
        hostmpn._check(lib.bzk_mpn_prover_create(hostmpn._h, c, ct.c_void_p(1), _ptr(jj_d), _ptr(fee), ct.byref(p)))
        lib.bzk_mpn_circuit_free(c)                       # the prover keeps its own copies
        provers[kind] = p
    for i, work in works.items():
        kind = work["data"][0]
        blob = Wr.work_to_bytes(work)
        out = ct.create_string_buffer(391)
        st = lib.bzk_mpn_prover_prove_work(hostmpn._h, provers[kind], blob, len(blob), addr, _ptr(rs[0]), _ptr(rs[1]), 1, out)
        assert st == 0, (kind, st, lib.shim_last_unsat_row())
        assert out.raw[:4] == bytes(4) and len(out.raw) == 391
        # the claimed end state is a public input the circuit ties to the transitions
        bad = dict(work, public_inputs=dict(work["public_inputs"], next_state=work["public_inputs"]["next_state"] + 1))
        bb = Wr.work_to_bytes(bad)
        assert lib.bzk_mpn_prover_prove_work(hostmpn._h, provers[kind], bb, len(bb), addr, _ptr(rs[0]), _ptr(rs[1]), 1, out) == -7
        other = provers["update" if kind != "update" else "deposit"]
        assert lib.bzk_mpn_prover_prove_work(hostmpn._h, other, blob, len(blob), addr, _ptr(rs[0]), _ptr(rs[1]), 1, out) == -1
        assert lib.bzk_mpn_prover_prove_work(hostmpn._h, provers[kind], blob[:-2], len(blob) - 2, addr, _ptr(rs[0]), _ptr(rs[1]), 1, out) == -1
    for p in provers.values():
        lib.bzk_mpn_prover_free(hostmpn._h, p)


def _vec(items, enc):
    w = Wr.Writer()
    w.vec(items, enc)
    return bytes(w.b)


def test_native_prepare_works_equals_the_python_restatement_byte_for_byte(hostmpn):
    prepare_works_case(hostmpn)


def prepare_works_case(hostmpn):
    """(shared with the GPU tier: `hostmpn` is any context-like object)  bzk_mpn_prepare_works (csrc/mpn_wire.cu over the three native builders) against works.prepare_works on the same block:
    deposit -> withdraw -> update on one fork, the depositor's brand-new account spending in the update batch of the same block
    (`new_account_indices`), a rejected withdrawal (wrong calldata) and a deposit rejected through its L1 source; the
    GetMpnWorkResponse images are equal byte for byte, the forks agree, the original ledger did not move — and two update batches
    continue each other."""
    from test_native_host_cpu import _load
    from test_mpn_cpu import make_state, transfer
    lib = hostmpn._l
    st, keys, deposits, withdraws, wpay, updates = _scenario()
    # explicit L1 payments for every deposit; the third comes from the source of a deposit whose key does not decompress
    deposits = deposits + [D.MpnDeposit((6, False), 77, 1), D.MpnDeposit(N.jj_compress(keys[1][0]), 77, 3)]
    srcs = [bytes([1]) * 32, bytes([2]) * 32, bytes([3]) * 32, bytes([3]) * 32]
    dpay = {k: {"memo": "d%d" % k, "contract_id": 0x1234, "deposit_circuit_id": 0, "calldata": 0, "src": srcs[k],
                "amount": {"token_id": Wr.scalar_contract_id(d.token_id), "amount": d.amount}, "fee": {"token_id": "ziesha", "amount": 0}, "nonce": k + 1,
                "sig": bytes([9]) * 64 if k == 1 else None} for k, d in enumerate(deposits)}
    # a second withdrawal whose payment carries the wrong calldata
    w2 = D.MpnWithdraw(N.jj_compress(keys[2][0]), 1, amount=U.Money(U.ZIESHA, 7), fee=U.Money(U.ZIESHA, 1))
    pay2 = {"memo": "", "contract_id": None, "withdraw_circuit_id": 0, "calldata": 0, "dst": bytes(32), "amount": {"token_id": "ziesha", "amount": 7},
            "fee": {"token_id": "ziesha", "amount": 1}}
    w2.fingerprint = Wk.withdraw_fingerprint(pay2)
    w2.sign(keys[2][1])
    pay2["calldata"] = w2.expected_calldata() + 1
    withdraws, wpay = withdraws + [w2], {**wpay, 1: pay2}
    config = _config()
    rewards = {"deposit": 11, "withdraw": 22, "update": 33}
    led = _load(hostmpn, st, 3, 3)
    root0 = led.root
    want, fork_py = Wk.prepare_works(config, st, deposits, withdraws, updates, rewards, height=9, deposit_payments=dpay, withdraw_payments=wpay)
    assert [len(want[i]["data"][1]) for i in range(3)] == [2, 1, 2]              # the bad key took its source's next deposit with it; one withdrawal was refused

    def run(cfg, deps, wds, ups, state):
        cw = Wr.Writer()
        Wr.enc_config(cw, cfg)
        cb = bytes(cw.b)
        db = _vec([{"mpn_address": tuple(d.mpn_address), "payment": dpay[k]} for k, d in enumerate(deps)], Wr.enc_mpn_deposit)
        wb = _vec([{"mpn_address": tuple(w.mpn_address), "mpn_withdraw_nonce": w.mpn_withdraw_nonce, "mpn_sig": {"r": tuple(w.mpn_sig["r"]), "s": w.mpn_sig["s"]},
                    "payment": wpay[k]} for k, w in enumerate(wds)], Wr.enc_mpn_withdraw)
        ub = _vec([{"nonce": t.nonce, "src_pub_key": tuple(t.src_pub_key), "dst_pub_key": tuple(t.dst_pub_key), "amount": Wk._money_w(t.amount),
                    "fee": Wk._money_w(t.fee), "sig": {"r": tuple(t.sig["r"]), "s": t.sig["s"]}} for t in ups], Wr.enc_mpn_tx)
        rw = np.array([11, 22, 33], np.uint64)
        fee = _canon(U.ZIESHA)
        fork, buf, ln, n = ct.c_void_p(), ct.c_void_p(), ct.c_size_t(), ct.c_uint64()
        hostmpn._check(lib.bzk_mpn_prepare_works(hostmpn._h, state._h, cb, len(cb), db, len(db), wb, len(wb), ub, len(ub), _ptr(rw), 9, _ptr(fee),
                                                 ct.byref(fork), ct.byref(buf), ct.byref(ln), ct.byref(n)))
        out = ct.string_at(buf, ln.value)
        lib.bzk_buffer_free(buf)
        return out, fork, n.value

    got, fork, n = run(config, deposits, withdraws, updates, led)
    assert n == 3 and got == Wr.get_mpn_work_response_to_bytes(want)
    assert led.root == root0 == st.root                                          # built on a fork
    root, size, cnt, pend = np.zeros(4, np.uint64), ct.c_uint64(), ct.c_uint64(), ct.c_uint64()
    assert lib.bzk_mpn_state_info(fork, _ptr(root), ct.byref(size), ct.byref(cnt), ct.byref(pend)) == 0
    assert (_int(root), size.value, pend.value) == (fork_py.root, fork_py.state_size, len(fork_py.new_account_indices))
    # `final_delta`: the changed scalar leaves as ZkDeltaPairs, equal to the restatement's (same entries, same order, same bytes)
    dbuf, dlen, dn = ct.c_void_p(), ct.c_size_t(), ct.c_uint64()
    assert lib.bzk_mpn_state_delta(led._h, fork, ct.byref(dbuf), ct.byref(dlen), ct.byref(dn)) == 0
    dw = Wr.Writer()
    delta = Wk.final_delta(st, fork_py)
    Wk.enc_delta(dw, delta)
    assert dn.value == len(delta) > 10 and ct.string_at(dbuf, dlen.value) == bytes(dw.b)
    lib.bzk_buffer_free(dbuf)
    lib.bzk_mpn_state_free(fork)
    # several batches of one kind continue each other (mod.rs:396-414)
    st2, keys2 = make_state(3, 3, 3)
    ups = [transfer(keys2, 0, 1, 1), transfer(keys2, 1, 2, 1), transfer(keys2, 2, 0, 1), transfer(keys2, 0, 2, 2), transfer(keys2, 1, 0, 2),
           transfer(keys2, 2, 1, 2)]
    led2 = _load(hostmpn, st2, 3, 3)
    cfg2 = _config(num=(0, 0, 2))
    want2, _ = Wk.prepare_works(cfg2, st2, [], [], ups, rewards, height=9)
    got2, fork2, n2 = run(cfg2, [], [], ups, led2)
    assert n2 == 2 and got2 == Wr.get_mpn_work_response_to_bytes(want2)
    lib.bzk_mpn_state_free(fork2)
    # malformed inputs are refused
    fork, buf, ln, n = ct.c_void_p(), ct.c_void_p(), ct.c_size_t(), ct.c_uint64()
    rw, fee = np.array([1, 2, 3], np.uint64), _canon(U.ZIESHA)
    assert lib.bzk_mpn_prepare_works(hostmpn._h, led._h, b"\x03\x03", 2, None, 0, None, 0, None, 0, _ptr(rw), 0, _ptr(fee), ct.byref(fork), ct.byref(buf),
                                     ct.byref(ln), ct.byref(n)) == -1
    # a ledger of another shape than the config's is refused before any row is written
    from bazuka_b200.mpn.ledger import NativeLedger
    odd = NativeLedger(hostmpn, 4, 2)
    cw = Wr.Writer()
    Wr.enc_config(cw, _config())
    assert lib.bzk_mpn_prepare_works(hostmpn._h, odd._h, bytes(cw.b), len(cw.b), None, 0, None, 0, None, 0, _ptr(rw), 0, _ptr(fee), ct.byref(fork),
                                     ct.byref(buf), ct.byref(ln), ct.byref(n)) == -1
    led.free(); led2.free(); odd.free()


def test_native_mpn_work_verify_with_a_real_proof(hostmpn, works, cref):
    """`MpnWork::verify` natively (bzk_mpn_work_verify: commitment from (prover, reward), then check_proof against the work's
    own verifying key).  The key and the proof come from the C oracle's setup / prover on a small circuit with the five public
    inputs of an MPN proof, proved for the values this work and this prover address imply: accepted; another address, another
    reward, or another work's inputs are not."""
    from oracle import groth16_c as GC
    from oracle.py import groth16 as G
    from bazuka_b200 import groth16 as BG
    from conftest import fr_arr
    from test_groth16_cpu import to_csr
    lib = hostmpn._l
    cs = G.R1CS(num_inputs=6, num_aux=2)             # inputs 1..5 = commitment, height, state, aux_data, next_state
    cs.enforce([(6, 1)], [(6, 1)], [(7, 1)])          # w * w = u
    cs.enforce([(7, 1), (1, 1)], [(0, 1)], [(7, 1), (1, 1)])   # a row that reads a public input
    mats = to_csr(cs)
    pk = GC.setup(cs.num_inputs, cs.num_aux, mats, cref.fr_random(15, 5))
    vk_blob = bytes(BG.vk_to_bincode(pk["vk"]))
    me, other = bytes(range(32)), bytes(range(1, 33))
    work = dict(works[2], config=dict(works[2]["config"], update_vk=vk_blob))
    pub = Wk.work_public_inputs(work, me)
    z = fr_arr([1] + pub + [3, 9])
    r, s = cref.fr_random(16, 2)
    proof = bytes(GC.proof_bytes(*GC.prove(cs.num_inputs, cs.num_aux, mats, pk, z[:6], z[6:], r, s)))
    assert Wk.verify_work(work, me, np.frombuffer(proof, dtype=np.uint8))

    def check(w, addr):
        h = _Work(lib, Wr.work_to_bytes(w))
        st = lib.bzk_mpn_work_verify(h.h, addr, proof)
        h.free()
        return st

    assert check(work, me) == 1
    assert check(work, other) == 0                                   # the commitment binds the proof to its prover
    assert check(dict(work, reward=work["reward"] + 1), me) == 0     # ... and to the reward
    assert check(dict(work, public_inputs=dict(work["public_inputs"], height=10)), me) == 0
    assert check(works[2], me) == 0                                  # another key (here: not even valid points) never accepts


def test_worker_loop_over_the_native_prover_and_codec(hostmpn, works):
    """`WorkerClient.run_once` with `NativeMpnProver` against an in-process node that speaks only through the NATIVE codec: the
    node serves the GetMpnWorkResponse image, the worker proves each work from its bytes (here over the host build, whose prove
    call checks satisfiability and returns identity points), posts a PostMpnSolutionRequest; the node decodes it, finds one
    387-byte proof per work and runs `MpnWork::verify` on each (identity points are not a valid proof: 0 accepted)."""
    import json, os
    lib = hostmpn._l
    me = bytes(range(32))
    nat = Wk.NativeMpnProver(hostmpn)
    # the node checks against a REAL key (a production MPN key, tests/golden/mpn_vks.json): the fixture's works carry placeholder
    # key images whose "points" all have the infinity byte set — under such a key e(alpha, beta) = 1 and identity points verify
    prod_vk = bytes.fromhex(next(iter(json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mpn_vks.json")))["vks"].values())))

    class _Circ:            # what NativeMpnProver.add_circuit needs of mpn.native_circuit.Native*Circuit / groth16.ProvingKey
        def __init__(self, h): self._h = h
    for kind in ("deposit", "withdraw", "update"):
        c, _ = _compile(lib, kind, 3, 3, 1)
        nat.add_circuit(kind, _Circ(c), _Circ(ct.c_void_p(1)))
        lib.bzk_mpn_circuit_free(c)
    log = []

    def node(method, url, body):
        if url.endswith("/bincode/mpn/work"):
            want = ct.create_string_buffer(40)
            assert lib.bzk_mpn_get_work_request_encode(me, want) == 0 and body == want.raw
            return Wr.get_mpn_work_response_to_bytes(works)
        prover, proofs = Wr.post_mpn_solution_request_from_bytes(body)
        ok = 0
        for wid, p in proofs.items():
            assert len(p) == 387
            kind = works[wid]["data"][0]
            h = _Work(lib, Wr.work_to_bytes(dict(works[wid], config=dict(works[wid]["config"], **{kind + "_vk": prod_vk}))))
            ok += lib.bzk_mpn_work_verify(h.h, prover, p) == 1
            h.free()
        log.append((prover, sorted(proofs)))
        return ok.to_bytes(8, "little")

    client = Wk.WorkerClient("127.0.0.1:1", me, nat, opener=node)
    zero = np.zeros(4, np.uint64)
    assert client.run_once(lambda: (zero, zero)) == (3, 0)
    assert log == [(me, [0, 1, 2])]
    nat.free()


def test_transactions_to_accepted_proofs_through_the_native_path(hostmpn, cref):
    """the whole protocol on the CPU tier with REAL proofs: keys from the C oracle's setup on the natively compiled circuits;
    a block's traffic -> bzk_mpn_prepare_works (native builders) -> GetMpnWorkResponse bytes -> per work bzk_mpn_prover_prove_work
    (native codec, rows, witness drivers; this tier's prove call checks satisfiability and keeps z) -> the C ORACLE's prover on
    that z and key -> bzk_mpn_work_verify (`MpnWork::verify`, libbzk's host pairing) accepts the proof for the prover it was made
    for and for no other.  What the GPU tier adds is the MSM / NTT half of the prove call, bit-exact against this same oracle."""
    from oracle import groth16_c as GC
    from bazuka_b200 import groth16 as BG
    from test_native_host_cpu import _load
    from test_mpn_cpu import make_state, transfer
    lib = hostmpn._l
    lib.shim_last_z.argtypes = [ct.c_void_p, ct.c_uint64, ct.c_void_p, ct.c_uint64]
    A = T = 3
    B = 0                                                        # one slot per batch keeps the oracle's setup to seconds
    keys_c, provers, r1 = {}, {}, {}
    jj_d, fee = _canon(N.JJ_D), _canon(U.ZIESHA)
    for k, kind in enumerate(("deposit", "withdraw", "update")):
        c, blob = _compile(lib, kind, A, T, B)
        shape = np.zeros(12, np.uint64)
        assert lib.bzk_mpn_circuit_shape(c, _ptr(shape)) == 0
        ni, na, ncons = int(shape[0]), int(shape[1]), int(shape[2])
        mats = []
        for side in range(3):
            nnz = int(shape[3 + side])
            rp, col, val = np.zeros(ncons + 1, np.uint64), np.zeros(max(nnz, 1), np.uint32), np.zeros((max(nnz, 1), 4), np.uint64)
            assert lib.bzk_mpn_circuit_matrix(c, side, _ptr(rp), _ptr(col), _ptr(val)) == 0
            mats.append((rp, col[:nnz], val[:nnz]))
        r1[kind] = (ni, na, mats)
        keys_c[kind] = GC.setup(ni, na, mats, cref.fr_random(40 + k, 5))
        p = ct.c_void_p()
        hostmpn._check(lib.bzk_mpn_prover_create(hostmpn._h, c, ct.c_void_p(1), _ptr(jj_d), _ptr(fee), ct.byref(p)))
        lib.bzk_mpn_circuit_free(c)
        provers[kind] = p
    cfg = dict(_config(), log4_deposit_batch_size=B, log4_withdraw_batch_size=B, log4_update_batch_size=B,
               **{kind + "_vk": bytes(BG.vk_to_bincode(keys_c[kind]["vk"])) for kind in keys_c})
    # the block: a newcomer deposits, an old account withdraws, the newcomer pays an old account — each batch one transaction
    st, keys = make_state(A, T, 3)
    keys.append(N.eddsa_keys(b"dep-new"))
    dep = {"mpn_address": tuple(N.jj_compress(keys[3][0])),
           "payment": {"memo": "hello", "contract_id": 0x1234, "deposit_circuit_id": 0, "calldata": 0, "src": bytes([7]) * 32,
                       "amount": {"token_id": "ziesha", "amount": 5000}, "fee": {"token_id": "ziesha", "amount": 0}, "nonce": 1, "sig": None}}
    w = D.MpnWithdraw(N.jj_compress(keys[1][0]), 1, amount=U.Money(U.ZIESHA, 100), fee=U.Money(U.ZIESHA, 2))
    pay = {"memo": "rent", "contract_id": 0x1234, "withdraw_circuit_id": 0, "calldata": 0, "dst": bytes(range(32)),
           "amount": {"token_id": "ziesha", "amount": 100}, "fee": {"token_id": "ziesha", "amount": 2}}
    w.fingerprint = Wk.withdraw_fingerprint(pay)
    w.sign(keys[1][1])
    pay["calldata"] = w.expected_calldata()
    wd = {"mpn_address": tuple(w.mpn_address), "mpn_withdraw_nonce": 1, "mpn_sig": {"r": tuple(w.mpn_sig["r"]), "s": w.mpn_sig["s"]}, "payment": pay}
    t = transfer(keys, 3, 0, 1, amount=40, fee=1)
    up = {"nonce": t.nonce, "src_pub_key": tuple(t.src_pub_key), "dst_pub_key": tuple(t.dst_pub_key), "amount": Wk._money_w(t.amount),
          "fee": Wk._money_w(t.fee), "sig": {"r": tuple(t.sig["r"]), "s": t.sig["s"]}}
    led = _load(hostmpn, st, A, T)
    cw = Wr.Writer()
    Wr.enc_config(cw, cfg)
    cb, db, wb, ub = bytes(cw.b), _vec([dep], Wr.enc_mpn_deposit), _vec([wd], Wr.enc_mpn_withdraw), _vec([up], Wr.enc_mpn_tx)
    rw = np.array([11, 22, 33], np.uint64)
    fork, buf, ln, n = ct.c_void_p(), ct.c_void_p(), ct.c_size_t(), ct.c_uint64()
    hostmpn._check(lib.bzk_mpn_prepare_works(hostmpn._h, led._h, cb, len(cb), db, len(db), wb, len(wb), ub, len(ub), _ptr(rw), 9, _ptr(fee), ct.byref(fork),
                                             ct.byref(buf), ct.byref(ln), ct.byref(n)))
    resp = ct.string_at(buf, ln.value)
    lib.bzk_buffer_free(buf)
    assert n.value == 3
    served = Wr.get_mpn_work_response_from_bytes(resp)
    assert [len(served[i]["data"][1]) for i in range(3)] == [1, 1, 1]           # every batch took its transaction
    me, other = bytes(range(32)), bytes(range(1, 33))
    for wid, work in served.items():
        kind = work["data"][0]
        blob = Wr.work_to_bytes(work)
        out = ct.create_string_buffer(391)
        r, s = cref.fr_random(60 + wid, 2)
        assert lib.bzk_mpn_prover_prove_work(hostmpn._h, provers[kind], blob, len(blob), me, _ptr(r), _ptr(s), 1, out) == 0
        ni, na, mats = r1[kind]
        z_in, z_aux = np.zeros((ni, 4), np.uint64), np.zeros((na, 4), np.uint64)
        assert lib.shim_last_z(_ptr(z_in), ni, _ptr(z_aux), na) == 0
        from bazuka_b200.mpn.cs import to_mont
        assert (z_in[1:] == to_mont(Wk.work_public_inputs(work, me))).all()     # the public inputs the node will check against
        proof = bytes(GC.proof_bytes(*GC.prove(ni, na, mats, keys_c[kind], z_in, z_aux, r, s)))
        h = _Work(lib, blob)
        assert lib.bzk_mpn_work_verify(h.h, me, proof) == 1
        assert lib.bzk_mpn_work_verify(h.h, other, proof) == 0
        h.free()
    for p in provers.values():
        lib.bzk_mpn_prover_free(hostmpn._h, p)
    lib.bzk_mpn_state_free(fork)
    led.free()


def test_decoder_survives_mutated_images(hostmpn, works, host_hasher):
    """truncations, bit flips and spliced length prefixes of valid works: every image is either refused (BZK_ERR_BAD_ARG) or
    decodes to something that re-encodes and whose rows can be asked for — never a crash or an out-of-bounds read (the length
    prefixes of vectors, maps, strings and keys are all bounded before anything is allocated or copied)."""
    import random
    lib = hostmpn._l
    rng = random.Random(20260923)
    jj_d, fee = _canon(N.JJ_D), _canon(U.ZIESHA)
    refused = accepted = 0
    for work in works.values():
        blob = Wr.work_to_bytes(work)
        for trial in range(400):
            b = bytearray(blob)
            mode = trial % 4
            if mode == 0:
                b = b[:rng.randrange(len(b))]
            elif mode == 1:
                for _ in range(rng.randint(1, 4)):
                    b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            elif mode == 2:                                   # a huge length prefix somewhere
                o = rng.randrange(len(b) - 8)
                b[o:o + 8] = struct.pack("<Q", rng.choice([2**63, 2**32, 10**6, 65]))
            else:                                             # garbage tail after a valid head
                cut = rng.randrange(len(b))
                b = b[:cut] + bytes(rng.randrange(256) for _ in range(rng.randint(1, 64)))
            h = ct.c_void_p()
            st = lib.bzk_mpn_work_decode(bytes(b), len(b), ct.byref(h), None)
            assert st in (0, -1)
            if st != 0:
                refused += 1
                continue
            accepted += 1
            n = ct.c_size_t()
            assert lib.bzk_mpn_work_encode(h, None, 0, ct.byref(n)) == 0 and n.value == len(b)    # canonical: same length back
            info = np.zeros(64, np.uint32)
            lib.bzk_mpn_work_get_info(h, _ptr(info))
            kind, A, T, B = (int(x) for x in info[:4])
            if (A, T, B) == (3, 3, 1):                        # buffers sized for the original shape
                if kind == 2:
                    raws, ext = np.zeros((4, 32 + 9 * T + 6 * A, 4), np.uint64), np.zeros((4, 2, 4), np.uint64)
                    assert lib.bzk_mpn_work_update_rows(h, host_hasher, _ptr(jj_d), _ptr(fee), _ptr(raws), _ptr(ext)) in (0, -1, -4)
                else:
                    w1, w2, wr = (5, 9 + 3 * T + 3 * A, 4) if kind == 0 else (12, 12 + 6 * T + 3 * A, 7)
                    r1, r2, ro, rv = (np.zeros((4, w, 4), np.uint64) for w in (w1, w2, 1, wr))
                    assert lib.bzk_mpn_work_dw_rows(h, host_hasher, _ptr(jj_d), _ptr(r1), _ptr(r2), _ptr(ro), _ptr(rv)) in (0, -1, -4)
            lib.bzk_mpn_work_free(h)
    assert refused > 600 and accepted > 50, (refused, accepted)


@pytest.mark.parametrize("kind", ["deposit", "withdraw", "update"])
def test_reference_circuit_test_shape_all_null_batches(hostmpn, cref, kind):
    """the reference's own circuit tests (/root/reference/src/mpn/circuits/test.rs:117-229): each MPN circuit at log4 sizes 3 / 3 / 1
    with four NULL transitions -> setup -> prove -> verify.  Here on the native path: the circuit compiled by C++, an `MpnWork`
    without transitions through bzk_mpn_prover_prove_work (all four slots padded like `..Transition::null`), setup and proof by
    the C oracle on the resulting assignment, `MpnWork::verify` by libbzk's pairing — accepted; a wrong height is not."""
    from oracle import groth16_c as GC
    from bazuka_b200 import groth16 as BG
    lib = hostmpn._l
    lib.shim_last_z.argtypes = [ct.c_void_p, ct.c_uint64, ct.c_void_p, ct.c_uint64]
    A, T, B = 3, 3, 1
    st = U.MpnState(A, T)
    pub = (D.deposit(st, [], B)[0] if kind == "deposit" else D.withdraw(st, [], B)[0] if kind == "withdraw" else U.update(st, [], B)[0])
    c, blob = _compile(lib, kind, A, T, B)
    shape = np.zeros(12, np.uint64)
    lib.bzk_mpn_circuit_shape(c, _ptr(shape))
    ni, na, ncons = int(shape[0]), int(shape[1]), int(shape[2])
    mats = []
    for side in range(3):
        nnz = int(shape[3 + side])
        rp, col, val = np.zeros(ncons + 1, np.uint64), np.zeros(max(nnz, 1), np.uint32), np.zeros((max(nnz, 1), 4), np.uint64)
        lib.bzk_mpn_circuit_matrix(c, side, _ptr(rp), _ptr(col), _ptr(val))
        mats.append((rp, col[:nnz], val[:nnz]))
    key = GC.setup(ni, na, mats, cref.fr_random(90, 5))
    jj_d, fee = _canon(N.JJ_D), _canon(U.ZIESHA)
    p = ct.c_void_p()
    hostmpn._check(lib.bzk_mpn_prover_create(hostmpn._h, c, ct.c_void_p(1), _ptr(jj_d), _ptr(fee), ct.byref(p)))
    lib.bzk_mpn_circuit_free(c)
    cfg = dict(_config(), **{kind + "_vk": bytes(BG.vk_to_bincode(key["vk"]))})
    work = {"config": cfg, "public_inputs": dict(pub, height=3), "data": (kind, []), "new_root": {"state_hash": st.root, "state_size": 0}, "reward": 5}
    wb = Wr.work_to_bytes(work)
    me = bytes(range(32))
    r, s = cref.fr_random(91, 2)
    out = ct.create_string_buffer(391)
    assert lib.bzk_mpn_prover_prove_work(hostmpn._h, p, wb, len(wb), me, _ptr(r), _ptr(s), 1, out) == 0
    z_in, z_aux = np.zeros((ni, 4), np.uint64), np.zeros((na, 4), np.uint64)
    assert lib.shim_last_z(_ptr(z_in), ni, _ptr(z_aux), na) == 0
    proof = bytes(GC.proof_bytes(*GC.prove(ni, na, mats, key, z_in, z_aux, r, s)))
    h = _Work(lib, wb)
    assert lib.bzk_mpn_work_verify(h.h, me, proof) == 1
    h.free()
    h = _Work(lib, Wr.work_to_bytes(dict(work, public_inputs=dict(pub, height=4))))
    assert lib.bzk_mpn_work_verify(h.h, me, proof) == 0
    h.free()
    lib.bzk_mpn_prover_free(hostmpn._h, p)


def test_single_update_at_the_production_tree_shape_through_the_native_prover(hostmpn):
    """BASELINE configs[0] — UpdateCircuit A = 15, T = 3, B = 0 (56 776 constraints), one signed transfer between two funded
    accounts — from the wire image of its work through bzk_mpn_prover_prove_work on the host build: the natively compiled
    circuit is satisfied by the assignment the native rows + witness driver produce (depth-15 Merkle paths, both EdDSA ladders)."""
    from test_mpn_cpu import transfer
    lib = hostmpn._l
    lib.shim_last_unsat_row.restype = ct.c_uint64
    A, T, B = 15, 3, 0
    st, keys = U.MpnState(A, T), []
    for i, seed in enumerate((b"ABC", b"DEF")):                      # seeds of `TxBuilder::new`
        pk, sk = N.eddsa_keys(seed)
        keys.append((pk, sk))
        st.set(i, U.MpnAccount(0, 0, pk, {0: U.Money(U.ZIESHA, 10 ** 12)}))
    pub, trans, rej = U.update(st, [transfer(keys, 0, 1, 1, amount=1000, fee=10)], B)
    assert len(trans) == 1 and not rej
    cfg = dict(_config(), log4_tree_size=A, log4_token_tree_size=T, log4_update_batch_size=B)
    work = {"config": cfg, "public_inputs": dict(pub, height=1), "data": ("update", Wk.transitions_to_wire("update", trans)),
            "new_root": {"state_hash": st.root, "state_size": st.state_size}, "reward": 7}
    wb = Wr.work_to_bytes(work)
    c, blob = _compile(lib, "update", A, T, B)
    shape = np.zeros(12, np.uint64)
    lib.bzk_mpn_circuit_shape(c, _ptr(shape))
    assert int(shape[2]) == 56776                                     # ≈ 56.8 k constraints for one transaction
    jj_d, fee = _canon(N.JJ_D), _canon(U.ZIESHA)
    p = ct.c_void_p()
    hostmpn._check(lib.bzk_mpn_prover_create(hostmpn._h, c, ct.c_void_p(1), _ptr(jj_d), _ptr(fee), ct.byref(p)))
    lib.bzk_mpn_circuit_free(c)
    zero, out = np.zeros(4, np.uint64), ct.create_string_buffer(391)
    st_ = lib.bzk_mpn_prover_prove_work(hostmpn._h, p, wb, len(wb), bytes(range(32)), _ptr(zero), _ptr(zero), 1, out)
    assert st_ == 0, (st_, lib.shim_last_unsat_row())
    # a forged signature is caught by the circuit, not proved
    bad = Wk.transitions_to_wire("update", trans)
    bad[0]["tx"]["sig"]["s"] = (bad[0]["tx"]["sig"]["s"] + 1) % N.R
    bb = Wr.work_to_bytes(dict(work, data=("update", bad)))
    assert lib.bzk_mpn_prover_prove_work(hostmpn._h, p, bb, len(bb), bytes(range(32)), _ptr(zero), _ptr(zero), 1, out) == -7
    lib.bzk_mpn_prover_free(hostmpn._h, p)
