"""GPU tier: real MPN update proofs.  UpdateCircuit (reference test shape A=3,T=3,B=1 with signed
transfers, and the production tree shape A=15,T=3 with one transaction = BASELINE configs[0]) is
synthesised by the host layer, proved on the GPU, and checked three ways: proof bytes equal the CPU
prover's on the same parameters, the pairing verifier accepts with the five public inputs in
`groth16_verify` order (/root/reference/src/zk/groth16/mod.rs:109-118), and rejects a wrong input."""
import numpy as np
import pytest

from test_mpn_cpu import make_state, transfer

pytestmark = pytest.mark.gpu


def _prove_and_check(ctx, cref, circ, seed):
    from bazuka_b200 import groth16 as BG
    from bazuka_b200.mpn import cs as C
    from oracle import groth16_c as GC
    cs = circ.synthesize(C.ConstraintSystem())
    assert cs.is_satisfied()[0]
    ni, na, mats, inputs, aux = cs.to_csr()
    pr = BG.Prover(ctx, BG.R1CS(ni, na, *mats))
    tox = cref.fr_random(seed, 5)
    pk, vk = BG.setup_gpu(ctx, pr.r1cs, tox, cref.g1_generator(), cref.g2_generator())
    r, s = cref.fr_random(seed + 1, 2)
    blob, pts = pr.prove(pk, inputs, aux, r, s)
    a_idx, b_idx = GC.density(ni, na, mats)
    cpk = {"log_m": pr.log_m, "vk": vk, "a_idx": a_idx, "b_idx": b_idx}
    for k in ("h", "l", "a", "b_g1", "b_g2"):
        cpk[k] = pk.device_images[k].cpu().numpy()
    assert (blob == GC.proof_bytes(*GC.prove(ni, na, mats, cpk, inputs, aux, r, s))).all()
    # the product's own verifier (host pairing in libbzk) and the big-integer oracle both accept
    assert BG.verify(vk, inputs[1:], pts)
    assert BG.verify_bytes(BG.vk_to_bincode(vk), inputs[1:], blob)
    assert GC.verify_py(vk, inputs[1:], pts)
    wrong = inputs[1:].copy()
    wrong[4] = wrong[3]  # claim a different next_state
    assert not BG.verify(vk, wrong, pts)
    return cs, pr


def test_update_circuit_reference_shape_real_transfers(ctx, cref):
    from bazuka_b200.mpn import native as N, update as U
    st, keys = make_state(3, 3, 3)
    keys.append(N.eddsa_keys(b"newcomer"))
    txs = [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1, amount=5), transfer(keys, 0, 3, 2, amount=77, fee=3)]
    pub, trans, rej = U.update(st, txs, 1)
    assert len(trans) == 3 and not rej
    cs, pr = _prove_and_check(ctx, cref, U.UpdateCircuit(3, 3, 1, commitment=42, height=7, transitions=trans, **pub), 61)
    assert pr.log_m == 17


def test_update_circuit_production_tree_single_tx(ctx, cref):
    from bazuka_b200.mpn import update as U
    st, keys = make_state(15, 3, 2)
    pub, trans, _ = U.update(st, [transfer(keys, 0, 1, 1)], 0)
    cs, pr = _prove_and_check(ctx, cref, U.UpdateCircuit(15, 3, 0, commitment=1, height=0, transitions=trans, **pub), 71)
    assert pr.log_m == 16 and abs(cs.num_constraints - 56800) < 200


def test_null_witness_proves_with_same_key_shape(ctx, cref):
    """the circuit shape is witness-independent: an all-null batch (the reference's own circuit test,
    /root/reference/src/mpn/circuits/test.rs:117-149) has the same R1CS as a real one."""
    from bazuka_b200.mpn import cs as C, native as N, update as U
    a = U.UpdateCircuit(3, 3, 1, state=5, next_state=5, aux_data=N.poseidon([U.ZIESHA, 0])).synthesize(C.ConstraintSystem())
    st, keys = make_state(3, 3, 2)
    pub, trans, _ = U.update(st, [transfer(keys, 0, 1, 1)], 1)
    b = U.UpdateCircuit(3, 3, 1, transitions=trans, **pub).synthesize(C.ConstraintSystem())
    ma, mb = a.to_csr()[2], b.to_csr()[2]
    for (rp1, c1, v1), (rp2, c2, v2) in zip(ma, mb):
        assert (rp1 == rp2).all() and (c1 == c2).all() and (v1 == v2).all()


def test_deposit_and_withdraw_circuits_prove(ctx, cref):
    from bazuka_b200.mpn import dw as D, native as N, update as U
    st, keys = make_state(3, 3, 2)
    deps = [D.MpnDeposit(N.jj_compress(keys[0][0]), U.ZIESHA, 500), D.MpnDeposit(N.jj_compress(N.eddsa_keys(b"dep-new")[0]), 77, 9)]
    pub, tr = D.deposit(st, deps, 1)
    _prove_and_check(ctx, cref, D.DepositCircuit(3, 3, 1, commitment=3, height=1, transitions=tr, **pub), 81)
    w = D.MpnWithdraw(N.jj_compress(keys[1][0]), 1, amount=U.Money(U.ZIESHA, 100), fee=U.Money(U.ZIESHA, 2), fingerprint=4242)
    w.sign(keys[1][1])
    pub, tr = D.withdraw(st, [w], 1)
    assert len(tr) == 1
    _prove_and_check(ctx, cref, D.WithdrawCircuit(3, 3, 1, commitment=4, height=2, transitions=tr, **pub), 91)


def test_gpu_witness_equals_host_synthesis_and_proves(ctx, cref):
    """csrc/witness.cu (one thread per slot interpreting the compiled slot program) writes exactly the aux
    vector `UpdateCircuit.synthesize` assigns, and the prover fed from device memory yields the same proof
    bytes as the host-witness call."""
    import torch
    from bazuka_b200 import groth16 as BG
    from bazuka_b200.mpn import cs as C, native as N, update as U
    from bazuka_b200.mpn.gpu_witness import UpdateWitnessGpu
    st, keys = make_state(3, 3, 3)
    keys.append(N.eddsa_keys(b"newcomer"))
    txs = [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1, amount=5), transfer(keys, 0, 3, 2, amount=77, fee=3)]
    pub, trans, _ = U.update(st, txs, 1)
    circ = U.UpdateCircuit(3, 3, 1, commitment=42, height=7, transitions=trans, **pub)
    ni, na, mats, inputs, aux = circ.synthesize(C.ConstraintSystem()).to_csr()
    gw = UpdateWitnessGpu(ctx, 3, 3)
    d_in, d_aux = gw.witness(circ)
    got = d_aux.cpu().numpy().view(np.uint64)
    assert got.shape == aux.shape
    bad = np.nonzero((got != aux).any(axis=1))[0]
    assert len(bad) == 0, (len(bad), bad[:8], [gw.prog.ops[(b - gw.prog.p_aux) % gw.prog.n_ops].tolist() for b in bad[:8]])
    assert (d_in.cpu().numpy().view(np.uint64) == inputs).all()
    pr = BG.Prover(ctx, BG.R1CS(ni, na, *mats))
    pk, vk = BG.setup_gpu(ctx, pr.r1cs, cref.fr_random(91, 5), cref.g1_generator(), cref.g2_generator())
    r, s = cref.fr_random(92, 2)
    blob_host, _ = pr.prove(pk, inputs, aux, r, s)
    blob_dev, pts = pr.prove_dev(pk, d_in, d_aux, r, s)
    assert (blob_host == blob_dev).all()
    assert BG.verify(vk, inputs[1:], pts)
    # a null batch (all slots disabled) through the same program
    null = U.UpdateCircuit(3, 3, 1, state=st.root, next_state=st.root, aux_data=N.poseidon([U.ZIESHA, 0]))
    _, _, _, _, aux0 = null.synthesize(C.ConstraintSystem()).to_csr()
    _, d_aux0 = gw.witness(null)
    assert (d_aux0.cpu().numpy().view(np.uint64) == aux0).all()
    gw.free(); pk.free(); pr.free()


def test_versioned_tree_update_kernel_vs_sequential_reference(ctx):
    """k_tree4_versioned_level (one launch per level for the whole batch) against the write-by-write loop:
    random forests with heavy collisions (same leaves rewritten, neighbours under one parent, several trees),
    depths 1, 3, 15 and 17."""
    import random
    from bazuka_b200.mpn import native as N
    from bazuka_b200.mpn.batch_update import GpuTreeHasher
    from oracle.py.state import sequential_tree_updates
    h = GpuTreeHasher(ctx)
    rng = random.Random(5)
    for depth, n, ntrees, span in ((1, 9, 2, 4), (3, 60, 3, 64), (15, 150, 2, 40), (17, 40, 1, 1 << 34)):
        tree_ids = [rng.randrange(ntrees) for _ in range(n)]
        indices = [rng.randrange(min(span, 1 << (2 * depth))) for _ in range(n)]
        leaves = [rng.randrange(N.R) for _ in range(n)]
        # pre-batch trees: sparse trees with a few random leaves, proofs read from them
        trees = [N.SparseTree4(depth, rng.randrange(N.R)) for _ in range(ntrees)]
        for t in trees:
            for _ in range(5):
                t.set_leaf(rng.randrange(min(span, 1 << (2 * depth))), rng.randrange(N.R))
        init = [trees[t].prove(i) for t, i in zip(tree_ids, indices)]
        want_vals, want_proofs = sequential_tree_updates(depth, tree_ids, indices, leaves, init, N.poseidon)
        got_vals, got_proofs = h.tree_update(depth, tree_ids, indices, leaves, init)
        assert got_proofs == want_proofs, depth
        assert got_vals == want_vals, depth
        # and the sequential reference is the plain sparse tree: roots agree write by write
        for e in range(n):
            trees[tree_ids[e]].set_leaf(indices[e], leaves[e])
            assert trees[tree_ids[e]].root == want_vals[depth][e]


def test_batched_transition_builder_on_gpu_equals_update(ctx):
    """update_batched with the GPU primitives == update(): transitions, public inputs, rejections, final state."""
    import copy
    from bazuka_b200.mpn import batch_update as BU, update as U
    from test_mpn_cpu import _batch_scenario, _assert_same_transitions
    st1, txs = _batch_scenario()
    st2 = copy.deepcopy(st1)
    pub1, tr1, rej1 = U.update(st1, txs, 2)
    pub2, tr2, rej2 = BU.update_batched(BU.GpuTreeHasher(ctx), st2, txs, 2)
    assert pub1 == pub2 and rej1 == rej2 and len(tr1) == 6
    _assert_same_transitions(tr1, tr2)
    assert st1.root == st2.root and st1.tree.levels == st2.tree.levels


@pytest.mark.parametrize("kind", ["deposit", "withdraw"])
def test_deposit_withdraw_gpu_witness_and_proof(ctx, cref, kind):
    """the two-phase slot programs on the GPU (+ host reveal) give `synthesize`'s aux vector for the deposit and
    withdraw circuits, and the proof from the resident witness verifies with the five public inputs."""
    from bazuka_b200 import groth16 as BG
    from bazuka_b200.mpn import cs as C
    from bazuka_b200.mpn.dw_witness import TwoPhaseWitnessGpu
    from test_mpn_cpu import _dw_scenario
    circ = _dw_scenario(kind)
    ni, na, mats, inputs, aux = circ.synthesize(C.ConstraintSystem()).to_csr()
    gw = TwoPhaseWitnessGpu(ctx, kind, 3, 3)
    d_in, d_aux = gw.witness(circ)
    got = d_aux.cpu().numpy().view(np.uint64)
    assert got.shape == aux.shape
    bad = np.nonzero((got != aux).any(axis=1))[0]
    assert len(bad) == 0, (len(bad), bad[:8])
    assert (d_in.cpu().numpy().view(np.uint64) == inputs).all()
    pr = BG.Prover(ctx, BG.R1CS(ni, na, *mats))
    pk, vk = BG.setup_gpu(ctx, pr.r1cs, cref.fr_random(95, 5), cref.g1_generator(), cref.g2_generator())
    r, s = cref.fr_random(96, 2)
    blob, pts = pr.prove_dev(pk, d_in, d_aux, r, s)
    assert (blob == pr.prove(pk, inputs, aux, r, s)[0]).all()
    assert BG.verify(vk, inputs[1:], pts)
    gw.free(); pk.free(); pr.free()


def test_batched_deposit_withdraw_builders_on_gpu(ctx):
    """deposit_batched / withdraw_batched with the GPU primitives == the sequential builders."""
    import copy
    from bazuka_b200.mpn import batch_update as BU, dw as D, native as N, update as U
    from test_mpn_cpu import _assert_same_transitions
    h = BU.GpuTreeHasher(ctx)
    st1, keys = make_state(3, 3, 2)
    newpk, _ = N.eddsa_keys(b"dep-new")
    deps = [D.MpnDeposit(N.jj_compress(keys[0][0]), U.ZIESHA, 500), D.MpnDeposit((6, False), 77, 1, "carol"),
            D.MpnDeposit(N.jj_compress(newpk), 77, 9), D.MpnDeposit(N.jj_compress(keys[1][0]), 77, 1, "carol"),
            D.MpnDeposit(N.jj_compress(keys[1][0]), 77, 1), D.MpnDeposit(N.jj_compress(keys[0][0]), 77, 4)]
    st2 = copy.deepcopy(st1)
    pub1, tr1 = D.deposit(st1, deps, 1)
    pub2, tr2 = BU.deposit_batched(h, st2, deps, 1)
    assert pub1 == pub2 and st1.tree.levels == st2.tree.levels and len(tr1) == 4
    _assert_same_transitions(tr1, tr2)
    ws = []
    for i, amt, nonce in ((0, 100, 1), (1, 5, 1), (0, 30, 2)):
        w = D.MpnWithdraw(N.jj_compress(keys[i][0]), nonce, amount=U.Money(U.ZIESHA, amt), fee=U.Money(U.ZIESHA, 2), fingerprint=1000 + amt)
        w.sign(keys[i][1])
        ws.append(w)
    ws[0].calldata = ws[0].expected_calldata()
    bad = D.MpnWithdraw(N.jj_compress(keys[1][0]), 2, amount=U.Money(U.ZIESHA, 1), fee=U.Money(U.ZIESHA, 0), fingerprint=7)
    bad.sign(keys[1][1])
    bad.calldata = bad.expected_calldata() + 1
    ws.insert(2, bad)
    pub1, tr1 = D.withdraw(st1, ws, 1)
    pub2, tr2 = BU.withdraw_batched(h, st2, ws, 1)
    assert pub1 == pub2 and st1.tree.levels == st2.tree.levels and len(tr1) == 3
    _assert_same_transitions(tr1, tr2)


@pytest.mark.parametrize("kind", ["deposit", "withdraw"])
def test_native_deposit_withdraw_builders_and_witness(ctx, cref, kind):
    """csrc/mpn_host.cu bzk_mpn_{deposit,withdraw}_build + bzk_mpn_dw_witness against the Python restatement: same
    accepted set over two consecutive batches (new account, repeated account, bad signature / nonce / balance /
    unknown key), rows equal to `deposit_raws` / `withdraw_raws` of the sequential builder's transitions, same entering
    roots, reveal rows and public values; the resident witness equals `synthesize`'s and its proof verifies."""
    import copy
    from bazuka_b200 import groth16 as BG
    from bazuka_b200.mpn import cs as C, dw as D, dw_witness as DW, native as N, update as U
    from bazuka_b200.mpn.gpu_witness import _canon_rows
    from bazuka_b200.mpn.ledger import NativeLedger
    from bazuka_b200.mpn.native_circuit import NativeTwoPhaseCircuit
    A = T = 3
    B = 1
    st, keys = make_state(A, T, 3)
    led = NativeLedger(ctx, A, T)
    for i, a in st.accounts.items():
        led.set_account(i, a)
    assert led.root == st.root
    new1, new2 = N.eddsa_keys(b"dep-new")[0], N.eddsa_keys(b"dep-new-2")[0]
    if kind == "deposit":
        mk = lambda pk, tok, amt, src=None: D.MpnDeposit(N.jj_compress(pk), tok, amt, src)
        # batch 2: a key that does not decompress puts its L1 source on the rejected list (deposit.rs:33,68-83), which takes the
        # source's next deposit with it; another source's deposit to the same account goes through
        batches = [[mk(keys[0][0], U.ZIESHA, 500, "a"), mk(new1, 77, 9), mk(keys[1][0], 77, 1, "b"), mk(keys[0][0], 77, 4, "a"), mk(new2, 5, 5)],
                   [mk(new1, 78, 3), D.MpnDeposit((6, False), 77, 1, "carol"), mk(new2, 5, 1, "carol"), mk(new2, 5, 2, "dave")]]
        seq, build, raws_of, w_rev = D.deposit, led.deposit_build, DW.deposit_raws, 4
    else:
        def mk(i, amt, nonce, fee=2, sk=None, tok=U.ZIESHA):
            w = D.MpnWithdraw(N.jj_compress(keys[i][0]), nonce, amount=U.Money(tok, amt), fee=U.Money(U.ZIESHA, fee), fingerprint=1000 + amt)
            w.sign(sk or keys[i][1])
            return w
        stranger = D.MpnWithdraw(N.jj_compress(new1), 1, amount=U.Money(U.ZIESHA, 1), fee=U.Money(U.ZIESHA, 0), fingerprint=1)
        good_cd, bad_cd = mk(0, 100, 1), mk(2, 3, 1)         # `verify_calldata` (withdraw.rs:77): the payment's calldata is checked when given
        good_cd.calldata, bad_cd.calldata = good_cd.expected_calldata(), bad_cd.expected_calldata() + 1
        batches = [[good_cd, mk(1, 5, 1, sk=keys[0][1]), mk(1, 5, 1), mk(0, 30, 2), mk(2, 7, 2), stranger, bad_cd],
                   [mk(2, 10**15, 1), mk(2, 7, 1, tok=12345), mk(0, 1, 3), mk(1, 1, 2, fee=10**15), mk(1, 1, 2)]]
        seq, build, raws_of, w_rev = D.withdraw, led.withdraw_build, DW.withdraw_raws, 7
    nc = NativeTwoPhaseCircuit(kind, A, T, B)
    nw = DW.NativeTwoPhaseWitness(ctx, nc)
    circ_cls = D.DepositCircuit if kind == "deposit" else D.WithdrawCircuit
    last = None
    for items in batches:
        before = copy.deepcopy(st)
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
        d_in, d_aux = nw.witness(rows, 3, 1)
        got = d_aux.cpu().numpy().view(np.uint64)
        assert got.shape == aux.shape
        bad = np.nonzero((got != aux).any(axis=1))[0]
        assert len(bad) == 0, (len(bad), bad[:8])
        assert (d_in.cpu().numpy().view(np.uint64) == inputs).all()
        last = (ni, na, mats, inputs, d_in, d_aux)
    ni, na, mats, inputs, d_in, d_aux = last
    n_ni, n_na, n_mats = nc.r1cs()
    assert (n_ni, n_na) == (ni, na)
    pr = BG.Prover(ctx, BG.R1CS(n_ni, n_na, *n_mats))
    pk, vk = BG.setup_gpu(ctx, pr.r1cs, cref.fr_random(195, 5), cref.g1_generator(), cref.g2_generator())
    r, s = cref.fr_random(196, 2)
    blob, pts = pr.prove_dev(pk, d_in, d_aux, r, s)
    assert BG.verify(vk, inputs[1:], pts)
    nw.free(); nc.free(); led.free(); pk.free(); pr.free()


def test_worker_object_end_to_end(ctx, cref):
    """MpnUpdateWorker: transactions -> batched builder -> GPU witness -> resident proof -> 391-byte ZkProof that the
    validator-side byte-image check accepts; byte-equal to the proof of the host-synthesised witness of the
    sequentially built work under the same key; a different claimed state is rejected."""
    import copy
    from bazuka_b200.mpn import cs as C, native as N, update as U
    from bazuka_b200.mpn.worker import MpnUpdateWorker
    st, keys = make_state(3, 3, 3)
    keys.append(N.eddsa_keys(b"newcomer"))
    txs = [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1, amount=5), transfer(keys, 0, 3, 2, amount=77, fee=3)]
    st_seq = copy.deepcopy(st)
    worker = MpnUpdateWorker(ctx, 3, 3, 1, cref.fr_random(77, 5))
    work = worker.build(st, txs, commitment=42, height=7)
    assert work.accepted == 3 and not work.rejected
    r, s = cref.fr_random(78, 2)
    zk = worker.prove(work, r, s)
    assert zk.shape == (391,) and worker.verify(work, zk)
    pub, trans, _ = U.update(st_seq, txs, 1)
    assert st_seq.root == st.root
    ni, na, mats, inputs, aux = U.UpdateCircuit(3, 3, 1, commitment=42, height=7, transitions=trans, **pub).synthesize(C.ConstraintSystem()).to_csr()
    assert (inputs[1:] == work.public_inputs).all()
    blob, _ = worker.prover.prove(worker.pk, inputs, aux, r, s)
    assert (zk[4:] == blob).all()
    bad = copy.copy(work)
    bad.public_inputs = work.public_inputs.copy()
    bad.public_inputs[4] = bad.public_inputs[2]
    assert not worker.verify(bad, zk)
    worker.free()


def test_witness_program_upload_rejects_malformed_programs(ctx):
    """the device interpreter trusts its program, so the upload validates it: unknown opcode, an operand that
    reads a variable defined later, a RAW index past n_raw, a JJ without its NOP half -> BZK_ERR_BAD_ARG."""
    import ctypes as ct
    from bazuka_b200 import _lib
    from bazuka_b200.api import _host_ptr
    from bazuka_b200.mpn.cs import to_mont
    from bazuka_b200.mpn import native as N
    coefs, jj_d = to_mont([1]), to_mont([N.JJ_D])
    lc_ptr = np.array([0, 1, 2], dtype=np.int32)      # lc 0 = V[0] (ONE), lc 1 = V[2] (block variable 1)
    lc_slot = np.array([0, 2], dtype=np.int32)
    lc_coef = np.array([0, 0], dtype=np.int32)

    def upload(ops, n_raw=1):
        ops = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 6)
        h = ct.c_void_p()
        st = ctx._l.bzk_witness_program_upload(ctx._h, _host_ptr(ops), len(ops), _host_ptr(lc_ptr), 2, _host_ptr(lc_slot), _host_ptr(lc_coef), 2,
                                               _host_ptr(coefs), 1, n_raw, 0, _host_ptr(jj_d), ct.byref(h))
        if st == 0:
            ctx._l.bzk_witness_program_free(ctx._h, h)
        return st

    assert upload([[0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0]]) == 0                     # RAW; MUL(ONE, ONE)
    assert upload([[0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0]]) == 0  # lc 1 reads block variable 1, defined earlier
    assert upload([[9, 0, 0, 0, 0, 0]]) == -1
    assert upload([[0, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0]]) == -1   # reads block variable 1 while defining it
    assert upload([[0, 0, 0, 0, 0, 5]]) == -1
    assert upload([[6, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0]]) == -1
    assert upload([[0, 0, 0, 0, 0, 0], [7, 0, 0, 0, 0, 0]]) == -1   # a NOP that is not the second half of a JJ


@pytest.mark.parametrize("kind", ["deposit", "withdraw"])
def test_deposit_withdraw_worker_end_to_end(ctx, cref, kind):
    """deposits / withdrawals -> batched builder -> two-phase GPU witness -> proof blob accepted by the byte-image
    verifier, rejected for another claimed next state; the builder's work equals the sequential builder's."""
    import copy
    from bazuka_b200.mpn import dw as D, native as N, update as U
    from bazuka_b200.mpn.worker import MpnDepositWithdrawWorker
    st, keys = make_state(3, 3, 2)
    if kind == "deposit":
        newpk, _ = N.eddsa_keys(b"dep-new")
        items = [D.MpnDeposit(N.jj_compress(keys[0][0]), U.ZIESHA, 500), D.MpnDeposit(N.jj_compress(newpk), 77, 9),
                 D.MpnDeposit(N.jj_compress(keys[1][0]), 77, 1)]
        seq = lambda s: D.deposit(s, items, 1)
    else:
        items = []
        for i, amt in enumerate([100, 5]):
            w = D.MpnWithdraw(N.jj_compress(keys[i][0]), 1, amount=U.Money(U.ZIESHA, amt), fee=U.Money(U.ZIESHA, 2), fingerprint=1000 + i)
            w.sign(keys[i][1])
            items.append(w)
        seq = lambda s: D.withdraw(s, items, 1)
    st_seq = copy.deepcopy(st)
    worker = MpnDepositWithdrawWorker(ctx, kind, 3, 3, 1, cref.fr_random(81, 5))
    work = worker.build(st, items, commitment=11, height=3)
    pub, trans = seq(st_seq)
    assert work.accepted == len(trans) and st.root == st_seq.root
    assert {k: getattr(work.circuit, k) for k in ("state", "aux_data", "next_state")} == pub
    r, s = cref.fr_random(82, 2)
    zk = worker.prove(work, r, s)
    assert worker.verify(work, zk)
    bad = copy.copy(work)
    bad.public_inputs = work.public_inputs.copy()
    bad.public_inputs[4] = bad.public_inputs[2]
    assert not worker.verify(bad, zk)
    worker.free()


def test_native_ledger_update_build_equals_python_builder(ctx, cref):
    """csrc/mpn_host.cu (C++ ledger logic + batched GPU hashing) against the Python restatement of `update()`:
    same accepted set, same rows of circuit inputs (raw_values order), same entering roots, public inputs and final
    state, over two consecutive batches (new account, self-transfer, same-token fee, four kinds of rejection); and
    the proof made from the native rows equals the proof made from the Python transitions."""
    import copy
    from bazuka_b200 import groth16 as BG
    from bazuka_b200.mpn import cs as C, native as N, update as U, witness_program as W
    from bazuka_b200.mpn.gpu_witness import UpdateWitnessGpu, _canon_rows
    from bazuka_b200.mpn.ledger import NativeLedger
    from test_mpn_cpu import _batch_scenario
    st, txs = _batch_scenario()
    led = NativeLedger(ctx, 3, 3)
    assert led.n_raw == len(W.raw_values(U.UpdateTransition.null(3, 3), 3, 3))
    for i, a in st.accounts.items():
        led.set_account(i, a)
    assert led.root == st.root
    assert NativeLedger(ctx, 30, 1).root == int("501a18871f186db1437e77e2c33acfa81405608cc60806399347215dbe98f714", 16)  # reference KAT
    circs = []
    for batch, B in ((txs, 2), (txs[:0] + [t for t in txs[6:]], 1)):
        pub, trans, rej = U.update(st, batch, B)
        circ = U.UpdateCircuit(3, 3, B, commitment=5, height=1, transitions=trans, **pub)
        raws, ext, acc, public, n_acc = led.update_build(batch, B)
        assert n_acc == len(trans) and public == pub and led.root == st.root
        assert [t for t, a in zip(batch, acc) if not a][:len(rej)] == rej or len(trans) == (1 << (2 * B))
        want_raws = np.stack([_canon_rows(W.raw_values(tr, 3, 3)) for tr in circ.transitions])
        want_ext = np.stack([_canon_rows([circ.fee_token, r]) for r in W.slot_roots(circ)])
        assert (raws == want_raws).all(), np.nonzero((raws != want_raws).any(axis=2))
        assert (ext == want_ext).all()
        circs.append((circ, raws, ext))
    circ, raws, ext = circs[0]
    gw = UpdateWitnessGpu(ctx, 3, 3)
    d_in1, d_aux1 = gw.witness(circ)
    d_in2, d_aux2 = gw.witness_rows(raws, ext, circ)
    assert (d_aux1 == d_aux2).all() and (d_in1 == d_in2).all()
    gw.free(); led.free()


def test_native_per_batch_path_equals_python_path(ctx, cref):
    """transactions -> proof with only native calls in between (bzk_mpn_update_build, bzk_mpn_update_witness,
    bzk_groth16_prove_dev) gives the same 391 bytes as the Python builder + witness glue under the same key, and the
    byte-image verifier accepts it."""
    import copy
    from bazuka_b200 import groth16 as BG
    from bazuka_b200.mpn import native as N
    from bazuka_b200.mpn.ledger import NativeLedger
    from bazuka_b200.mpn.worker import MpnUpdateWorker
    st, keys = make_state(3, 3, 3)
    keys.append(N.eddsa_keys(b"newcomer"))
    txs = [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1, amount=5), transfer(keys, 0, 3, 2, amount=77, fee=3)]
    led = NativeLedger(ctx, 3, 3)
    for i, a in st.accounts.items():
        led.set_account(i, a)
    worker = MpnUpdateWorker(ctx, 3, 3, 1, cref.fr_random(77, 5))
    r, s = cref.fr_random(78, 2)
    zk_native, pub_native, accepted = worker.prove_native(led, txs, r, s, commitment=42, height=7)
    work = worker.build(st, txs, commitment=42, height=7)
    zk_py = worker.prove(work, r, s)
    assert accepted.all() and (pub_native == work.public_inputs).all() and led.root == st.root
    assert (zk_native == zk_py).all() and worker.verify(work, zk_native)
    worker.free(); led.free()


def test_worker_on_natively_compiled_circuit(ctx, cref):
    """everything native: circuit (R1CS + witness programs) compiled by C++, ledger + builder in C++, witness driver,
    prover — the 391-byte proof equals the one from the Python-defined circuit under the same toxic waste."""
    from bazuka_b200.mpn import native as N
    from bazuka_b200.mpn.ledger import NativeLedger
    from bazuka_b200.mpn.worker import MpnUpdateWorker
    st, keys = make_state(3, 3, 3)
    keys.append(N.eddsa_keys(b"newcomer"))
    txs = [transfer(keys, 0, 1, 1), transfer(keys, 1, 2, 1, amount=5), transfer(keys, 0, 3, 2, amount=77, fee=3)]
    tox, (r, s) = cref.fr_random(77, 5), cref.fr_random(78, 2)
    proofs = []
    for compiler in ("native", "python"):
        led = NativeLedger(ctx, 3, 3)
        for i, a in st.accounts.items():
            led.set_account(i, a)
        worker = MpnUpdateWorker(ctx, 3, 3, 1, tox, compiler=compiler)
        zk, pub, accepted = worker.prove_native(led, txs, r, s, commitment=42, height=7)
        assert accepted.all()
        from bazuka_b200 import groth16 as BG
        assert BG.verify_bytes(worker.vk_blob, pub, zk[4:])
        proofs.append(zk)
        worker.free(); led.free()
    assert (proofs[0] == proofs[1]).all()


def test_native_ledger_follows_the_reference_rules_like_the_python_builder(ctx):
    """csrc/mpn_host.cu against the Python restatement on the rules round 1 got wrong (/root/reference/src/mpn/update.rs):
    keys that do not decompress are filtered instead of aborting the batch (:31-38); a new receiver gets index
    `mpn_account_count + |new_account_indices|` and the map threads across batches on one fork (:47-70); `state_size`
    is tracked (:29,256-266); a fork is independent of the ledger it was cloned from (`fork_on_ram`, mod.rs:313)."""
    from bazuka_b200.mpn import native as N, update as U
    from bazuka_b200.mpn.ledger import NativeLedger
    from test_mpn_cpu import _off_curve_key
    st, keys = make_state(3, 3, 3)
    led = NativeLedger(ctx, 3, 3)
    for i, a in st.accounts.items():
        led.set_account(i, a)
    # the chain has indexed ten accounts: load an empty-but-indexed slot 9 on both sides
    st.account_count = 10
    led.set_account(9, U.MpnAccount())
    assert led.info() == {"state_hash": st.root, "state_size": st.state_size, "account_count": 10, "pending_accounts": 0}
    new1, new2 = N.eddsa_keys(b"newcomer"), N.eddsa_keys(b"second")
    keys += [new1, new2]
    bad = transfer(keys, 0, 1, 1)
    bad.dst_pub_key = _off_curve_key()
    batch1 = [bad, transfer(keys, 0, 3, 1, amount=500), transfer(keys, 1, 4, 1, amount=7)]
    fork = led.fork()
    pub_py, trans, rej = U.update(st, batch1, 1)
    raws, ext, acc, pub, n_acc = fork.update_build(batch1, 1)
    assert acc.tolist() == [False, True, True] and n_acc == 2 and pub == pub_py and rej == [bad]
    assert [t.dst_index for t in trans] == [10, 11]
    assert fork.info() == {"state_hash": st.root, "state_size": st.state_size, "account_count": 10, "pending_accounts": 2}
    assert led.info()["state_hash"] != st.root and led.info()["pending_accounts"] == 0       # the original did not move
    # second batch on the same fork: the newcomer is a sender, a third new address continues the numbering
    keys.append(N.eddsa_keys(b"third"))
    batch2 = [transfer(keys, 3, 0, 1, amount=50, fee=1), transfer(keys, 0, 5, 2, amount=1)]
    pub_py, trans, rej = U.update(st, batch2, 1)
    raws, ext, acc, pub, n_acc = fork.update_build(batch2, 1)
    assert acc.all() and pub == pub_py and [(t.src_index, t.dst_index) for t in trans] == [(10, 0), (0, 12)]
    # a fork of the ORIGINAL never saw the map: the newcomer is an unknown sender there
    other = led.fork()
    _, _, acc_o, _, n_o = other.update_build(batch2[:1], 1)
    assert n_o == 0 and not acc_o.any()
    st.commit_accounts(); fork.commit_accounts()
    assert fork.info() == {"state_hash": st.root, "state_size": st.state_size, "account_count": 13, "pending_accounts": 0}
    batch3 = [transfer(keys, 4, 3, 1, amount=2, fee=1)]
    pub_py, trans, _ = U.update(st, batch3, 1)
    _, _, acc, pub, _ = fork.update_build(batch3, 1)
    assert acc.all() and pub == pub_py and fork.info()["state_size"] == st.state_size
    # set_account: a slot whose token id is zero is dropped; overwriting an account releases its address
    z = NativeLedger(ctx, 3, 3)
    a0 = U.MpnAccount(0, 0, keys[0][0], {0: U.Money(U.ZIESHA, 5), 1: U.Money(0, 9)})
    z.set_account(0, a0)
    ref = U.MpnState(3, 3)
    ref.set(0, U.MpnAccount(0, 0, keys[0][0], {0: U.Money(U.ZIESHA, 5)}))
    assert z.root == ref.root and z.info()["state_size"] == ref.state_size
    z.set_account(0, U.MpnAccount(0, 0, keys[1][0], {0: U.Money(U.ZIESHA, 5)}))
    _, _, acc, _, _ = z.update_build([transfer(keys, 0, 1, 1, amount=1)], 0)     # key 0 no longer owns an account
    assert not acc.any()
    for l in (led, fork, other, z):
        l.free()


def test_update_witness_rejects_rows_of_the_wrong_shape(ctx):
    """bzk_mpn_update_witness checks its arguments against the uploaded programs (a program compiled for another (A, T, B)
    would otherwise read and write out of bounds): wrong n_raw / slot_vars / slot count -> BZK_ERR_BAD_ARG."""
    import ctypes as ct
    import torch
    from bazuka_b200.api import _dev_ptr, _host_ptr
    from bazuka_b200.mpn import witness_program as W
    from bazuka_b200.mpn.gpu_witness import UpdateWitnessGpu, upload_program, _canon_rows
    gw = UpdateWitnessGpu(ctx, 3, 3)
    p = gw.prog
    ep = W.compile_update_epilogue(p, 1)
    eh = upload_program(ctx, ep)
    n = 4
    raws, ext = np.zeros((n, p.n_raw, 4), np.uint64), np.zeros((n, 2, 4), np.uint64)
    pro = _canon_rows([0, 0, 0, 1, 0, 0])
    d_in = torch.empty((6, 4), dtype=torch.int64, device="cuda")
    d_aux = torch.empty((p.p_aux + n * p.n_ops + ep.n_ops, 4), dtype=torch.int64, device="cuda")

    def call(n_slots, slot_vars, epi_vars, n_raw):
        return ctx._l.bzk_mpn_update_witness(ctx._h, gw._h, eh, n_slots, 3, slot_vars, epi_vars, _host_ptr(raws), _host_ptr(ext), n_raw,
                                             _host_ptr(pro), _dev_ptr(d_in), _dev_ptr(d_aux))
    assert call(n, p.n_ops, ep.n_ops, p.n_raw) == 0
    assert call(n, p.n_ops, ep.n_ops, p.n_raw - 1) == -1
    assert call(n, p.n_ops + 1, ep.n_ops, p.n_raw) == -1
    assert call(n, p.n_ops, ep.n_ops + 1, p.n_raw) == -1
    assert call(16, p.n_ops, ep.n_ops, p.n_raw) == -1          # the epilogue was compiled for 4 slots
    ctx.synchronize()
    ctx._l.bzk_witness_program_free(ctx._h, eh)
    gw.free()


def test_worker_protocol_end_to_end_against_an_in_process_node(ctx, cref):
    """a block's worth of traffic -> `prepare_works` (deposit, withdraw, update on one fork) -> bincode `GetMpnWorkResponse`
    -> the worker proves every work from its WIRE image on the GPU -> bincode `PostMpnSolutionRequest` -> the node side runs
    `MpnWork::verify` (commitment from (prover, reward) + check_proof) and counts the accepted proofs; a proof posted under
    another prover address is refused (/root/reference/src/mpn/mod.rs:108-129,281-295; client/messages.rs:368-397)."""
    from bazuka_b200.mpn import wire as Wr, works as Wk
    from bazuka_b200.mpn.worker import MpnDepositWithdrawWorker, MpnUpdateWorker
    from test_wire_cpu import _scenario
    st, keys, deposits, withdraws, wpay, updates = _scenario()
    A, T, B = 3, 3, 1
    wu = MpnUpdateWorker(ctx, A, T, B, cref.fr_random(301, 5))
    wd = MpnDepositWithdrawWorker(ctx, "deposit", A, T, B, cref.fr_random(302, 5))
    ww = MpnDepositWithdrawWorker(ctx, "withdraw", A, T, B, cref.fr_random(303, 5))
    config = {"log4_tree_size": A, "log4_token_tree_size": T, "log4_deposit_batch_size": B, "log4_withdraw_batch_size": B, "log4_update_batch_size": B,
              "mpn_contract_id": 0x1234, "mpn_num_update_batches": 1, "mpn_num_deposit_batches": 1, "mpn_num_withdraw_batches": 1,
              "deposit_vk": bytes(wd.vk_blob), "withdraw_vk": bytes(ww.vk_blob), "update_vk": bytes(wu.vk_blob)}
    works, fork = Wk.prepare_works(config, st, deposits, withdraws, updates, {"deposit": 11, "withdraw": 22, "update": 33}, height=9,
                                   withdraw_payments=wpay)
    me, other = bytes(range(32)), bytes(range(1, 33))
    served, accepted_log = [], []

    def node(method, url, body):   # the three endpoints of /root/reference/src/node/mod.rs:393-413
        served.append((method, url.rsplit("/", 1)[1]))
        if url.endswith("/bincode/mpn/worker"):
            return b"\x01"
        if url.endswith("/bincode/mpn/work"):
            assert Wr.dec_address(Wr.Reader(body)) == me
            return Wr.get_mpn_work_response_to_bytes(works)
        prover, proofs = Wr.post_mpn_solution_request_from_bytes(body)
        ok = sum(1 for wid, p in proofs.items() if Wk.verify_work(works[wid], prover, np.frombuffer(p, dtype=np.uint8)))
        accepted_log.append((prover, ok))
        return (ok).to_bytes(8, "little")

    prover = Wk.MpnProver(ctx)
    prover.add_circuit("update", wu.prover, wu.pk, wu.witness)
    prover.add_circuit("deposit", wd.prover, wd.pk, wd.witness)
    prover.add_circuit("withdraw", ww.prover, ww.pk, ww.witness)
    client = Wk.WorkerClient("127.0.0.1:8765", me, prover, opener=node)
    assert client.register()
    seeds = iter(range(400, 500))
    n_works, n_ok = client.run_once(lambda: tuple(cref.fr_random(next(seeds), 2)))
    assert (n_works, n_ok) == (3, 3) and served == [("POST", "worker"), ("GET", "work"), ("POST", "solution")]
    # the commitment binds the proof to the prover: the same proof under another address is not accepted
    r, s = cref.fr_random(77, 2)
    p = prover.prove(works[2], me, r, s)
    assert Wk.verify_work(works[2], me, np.frombuffer(p, dtype=np.uint8))
    assert not Wk.verify_work(works[2], other, np.frombuffer(p, dtype=np.uint8))
    assert not Wk.verify_work(dict(works[2], reward=34), me, np.frombuffer(p, dtype=np.uint8))
    for w in (wu, wd, ww):
        w.free()


def test_prepare_works_with_the_batched_gpu_builders_equals_the_sequential_ones(ctx):
    """`prepare_works` over batch_update.{deposit,withdraw,update}_batched (all hashing in batched GPU launches) produces the
    same works, byte for byte, and the same final ledger as over the sequential builders."""
    from bazuka_b200.mpn import batch_update as BU, wire as Wr, works as Wk
    from test_wire_cpu import _scenario, _config
    st, keys, deposits, withdraws, wpay, updates = _scenario()
    h = BU.GpuTreeHasher(ctx)
    rewards = {"deposit": 11, "withdraw": 22, "update": 33}
    seq, fork_s = Wk.prepare_works(_config(), st, deposits, withdraws, updates, rewards, height=9, withdraw_payments=wpay)
    gpu, fork_g = Wk.prepare_works(_config(), st, deposits, withdraws, updates, rewards, height=9, withdraw_payments=wpay,
                                   builders=(lambda s, d, b: BU.deposit_batched(h, s, d, b), lambda s, w, b: BU.withdraw_batched(h, s, w, b),
                                             lambda s, u, b: BU.update_batched(h, s, u, b)))
    assert [Wr.work_to_bytes(gpu[i]) for i in range(3)] == [Wr.work_to_bytes(seq[i]) for i in range(3)]
    assert fork_g.compressed == fork_s.compressed and fork_g.new_account_indices == fork_s.new_account_indices
    assert Wk.final_delta(st, fork_g) == Wk.final_delta(st, fork_s)
