"""GPU tier, last file of the tier: the external prover as ONE native call — `bincode(MpnWork)` in, `ZkProof::Groth16` out
(csrc/mpn_prover.cu, csrc/mpn_wire.cu).  Written after this round's GPU budget was spent: the composition is exercised in the
CPU tier over the host build (tests/test_wire_native_cpu.py: work bytes -> rows -> witness drivers -> a satisfying assignment of
the natively compiled circuit), every native call it strings together has its own GPU test in tests/test_gpu_mpn.py; this file
is the same composition over the real kernels and has not been run on a GPU by its author."""
import ctypes as ct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_native_prover_turns_wire_works_into_accepted_proofs(ctx, cref):
    """three works of one block (deposit, withdraw, update; `prepare_works` on one fork) as bincode -> bzk_mpn_prover_prove_work
    per work -> the 391 bytes equal the Python prover's for the same key and (r, s); `MpnWork::verify` (native, from the work's
    own verifying key and the commitment of (prover, reward)) accepts them for the address they were made for and for no other."""
    from bazuka_b200.mpn import wire as Wr, works as Wk
    from bazuka_b200.mpn.native_circuit import NativeTwoPhaseCircuit, NativeUpdateCircuit
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
    works, _ = Wk.prepare_works(config, st, deposits, withdraws, updates, {"deposit": 11, "withdraw": 22, "update": 33}, height=9, withdraw_payments=wpay)
    py = Wk.MpnProver(ctx)
    py.add_circuit("update", wu.prover, wu.pk, wu.witness)
    py.add_circuit("deposit", wd.prover, wd.pk, wd.witness)
    py.add_circuit("withdraw", ww.prover, ww.pk, ww.witness)
    circuits = {"update": NativeUpdateCircuit(A, T, B), "deposit": NativeTwoPhaseCircuit("deposit", A, T, B), "withdraw": NativeTwoPhaseCircuit("withdraw", A, T, B)}
    nat = Wk.NativeMpnProver(ctx)
    for kind, w in (("update", wu), ("deposit", wd), ("withdraw", ww)):
        nat.add_circuit(kind, circuits[kind], w.pk)
        circuits[kind].free()
    me, other = bytes(range(32)), bytes(range(1, 33))
    lib = ctx._l
    for wid, work in works.items():
        blob = Wr.work_to_bytes(work)
        r, s = cref.fr_random(500 + wid, 2)
        zk = nat.prove(blob, me, r, s)
        assert len(zk) == 391 and zk[:4] == bytes(4)
        assert zk[4:] == py.prove(work, me, r, s)                       # same key, same (r, s), same witness -> same bytes
        h = ct.c_void_p()
        assert lib.bzk_mpn_work_decode(blob, len(blob), ct.byref(h), None) == 0
        assert lib.bzk_mpn_work_verify(h, me, zk[4:]) == 1              # `MpnWork::verify`
        assert lib.bzk_mpn_work_verify(h, other, zk[4:]) == 0           # the commitment binds the proof to its prover
        lib.bzk_mpn_work_free(h)
        assert Wk.verify_work(work, me, np.frombuffer(zk[4:], dtype=np.uint8))
        # a work whose claimed end state does not follow from its transitions is refused, not proved
        bad = Wr.work_to_bytes(dict(work, public_inputs=dict(work["public_inputs"], next_state=work["public_inputs"]["next_state"] + 1)))
        with pytest.raises(Exception):
            nat.prove(bad, me, r, s)
    nat.free()
    for w in (wu, wd, ww):
        w.free()


def test_native_prepare_works_on_the_gpu_builders(ctx):
    """bzk_mpn_prepare_works over the real kernels (batched Poseidon, versioned tree update): the same block as the CPU tier's
    case — GetMpnWorkResponse image equal to works.prepare_works' byte for byte, forks agree, several batches continue."""
    from test_wire_native_cpu import prepare_works_case
    prepare_works_case(ctx)
