"""The prover worker's job as one object: transactions in, `ZkProof::Groth16` blob out, all heavy steps on the GPU.

In the reference a validator turns a batch into an `MpnWork` (`prepare_works`, /root/reference/src/mpn/mod.rs:298-424:
`update()` builds the transitions and public inputs), an external worker proves it (bellman `create_proof` on the
`UpdateCircuit`) and posts a `ZkProof` that `MpnWork::verify` checks (`mod.rs:281-295` -> `check_proof`,
/root/reference/src/zk/mod.rs:157-193).  `MpnUpdateWorker` is that path for one circuit shape (A, T, B):

    worker = MpnUpdateWorker(ctx, A, T, B, toxic)        # once: R1CS template, proving key, witness program
    work   = worker.build(state, txs, commitment, height)  # batched transition builder  (batch_update.py)
    blob   = worker.prove(work, r, s)                      # GPU witness -> resident z -> proof (391 bytes)
    worker.verify(work, blob)                              # the validator's check_proof

The trusted setup is bellman's `generate_parameters` with explicit toxic waste (the reference's production keys
come from an external ceremony; /root/reference/src/config/blockchain.rs:372-400 is its test-key path)."""
from dataclasses import dataclass

import numpy as np

from .. import groth16 as BG
from . import batch_update as BU
from . import fastsynth as FS
from . import update as U
from .cs import to_mont
from .gpu_witness import UpdateWitnessGpu


@dataclass
class UpdateWork:
    circuit: U.UpdateCircuit
    public_inputs: np.ndarray      # [5,4] Montgomery: commitment, height, state, aux_data, next_state
    accepted: int
    rejected: list


class MpnUpdateWorker:
    def __init__(self, ctx, A, T, B, toxic, g1=None, g2=None, compiler="python"):
        """compiler: "python" = R1CS template and witness programs from the Python circuit definition;
        "native" = from libbzk's C++ definition (csrc/mpn_circuit.cu) — the emitted arrays are identical
        (tests/test_mpn_cpu.py::test_native_circuit_compiler_equals_python)."""
        self.ctx, self.A, self.T, self.B = ctx, A, T, B
        prog = epilogues = None
        if compiler == "native":
            from .native_circuit import NativeUpdateCircuit
            nc = NativeUpdateCircuit(A, T, B)
            ni, na, mats = nc.r1cs()
            prog, epilogues = nc.program(0), {B: nc.program(1)}
            nc.free()
        else:
            shape = U.UpdateCircuit(A, T, B, fee_token=U.ZIESHA)   # all-null batch: the R1CS does not depend on values
            ni, na, mats, _, _ = FS.synthesize_update(shape, structure_only=True)
        self.prover = BG.Prover(ctx, BG.R1CS(ni, na, *mats))
        self.pk, self.vk = BG.setup_gpu(ctx, self.prover.r1cs, toxic, BG.G1_GENERATOR if g1 is None else g1, BG.G2_GENERATOR if g2 is None else g2)
        self.vk_blob = BG.vk_to_bincode(self.vk)
        self.witness = UpdateWitnessGpu(ctx, A, T, prog, epilogues)
        self.hasher = BU.GpuTreeHasher(ctx)

    def build(self, state, txs, commitment=0, height=0, fee_token=U.ZIESHA) -> UpdateWork:
        pub, trans, rejected = BU.update_batched(self.hasher, state, txs, self.B, fee_token)
        circ = U.UpdateCircuit(self.A, self.T, self.B, commitment=commitment, height=height, fee_token=fee_token, transitions=trans, **pub)
        inputs = to_mont([commitment, height, pub["state"], pub["aux_data"], pub["next_state"]])
        return UpdateWork(circ, inputs, len(trans), rejected)

    def prove(self, work: UpdateWork, r, s, check_satisfied=True):
        """-> 391-byte bincode image of `ZkProof::Groth16` for the work."""
        d_in, d_aux = self.witness.witness(work.circuit)
        blob, _ = self.prover.prove_dev(self.pk, d_in, d_aux, r, s, check_satisfied=check_satisfied)
        return BG.zkproof_blob(blob)

    def prove_native(self, ledger, txs, r, s, commitment=0, height=0, fee_token=U.ZIESHA, check_satisfied=True):
        """the per-batch path with nothing but native calls between the transactions and the proof:
        bzk_mpn_update_build (C++ ledger + batched GPU hashing) -> bzk_mpn_update_witness (slot + epilogue
        programs, z resident) -> bzk_groth16_prove_dev.  `ledger` = mpn.ledger.NativeLedger.
        -> (391-byte ZkProof image, public inputs [5,4] Montgomery, accepted mask)."""
        raws, ext, accepted, pub, _ = ledger.update_build(txs, self.B, fee_token)
        d_in, d_aux = self.witness.witness_native(raws, ext, [commitment, height, pub["state"], fee_token, pub["aux_data"], pub["next_state"]], self.B)
        blob, _ = self.prover.prove_dev(self.pk, d_in, d_aux, r, s, check_satisfied=check_satisfied)
        return BG.zkproof_blob(blob), to_mont([commitment, height, pub["state"], pub["aux_data"], pub["next_state"]]), accepted

    def verify(self, work: UpdateWork, zkproof) -> bool:
        """`check_proof(vk, commitment, height, state, aux_data, next_state, proof)` on the reference's byte images"""
        z = np.asarray(zkproof, dtype=np.uint8)
        if z.size != 391 or z[:4].any():
            return False
        return BG.verify_bytes(self.vk_blob, work.public_inputs, z[4:])

    def free(self):
        self.witness.free(); self.pk.free(); self.prover.free()


class MpnDepositWithdrawWorker:
    """the same worker for the deposit / withdraw circuits (`kind` = "deposit" | "withdraw"): batched builder
    (batch_update.deposit_batched / withdraw_batched), two-phase GPU witness (dw_witness.py), resident prover.
    The R1CS comes from one host synthesis of the all-null batch (value-independent; one-off per shape)."""

    def __init__(self, ctx, kind, A, T, B, toxic, g1=None, g2=None):
        from . import dw as D
        from .cs import ConstraintSystem
        from .dw_witness import TwoPhaseWitnessGpu
        self.ctx, self.kind, self.A, self.T, self.B = ctx, kind, A, T, B
        self.circ_cls = D.DepositCircuit if kind == "deposit" else D.WithdrawCircuit
        ni, na, mats, _, _ = self.circ_cls(A, T, B).synthesize(ConstraintSystem()).to_csr()
        self.prover = BG.Prover(ctx, BG.R1CS(ni, na, *mats))
        self.pk, self.vk = BG.setup_gpu(ctx, self.prover.r1cs, toxic, BG.G1_GENERATOR if g1 is None else g1, BG.G2_GENERATOR if g2 is None else g2)
        self.vk_blob = BG.vk_to_bincode(self.vk)
        self.witness = TwoPhaseWitnessGpu(ctx, kind, A, T)
        self.hasher = BU.GpuTreeHasher(ctx)

    def build(self, state, items, commitment=0, height=0) -> UpdateWork:
        fn = BU.deposit_batched if self.kind == "deposit" else BU.withdraw_batched
        pub, trans = fn(self.hasher, state, items, self.B)
        circ = self.circ_cls(self.A, self.T, self.B, commitment=commitment, height=height, transitions=trans, **pub)
        inputs = to_mont([commitment, height, pub["state"], pub["aux_data"], pub["next_state"]])
        return UpdateWork(circ, inputs, len(trans), [])

    def prove(self, work: UpdateWork, r, s, check_satisfied=True):
        d_in, d_aux = self.witness.witness(work.circuit)
        blob, _ = self.prover.prove_dev(self.pk, d_in, d_aux, r, s, check_satisfied=check_satisfied)
        return BG.zkproof_blob(blob)

    def verify(self, work: UpdateWork, zkproof) -> bool:
        z = np.asarray(zkproof, dtype=np.uint8)
        if z.size != 391 or z[:4].any():
            return False
        return BG.verify_bytes(self.vk_blob, work.public_inputs, z[4:])

    def free(self):
        self.witness.free(); self.pk.free(); self.prover.free()
