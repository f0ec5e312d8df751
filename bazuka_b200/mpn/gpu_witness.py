"""UpdateCircuit witness on the GPU: host glue around csrc/witness.cu.

    gw = UpdateWitnessGpu(ctx, A, T)               # compile + upload the slot program once per (A, T)
    d_inputs, d_aux = gw.witness(circ)             # CUDA tensors [ni,4] / [na,4] (Montgomery), z = inputs ++ aux
    prover.prove_dev(pk, d_inputs, d_aux, r, s)

What stays on the host is what bellman's synthesize does OUTSIDE the per-slot loop
(/root/reference/src/mpn/circuits/update_circuit.rs:52-80 and 470-493): the six prologue variables and the
epilogue's one Poseidon(fee_token, fee_sum) gadget — a few hundred values."""
import ctypes as ct

import numpy as np

from . import native as N
from . import witness_program as W
from .cs import ConstraintSystem, AllocatedNum, R, to_mont
from .gadgets import Number

_RINV = pow(1 << 256, -1, R)


def _canon_rows(values):
    buf = b"".join((v % R).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def upload_program(ctx, p):
    from ..api import _host_ptr
    ops = np.ascontiguousarray(p.ops, dtype=np.int32)
    coefs = np.ascontiguousarray(p.coefs_mont())
    jj_d = to_mont([N.JJ_D])
    h = ct.c_void_p()
    ctx._check(ctx._l.bzk_witness_program_upload(
        ctx._h, _host_ptr(ops), len(ops), _host_ptr(p.lc_ptr), len(p.lc_ptr) - 1, _host_ptr(p.lc_slot), _host_ptr(p.lc_coef),
        len(p.lc_slot), _host_ptr(coefs), len(coefs), p.n_raw, p.n_ext, _host_ptr(jj_d), ct.byref(h)))
    return h


class UpdateWitnessGpu:
    def __init__(self, ctx, A, T, prog=None, epilogues=None):
        """prog / epilogues ({log4_batch: program}): ready-made programs, e.g. from the native compiler
        (mpn/native_circuit.py); by default the Python definition is compiled here."""
        self.ctx, self.A, self.T = ctx, A, T
        self.prog = prog if prog is not None else W.compile_update_block(A, T)
        self._h = upload_program(ctx, self.prog)
        self._epi = {b: (ep, upload_program(ctx, ep)) for b, ep in (epilogues or {}).items()}   # log4_batch -> (program, handle)

    def free(self):
        if self._h:
            self.ctx._l.bzk_witness_program_free(self.ctx._h, self._h)
            self._h = None
        for _, h in self._epi.values():
            self.ctx._l.bzk_witness_program_free(self.ctx._h, h)
        self._epi = {}

    def witness_native(self, raws, ext, prologue, log4_batch):
        """the whole batch witness through bzk_mpn_update_witness (no Python between the builder's rows and z):
        raws [4^B, n_raw, 4], ext [4^B, 2, 4] canonical (mpn/ledger.py), prologue = [commitment, height, state,
        fee_token, aux_data, next_state] as ints.  -> (d_inputs, d_aux) CUDA tensors."""
        import torch
        from ..api import _dev_ptr, _host_ptr
        ctx, p = self.ctx, self.prog
        if log4_batch not in self._epi:
            ep = W.compile_update_epilogue(p, log4_batch)
            self._epi[log4_batch] = (ep, upload_program(ctx, ep))
        ep, eh = self._epi[log4_batch]
        n = 1 << (2 * log4_batch)
        raws = np.ascontiguousarray(raws, dtype=np.uint64).reshape(n, p.n_raw, 4)
        ext = np.ascontiguousarray(ext, dtype=np.uint64).reshape(n, 2, 4)
        pro = _canon_rows(prologue)
        dev = torch.device("cuda", ctx.device)
        d_in = torch.empty((6, 4), dtype=torch.int64, device=dev)
        d_aux = torch.empty((p.p_aux + n * p.n_ops + ep.n_ops, 4), dtype=torch.int64, device=dev)
        ctx._check(ctx._l.bzk_mpn_update_witness(ctx._h, self._h, eh, n, self.T, p.n_ops, ep.n_ops, _host_ptr(raws), _host_ptr(ext), p.n_raw,
                                                 _host_ptr(pro), _dev_ptr(d_in), _dev_ptr(d_aux)))
        return d_in, d_aux

    def witness(self, circ):
        """-> (d_inputs [ni,4], d_aux [na,4]) int64 CUDA tensors holding Montgomery images."""
        assert (circ.A, circ.T) == (self.A, self.T)
        raws = np.concatenate([_canon_rows(W.raw_values(tr, self.A, self.T)) for tr in circ.transitions])
        ext = _canon_rows([v for root in W.slot_roots(circ) for v in (circ.fee_token, root)])
        return self.witness_rows(raws, ext, circ)

    def witness_rows(self, raws, ext, circ):
        """same from ready-made rows (e.g. the native builder's, mpn/ledger.py): raws [n, n_raw, 4], ext [n, 2, 4]
        canonical; `circ` only supplies the six public / prologue values (its transitions are not read)."""
        import torch
        from ..api import _dev_ptr, _host_ptr
        p, ctx = self.prog, self.ctx
        raws = np.ascontiguousarray(raws, dtype=np.uint64).reshape(-1, 4)
        ext = np.ascontiguousarray(ext, dtype=np.uint64).reshape(-1, 4)
        n, a_tx = len(ext) // 2, self.prog.n_ops
        assert len(raws) == n * p.n_raw
        # ---- prologue on the host
        cs = ConstraintSystem()
        state_wit, fee_tok, aux_wit, claimed = circ._prologue(cs)
        assert len(cs.aux) == p.p_aux
        # ---- slots on the device
        # epilogue size is value-independent: synthesise it once with placeholders to learn it
        probe = ConstraintSystem()
        circ._epilogue(probe, AllocatedNum(probe.alloc(0), 0), AllocatedNum(probe.alloc(0), 0), AllocatedNum(probe.alloc(0), 0),
                       AllocatedNum(probe.alloc(0), 0), Number.zero())
        n_epi = len(probe.aux) - 4
        na = p.p_aux + n * a_tx + n_epi
        dev = torch.device("cuda", ctx.device) if hasattr(ctx, "device") else torch.device("cuda")
        d_aux = torch.empty((na, 4), dtype=torch.int64, device=dev)
        block = d_aux[p.p_aux:p.p_aux + n * a_tx]
        ctx._check(ctx._l.bzk_witness_run_dev(ctx._h, self._h, _host_ptr(raws), _host_ptr(ext), n, _dev_ptr(block)))
        # ---- epilogue on the host: needs the fee sum and the last state root (read back: n + 1 elements)
        idx = torch.tensor([k * a_tx + p.final_fee for k in range(n)] + [(n - 1) * a_tx + p.state_out], device=dev)
        back = block[idx].cpu().numpy().view(np.uint64)
        vals = [int.from_bytes(row.tobytes(), "little") * _RINV % R for row in back]
        fee_sum = Number.zero()
        for k in range(n):
            fee_sum = Number(fee_sum.lc.add_term(1, 2 * (p.p_aux + k * a_tx + p.final_fee) + 1), fee_sum.value + vals[k])
        cs.aux.extend([0] * (n * a_tx))
        last_state = AllocatedNum(2 * (p.p_aux + (n - 1) * a_tx + p.state_out) + 1, vals[n])
        circ._epilogue(cs, last_state, fee_tok, aux_wit, claimed, fee_sum)
        epi = cs.aux[p.p_aux + n * a_tx:]
        assert len(epi) == n_epi
        d_aux[:p.p_aux] = torch.from_numpy(to_mont(cs.aux[:p.p_aux]).view(np.int64)).to(dev)
        d_aux[p.p_aux + n * a_tx:] = torch.from_numpy(to_mont(epi).view(np.int64)).to(dev)
        d_inputs = torch.from_numpy(to_mont(cs.inputs).view(np.int64)).to(dev)
        torch.cuda.current_stream(dev).synchronize()  # stream-level: other contexts (the prover) keep running
        return d_inputs, d_aux
