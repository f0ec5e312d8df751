"""Deposit / Withdraw circuit witnesses on the GPU (the same slot-program interpreter as the update circuit).

These circuits walk the batch twice (/root/reference/src/mpn/circuits/deposit_circuit.rs:47-293,
withdraw_circuit.rs:49-413): a first loop allocates each transaction's fields and its calldata hash, then the
batch is `reveal`ed into one root (/root/reference/src/zk/groth16/gadgets/reveal/mod.rs:13-64), then a second
loop applies the transactions to the state.  So there are TWO slot programs — phase 1 and phase 2, the
second reading phase-1 variables of its own slot (all of them raw transaction fields) plus the state root
entering the slot as externals — and the reveal tree (85 Poseidon gadgets for 64 slots), which depends on every
slot, is evaluated on the host between the two launches.

aux layout = synthesize's: [5 public-input copies][phase 1 x n][reveal][phase 2 x n]."""
import ctypes as ct

import numpy as np

from . import dw as D
from . import native as N
from . import witness_program as W
from .cs import ConstraintSystem, AllocatedNum, R, to_mont
from .fastsynth import FAKE_STATE_VAR
from .gadgets import Number
from .gpu_witness import _canon_rows

_flat = lambda proof: [s for level in proof for s in level]


def deposit_raws(tr, A, T):
    pk = N.jj_decompress(tr.tx.mpn_address)
    p1 = [1 if tr.enabled else 0, tr.tx.token_id, tr.tx.amount, pk[0], pk[1]]
    p2 = [tr.account_index, tr.token_index, tr.before.tx_nonce, tr.before.withdraw_nonce, tr.before.address[0], tr.before.address[1],
          tr.before_balances_hash, tr.before_balance.token_id, tr.before_balance.amount] + _flat(tr.balance_proof) + _flat(tr.proof)
    return [v % R for v in p1], [v % R for v in p2]


def withdraw_raws(tr, A, T):
    pk = N.jj_decompress(tr.tx.mpn_address)
    p1 = [1 if tr.enabled else 0, tr.tx.amount.token_id, tr.tx.amount.amount, tr.tx.fee.token_id, tr.tx.fee.amount,
          tr.tx.fingerprint if tr.enabled else 0, pk[0], pk[1], tr.tx.mpn_withdraw_nonce, tr.tx.mpn_sig["r"][0], tr.tx.mpn_sig["r"][1],
          tr.tx.mpn_sig["s"]]
    p2 = ([tr.account_index, tr.token_index, tr.fee_token_index, tr.before.tx_nonce, tr.before.withdraw_nonce, tr.before.address[0],
           tr.before.address[1], tr.before_token_hash, tr.before_token_balance.token_id, tr.before_token_balance.amount]
          + _flat(tr.token_balance_proof) + [tr.before_fee_balance.token_id, tr.before_fee_balance.amount] + _flat(tr.fee_balance_proof)
          + _flat(tr.proof))
    return [v % R for v in p1], [v % R for v in p2]


KINDS = {"deposit": (D.DepositCircuit, D.DepositTransition, deposit_raws),
         "withdraw": (D.WithdrawCircuit, D.WithdrawTransition, withdraw_raws)}


class TwoPhasePrograms:
    """the two compiled slot programs of a deposit / withdraw circuit plus the bookkeeping to place them"""

    def __init__(self, kind, A, T):
        circ_cls, tr_cls, self.raws_of = KINDS[kind]
        self.kind, self.A, self.T = kind, A, T
        tr = tr_cls.null(A, T)
        circ = circ_cls(A, T, 0, transitions=[tr])
        cs = ConstraintSystem(record=True)
        D._public_inputs(cs, circ)
        self.p_aux = p0 = len(cs.aux)
        wits, row = circ._phase1(cs, tr)
        self.n1 = len(cs.aux) - p0
        # the revealed row: each entry is one phase-1 variable; where it sits in the block
        self.row_local = []
        for num in row:
            (var, coef), = num.lc.t.items()
            assert coef == 1 and var % 2 == 1
            self.row_local.append((var >> 1) - p0)
        start2 = len(cs.aux)
        state_out = circ._phase2(cs, tr, wits, AllocatedNum(FAKE_STATE_VAR, 0))
        self.n2 = len(cs.aux) - start2
        rec1, rec2 = cs.recipes[p0:p0 + self.n1], cs.recipes[start2:]
        self.prog1 = W.compile_block(rec1, p0, [])
        # externals of phase 2: phase-1 variables of the same slot (must be raw fields) and the entering state
        ext = []
        for rec in rec2:
            for lc in rec[1:]:
                if hasattr(lc, "t"):
                    for v in lc.t:
                        if v != 0 and not (v % 2 == 1 and start2 <= (v >> 1) < start2 + self.n2) and v not in ext:
                            ext.append(v)
        raw_index, k = {}, 0
        for j, rec in enumerate(rec1):
            if rec[0] == "raw":
                raw_index[2 * (p0 + j) + 1] = k
                k += 1
        self.ext_src = []  # per external: ('raw1', k) or ('state',)
        for v in ext:
            if v == FAKE_STATE_VAR:
                self.ext_src.append(("state",))
            else:
                assert v in raw_index, "phase 2 reads a derived phase-1 variable"
                self.ext_src.append(("raw1", raw_index[v]))
        self.prog2 = W.compile_block(rec2, start2, ext)
        self.state_out = (state_out.var >> 1) - start2
        r1, r2 = self.raws_of(tr, A, T)
        assert (len(r1), len(r2)) == (self.prog1.n_raw, self.prog2.n_raw)

    def ext_values(self, raws1, state_in):
        return [state_in if s[0] == "state" else raws1[s[1]] for s in self.ext_src]


def slot_roots(circ):
    """state root entering each slot: pre_root for real transitions, the final state for the null padding"""
    last = max([k for k, t in enumerate(circ.transitions) if t.enabled], default=-1)
    out = []
    for k, t in enumerate(circ.transitions):
        if t.enabled:
            out.append(t.pre_root)
        elif k > last:
            out.append(circ.next_state if last >= 0 else circ.state)
        else:
            out.append(next(x for x in circ.transitions[k:] if x.enabled).pre_root)
    return out


def reveal_rows_native(kind, circ):
    """values of each slot's revealed row (the calldata hash needs one native Poseidon per enabled slot)"""
    rows = []
    for tr in circ.transitions:
        if kind == "deposit":
            pk = N.jj_decompress(tr.tx.mpn_address)
            cd = N.poseidon([pk[0], pk[1]]) if tr.enabled else 0
            rows.append([1 if tr.enabled else 0, tr.tx.token_id, tr.tx.amount, cd])
        else:
            w = tr.tx
            pk = N.jj_decompress(w.mpn_address)
            cd = N.poseidon([pk[0], pk[1], w.mpn_withdraw_nonce, w.mpn_sig["r"][0], w.mpn_sig["r"][1], w.mpn_sig["s"]]) if tr.enabled else 0
            rows.append([1 if tr.enabled else 0, w.amount.token_id, w.amount.amount, w.fee.token_id, w.fee.amount,
                         w.fingerprint if tr.enabled else 0, cd])
    return rows


def host_parts(progs: TwoPhasePrograms, circ):
    """-> (inputs, prologue aux, reveal aux) as canonical ints; the reveal is synthesised on placeholders for
    the slot variables (only its own allocations are kept)."""
    n = len(circ.transitions)
    cs = ConstraintSystem()
    D._public_inputs(cs, circ)
    p0 = len(cs.aux)
    cs.aux.extend([0] * (n * progs.n1))
    rows = []
    for k, vals in enumerate(reveal_rows_native(progs.kind, circ)):
        rows.append([Number.of(AllocatedNum(2 * (p0 + k * progs.n1 + j) + 1, v)) for j, v in zip(progs.row_local, vals)])
    D.reveal_list_of_structs(cs, circ.B, rows)
    return cs.inputs, cs.aux[:p0], cs.aux[p0 + n * progs.n1:]


class TwoPhaseWitnessGpu:
    def __init__(self, ctx, kind, A, T):
        from .gpu_witness import upload_program
        self.ctx, self.progs = ctx, TwoPhasePrograms(kind, A, T)
        self._h = [upload_program(ctx, p) for p in (self.progs.prog1, self.progs.prog2)]

    def free(self):
        for h in self._h:
            self.ctx._l.bzk_witness_program_free(self.ctx._h, h)
        self._h = []

    def witness(self, circ):
        """-> (d_inputs [ni,4], d_aux [na,4]) int64 CUDA tensors (Montgomery), z = inputs ++ aux of `synthesize`."""
        import torch
        from ..api import _dev_ptr, _host_ptr
        pg, ctx = self.progs, self.ctx
        n = len(circ.transitions)
        raws = [pg.raws_of(tr, pg.A, pg.T) for tr in circ.transitions]
        roots = slot_roots(circ)
        inputs, pro, rev = host_parts(pg, circ)
        p0, n_rev = len(pro), len(rev)
        na = p0 + n * pg.n1 + n_rev + n * pg.n2
        dev = torch.device("cuda", ctx.device)
        d_aux = torch.empty((na, 4), dtype=torch.int64, device=dev)
        r1 = _canon_rows([v for a, _ in raws for v in a])
        ctx._check(ctx._l.bzk_witness_run_dev(ctx._h, self._h[0], _host_ptr(r1), None, n, _dev_ptr(d_aux[p0:p0 + n * pg.n1])))
        r2 = _canon_rows([v for _, b in raws for v in b])
        ext = _canon_rows([v for k in range(n) for v in pg.ext_values(raws[k][0], roots[k])])
        o2 = p0 + n * pg.n1 + n_rev
        ctx._check(ctx._l.bzk_witness_run_dev(ctx._h, self._h[1], _host_ptr(r2), _host_ptr(ext), n, _dev_ptr(d_aux[o2:])))
        d_aux[:p0] = torch.from_numpy(to_mont(pro).view(np.int64)).to(dev)
        d_aux[p0 + n * pg.n1:o2] = torch.from_numpy(to_mont(rev).view(np.int64)).to(dev)
        d_inputs = torch.from_numpy(to_mont(inputs).view(np.int64)).to(dev)
        torch.cuda.current_stream(dev).synchronize()
        return d_inputs, d_aux


def compile_reveal_program(progs: TwoPhasePrograms, B):
    """the `reveal` of the batch (reveal/mod.rs:13-64: one Poseidon per row, then a 4-ary Poseidon tree) as a
    witness program of ONE instance whose externals are every slot's revealed row, slot-major — the piece that lets
    the whole two-phase witness run through the interpreter (host_parts() evaluates it with the Python gadget)."""
    circ_cls = KINDS[progs.kind][0]
    n = 1 << (2 * B)
    circ = circ_cls(progs.A, progs.T, B)
    cs = ConstraintSystem(record=True)
    D._public_inputs(cs, circ)
    p0 = len(cs.aux)
    cs.aux.extend([0] * (n * progs.n1))
    cs.recipes.extend([("raw",)] * (n * progs.n1))
    rows, ext = [], []
    for k in range(n):
        row_vars = [2 * (p0 + k * progs.n1 + j) + 1 for j in progs.row_local]
        ext += row_vars
        rows.append([Number.of(AllocatedNum(v, 0)) for v in row_vars])
    start = len(cs.aux)
    D.reveal_list_of_structs(cs, B, rows)
    return W.compile_block(cs.recipes[start:], start, ext)


class NativeTwoPhaseWitness:
    """the witness of a deposit / withdraw batch with no Python between the native builder's rows
    (ledger.NativeLedger.{deposit,withdraw}_build) and z: the three programs of the native compiler
    (native_circuit.NativeTwoPhaseCircuit: phase 1, phase 2, reveal) through bzk_mpn_dw_witness."""

    def __init__(self, ctx, circuit):
        from .gpu_witness import upload_program
        self.ctx, self.circuit = ctx, circuit
        self.progs = [circuit.program(k) for k in range(3)]
        self._h = [upload_program(ctx, p) for p in self.progs]
        self.ext_src = np.array([-1 if s[0] == "state" else s[1] for s in circuit.ext_src], dtype=np.int32)

    def free(self):
        for h in self._h:
            self.ctx._l.bzk_witness_program_free(self.ctx._h, h)
        self._h = []

    def witness(self, rows, commitment, height):
        """rows = the builder's dict -> (d_inputs [6,4], d_aux [na,4]) int64 CUDA tensors (Montgomery)"""
        import torch
        from ..api import _dev_ptr, _host_ptr
        ctx, c = self.ctx, self.circuit
        n = 1 << (2 * c.B)
        pub = rows["public"]
        head = _canon_rows([commitment, height, pub["state"], pub["aux_data"], pub["next_state"]])
        dev = torch.device("cuda", ctx.device)
        d_in = torch.empty((6, 4), dtype=torch.int64, device=dev)
        d_aux = torch.empty((c.num_aux, 4), dtype=torch.int64, device=dev)
        assert c.num_aux == 5 + n * (c.n1 + c.n2) + c.reveal_vars
        ctx._check(ctx._l.bzk_mpn_dw_witness(ctx._h, self._h[0], self._h[1], self._h[2], n, _host_ptr(rows["raws1"]), _host_ptr(rows["raws2"]),
                                             _host_ptr(rows["roots"]), _host_ptr(self.ext_src), len(self.ext_src), _host_ptr(rows["reveal"]),
                                             _host_ptr(head), _dev_ptr(d_in), _dev_ptr(d_aux)))
        return d_in, d_aux
