"""Parallel synthesis of UpdateCircuit for large batches.

`UpdateCircuit.synthesize` is a loop of structurally identical transition blocks chained only through
`state_wit` (/root/reference/src/mpn/circuits/update_circuit.rs:81-469).  So the R1CS of an n-slot batch is
  prologue (5 inputize rows) | n x block template (column indices shifted per slot) | epilogue
and the witness of slot k depends only on its transition and the state root entering it.  This module
builds the template once, replicates it with numpy, and computes the per-slot witness values in worker
processes — the same constraint system and the same witness as the sequential synthesiser
(tests/test_mpn_cpu.py checks equality), in seconds instead of minutes for the production batch
(B=4: 256 slots, 14.4 M constraints)."""
import multiprocessing as mp
import os

import numpy as np

from . import update as U
from .cs import LC, ConstraintSystem, AllocatedNum, to_mont, R
from .gadgets import Number

FAKE_STATE_VAR = 2 * (10 ** 12) + 1  # stand-in id for the state variable entering a block


def _block(args):
    A, T, B, pro_vals, tr, pre_root, want_rows = args
    circ = U.UpdateCircuit(A, T, 0, commitment=pro_vals[0], height=pro_vals[1], state=pro_vals[2], aux_data=pro_vals[4],
                           next_state=pro_vals[5], fee_token=pro_vals[3], transitions=[tr])
    cs = ConstraintSystem()
    state_wit, fee_tok, aux_wit, claimed = circ._prologue(cs)
    P_aux, P_rows = len(cs.aux), len(cs.rows)
    state_in = AllocatedNum(FAKE_STATE_VAR, pre_root % R)
    state_out, fee_sum = circ._tx_block(cs, tr, state_in, fee_tok, Number.zero())
    vals = b"".join(v.to_bytes(32, "little") for v in cs.aux[P_aux:])
    info = (state_out.var, circ._last_final_fee.var, state_out.value, circ._last_final_fee.value)
    rows = cs.rows[P_rows:] if want_rows else None
    return vals, info, rows, P_aux, P_rows


def synthesize_update(circ: U.UpdateCircuit, workers=None, structure_only=False):
    """-> (num_inputs, num_aux, mats, inputs [ni,4] Montgomery, aux_canonical [na,4] uint64 canonical).
    The caller converts aux to Montgomery (GPU: one elementwise product by R^2).
    structure_only: build the R1CS from ONE synthesised slot and skip the other slots' values (the returned
    aux is then meaningless) — the shape the prover/setup need when the witness comes from the GPU
    (gpu_witness.py)."""
    n = len(circ.transitions)
    pro_vals = [circ.commitment, circ.height, circ.state, circ.fee_token, circ.aux_data, circ.next_state]
    # state root entering every slot: recorded by update() for real transitions; disabled slots keep the state
    roots, cur = [], circ.state
    for tr in circ.transitions:
        if tr.enabled:
            cur = tr.pre_root
        roots.append(cur)
        if tr.enabled:
            cur = None  # filled from the block result below
    # roots of slots following an enabled one come from that block's state_out: resolve sequentially
    # using update()'s bookkeeping: slot k+1's pre_root (if enabled) or the final next_state
    for k in range(n):
        if roots[k] is None:
            roots[k] = circ.transitions[k].pre_root if circ.transitions[k].enabled else circ.next_state
    # after the last enabled slot the state is next_state
    last_enabled = max([k for k, t in enumerate(circ.transitions) if t.enabled], default=-1)
    for k in range(last_enabled + 1, n):
        roots[k] = circ.next_state if last_enabled >= 0 else circ.state
    jobs = [(circ.A, circ.T, circ.B, pro_vals, tr, roots[k], k == 0) for k, tr in enumerate(circ.transitions)]
    workers = workers or min(os.cpu_count() or 1, n)
    if structure_only:
        res = [_block(jobs[0])] * n
    elif workers > 1 and n > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_block, jobs, chunksize=max(1, n // (workers * 2)))
    else:
        res = [_block(j) for j in jobs]
    vals0, info0, rows_t, P_aux, P_rows = res[0]
    A_tx = len(vals0) // 32
    R_tx = len(rows_t)
    s_out_local = (info0[0] >> 1) - P_aux
    f_local = (info0[1] >> 1) - P_aux
    # ---- prologue + epilogue on a real constraint system (variable ids continue after the blocks)
    cs = ConstraintSystem()
    state_wit, fee_tok, aux_wit, claimed = circ._prologue(cs)
    assert len(cs.aux) == P_aux and len(cs.rows) == P_rows
    pro_rows = list(cs.rows)
    cs.aux.extend([0] * (n * A_tx))  # placeholders; real values are assembled from the workers
    fee_sum = Number.zero()
    for k in range(n):
        ff_var = 2 * (P_aux + k * A_tx + f_local) + 1
        fee_sum = Number(fee_sum.lc.add_term(1, ff_var), fee_sum.value + res[k][1][3])
    last_state = AllocatedNum(2 * (P_aux + (n - 1) * A_tx + s_out_local) + 1, res[n - 1][1][2])
    cs.rows = []
    circ._epilogue(cs, last_state, fee_tok, aux_wit, claimed, fee_sum)
    epi_rows = cs.rows
    epi_aux = cs.aux[P_aux + n * A_tx:]
    ni = len(cs.inputs)
    na = P_aux + n * A_tx + len(epi_aux)

    def zidx_arr(var, k):
        """template variable ids -> z indices for slot k (vectorised)."""
        var = np.asarray(var, dtype=np.int64)
        is_in = (var % 2 == 0)
        aidx = var >> 1
        out = np.where(is_in, aidx, 0)
        pro = (~is_in) & (aidx < P_aux) & (var != FAKE_STATE_VAR)
        out = np.where(pro, ni + aidx, out)
        loc = (~is_in) & (aidx >= P_aux) & (var != FAKE_STATE_VAR)
        out = np.where(loc, ni + P_aux + k * A_tx + (aidx - P_aux), out)
        fake = (var == FAKE_STATE_VAR)
        prev = (state_wit.var >> 1) + ni if k == 0 else ni + P_aux + (k - 1) * A_tx + s_out_local
        return np.where(fake, prev, out)

    def plain_rows(rows):
        ni_ = ni
        mats = []
        for side in range(3):
            counts, cols, vals = [], [], []
            for row in rows:
                c = 0
                for v, co in row[side].t.items():
                    if co:
                        cols.append((v >> 1) if v % 2 == 0 else ni_ + (v >> 1))
                        vals.append(co)
                        c += 1
                counts.append(c)
            mats.append((np.array(counts, dtype=np.int64), np.array(cols, dtype=np.int64), to_mont(vals)))
        return mats

    # template as flat arrays per side
    tmpl = []
    for side in range(3):
        counts, tvars, vals = [], [], []
        for row in rows_t:
            c = 0
            for v, co in row[side].t.items():
                if co:
                    tvars.append(v)
                    vals.append(co)
                    c += 1
            counts.append(c)
        tmpl.append((np.array(counts, dtype=np.int64), np.array(tvars, dtype=np.int64), to_mont(vals)))
    pro_m, epi_m = plain_rows(pro_rows), plain_rows(epi_rows)
    mats = []
    for side in range(3):
        tc, tv, tval = tmpl[side]
        counts = np.concatenate([pro_m[side][0], np.tile(tc, n), epi_m[side][0]])
        cols = np.concatenate([pro_m[side][1]] + [zidx_arr(tv, k) for k in range(n)] + [epi_m[side][1]])
        vals = np.concatenate([pro_m[side][2], np.tile(tval, (n, 1)), epi_m[side][2]])
        rp = np.zeros(len(counts) + 1, dtype=np.uint64)
        np.cumsum(counts, out=rp[1:])
        mats.append((rp, cols.astype(np.uint32), np.ascontiguousarray(vals)))
    aux_bytes = b"".join([b"".join(v.to_bytes(32, "little") for v in cs.aux[:P_aux])] + [r[0] for r in res] +
                         [b"".join(v.to_bytes(32, "little") for v in epi_aux)])
    aux_canon = np.frombuffer(aux_bytes, dtype=np.uint64).reshape(-1, 4).copy()
    assert len(aux_canon) == na
    return ni, na, mats, to_mont(cs.inputs), aux_canon


def canon_to_mont_host(aux_canon):
    """host conversion (tests / small circuits); large batches use the GPU product by R^2."""
    ints = [int.from_bytes(row.tobytes(), "little") for row in aux_canon]
    return to_mont(ints)
