#!/usr/bin/env python3
"""Steady-state MPN update proving: batches of 4^B signed transfers flow through
    batched transition builder                     (own context, 1 thread)
    GPU witness                                    (own contexts, 2 threads taking alternate batches: a slot is one
                                                    latency-bound chain, slower still next to the prover's kernels)
    Groth16 prove from the resident witness        (context 1, main thread)
so the transitions of batch k+2 are built and the witness of batch k+1 is computed while batch k is being proved
(the witness kernel occupies a handful of warps).  Prints one JSON line: wall time per batch in steady state, proofs/s, transactions/s, and the stage times.
usage: bench_mpn_pipeline.py A,T,B [n_batches] [--python-host]
Default: the native per-batch path (bzk_mpn_update_build -> bzk_mpn_update_witness -> bzk_groth16_prove_dev; Python only
moves buffers between the three calls); --python-host uses the Python builder and witness glue instead."""
import json, os, queue, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bazuka_b200 as B
from bazuka_b200 import groth16 as BG
from bazuka_b200.mpn import batch_update as BU, native as N, update as U
from bazuka_b200.mpn.gpu_witness import UpdateWitnessGpu
from bazuka_b200.mpn.worker import MpnUpdateWorker


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    native = "--python-host" not in sys.argv
    A, T, Bb = (int(v) for v in (argv[0] if argv else "15,3,2").split(","))
    n_batches = int(argv[1]) if len(argv) > 1 else 4
    ntx = 1 << (2 * Bb)
    N_WIT = 2
    ctx1, ctx3 = B.Context(0), B.Context(0)
    wctx = [B.Context(0) for _ in range(N_WIT)]
    # the builder's and the witness kernels are small and latency-bound; the prover's saturate the SMs.  Give the
    # small ones high-priority streams so their CTAs are placed first whenever an MSM CTA retires.
    hp = [torch.cuda.Stream(priority=-1) for _ in range(N_WIT + 1)]
    for c, s_ in zip(wctx + [ctx3], hp):
        with torch.cuda.stream(s_):
            c.use_torch_stream()
    d = torch.empty((7, 4), dtype=torch.int64, device="cuda"); ctx1.fr_random_dev(99, 7, d); ctx1.synchronize()
    rnd = d.cpu().numpy().view(np.uint64)
    t0 = time.time()
    worker = MpnUpdateWorker(ctx1, A, T, Bb, rnd[:5])
    t_setup = time.time() - t0
    wits, hasher3 = [UpdateWitnessGpu(c, A, T) for c in wctx], BU.GpuTreeHasher(ctx3)
    # ledger + signed transfers for all batches (host signing is outside the measured pipeline: it is the wallets' work)
    nacc = max(2, min(ntx + 1, 64))
    st, keys = U.MpnState(A, T), []
    for i in range(nacc):
        pk, sk = N.eddsa_keys(b"acct%d" % i); keys.append((pk, sk))
        st.set(i, U.MpnAccount(0, 0, pk, {0: U.Money(U.ZIESHA, 10 ** 12)}))
    led = None
    if native:
        from bazuka_b200.mpn.ledger import NativeLedger, pack_txs
        led = NativeLedger(ctx3, A, T)
        for i, a in st.accounts.items():
            led.set_account(i, a)
        assert led.root == st.root
    nonces, batches = [0] * nacc, []
    t0 = time.time()
    for b in range(n_batches + 1):
        txs = []
        for k in range(ntx):
            s, dd = k % nacc, (k + 1) % nacc
            nonces[s] += 1
            tx = U.MpnTransaction(nonces[s], N.jj_compress(keys[s][0]), N.jj_compress(keys[dd][0]), U.Money(U.ZIESHA, 1000 + k), U.Money(U.ZIESHA, 10))
            tx.sign(keys[s][1]); txs.append(tx)
        batches.append(pack_txs(txs) if native else txs)
    t_sign = time.time() - t0
    q = queue.Queue(maxsize=1)
    stage = {"build": [], "witness": [], "prove": []}
    blobs, works = [], []

    qb = queue.Queue(maxsize=1)

    def builder():
        for b, txs in enumerate(batches):
            t = time.perf_counter()
            if native:
                raws, ext, acc, pub, n_acc = led.update_build(txs, Bb)
                assert n_acc == ntx
                circ = (raws, ext)
            else:
                pub, trans, rej = BU.update_batched(hasher3, st, txs, Bb)
                assert len(trans) == ntx and not rej
                circ = U.UpdateCircuit(A, T, Bb, commitment=b + 1, height=b, transitions=trans, **pub)
            stage["build"].append(time.perf_counter() - t)
            qb.put((b, circ, pub))
        qb.put(None)

    def witnesser(w):
        while True:
            item = qb.get()
            if item is None:
                qb.put(None)      # let the other witness thread see the end marker too
                break
            b, circ, pub = item
            t = time.perf_counter()
            if native:
                wit = wits[w].witness_native(circ[0], circ[1], [b + 1, b, pub["state"], U.ZIESHA, pub["aux_data"], pub["next_state"]], Bb)
            else:
                wit = wits[w].witness(circ)
            stage["witness"].append(time.perf_counter() - t)
            q.put((b, circ, pub, wit))
        q.put(None)

    marks = []
    th = threading.Thread(target=builder)
    th2 = [threading.Thread(target=witnesser, args=(w,)) for w in range(N_WIT)]
    th.start()
    for t_ in th2:
        t_.start()
    ended = 0
    while True:
        item = q.get()
        if item is None:
            ended += 1
            if ended == N_WIT:
                break
            continue
        b, circ, pub, (d_in, d_aux) = item
        t = time.perf_counter()
        blob, pts = worker.prover.prove_dev(worker.pk, d_in, d_aux, rnd[5], rnd[6], check_satisfied=(b == 0))
        stage["prove"].append(time.perf_counter() - t)
        marks.append(time.perf_counter())
        blobs.append(blob); works.append((b, pub))
    th.join()
    for t_ in th2:
        t_.join()
    from bazuka_b200.mpn.cs import to_mont
    ok = all(BG.verify_bytes(worker.vk_blob, to_mont([b + 1, b, pub["state"], pub["aux_data"], pub["next_state"]]), blob) for (b, pub), blob in zip(works, blobs))
    per_batch = (marks[-1] - marks[0]) / (len(marks) - 1)   # batch 0 is the warm-up
    med = lambda v: float(np.median(v[1:]))
    print(json.dumps({"circuit": "UpdateCircuit", "A": A, "T": T, "B": Bb, "tx_per_batch": ntx, "batches_timed": len(marks) - 1,
                      "steady_state_s_per_batch": round(per_batch, 4), "proofs_per_s": round(1 / per_batch, 3), "tx_per_s": round(ntx / per_batch, 1),
                      "stage_s_median": {k: round(med(v), 4) for k, v in stage.items()}, "all_proofs_verify": bool(ok),
                      "host_path": "native (libbzk builder + witness driver)" if native else "python builder + witness glue",
                      "one_off_s": {"r1cs_template_key_setup": round(t_setup, 1), "host_signing_all_batches": round(t_sign, 1)},
                      "pipeline": "builder, 2 witness workers (alternate batches) and the prover run on their own threads and contexts; wall clock between consecutive proofs"}), flush=True)


if __name__ == "__main__":
    main()
