#!/usr/bin/env python3
"""Proof THROUGHPUT for the single-update MPN circuit: k host threads, each with its own bzk_ctx on the
same GPU, proving back to back (a worker serving several `MpnWork`s).  Latency-bound phases of one
proof overlap the others'.  usage: bench_mpn_throughput.py [threads...]"""
import json, os, sys, time, threading
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bazuka_b200 as B
from bazuka_b200 import groth16 as BG
from bazuka_b200.mpn import cs as C, native as N, update as U
from bench_groth16 import G1_GEN, G2_GEN

def main():
    ks = [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8]
    st, keys = U.MpnState(15, 3), []
    for i in range(2):
        pk, sk = N.eddsa_keys(b"acct%d" % i); keys.append((pk, sk))
        st.set(i, U.MpnAccount(0, 0, pk, {0: U.Money(U.ZIESHA, 10 ** 12)}))
    tx = U.MpnTransaction(1, N.jj_compress(keys[0][0]), N.jj_compress(keys[1][0]), U.Money(U.ZIESHA, 1000), U.Money(U.ZIESHA, 10)); tx.sign(keys[0][1])
    pub, trans, _ = U.update(st, [tx], 0)
    cs = U.UpdateCircuit(15, 3, 0, commitment=1, height=0, transitions=trans, **pub).synthesize(C.ConstraintSystem())
    ni, na, mats, inputs, aux = cs.to_csr()
    for k in ks:
        ctxs = [B.Context(0) for _ in range(k)]
        provers, pks = [], []
        for c in ctxs:
            pr = BG.Prover(c, BG.R1CS(ni, na, *mats))
            d = torch.empty((7, 4), dtype=torch.int64, device="cuda"); c.fr_random_dev(99, 7, d); c.synchronize(); rnd = d.cpu().numpy().view(np.uint64)
            pk, vk = BG.setup_gpu(c, pr.r1cs, rnd[:5], G1_GEN, G2_GEN)
            provers.append(pr); pks.append(pk)
        ref, _ = provers[0].prove(pks[0], inputs, aux, rnd[5], rnd[6])
        reps = 20
        def work(i):
            for _ in range(reps):
                b, _ = provers[i].prove(pks[i], inputs, aux, rnd[5], rnd[6], check_satisfied=False)
            assert (b == ref).all()
        for i in range(k): work.__call__  # no-op
        ths = [threading.Thread(target=work, args=(i,)) for i in range(k)]
        t0 = time.perf_counter()
        for t in ths: t.start()
        for t in ths: t.join()
        dt = time.perf_counter() - t0
        print(json.dumps({"circuit": "UpdateCircuit A=15,T=3,B=0", "threads": k, "proofs": k * reps, "wall_s": round(dt, 3), "proofs_per_s": round(k * reps / dt, 1)}), flush=True)
        for p in pks: p.free()
        for p in provers: p.free()
        for c in ctxs: c.close()

if __name__ == "__main__":
    main()
