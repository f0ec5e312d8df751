#!/usr/bin/env python3
"""Small driver for ncu captures: NTT 2^24, Poseidon-4 2^20, G2 MSM 2^20 (also prints its stage split), the
versioned tree update and the witness interpreter on a 16-transfer batch.
usage: ncu ... python tools/ncu_probe.py ntt|poseidon|msm_g2|tree|witness"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bazuka_b200 as B
ctx = B.Context(0)
what = sys.argv[1] if len(sys.argv) > 1 else "ntt"
if what == "ntt":
    n = 1 << 24
    a = torch.empty((n, 4), dtype=torch.int64, device="cuda"); ctx.fr_random_dev(3, n, a)
    for _ in range(2):
        ctx.ntt_dev(a, 24, 0)
    ctx.synchronize()
elif what == "poseidon":
    n = 1 << 20
    a = torch.empty((n, 4, 4), dtype=torch.int64, device="cuda"); ctx.fr_random_dev(5, n * 4, a)
    o = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    for _ in range(2):
        ctx.poseidon_dev(a, 4, o)
    ctx.synchronize()
elif what in ("witness", "tree", "msm_g2"):
    import numpy as np
    if what == "msm_g2":
        n = 1 << 20
        b = torch.empty((n, 200), dtype=torch.uint8, device="cuda"); ctx.g2_random_bases_dev(2, n, b)
        s = torch.empty((n, 4), dtype=torch.int64, device="cuda"); ctx.fr_random_dev(1, n, s)
        bases = ctx.g2_bases_from_dev(b, n)
        ctx.set_timing(True)
        for _ in range(3):
            ctx.msm_g2_resident(bases, s)
        print("msm_g2 2^20 stage ms:", ctx.stage_ms())
    else:
        from bazuka_b200.mpn import batch_update as BU, native as N, update as U
        from bazuka_b200.mpn.gpu_witness import UpdateWitnessGpu
        A, T, Bb, ntx = 15, 3, 2, 16
        st, keys = U.MpnState(A, T), []
        for i in range(17):
            pk, sk = N.eddsa_keys(b"acct%d" % i); keys.append((pk, sk))
            st.set(i, U.MpnAccount(0, 0, pk, {0: U.Money(U.ZIESHA, 10 ** 12)}))
        txs = []
        for k in range(ntx):
            tx = U.MpnTransaction(1, N.jj_compress(keys[k][0]), N.jj_compress(keys[k + 1][0]), U.Money(U.ZIESHA, 1000 + k), U.Money(U.ZIESHA, 10))
            tx.sign(keys[k][1]); txs.append(tx)
        pub, trans, rej = BU.update_batched(BU.GpuTreeHasher(ctx), st, txs, Bb)   # 'tree': k_tree4_versioned_level launches
        if what == "witness":
            gw = UpdateWitnessGpu(ctx, A, T)
            gw.witness(U.UpdateCircuit(A, T, Bb, commitment=1, height=0, transitions=trans, **pub))
