#!/usr/bin/env python3
"""MPN update proofs on the GPU: real UpdateCircuit instances (signed transfers on a Poseidon state),
prints one JSON line per shape with constraint count, witness/CSR build time (host, Python) and the
GPU proving time.  usage: bench_mpn.py [--host-witness] [--host-builder] A,T,B[,ntx] ...
Batches of >= 16 slots take the witness from the GPU (csrc/witness.cu via mpn/gpu_witness.py; R1CS from one
synthesised slot); --host-witness selects the worker-process synthesiser instead."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bazuka_b200 as B
from bazuka_b200 import groth16 as BG
from bazuka_b200.mpn import cs as C, native as N, update as U
from bench_groth16 import G1_GEN, G2_GEN

def main():
    ctx = B.Context(0)
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    host_witness = "--host-witness" in sys.argv
    host_builder = "--host-builder" in sys.argv
    shapes = [tuple(int(v) for v in a.split(",")) for a in (argv or ["3,3,1", "15,3,0", "15,3,1"])]
    for shp in shapes:
        A, T, Bb = shp[:3]
        ntx = shp[3] if len(shp) > 3 else 1 << (2 * Bb)
        t0 = time.time()
        st, keys = U.MpnState(A, T), []
        nacc = max(2, min(ntx + 1, 64))
        for i in range(nacc):
            pk, sk = N.eddsa_keys(b"acct%d" % i); keys.append((pk, sk))
            st.set(i, U.MpnAccount(0, 0, pk, {0: U.Money(U.ZIESHA, 10 ** 12)}))
        nonces = [0] * nacc
        txs = []
        for k in range(ntx):
            s, d = k % nacc, (k + 1) % nacc
            nonces[s] += 1
            tx = U.MpnTransaction(nonces[s], N.jj_compress(keys[s][0]), N.jj_compress(keys[d][0]), U.Money(U.ZIESHA, 1000 + k), U.Money(U.ZIESHA, 10))
            tx.sign(keys[s][1]); txs.append(tx)
        t_sign = time.time() - t0
        t0 = time.time()
        if host_builder or ntx < 16:
            pub, trans, rej = U.update(st, txs, Bb)
        else:
            # batched transition builder: ledger logic on the host, all hashing in ~20 GPU launches
            from bazuka_b200.mpn import batch_update as BU
            pub, trans, rej = BU.update_batched(BU.GpuTreeHasher(ctx), st, txs, Bb)
        t_build = time.time() - t0
        t0 = time.time()
        circ = U.UpdateCircuit(A, T, Bb, commitment=1, height=0, transitions=trans, **pub)
        d_wit, t_wit = None, None
        if (1 << (2 * Bb)) >= 16 and not host_witness:
            from bazuka_b200.mpn import fastsynth as FS
            from bazuka_b200.mpn.gpu_witness import UpdateWitnessGpu
            ni, na, mats, inputs, _ = FS.synthesize_update(circ, structure_only=True)
            ncons = len(mats[0][0]) - 1
            t1 = time.time(); gw = UpdateWitnessGpu(ctx, A, T); t_compile = time.time() - t1
            gw.witness(circ)  # warm-up (module load, arena growth)
            t1 = time.time(); d_wit = gw.witness(circ); t_wit = time.time() - t1
            gw.free()
            a_host = d_wit[1].cpu().numpy().view(np.uint64)
            one = np.array([0x00000001fffffffe, 0x5884b7fa00034802, 0x998c4fefecbc4ff5, 0x1824b159acc5056f], dtype=np.uint64)
            ones = float(((a_host == 0).all(axis=1) | (a_host == one).all(axis=1)).mean())
            aux = None
            del a_host
        elif (1 << (2 * Bb)) >= 16:
            # production-size batches: template + worker processes (bazuka_b200/mpn/fastsynth.py), witness
            # converted to Montgomery form on the GPU (one elementwise product by R^2)
            from bazuka_b200.mpn import fastsynth as FS
            ni, na, mats, inputs, aux_canon = FS.synthesize_update(circ)
            d = torch.from_numpy(aux_canon.view(np.int64)).cuda()
            r2 = torch.from_numpy(np.repeat(np.array([[0xc999e990f3f29c6d, 0x2b6cedcb87925c23, 0x05d314967254398f, 0x0748d9d99f59ff11]], dtype=np.uint64), na, axis=0).view(np.int64)).cuda()
            o = torch.empty_like(d)
            ctx.fr_binop_dev(2, d, r2, o, na); ctx.synchronize()
            aux = o.cpu().numpy().view(np.uint64)
            ones = float(((aux_canon[:, 1:] == 0).all(axis=1) & (aux_canon[:, 0] <= 1)).mean())
            ncons = len(mats[0][0]) - 1
            del d, r2, o
        else:
            cs = circ.synthesize(C.ConstraintSystem())
            ni, na, mats, inputs, aux = cs.to_csr()
            ones = sum(1 for v in cs.aux if v in (0, 1)) / len(cs.aux)
            ncons = cs.num_constraints
        t_syn = time.time() - t0
        pr = BG.Prover(ctx, BG.R1CS(ni, na, *mats))
        d = torch.empty((7, 4), dtype=torch.int64, device="cuda"); ctx.fr_random_dev(99, 7, d); ctx.synchronize(); rnd = d.cpu().numpy().view(np.uint64)
        t0 = time.time(); pk, vk = BG.setup_gpu(ctx, pr.r1cs, rnd[:5], G1_GEN, G2_GEN); t_setup = time.time() - t0
        if d_wit is not None:
            prove = lambda chk: pr.prove_dev(pk, d_wit[0], d_wit[1], rnd[5], rnd[6], check_satisfied=chk)
        else:
            prove = lambda chk: pr.prove(pk, inputs, aux, rnd[5], rnd[6], check_satisfied=chk)
        blob, pts = prove(True)   # check_satisfied: a*b == c on every constraint, i.e. the witness is valid
        assert BG.verify(vk, inputs[1:], pts)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); b2, _ = prove(False); ts.append(time.perf_counter() - t0)
        assert (b2 == blob).all()
        print(json.dumps({"circuit": "UpdateCircuit", "A": A, "T": T, "B": Bb, "tx_slots": 1 << (2 * Bb), "accepted": len(trans), "constraints": ncons,
                          "log_m": pr.log_m, "aux": na, "witness_0_1_fraction": round(ones, 3), "ledger_setup_and_signing_s": round(t_sign, 2), "transition_builder": "host sequential" if (host_builder or ntx < 16) else "gpu batched",
                          "transition_build_s": round(t_build, 3),
                          "synthesize_s": round(t_syn, 2), "witness": "gpu" if d_wit is not None else "host",
                          "gpu_witness_s": None if t_wit is None else round(t_wit, 3), "gpu_setup_s": round(t_setup, 2), "prove_ms_best": round(min(ts) * 1e3, 2),
                          "proofs_per_s": round(1 / min(ts), 2), "tx_per_s": round(len(trans) / min(ts), 1)}), flush=True)
        pk.free(); pr.free()

if __name__ == "__main__":
    main()
