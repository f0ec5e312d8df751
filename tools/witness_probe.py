"""Witness generation of one MPN update batch, timed alone (no proving key): native circuit compiler -> native ledger
builder -> bzk_mpn_update_witness.  `BZK_WITNESS_SERIAL=1` selects the round-1 one-thread-per-slot kernel.

    python tools/witness_probe.py [B=4] [signed=32] [reps=5]

Only `signed` transactions are real (signing in Python is ~60 ms each); the remaining slots are the null padding, which
runs the same program."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import bazuka_b200 as bzk
    from bazuka_b200.mpn import native as N, update as U
    from bazuka_b200.mpn.gpu_witness import UpdateWitnessGpu
    from bazuka_b200.mpn.ledger import NativeLedger, pack_txs
    from bazuka_b200.mpn.native_circuit import NativeUpdateCircuit
    A, T = 15, 3
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    signed = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    ctx = bzk.Context(0)
    t0 = time.perf_counter()
    nc = NativeUpdateCircuit(A, T, B)
    prog, epi = nc.program(0), nc.program(1)
    out = {"B": B, "slots": 1 << (2 * B), "slot_ops": int(prog.n_ops), "epilogue_ops": int(epi.n_ops), "compile_s": time.perf_counter() - t0,
           "kernel": "serial" if os.environ.get("BZK_WITNESS_SERIAL", "0") not in ("", "0") else "levels"}
    nc.free()
    wit = UpdateWitnessGpu(ctx, A, T, prog, {B: epi})
    led = NativeLedger(ctx, A, T)
    nacc = min(signed, 16) or 1
    keys = [N.eddsa_keys(b"acct%d" % i) for i in range(nacc)]
    for i, (pk, _) in enumerate(keys):
        led.set_account(i, U.MpnAccount(0, 0, pk, {0: U.Money(U.ZIESHA, 10 ** 12)}))
    nonces, times = [0] * nacc, []
    for rep in range(reps + 1):
        txs = []
        for k in range(signed):
            s, d = k % nacc, (k + 1) % nacc
            nonces[s] += 1
            tx = U.MpnTransaction(nonces[s], N.jj_compress(keys[s][0]), N.jj_compress(keys[d][0]), U.Money(U.ZIESHA, 1000 + k), U.Money(U.ZIESHA, 10))
            tx.sign(keys[s][1])
            txs.append(tx)
        raws, ext, acc, pub, n_acc = led.update_build(pack_txs(txs), B)
        assert n_acc == signed
        ctx.synchronize()
        t0 = time.perf_counter()
        d_in, d_aux = wit.witness_native(raws, ext, [rep + 1, rep, pub["state"], U.ZIESHA, pub["aux_data"], pub["next_state"]], B)
        ctx.synchronize()
        if rep:
            times.append(time.perf_counter() - t0)
    out["witness_s"] = {"min": min(times), "median": float(np.median(times)), "all": times}
    out["checksum"] = int(d_aux.view(-1)[::997].sum().item())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
