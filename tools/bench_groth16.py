#!/usr/bin/env python3
"""Groth16 proofs/s on synthetic MPN-like circuits (development probe; prints one JSON per size)."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bazuka_b200 as B
from bazuka_b200 import groth16 as BG, synth

def main():
    ctx = B.Context(0)
    sizes = [(int(a), int(b)) for a, b in (x.split("x") for x in (sys.argv[1:] or ["1024x100", "4096x150"]))]
    for lanes, rounds in sizes:
        t0 = time.time()
        ni, na, mats, inputs, aux = synth.build(lanes, rounds, seed=17, ops=synth.GpuOps(ctx))
        t_syn = time.time() - t0
        r1 = BG.R1CS(ni, na, *mats)
        pr = BG.Prover(ctx, r1)
        tox = np.zeros((5, 4), np.uint64); d = torch.empty((5, 4), dtype=torch.int64, device="cuda"); ctx.fr_random_dev(99, 5, d); ctx.synchronize(); tox[:] = d.cpu().numpy().view(np.uint64)
        g1 = np.zeros(104, np.uint8); g2 = np.zeros(200, np.uint8)
        # generators: [1]G via fixed-base kernels would be circular; take them from the random-base kernels' convention: P = [k]G with k=1 is not exposed, so use oracle-free constants
        from bazuka_b200.groth16 import _fr_one
        one = torch.from_numpy(_fr_one().reshape(1, 4).view(np.int64)).cuda()
        t0 = time.time()
        pk, vk = BG.setup_gpu(ctx, r1, tox, G1_GEN, G2_GEN)
        t_setup = time.time() - t0
        rs = torch.empty((2, 4), dtype=torch.int64, device="cuda"); ctx.fr_random_dev(5, 2, rs); ctx.synchronize(); rs = rs.cpu().numpy().view(np.uint64)
        blob, _ = pr.prove(pk, inputs, aux, rs[0], rs[1])
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); blob2, _ = pr.prove(pk, inputs, aux, rs[0], rs[1], check_satisfied=False); ts.append(time.perf_counter() - t0)
        assert (blob == blob2).all()
        print(json.dumps({"lanes": lanes, "rounds": rounds, "constraints": r1.num_constraints, "log_m": pr.log_m, "num_aux": na,
                          "a_len": pr.a_len, "b_len": pr.b_len, "synth_s": t_syn, "setup_s": t_setup, "prove_ms_best": min(ts) * 1e3,
                          "proofs_per_s": 1 / min(ts)}), flush=True)
        pk.free(); pr.free()

def gen_img(xs, ys):
    pass

# BLS12-381 generators as wire images (Montgomery), computed once from the standard affine coordinates
def _mont_fp(x):
    P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
    return ((x << 384) % P).to_bytes(48, "little")
G1_GEN = np.frombuffer(_mont_fp(0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB) + _mont_fp(0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1) + bytes(8), dtype=np.uint8).copy()
G2_GEN = np.frombuffer(_mont_fp(0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8) + _mont_fp(0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E) + _mont_fp(0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801) + _mont_fp(0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE) + bytes(8), dtype=np.uint8).copy()

if __name__ == "__main__":
    main()
