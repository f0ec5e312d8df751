#!/usr/bin/env python3
"""MSM timing probe: G1 / G2 sums of 2^k random terms over resident (tabled) bases, CUDA-event time per call and the
stage split — for tuning environment knobs (BZK_TABLE_C, BZK_AFFINE_ROUNDS, BZK_AFFINE_ROUNDS_G2, BZK_RED_BLOCKS).
usage: msm_probe.py [g1|g2] [log_n] [levels]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bazuka_b200 as B

kind = sys.argv[1] if len(sys.argv) > 1 else "g1"
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
levels = int(sys.argv[3]) if len(sys.argv) > 3 else 16
n = 1 << log_n
ctx = B.Context(0)
ctx.use_torch_stream()
img = torch.empty((n, 104 if kind == "g1" else 200), dtype=torch.uint8, device="cuda")
(ctx.g1_random_bases_dev if kind == "g1" else ctx.g2_random_bases_dev)(2, n, img)
bases = (ctx.g1_bases_from_dev if kind == "g1" else ctx.g2_bases_from_dev)(img, n)
lv = bases.precompute(levels) if levels > 1 else 1
s = torch.empty((n, 4), dtype=torch.int64, device="cuda")
ctx.fr_random_dev(1, n, s)
run = ctx.msm_g1_resident if kind == "g1" else ctx.msm_g2_resident
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ref = run(bases, s)
ctx.set_timing(True)
ts = []
for _ in range(5):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = run(bases, s); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
    assert (out == ref).all()
runs, last, tot = ctx.stage_ms()
print(json.dumps({"kind": kind, "log_n": log_n, "levels": lv, "ms": round(sorted(ts)[len(ts) // 2], 3),
                  "stages_ms": {k: round(float(tot[i] / max(runs, 1)), 3) for i, k in enumerate(B.Context.MSM_STAGES)},
                  "env": {k: v for k, v in os.environ.items() if k.startswith("BZK_")}}))
