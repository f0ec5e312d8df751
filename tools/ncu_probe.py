#!/usr/bin/env python3
"""Small driver for ncu captures of the NTT / Poseidon / SpMV kernels (one invocation of each at the
BASELINE sizes).  usage: ncu ... python tools/ncu_probe.py ntt|poseidon"""
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
