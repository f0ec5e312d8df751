#!/usr/bin/env python3
"""Quick GPU timing probe (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bazuka_b200 as B

ctx = B.Context(0)
ctx.use_torch_stream()

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)

for lg in (16, 20, 22):
    n = 1 << lg
    img = torch.empty((n, 104), dtype=torch.uint8, device="cuda")
    t0 = time.time(); ctx.g1_random_bases_dev(2, n, img); torch.cuda.synchronize(); tg = time.time() - t0
    rb = ctx.g1_bases_from_dev(img, n)
    s = torch.empty((n, 4), dtype=torch.int64, device="cuda"); ctx.fr_random_dev(1, n, s)
    best, avg = timeit(lambda: ctx.msm_g1_resident(rb, s))
    print(f"msm_g1 2^{lg}: best {best:.3f} ms avg {avg:.3f} ms  -> {n/best/1e3:.1f} M scalars/s, {128*n/best/1e6:.1f} GB/s algorithmic (gen bases {tg:.2f}s)", flush=True)
    rb.free(); del img, s
for lg in (20, 24):
    n = 1 << lg
    a = torch.empty((n, 4), dtype=torch.int64, device="cuda"); ctx.fr_random_dev(3, n, a)
    for op in (0, 1, 2, 3):
        best, avg = timeit(lambda: ctx.ntt_dev(a, lg, op))
        print(f"ntt 2^{lg} op{op}: best {best:.3f} ms -> {64*n/best/1e6:.1f} GB/s algorithmic", flush=True)
for ar in (2, 4, 5, 7):
    n = 1 << 20
    a = torch.empty((n, ar, 4), dtype=torch.int64, device="cuda"); ctx.fr_random_dev(5, n * ar, a)
    o = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    best, avg = timeit(lambda: ctx.poseidon_dev(a, ar, o))
    print(f"poseidon-{ar} 2^20 hashes: best {best:.3f} ms -> {n/best/1e3:.2f} M hash/s", flush=True)
n = 1 << 24
a = torch.empty((n, 6), dtype=torch.int64, device="cuda"); a.random_(0, 1 << 60)
b = a.clone(); o = torch.empty_like(a)
best, _ = timeit(lambda: ctx.fp_mul_dev(a, b, o, n))
print(f"fp_mul 2^24: {best:.3f} ms -> {n/best/1e6:.2f} G mul/s (incl. 144 B/elt traffic = {144*n/best/1e6:.0f} GB/s)")
a4 = torch.empty((n, 4), dtype=torch.int64, device="cuda"); ctx.fr_random_dev(1, n, a4); o4 = torch.empty_like(a4)
best, _ = timeit(lambda: ctx.fr_binop_dev(2, a4, a4, o4, n))
print(f"fr_mul 2^24: {best:.3f} ms -> {n/best/1e6:.2f} G mul/s ({96*n/best/1e6:.0f} GB/s)")
