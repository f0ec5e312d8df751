"""CPU tier: the N>1 path's host logic on world_size=2 with the gloo backend — each rank holds the
partial sum of its base shard (computed by the oracle here; by the GPU in production), one
all-gather of 104-byte partials, local fold; result == MSM over the whole vector."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bazuka_b200 import dist as bd
    from oracle import cref
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bases, scalars = cref.g1_random_bases(2, n, 2), cref.fr_random(1, n)
    lo, hi = bd.shard_range(n, rank, world)
    partial = cref.msm_g1(bases[lo:hi], scalars[lo:hi], 2)
    total = bd.allgather_fold(partial, "g1", device="cpu")
    want = cref.msm_g1(bases, scalars, 2)
    b2, s2 = cref.g2_random_bases(3, 64, 2), cref.fr_random(4, 64)
    lo2, hi2 = bd.shard_range(64, rank, world)
    t2 = bd.allgather_fold(cref.msm_g2(b2[lo2:hi2], s2[lo2:hi2], 2), "g2", device="cpu")
    ok = bool((total == want).all()) and bool((t2 == cref.msm_g2(b2, s2, 2)).all())
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_msm_allgather_fold_gloo():
    world, n = 2, 1500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_ranges_cover():
    from bazuka_b200 import dist as bd
    for n in (0, 1, 7, 1 << 20):
        for w in (1, 2, 3, 8):
            r = [bd.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def test_host_group_add_matches_oracle(cref):
    from bazuka_b200 import dist as bd
    b = cref.g1_random_bases(5, 3)
    assert (bd.g1_add(b[0], b[1]) == cref.g1_add(b[0], b[1])).all()
    assert (bd.g1_add(b[0], b[0]) == cref.g1_add(b[0], b[0])).all()
    inf = np.zeros(104, np.uint8); inf[96] = 1
    assert (bd.g1_add(b[0], inf)[:96] == b[0][:96]).all()
    assert (bd.fold(b) == cref.g1_add(cref.g1_add(b[0], b[1]), b[2])).all()
    b2 = cref.g2_random_bases(6, 2)
    assert (bd.g2_add(b2[0], b2[1]) == cref.g2_add(b2[0], b2[1])).all()


def _proof_worker(rank, world, port, q):
    """base-sharded Groth16 (SURVEY.md §8e schedule S) on the host: each rank sums its contiguous slice of the
    five base vectors (oracle multiexp standing in for the GPU MSMs), ONE all-gather of 512 B per rank, local
    folds, `bzk_groth16_finalize` — must equal the unsharded oracle proof byte for byte."""
    sys.path.insert(0, ROOT)
    import ctypes as ct
    import torch.distributed as dist
    from bazuka_b200 import dist as bd, groth16 as BG, synth
    from oracle import cref, groth16_c as GC
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ni, na, mats, inputs, aux = synth.build(lanes=2, rounds=3, seed=9, ops=GC.CpuOps)
    pk = GC.setup(ni, na, mats, cref.fr_random(21, 5), threads=2)
    r, s = cref.fr_random(22, 2)
    want = GC.proof_bytes(*GC.prove(ni, na, mats, pk, inputs, aux, r, s, threads=2))
    z = np.ascontiguousarray(np.concatenate([inputs, aux]))
    m = 1 << pk["log_m"]
    ncons = len(mats[0][0]) - 1
    h = np.zeros((m - 1, 4), np.uint64)
    cm = [(np.ascontiguousarray(a, np.uint64), np.ascontiguousarray(b, np.uint32), np.ascontiguousarray(c, np.uint64)) for a, b, c in mats]
    cref.lib().bzko_groth16_h(ct.c_uint64(ni), ct.c_uint64(na), ct.c_uint64(ncons), *GC._csr_args(cm), GC._p(z), GC._p(h), ct.c_int(2))

    def part(bases, scalars, g2=False):
        lo, hi = bd.shard_range(len(scalars), rank, world)
        f = cref.msm_g2 if g2 else cref.msm_g1
        return f(bases[lo:hi], np.ascontiguousarray(scalars[lo:hi]), 2)

    zb = z[pk["b_idx"]]
    hl = bd.g1_add(part(pk["h"][: m - 1], h), part(pk["l"], z[ni:]))
    partials = (part(pk["a"], z[pk["a_idx"]]), part(pk["b_g1"], zb), part(pk["b_g2"], zb, True), hl)
    sums = BG.allgather_partials(partials, device="cpu")
    blob, pts = BG.finalize(pk["vk"], sums, r, s)
    q.put((rank, bool((blob == want).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_groth16_partials_allgather_finalize_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 500)
    procs = [ctx.Process(target=_proof_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=400) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
