#!/usr/bin/env python3
"""One Groth16 proof over N GPUs, base-sharded (SURVEY.md §8e schedule S), next to the replica schedule (R).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/bench_sharded.py [LANESxROUNDS ...]

Every rank synthesises the same circuit and derives the same proving key from the same toxic waste on its own
GPU, keeps only its contiguous 1/N of the five base vectors, runs the witness-side pipeline (SpMV, NTTs)
replicated, sums its shard, and one NCCL all-gather of 512 B per rank + host folds + finalize assemble the
proof on every rank.  Timing: CUDA-synchronised wall clock per proof, MAX over ranks.  Checks: the sharded
proof equals the single-GPU proof byte for byte, and verifies."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
import bazuka_b200 as B
from bazuka_b200 import groth16 as BG, synth


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = B.Context(local)
    dev = torch.device("cuda", local)
    sizes = [(int(a), int(b)) for a, b in (x.split("x") for x in (sys.argv[1:] or ["4096x150"]))]

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for lanes, rounds in sizes:
        ni, na, mats, inputs, aux = synth.build(lanes, rounds, seed=17, ops=synth.GpuOps(ctx))
        pr = BG.Prover(ctx, BG.R1CS(ni, na, *mats))
        d = torch.empty((7, 4), dtype=torch.int64, device=dev); ctx.fr_random_dev(99, 7, d); ctx.synchronize()
        rnd = d.cpu().numpy().view(np.uint64)
        pk, vk = BG.setup_gpu(ctx, pr.r1cs, rnd[:5], BG.G1_GENERATOR, BG.G2_GENERATOR)
        spk = BG.shard_proving_key(ctx, pk, pr.log_m, rank, world)
        want, _ = pr.prove(pk, inputs, aux, rnd[5], rnd[6])

        def sharded():
            parts = pr.prove_partial(spk, inputs, aux, check_satisfied=False)
            sums = BG.allgather_partials(parts, device=dev)
            return BG.finalize(vk, sums, rnd[5], rnd[6])

        blob, pts = sharded()
        assert (blob == want).all(), "sharded proof differs from the single-GPU proof"
        assert BG.verify(vk, inputs[1:], pts)
        t_sh, t_rep = [], []
        for _ in range(5):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter(); sharded(); torch.cuda.synchronize(dev); t_sh.append(max_over_ranks(time.perf_counter() - t0))
        for _ in range(5):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter(); pr.prove(pk, inputs, aux, rnd[5], rnd[6], check_satisfied=False); torch.cuda.synchronize(dev)
            t_rep.append(max_over_ranks(time.perf_counter() - t0))
        if rank == 0:
            sh, rep = min(t_sh), min(t_rep)
            print(json.dumps({"circuit": f"synthetic {lanes}x{rounds}", "constraints": pr.r1cs.num_constraints, "log_m": pr.log_m, "n_gpus": world,
                              "sharded_ms_per_proof": round(sh * 1e3, 2), "sharded_proofs_per_s": round(1 / sh, 2),
                              "replica_ms_per_proof": round(rep * 1e3, 2), "replica_proofs_per_s_all_gpus": round(world / rep, 2),
                              "exchange_bytes_per_rank": 512, "result_check": "sharded proof == single-GPU proof bytes; pairing verifier accepts",
                              "timing": "wall clock between device synchronisations, max over ranks, best of 5"}), flush=True)
        spk.free(); pk.free(); pr.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
