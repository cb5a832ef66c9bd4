"""BASELINE.json config 5: allreduce bandwidth sweep 1 KB - 1 GB fp32 at N GPUs (one process per GPU, torchrun).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sweep_allreduce.py

Each size: rank r fills r+1 (exact-sum check N(N+1)/2), 5 warm-up + 20 timed rounds; a round = stage kernel +
allreduce kernel (one-shot and two-shot both timed).  Time = CUDA events per rank, max over ranks (gloo allreduce of
the per-rank means).  algbw = S/t, busbw = algbw * 2(N-1)/N.  NCCL allreduce is timed beside it as a library baseline
only (it is not on the product path).  Rank 0 prints JSON lines and writes gpurun_out/sweep_allreduce_N.json.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moolib_b200 import _lib, peer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-bytes", type=int, default=1 << 30)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--nccl", type=int, default=1)
    ap.add_argument("--sizes", type=int, nargs="*", default=None)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    nccl_group = dist.new_group(backend="nccl") if args.nccl and world > 1 else None
    sizes = args.sizes
    if not sizes:
        sizes, s = [], 1024
        while s <= args.max_bytes:
            sizes.append(s)
            s *= 4
        sizes.append(4377904)  # the atari Net gradient set
        sizes = sorted(set(sizes))
    ctx = peer.make_context(max(sizes) + 64, nslots=1)
    results = []
    for S in sizes:
        n = S // 4
        src = torch.full((n,), float(rank + 1), device="cuda")
        dst = torch.empty(n, device="cuda")
        algos = [("oneshot", _lib.MB_AR_ALGO_ONESHOT)]
        if world > 1:
            algos.append(("twoshot", _lib.MB_AR_ALGO_TWOSHOT))
        for name, algo in algos:
            if name == "oneshot" and S * (world - 1) > (3 << 30):
                continue

            def one_round(stage=True):
                if stage:
                    ctx.stage([src])
                ctx.allreduce_flat(dst, scale=False, algo=algo)

            for _ in range(args.warmup):
                one_round()
            torch.cuda.synchronize()
            dist.barrier()
            # Steady state: `reps` rounds back to back inside ONE event pair (the ranks pace each other through the
            # kernel's own barrier, so host-side launch skew is amortised exactly as in a training loop).
            # (a) whole round = stage kernel + allreduce kernel, (b) allreduce kernel alone.
            def timed(stage):
                dist.barrier()
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
                evs[0].record()
                for i in range(args.reps):
                    one_round(stage)
                    evs[i + 1].record()
                evs[-1].synchronize()
                per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.reps))
                spread[0] = (per[0], per[len(per) // 2], per[-1])
                return evs[0].elapsed_time(evs[-1])
            spread = [None]
            t_round = timed(True)
            t_kernel = timed(False)
            k_spread = spread[0]
            one_round(True)  # leave a valid result behind for the exactness check
            torch.cuda.synchronize()
            ok = bool((dst == world * (world + 1) / 2).all().item())
            st = ctx.result()[1]
            t = torch.tensor([t_round / args.reps, t_kernel / args.reps], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            okt = torch.tensor([int(ok and st == 0)])
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            if rank == 0:
                bus = 2 * (world - 1) / world if world > 1 else 1.0
                rec = {"n_gpus": world, "bytes": S, "algo": name, "round_us": round(t[0].item() * 1e3, 2),
                       "kernel_us": round(t[1].item() * 1e3, 2),
                       "algbw_gbs": round(S / t[1].item() / 1e6, 2), "busbw_gbs": round(S / t[1].item() / 1e6 * bus, 2),
                       "ingress_gbs": round(S * (world - 1) / t[1].item() / 1e6, 2) if name == "oneshot" else None,
                       "exact": bool(okt.item()),
                       "kernel_us_min_med_max": [round(x * 1e3, 1) for x in k_spread]}
                results.append(rec)
                print(json.dumps(rec), flush=True)
        if nccl_group is not None:
            for _ in range(args.warmup):
                dist.all_reduce(src, group=nccl_group)
            torch.cuda.synchronize()
            dist.barrier()
            s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(args.reps):
                dist.all_reduce(src, group=nccl_group)
            e0.record()
            e0.synchronize()
            tt = s0.elapsed_time(e0)
            t = torch.tensor([tt / args.reps], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                bus = 2 * (world - 1) / world
                rec = {"n_gpus": world, "bytes": S, "algo": "nccl(library baseline)",
                       "kernel_us": round(t[0].item() * 1e3, 2), "algbw_gbs": round(S / t[0].item() / 1e6, 2),
                       "busbw_gbs": round(S / t[0].item() / 1e6 * bus, 2)}
                results.append(rec)
                print(json.dumps(rec), flush=True)
        del src, dst
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(results, open(f"gpurun_out/sweep_allreduce_{world}.json", "w"), indent=1)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
