"""BASELINE.json config 4: batch-stack throughput sweep, 64-4096 envs x 84x84x4 u8 (+ reward f32 + done u8).

For each env count B: T=21 time-stack of [B,...] steps into [T,B,...] (K-B2, one launch per step for all leaves),
then the cat re-tile [T,B,...] -> ceil(B/32) x [T,32,...] (K-B3).  Algorithmic bytes = 2 x payload per pass.
CUDA events on the launching stream, L2 flushed (256 MiB memset) between timed iterations, >=3 warm-up.
Prints one JSON object per line and writes gpurun_out/sweep_copy.json.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moolib_b200 import _lib  # noqa: E402

ROW = 4 * 84 * 84


def flush_l2(buf):
    # evict with CLEAN lines: a write flush would leave ~126 MB of dirty L2 that is written back during the timed
    # kernel and charged to it
    buf.sum()


def time_fn(fn, reps, warmup, flush):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    best = 1e30
    for _ in range(reps):
        if flush is not None:
            flush_l2(flush)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ms = s.elapsed_time(e)
        tot += ms
        best = min(best, ms)
    return tot / reps, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="*", default=[64, 128, 256, 512, 1024, 2048, 4096])
    ap.add_argument("--T", type=int, default=21)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/sweep_copy.json")
    ap.add_argument("--tag", default=os.environ.get("MB_COPY_IMPL", "auto"))
    args = ap.parse_args()
    dev = "cuda:0"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    results = []
    T = args.T
    for B in args.envs:
        # keep total memory bounded: T*B*ROW twice (time batch + learner batches) + one step
        step_state = torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8, device=dev)
        step_reward = torch.randn(B, device=dev)
        step_done = torch.rand(B, device=dev) < 0.01
        tb_state = torch.empty((T, B, 4, 84, 84), dtype=torch.uint8, device=dev)
        tb_reward = torch.empty((T, B), device=dev)
        tb_done = torch.empty((T, B), dtype=torch.bool, device=dev)
        payload_step = B * (ROW + 4 + 1)

        stack_jobs = [_lib.make_jobs([(step_state.data_ptr(), tb_state[t].data_ptr(), B * ROW, 1, 0, 0),
                                      (step_reward.data_ptr(), tb_reward[t].data_ptr(), B * 4, 1, 0, 0),
                                      (step_done.data_ptr(), tb_done[t].data_ptr(), B, 1, 0, 0)]) for t in range(T)]

        def stack_all():
            for t in range(T):
                _lib.copy2d_batch(stack_jobs[t])

        def stack_torch():
            for t in range(T):
                tb_state[t].copy_(step_state)
                tb_reward[t].copy_(step_reward)
                tb_done[t].copy_(step_done)

        Bl = 32
        nb = B // Bl
        lb_state = torch.empty((nb, T, Bl, 4, 84, 84), dtype=torch.uint8, device=dev)
        lb_reward = torch.empty((nb, T, Bl), device=dev)
        lb_done = torch.empty((nb, T, Bl), dtype=torch.bool, device=dev)

        def cat_jobs(k):
            return [(tb_state.data_ptr() + k * Bl * ROW, lb_state[k].data_ptr(), Bl * ROW, T, B * ROW, Bl * ROW),
                    (tb_reward.data_ptr() + k * Bl * 4, lb_reward[k].data_ptr(), Bl * 4, T, B * 4, Bl * 4),
                    (tb_done.data_ptr() + k * Bl, lb_done[k].data_ptr(), Bl, T, B, Bl)]

        per_batch = [_lib.make_jobs(cat_jobs(k)) for k in range(nb)]
        all_jobs = _lib.make_jobs([j for k in range(nb) for j in cat_jobs(k)])

        def cat_all():
            # one launch per learner batch covering all three leaves
            for k in range(nb):
                _lib.copy2d_batch(per_batch[k])

        def cat_one_launch():
            _lib.copy2d_batch(all_jobs)

        def cat_torch():
            for k in range(nb):
                lb_state[k].copy_(tb_state[:, k * Bl:(k + 1) * Bl])
                lb_reward[k].copy_(tb_reward[:, k * Bl:(k + 1) * Bl])
                lb_done[k].copy_(tb_done[:, k * Bl:(k + 1) * Bl])

        big_job = _lib.make_jobs([(tb_state.data_ptr(), lb_state.data_ptr(), T * B * ROW, 1, 0, 0)])

        def big_copy():
            # one contiguous copy of the whole time batch: the pure streaming rate of the kernel
            _lib.copy2d_batch(big_job)

        def big_copy_torch():
            lb_state.view(-1).copy_(tb_state.view(-1))

        payload = T * payload_step
        for name, fn, nbytes in [("stack", stack_all, 2 * payload), ("stack_torch", stack_torch, 2 * payload),
                                 ("cat", cat_all, 2 * payload), ("cat_1launch", cat_one_launch, 2 * payload),
                                 ("cat_torch", cat_torch, 2 * payload),
                                 ("contig", big_copy, 2 * T * B * ROW), ("contig_torch", big_copy_torch, 2 * T * B * ROW)]:
            mean_ms, best_ms = time_fn(fn, args.reps, args.warmup, flush)
            rec = {"impl": args.tag, "op": name, "envs": B, "T": T, "bytes": nbytes, "ms": round(mean_ms, 5),
                   "best_ms": round(best_ms, 5), "gbs": round(nbytes / mean_ms / 1e6, 1),
                   "best_gbs": round(nbytes / best_ms / 1e6, 1), "frac_hbm": round(nbytes / mean_ms / 1e6 / hbm, 4),
                   "frames_per_s": round(T * B / (mean_ms / 1e3))}
            results.append(rec)
            print(json.dumps(rec), flush=True)
        # correctness spot check of the last fn set
        stack_all()
        cat_one_launch()
        torch.cuda.synchronize()
        assert lb_state[0].equal(tb_state[:, :Bl]) and tb_state[T - 1].equal(step_state)
        del tb_state, lb_state
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    prev = []
    if os.path.exists(args.out):
        prev = json.load(open(args.out))
    json.dump(prev + results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
