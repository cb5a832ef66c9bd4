"""Isolated timing of the device-gated allreduce (K-A0 + K-A2) with all ranks in one process.
   python tools/r02/ar_micro.py [--sizes 4377904,...] [--algos oneshot,twoshot] [--iters 30]
Prints one JSON line per (size, algo): K-A2 time from the context's own events (median / p10 / p90), K-A0 time."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from moolib_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="4096,65536,1048576,4377904,16777216,67108864,268435456")
ap.add_argument("--algos", default="oneshot,twoshot")
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--world", type=int, default=0)
ap.add_argument("--tag", default="")
ap.add_argument("--inplace", type=int, default=1, help="two-shot: consume the result in the staging buffer (what the Accumulator does)")
a = ap.parse_args()
n = a.world or torch.cuda.device_count()
ALG = {"oneshot": _lib.MB_AR_ALGO_ONESHOT, "twoshot": _lib.MB_AR_ALGO_TWOSHOT, "auto": _lib.MB_AR_ALGO_AUTO}
for size in [int(x) for x in a.sizes.split(",")]:
    numel = (size // 4 + 3) // 4 * 4
    ctx = [_lib.ArContext(r, n, r, numel * 4) for r in range(n)]
    hs = [c.export() for c in ctx]
    for r, c in enumerate(ctx):
        for q in range(n):
            if q != r:
                c.import_peer(q, hs[q])
    dst = [torch.empty(numel, device=f"cuda:{r}") for r in range(n)]
    for r in range(n):
        for k in range(3):
            ctx[r].buffer(numel, ahead=k).fill_(float(r + 1))
    for algo in a.algos.split(","):
        if algo == "twoshot" and n < 2:
            continue
        red, gate = [], []
        for it in range(a.iters + 5):
            for r in range(n):
                with torch.cuda.device(r):
                    d = ctx[r].buffer(numel, ahead=0) if (algo == "twoshot" and a.inplace) else dst[r]
                    ctx[r].reduce_gated(1, flat_dst=d, hdr=(1, 0, 1, 1), scale=False, algo=ALG[algo])
            for r in range(n):
                torch.cuda.synchronize(r)
            ok = all(ctx[r].result()[1] == 0 for r in range(n))
            assert ok
            if it == 0:
                exp = float(n * (n + 1) // 2)
                # two-shot writes the reduced values back into the staging: re-fill before the next round
                outs = [ctx[r].buffer(numel, ahead=0) if (algo == "twoshot" and a.inplace) else dst[r] for r in range(n)]
                assert all((outs[r] == exp).all().item() for r in range(n)), (size, algo)
            ts = [ctx[r].round_times() for r in range(n)]
            if it >= 5:
                red.append(max(t[1] for t in ts))
                gate.append(max(t[0] for t in ts))
            for r in range(n):
                ctx[r].advance()
                ctx[r].buffer(numel, ahead=0).fill_(float(r + 1))
        red.sort(); gate.sort()
        med = red[len(red) // 2]
        bus = size * 2 * (n - 1) / n / med / 1e3 if n > 1 else 0
        print(json.dumps({"tag": a.tag, "n": n, "bytes": size, "algo": algo + ("(in place)" if algo == "twoshot" and a.inplace else ""), "reduce_us_p50": round(med, 2),
                          "p10": round(red[len(red) // 10], 2), "p90": round(red[len(red) * 9 // 10], 2),
                          "gate_us_p50": round(gate[len(gate) // 2], 2), "busbw_gbs": round(bus, 1),
                          "algbw_gbs": round(size / med / 1e3, 1)}), flush=True)
    for c in ctx:
        c.close()
