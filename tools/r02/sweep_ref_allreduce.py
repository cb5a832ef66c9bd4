"""BASELINE.json configs[4], CPU leg: the UNMODIFIED reference's group.all_reduce over its RPC transport (oracle/_ref),
N peers in one process on loopback exactly as test/test_reduce.py:97-104, blocking future.result() (no done() polling).
One JSON line per (N, size): median seconds per allreduce, algbw, busbw = algbw * 2(N-1)/N.  Host cores are stated."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import oracle

ap = argparse.ArgumentParser()
ap.add_argument("--worlds", type=int, nargs="*", default=[2, 4, 8])
ap.add_argument("--sizes", type=int, nargs="*", default=[1024, 65536, 1048576, 4377904, 16777216, 67108864, 268435456, 1073741824])
ap.add_argument("--budget-s", type=float, default=25.0, help="per (N, size): stop repeating after this many seconds")
a = ap.parse_args()
ref = oracle.load_reference()
import make_golden  # Peers: N reference Rpc peers + Broker in this process
make_golden.moolib = ref
port = 4700
for n in a.worlds:
    peers = make_golden.Peers(n, port); port += 1
    for S in a.sizes:
        numel = S // 4
        if S * n > (6 << 30):
            print(json.dumps({"n": n, "bytes": S, "skipped": "more than 6 GiB of host tensors"}), flush=True)
            continue
        ins = [torch.full((numel,), float(r + 1)) for r in range(n)]
        ts, t_begin, k = [], time.time(), 0
        while k < 12 and (time.time() - t_begin < a.budget_s or k < 2):
            xs = [t.clone() for t in ins]
            t0 = time.perf_counter()
            futs = [peers.groups[r].all_reduce(f"s{S}_{k}", xs[r]) for r in range(n)]
            for f in futs:
                f.result()
            ts.append(time.perf_counter() - t0)
            k += 1
        ok = bool((xs[0] == n * (n + 1) / 2).all().item())
        ts = sorted(ts[1:] if len(ts) > 2 else ts)
        t = ts[len(ts) // 2]
        print(json.dumps({"impl": "unmodified reference group.all_reduce over its RPC transport (oracle/_ref), N peers in one process",
                          "n": n, "bytes": S, "reps": len(ts), "seconds": round(t, 6), "us": round(t * 1e6, 1),
                          "algbw_gbs": round(S / t / 1e9, 4), "busbw_gbs": round(S / t / 1e9 * 2 * (n - 1) / n, 4),
                          "exact": ok, "cores": os.cpu_count()}), flush=True)
    del peers
