"""Isolated timing of the UnrollBatcher gather (one launch per unroll) and of the two-pass stack + cat it replaces.
   python tools/r02/gather_micro.py [--envs 256] [--reps 12]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import moolib_b200
from moolib_b200 import _C

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=256)
ap.add_argument("--reps", type=int, default=12)
ap.add_argument("--tag", default="")
a = ap.parse_args()
DEV, T, B, Bl = "cuda:0", 21, a.envs, 32
g = torch.Generator(device=DEV); g.manual_seed(1)

def item():
    return {"env_outputs": {"state": torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8, device=DEV, generator=g),
                            "reward": torch.randn(B, device=DEV, generator=g),
                            "done": torch.rand(B, device=DEV, generator=g) < 0.1,
                            "prev_action": torch.randint(0, 18, (B,), device=DEV, generator=g)},
            "actor_outputs": {"policy_logits": torch.randn(B, 18, device=DEV, generator=g),
                              "baseline": torch.randn(B, device=DEV, generator=g),
                              "action": torch.randint(0, 18, (B,), device=DEV, generator=g)}}

pool = [[item() for _ in range(T)] for _ in range(3)]   # 3 x 152 MB of sources: > L2
payload = sum(v.numel() * v.element_size() for grp in pool[0][0].values() for v in grp.values()) * T
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
ub = moolib_b200.UnrollBatcher(T, Bl, DEV, cat_dim=1)
tb, lb = moolib_b200.Batcher(T, DEV), moolib_b200.Batcher(Bl, DEV, dim=1)
res = {"tag": a.tag, "envs": B, "payload_mb": round(payload / 1e6, 2)}
ts = []
for r in range(a.reps):
    steps = pool[r % 3]
    for it in steps[:-1]:
        ub.stack(it)
    flush.zero_()
    torch.cuda._sleep(2_000_000)   # ~1 ms of GPU work queued: the host-side preparation of the launch is hidden, as in the loop
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _C.kernel_launches()
    import time as _t
    h0 = _t.perf_counter()
    s.record(); ub.stack(steps[-1]); e.record()
    host_us = (_t.perf_counter() - h0) * 1e6
    torch.cuda.synchronize()
    assert _C.kernel_launches() - l0 == 1
    while not ub.empty():
        ub.get()
    if r >= 2:
        ts.append(s.elapsed_time(e) * 1e3)
        res.setdefault("host_us", []).append(round(host_us, 1))
ts.sort()
res["host_us"] = sorted(res["host_us"])[len(res["host_us"]) // 2]
res["gather_us_p50"] = round(ts[len(ts) // 2], 1)
res["gather_gbs"] = round(2 * payload / ts[len(ts) // 2] / 1e3, 1)
# the two-pass path: T stack launches + one cat launch
t_stack, t_cat = [], []
for r in range(a.reps):
    steps = pool[r % 3]
    flush.zero_()
    torch.cuda._sleep(2_000_000)
    s, m, e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    s.record()
    for it in steps:
        tb.stack(it)
    m.record()
    data = tb.get()
    lb.cat(data)
    e.record()
    torch.cuda.synchronize()
    while not lb.empty():
        lb.get()
    if r >= 2:
        t_stack.append(s.elapsed_time(m) * 1e3); t_cat.append(m.elapsed_time(e) * 1e3)
t_stack.sort(); t_cat.sort()
res["stack_x21_us_p50"] = round(t_stack[len(t_stack) // 2], 1)
res["cat_us_p50"] = round(t_cat[len(t_cat) // 2], 1)
res["two_pass_gbs"] = round(4 * payload / (t_stack[len(t_stack) // 2] + t_cat[len(t_cat) // 2]) / 1e3, 1)
print(json.dumps(res), flush=True)
