"""BASELINE.json configs[3]: EnvPool batch-stack throughput sweep, 64-4096 envs x 84x84x4 u8 (+ reward f32 + done bool),
T=21 time steps gathered into 32-wide learner batches.

Per env count, on one B200 (CUDA events, >= 3 warm-up, L2 flushed by a 256 MiB memset between timed repetitions, GPU
kept busy in front of the timed launch so that host-side preparation is off the clock, as in the training loop):
  fused       UnrollBatcher: ONE launch per unroll, every byte moves once     (algorithmic bytes 2 x payload)
  two_pass    Batcher.stack x 21 + Batcher.cat                                 (algorithmic bytes 4 x payload)
  pinned_h2d  EnvStepperFuture.result(device=...) path: one launch per step reading the pinned host slab (PCIe-bound,
              reported against the link, not against HBM)
and on the box's host cores (--ref): the UNMODIFIED reference (oracle/_ref) moolib.Batcher("cpu") doing the same
stack x 21 + cat, plus torch.stack/cat as the floor.  One JSON line per env count.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, nargs="*", default=[64, 128, 256, 512, 1024, 2048, 4096])
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--ref", type=int, default=1)
ap.add_argument("--gpu", type=int, default=1)
a = ap.parse_args()
T, Bl = 21, 32


def item(B, dev, g, pinned=False):
    d = {"state": torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
         "reward": torch.randn(B, device=dev, generator=g), "done": torch.rand(B, device=dev, generator=g) < 0.1}
    return {k: v.pin_memory() for k, v in d.items()} if pinned else d


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


peak = 6575.1
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

for B in a.envs:
    payload = B * (4 * 84 * 84 + 4 + 1) * T
    rec = {"envs": B, "T": T, "learner_batch": Bl, "payload_mb": round(payload / 1e6, 2), "hbm_peak_gbs": peak}
    if a.gpu and torch.cuda.is_available():
        import moolib_b200
        from moolib_b200 import _C
        DEV = "cuda:0"
        g = torch.Generator(device=DEV); g.manual_seed(B)
        npool = 3 if B <= 1024 else 2
        pool = [[item(B, DEV, g) for _ in range(T)] for _ in range(npool)]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
        ub = moolib_b200.UnrollBatcher(T, Bl, DEV, cat_dim=1)
        tb, lb = moolib_b200.Batcher(T, DEV), moolib_b200.Batcher(Bl, DEV, dim=1)
        reps = a.reps if B <= 1024 else max(6, a.reps // 3)
        ts, t1, t2 = [], [], []
        for r in range(reps + 3):
            steps = pool[r % npool]
            for it in steps[:-1]:
                ub.stack(it)
            flush.zero_(); torch.cuda._sleep(1_500_000)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = _C.kernel_launches()
            s.record(); ub.stack(steps[-1]); e.record(); torch.cuda.synchronize()
            launches = _C.kernel_launches() - l0
            while not ub.empty():
                ub.get()
            if r >= 3:
                ts.append(s.elapsed_time(e) * 1e3)
        for r in range(reps + 3):
            steps = pool[r % npool]
            flush.zero_(); torch.cuda._sleep(1_500_000)
            s, m, e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            s.record()
            for it in steps:
                tb.stack(it)
            m.record(); lb.cat(tb.get()); e.record(); torch.cuda.synchronize()
            while not lb.empty():
                lb.get()
            if r >= 3:
                t1.append(s.elapsed_time(m) * 1e3); t2.append(m.elapsed_time(e) * 1e3)
        rec["fused"] = {"launches": launches, "us": round(med(ts), 1), "gbs": round(2 * payload / med(ts) / 1e3, 1),
                        "frac_of_hbm": round(2 * payload / med(ts) / 1e3 / peak, 3)}
        two = med(t1) + med(t2)
        rec["two_pass"] = {"launches": T + 1, "stack_us": round(med(t1), 1), "cat_us": round(med(t2), 1),
                           "gbs": round(4 * payload / two / 1e3, 1), "frac_of_hbm": round(4 * payload / two / 1e3 / peak, 3),
                           "speedup_fused": round(two / med(ts), 2)}
        del pool
        # pinned host source: one launch per step (all keys), PCIe-bound
        gh = torch.Generator().manual_seed(B)
        hpool = [item(B, "cpu", gh, pinned=True) for _ in range(4)]
        th = []
        for r in range(12):
            torch.cuda._sleep(500_000)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); out = moolib_b200.to_device(hpool[r % 4], DEV); e.record(); torch.cuda.synchronize()
            if r >= 3:
                th.append(s.elapsed_time(e) * 1e3)
        step_bytes = payload // T
        rec["pinned_h2d"] = {"launches_per_step": 1, "us_per_step": round(med(th), 1),
                             "link_gbs": round(step_bytes / med(th) / 1e3, 1), "bound": "PCIe Gen5 x16 (~55-64 GB/s/dir)"}
        del hpool, ub, tb, lb, flush
        torch.cuda.empty_cache()
    if a.ref:
        try:
            import oracle
            ref = oracle.load_reference()
            gc = torch.Generator().manual_seed(B)
            steps = [item(B, "cpu", gc) for _ in range(T)]
            reps = 3 if B >= 1024 else 6
            tr, tt = [], []
            for r in range(reps):
                tb, lb = ref.Batcher(T, "cpu"), ref.Batcher(Bl, "cpu", dim=1)
                t0 = time.perf_counter()
                for it in steps:
                    tb.stack(it)
                lb.cat(tb.get())
                n = 0
                while not lb.empty():
                    lb.get(); n += 1
                tr.append((time.perf_counter() - t0) * 1e6)
                t0 = time.perf_counter()
                full = {k: torch.stack([s_[k] for s_ in steps]) for k in steps[0]}
                outs = [{k: v[:, i * Bl:(i + 1) * Bl].contiguous() for k, v in full.items()} for i in range(B // Bl)]
                tt.append((time.perf_counter() - t0) * 1e6)
            rec["reference_cpu"] = {"impl": "unmodified reference moolib.Batcher('cpu') (oracle/_ref): stack x21 + cat",
                                    "us": round(med(tr), 1), "gbs": round(4 * payload / med(tr) / 1e3, 2),
                                    "torch_stack_cat_us": round(med(tt), 1), "cores": os.cpu_count(),
                                    "aten_threads": torch.get_num_threads()}
        except Exception as ex:  # noqa: BLE001
            rec["reference_cpu"] = {"unavailable": repr(ex)[:200]}
    print(json.dumps(rec), flush=True)
