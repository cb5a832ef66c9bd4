"""Pins the CPU oracle (oracle/) to the reference: against the golden fixtures generated from the unmodified
reference (tests/golden/make_golden.py) and, where oracle/_ref is present, against the reference run live."""
import ast

import numpy as np
import pytest
import torch

import oracle
from helpers import batcher_trials, gen_input, tree_masks


def _replay_batcher(make, trial):
    mode, size, dim, shape, dt, n, seed = trial
    b = make(size, dim)
    outs = []
    for j in range(n):
        x = torch.from_numpy(gen_input(seed * 100 + j, shape, dt))
        getattr(b, mode)(x)
        while not b.empty():
            outs.append(b.get())
    return outs


def test_oracle_batcher_matches_golden(golden_dir):
    g = np.load(f"{golden_dir}/batcher_golden.npz")
    for ti, trial in enumerate(batcher_trials(g)):
        outs = _replay_batcher(lambda size, dim: oracle.OracleBatcher(size, dim=dim), trial)
        assert len(outs) == int(g[f"t{ti}_nb"]), trial
        for k, o in enumerate(outs):
            ref = g[f"t{ti}_b{k}"]
            assert o.numpy().dtype == ref.dtype and o.numpy().shape == ref.shape
            assert o.numpy().tobytes() == ref.tobytes(), (trial, k)


def test_oracle_c_stack_and_cat_match_golden(golden_dir):
    """The C byte-movement functions (oracle_stack_slot / oracle_cat_narrow) replay the same trials."""
    g = np.load(f"{golden_dir}/batcher_golden.npz")
    for ti, (mode, size, dim, shape, dt, n, seed) in enumerate(batcher_trials(g)):
        nb = int(g[f"t{ti}_nb"])
        items = [gen_input(seed * 100 + j, shape, dt) for j in range(n)]
        if mode == "stack":
            for k in range(nb):
                bshape = shape[:dim] + [size] + shape[dim:]
                dst = np.zeros(bshape, dtype=items[0].dtype)
                for s in range(size):
                    oracle.stack_slot(dst, s, np.ascontiguousarray(items[k * size + s]), dim)
                assert dst.tobytes() == g[f"t{ti}_b{k}"].tobytes()
        else:
            # cat with carry (src/moolib.cc:767-811)
            bshape = list(shape)
            bshape[dim] = size
            dst = np.zeros(bshape, dtype=items[0].dtype)
            fill, k = 0, 0
            for it in items:
                off, m = 0, it.shape[dim]
                while off < m:
                    take = min(m - off, size - fill)
                    oracle.cat_narrow(dst, fill, np.ascontiguousarray(it), off, take, dim)
                    fill += take
                    off += take
                    if fill == size:
                        assert dst.tobytes() == g[f"t{ti}_b{k}"].tobytes(), (ti, k)
                        k += 1
                        fill = 0
                        dst = np.zeros(bshape, dtype=items[0].dtype)
            assert k == nb


def test_oracle_tree_allreduce_matches_golden(golden_dir):
    """group.all_reduce of the reference = oracle tree sum for one of the legal arrival orders, bit for bit."""
    g = np.load(f"{golden_dir}/allreduce_golden.npz")
    for n, rep, seed, numel in g["cases"]:
        ins = [gen_input(int(seed) * 16 + r, [int(numel)], "f32") for r in range(n)]
        ref = g[f"n{n}_r{rep}"]
        hdrs = [(1, 0, 1)] * n
        hit = False
        for mask in tree_masks(int(n)):
            out, _ = oracle.allreduce_tree(ins, hdrs, order=mask, scale=False)
            if out.tobytes() == ref.tobytes():
                hit = True
                break
        assert hit, f"no arrival order reproduces the reference for n={n} rep={rep}"
        # and the product's rank order stays inside the stated tolerance of the reference
        ours, _ = oracle.allreduce_rankorder(ins, hdrs, scale=False)
        assert (np.abs(ours.astype(np.float64) - ref) <= oracle.allreduce_tolerance(ins, ref)).all()


def _acc_round_inputs(n, plan):
    """Per-peer staged gradient (weight[31,32] ++ bias[31] in the flat layout) and header."""
    ins, hdrs = [], []
    numels = [31 * 32, 31]
    offs, total = oracle.flat_layout(numels)
    for i in range(n):
        if not plan[i]:
            ins.append(None)
            hdrs.append((0, 1, 0))
            continue
        staging = np.zeros(total, dtype=np.float32)
        for c, seed in enumerate(plan[i]):
            gw = gen_input(seed, [31, 32], "f32").reshape(-1).copy()
            gb = gen_input(seed + 1, [31], "f32").copy()
            oracle.stage(staging, [gw, gb], accumulate=c > 0, zero_src=True)
            assert not gw.any() and not gb.any()
        ins.append(staging)
        # every count round a peer takes part in without a gradient of its own is a skip
        hdrs.append((len(plan[i]), 0, 10 * len(plan[i])))
    return ins, hdrs, offs, numels


def test_oracle_accumulator_rounds_match_golden(golden_dir):
    g = np.load(f"{golden_dir}/accumulator_golden.npz")
    for r in g["rounds"]:
        tag, n, plan, vbs, ngrad, nskip, bsz = ast.literal_eval(str(r))
        ins, hdrs, offs, numels = _acc_round_inputs(n, plan)
        # skips: peers with fewer contributions than the longest plan skipped the extra count rounds
        maxc = max(len(p) for p in plan)
        hdrs = [(h[0], maxc - len(plan[i]), h[2]) for i, h in enumerate(hdrs)]
        ref = np.concatenate([g[f"{tag}_w"].reshape(-1), g[f"{tag}_b"].reshape(-1)])
        hit = False
        for mask in tree_masks(n):
            out, oh = oracle.allreduce_tree(ins, hdrs, order=mask, scale=True)
            flat = np.concatenate([out[offs[0]:offs[0] + numels[0]], out[offs[1]:offs[1] + numels[1]]])
            assert oh[:3] == (ngrad, nskip, bsz), (tag, oh)
            if flat.tobytes() == ref.tobytes():
                hit = True
                break
        assert hit, tag


def test_fill_batch_and_scatter_actions():
    slab = np.zeros((8, 4, 4), dtype=np.float32)
    rows = [gen_input(40 + i, [4, 4], "f32") for i in range(8)]
    for i in (3, 0, 7, 1, 2, 6, 5, 4):
        oracle.fill_batch(slab, i, rows[i])
    assert slab.tobytes() == np.stack(rows).tobytes()
    counters = np.array([0, 5, 0xFFFFFFFF, 7], dtype=np.uint32)
    acts = np.array([3, 0, 1, 17], dtype=np.int64)
    oracle.scatter_actions(counters, acts)
    assert counters.tolist() == [4, 6, 1, 25]  # prev + 1 + a, uint32 wrap (src/env.cc:340-345)


@pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built in this checkout")
def test_oracle_batcher_matches_live_reference():
    """256 random trials exactly as test/unit/test_batcher.py:13-52 draws them, reference vs restatement."""
    import random
    moolib = oracle.load_reference()
    rnd = random.Random(1234)
    for _ in range(64):
        size, dim = rnd.randint(1, 20), rnd.randint(0, 2)
        dims = rnd.randint(dim + 1, dim + 2)
        n = rnd.randint(20, 60)
        shape = [rnd.randint(1, 4) for _ in range(dims)]
        for mode in ("stack", "cat"):
            a, b = moolib.Batcher(size=size, dim=dim), oracle.OracleBatcher(size, dim=dim)
            for j in range(n):
                x = torch.randn(shape)
                getattr(a, mode)(x)
                getattr(b, mode)(x.clone())
                assert a.empty() == b.empty()
                while not a.empty():
                    assert a.get().equal(b.get())


@pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built in this checkout")
@pytest.mark.timeout(180)
def test_oracle_tree_allreduce_matches_live_reference():
    """group.all_reduce of the compiled reference, run live with N in-process peers (test/test_reduce.py layout), equals
    the C restatement of the tree for one of the legal arrival orders -- bit for bit, fresh random inputs."""
    import time
    moolib = oracle.load_reference()
    n = 4
    addr = "127.0.0.1:4871"
    broker_rpc = moolib.Rpc()
    broker_rpc.set_name("broker")
    broker = moolib.Broker(broker_rpc)
    broker_rpc.listen(addr)
    rpcs, groups = [], []
    for i in range(n):
        r = moolib.Rpc()
        r.set_name(f"peer{i}")
        r.set_timeout(30)
        r.connect(addr)
        g = moolib.Group(r, "live")
        g.set_timeout(30)
        g.set_sort_order(i)
        rpcs.append(r)
        groups.append(g)
    t0 = time.time()
    while not (all(g.active() and len(g.members()) == n for g in groups) and len({g.sync_id() for g in groups}) == 1):
        broker.update()
        for g in groups:
            g.update()
        time.sleep(0.02)
        assert time.time() - t0 < 90
    rng = np.random.default_rng(2026)
    for rep in range(3):
        ins = [rng.standard_normal(2049).astype(np.float32) for _ in range(n)]
        futs = [groups[r].all_reduce(f"live{rep}", torch.from_numpy(ins[r].copy())) for r in range(n)]
        res = [f.result(30).numpy() for f in futs]
        assert all(r.tobytes() == res[0].tobytes() for r in res)
        assert any(oracle.allreduce_tree(ins, [(1, 0, 1)] * n, order=m, scale=False)[0].tobytes() == res[0].tobytes()
                   for m in tree_masks(n))
        ours, _ = oracle.allreduce_rankorder(ins, [(1, 0, 1)] * n, scale=False)
        assert (np.abs(ours.astype(np.float64) - res[0]) <= oracle.allreduce_tolerance(ins, res[0])).all()
        broker.update()
        for g in groups:
            g.update()
