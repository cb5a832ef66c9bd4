"""HP-A parity on the GPU: stage + allreduce kernels through the C-ABI against the CPU oracle.

* bit-exact against the oracle in the product's summation order (ascending rank),
* within 1e-6 relative (SURVEY.md section 8c tolerance model, written out in oracle.allreduce_tolerance) of the
  reference's tree order -- including the golden outputs of the reference's own group.all_reduce / Accumulator,
* every rank ends with identical bits.

Multi-GPU cases run all ranks inside this one process (peer access instead of CUDA IPC); the cross-process path is
covered by tests/test_allreduce_ipc.py.  They are skipped when the box has fewer GPUs.
"""
import ast

import numpy as np
import pytest
import torch

import oracle
from helpers import gen_input, tree_masks
from moolib_b200 import _lib

pytestmark = pytest.mark.gpu

NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
ALGOS = {"oneshot": _lib.MB_AR_ALGO_ONESHOT, "twoshot": _lib.MB_AR_ALGO_TWOSHOT}


class World:
    def __init__(self, n, max_bytes, nslots=1):
        self.n = n
        self.ctx = [_lib.ArContext(r, n, r, max_bytes, nslots) for r in range(n)]
        hs = [c.export() for c in self.ctx]
        for r, c in enumerate(self.ctx):
            for q in range(n):
                if q != r:
                    c.import_peer(q, hs[q])

    def close(self):
        for c in self.ctx:
            c.close()

    def sync(self):
        for r in range(self.n):
            torch.cuda.synchronize(r)


def run_round(w, per_rank_tensors, hdrs, dst_like, algo, scale=True, slot=0, accumulate=False):
    """Stage each rank's tensor list, allreduce into fresh destination tensors, return them (cpu) + result header."""
    dsts = []
    for r in range(w.n):
        with torch.cuda.device(r):
            if per_rank_tensors[r] is not None:
                w.ctx[r].stage(per_rank_tensors[r], slot=slot, accumulate=accumulate, zero_src=True)
            dsts.append([torch.full_like(t, float("nan"), device=f"cuda:{r}") for t in dst_like])
    for r in range(w.n):
        with torch.cuda.device(r):
            h = hdrs[r] + (0 if per_rank_tensors[r] is None else 1,)
            w.ctx[r].allreduce(dsts[r], hdr=h, slot=slot, scale=scale, algo=algo, timeout_ms=20000)
    w.sync()
    res = [w.ctx[r].result(slot) for r in range(w.n)]
    return [[t.cpu() for t in d] for d in dsts], res


def flat(tensors, numels):
    offs, total = oracle.flat_layout(numels)
    out = np.zeros(total, dtype=np.float32)
    for t, o, n in zip(tensors, offs, numels):
        out[o:o + n] = t.reshape(-1).numpy() if isinstance(t, torch.Tensor) else t.reshape(-1)
    return out


def unflat(vec, numels):
    offs, _ = oracle.flat_layout(numels)
    return [vec[o:o + n] for o, n in zip(offs, numels)]


WORLDS = [n for n in (1, 2, 4, 8) if n <= max(NGPU, 1)]


@pytest.mark.parametrize("n", WORLDS)
@pytest.mark.parametrize("algo", list(ALGOS))
def test_atari_grad_list_matches_oracle(n, algo):
    """36 tensors / 1,094,476 floats (the atari Net's parameter shapes)."""
    shapes = atari_param_shapes()
    numels = [int(np.prod(s)) for s in shapes]
    assert sum(numels) == 1094476
    w = World(n, _lib.flat_numel(numels) * 4)
    try:
        ins_np = [[gen_input(100 * r + i, s, "f32") for i, s in enumerate(shapes)] for r in range(n)]
        ins = [[torch.from_numpy(a.copy()).to(f"cuda:{r}") for a in ins_np[r]] for r in range(n)]
        hdrs = [(1, 0, 32)] * n
        outs, res = run_round(w, ins, hdrs, [torch.empty(s) for s in shapes], ALGOS[algo])
        flat_in = [flat(ins_np[r], numels) for r in range(n)]
        exact, eh = oracle.allreduce_rankorder(flat_in, hdrs)
        tree, _ = oracle.allreduce_tree(flat_in, hdrs, order=0)
        tol = oracle.allreduce_tolerance(flat_in, tree, 1.0 / n)
        for r in range(n):
            got = flat(outs[r], numels)
            assert got.tobytes() == exact.tobytes(), f"rank {r} differs from the rank-order oracle"
            assert (np.abs(got.astype(np.float64) - tree) <= tol).all()
            assert res[r] == (eh, 0)
            # sources were zeroed by the stage kernel (accumulator.cc:410-418)
            assert all(not t.any().item() for t in ins[r])
    finally:
        w.close()


def atari_param_shapes():
    shapes, cin = [], 4
    for ch in (16, 32, 32):
        shapes += [(ch, cin, 3, 3), (ch,)]
        cin = ch
        for _ in range(2):
            shapes += [(ch, ch, 3, 3), (ch,), (ch, ch, 3, 3), (ch,)]
    shapes += [(256, 3872), (256,), (18, 256 + 18 + 1), (18,), (1, 256 + 18 + 1), (1,)]
    return shapes


@pytest.mark.parametrize("n", WORLDS)
def test_ragged_sizes_and_alignment(n):
    """numel 1..9, odd sizes, a tensor that is an unaligned view, zero-size tensor; flat (single tensor) mode."""
    numels = [1, 2, 3, 4, 5, 7, 9, 1023, 0, 4097, 31]
    w = World(n, 1 << 20)
    try:
        for algo in ALGOS.values():
            ins_np = [[gen_input(7 * r + i + 1, [m], "f32") for i, m in enumerate(numels)] for r in range(n)]
            ins = []
            for r in range(n):
                lst = []
                for a in ins_np[r]:
                    buf = torch.zeros(a.size + 1, device=f"cuda:{r}")
                    v = buf[1:]  # 4-byte aligned, not 16
                    v.copy_(torch.from_numpy(a))
                    lst.append(v)
                ins.append(lst)
            hdrs = [(r + 1, r, 10 * (r + 1)) for r in range(n)]
            dst_like = [torch.empty(m) for m in numels]
            outs, res = run_round(w, ins, hdrs, dst_like, algo)
            flat_in = [flat(ins_np[r], numels) for r in range(n)]
            exact, eh = oracle.allreduce_rankorder(flat_in, hdrs)
            for r in range(n):
                assert flat(outs[r], numels).tobytes() == exact.tobytes()
                assert res[r] == (eh, 0)
    finally:
        w.close()


@pytest.mark.parametrize("n", [m for m in WORLDS if m >= 2])
def test_skip_and_local_accumulation(n):
    """Rank n-1 only skips (empty gradient list, group.h:206-208); rank 0 contributes twice (accumulator.cc:959-975)."""
    numels = [992, 31]
    w = World(n, 1 << 16)
    try:
        for algo in ALGOS.values():
            g0a = [gen_input(500 + i, [m], "f32") for i, m in enumerate(numels)]
            g0b = [gen_input(600 + i, [m], "f32") for i, m in enumerate(numels)]
            ins_np, ins, hdrs = [], [], []
            with torch.cuda.device(0):
                w.ctx[0].stage([torch.from_numpy(a.copy()).cuda() for a in g0a], zero_src=True)
            stage0 = np.zeros(oracle.flat_layout(numels)[1], dtype=np.float32)
            oracle.stage(stage0, [a.copy() for a in g0a])
            oracle.stage(stage0, [a.copy() for a in g0b], accumulate=True)
            for r in range(n):
                if r == 0:
                    ins.append([torch.from_numpy(a.copy()).to("cuda:0") for a in g0b])
                    ins_np.append(stage0)
                    hdrs.append((2, 0, 20))
                elif r == n - 1:
                    ins.append(None)
                    ins_np.append(None)
                    hdrs.append((0, 2, 0))
                else:
                    a = [gen_input(700 + 10 * r + i, [m], "f32") for i, m in enumerate(numels)]
                    ins.append([torch.from_numpy(x.copy()).to(f"cuda:{r}") for x in a])
                    ins_np.append(flat(a, numels))
                    hdrs.append((1, 1, 10))
            # rank 0's second contribution accumulates onto the first
            dsts = []
            for r in range(n):
                with torch.cuda.device(r):
                    if ins[r] is not None:
                        w.ctx[r].stage(ins[r], accumulate=(r == 0), zero_src=True)
                    dsts.append([torch.full((m,), float("nan"), device=f"cuda:{r}") for m in numels])
            for r in range(n):
                with torch.cuda.device(r):
                    w.ctx[r].allreduce(dsts[r], hdr=hdrs[r] + (0 if ins[r] is None else 1,), algo=algo)
            w.sync()
            exact, eh = oracle.allreduce_rankorder(ins_np, hdrs, numel=stage0.size)
            for r in range(n):
                assert flat([t.cpu() for t in dsts[r]], numels).tobytes() == exact.tobytes()
                assert w.ctx[r].result() == (eh, 0)
    finally:
        w.close()


@pytest.mark.parametrize("n", WORLDS)
def test_all_ranks_skip_zeroes_gradients(n):
    w = World(n, 1 << 12)
    try:
        dsts = []
        for r in range(n):
            with torch.cuda.device(r):
                dsts.append([torch.ones(100, device=f"cuda:{r}"), torch.ones(3, device=f"cuda:{r}")])
                w.ctx[r].allreduce(dsts[r], hdr=(0, 1, 0, 0))
        w.sync()
        for r in range(n):
            assert all(not t.any().item() for t in dsts[r])  # accumulator.cc:426-428
            assert w.ctx[r].result() == ((0, n, 0, 0), 0)
    finally:
        w.close()


@pytest.mark.parametrize("n", [m for m in (2, 3, 4, 5, 8) if m <= NGPU])
def test_reference_golden_group_all_reduce(n, golden_dir):
    """The reference's own group.all_reduce outputs (tests/golden/allreduce_golden.npz): ours must be within the
    stated tolerance of them, and bit-identical to the oracle in our order."""
    g = np.load(f"{golden_dir}/allreduce_golden.npz")
    w = World(n, 1 << 16)
    try:
        for gn, rep, seed, numel in g["cases"]:
            if int(gn) != n:
                continue
            numel = int(numel)
            ins_np = [gen_input(int(seed) * 16 + r, [numel], "f32") for r in range(n)]
            ref = g[f"n{n}_r{rep}"]
            for algo in ALGOS.values():
                dsts = []
                for r in range(n):
                    with torch.cuda.device(r):
                        w.ctx[r].stage([torch.from_numpy(ins_np[r].copy()).cuda()])
                        dsts.append(torch.empty(numel, device=f"cuda:{r}"))
                for r in range(n):
                    with torch.cuda.device(r):
                        w.ctx[r].allreduce_flat(dsts[r], scale=False, algo=algo)
                w.sync()
                exact, _ = oracle.allreduce_rankorder(ins_np, [(1, 0, 1)] * n, scale=False)
                tol = oracle.allreduce_tolerance(ins_np, ref)
                for r in range(n):
                    got = dsts[r].cpu().numpy()
                    assert got.tobytes() == exact.tobytes()
                    assert (np.abs(got.astype(np.float64) - ref) <= tol).all()
    finally:
        w.close()


@pytest.mark.parametrize("n", [m for m in (2, 4) if m <= NGPU])
def test_reference_golden_accumulator_rounds(n, golden_dir):
    """Replays the Accumulator rounds recorded from the reference (plain, local accumulation, skipping peer)."""
    g = np.load(f"{golden_dir}/accumulator_golden.npz")
    numels = [31 * 32, 31]
    w = World(n, 1 << 14)
    try:
        for rec in g["rounds"]:
            tag, gn, plan, vbs, ngrad, nskip, bsz = ast.literal_eval(str(rec))
            if gn != n:
                continue
            maxc = max(len(p) for p in plan)
            dsts, hdrs, contrib, staged = [], [], [], []
            for r in range(n):
                with torch.cuda.device(r):
                    st_ = np.zeros(31 * 32 + 31, dtype=np.float32) if plan[r] else None
                    for c, seed in enumerate(plan[r]):
                        gw = torch.from_numpy(gen_input(seed, [31, 32], "f32")).cuda()
                        gb = torch.from_numpy(gen_input(seed + 1, [31], "f32")).cuda()
                        w.ctx[r].stage([gw, gb], accumulate=c > 0, zero_src=True)
                        st_ += np.concatenate([gen_input(seed, [992], "f32"), gen_input(seed + 1, [31], "f32")])
                    staged.append(st_)
                    dsts.append([torch.empty(31, 32, device=f"cuda:{r}"), torch.empty(31, device=f"cuda:{r}")])
                    hdrs.append((len(plan[r]), maxc - len(plan[r]), 10 * len(plan[r]), 1 if plan[r] else 0))
                    contrib.append(bool(plan[r]))
            for r in range(n):
                with torch.cuda.device(r):
                    w.ctx[r].allreduce(dsts[r], hdr=hdrs[r])
            w.sync()
            ref = np.concatenate([g[f"{tag}_w"].reshape(-1), g[f"{tag}_b"].reshape(-1)])
            # tolerance model of SURVEY.md section 8(c): 1e-6 * max(|ref|, sum_i |g_i| / numGradients)
            tol = oracle.allreduce_tolerance(staged, ref, 1.0 / max(ngrad, 1))
            for r in range(n):
                got = np.concatenate([t.cpu().numpy().reshape(-1) for t in dsts[r]])
                assert (np.abs(got.astype(np.float64) - ref) <= tol).all(), tag
                rh, st = w.ctx[r].result()
                assert rh[:3] == (ngrad, nskip, bsz) and st == 0
            got0 = np.concatenate([t.cpu().numpy().reshape(-1) for t in dsts[0]])
            for r in range(1, n):
                assert np.concatenate([t.cpu().numpy().reshape(-1) for t in dsts[r]]).tobytes() == got0.tobytes()
    finally:
        w.close()


@pytest.mark.parametrize("n", [m for m in WORLDS if m >= 2])
def test_many_rounds_two_slots_exact_integers(n):
    """Back-to-back rounds on alternating slots without host syncs in between: rank r contributes r+1 everywhere,
    so every element must be exactly N(N+1)/2 * k / N_gradients; exercises the epoch/parity protocol."""
    numel = 300000
    w = World(n, numel * 4, nslots=2)
    try:
        bufs = [torch.empty(numel, device=f"cuda:{r}") for r in range(n)]
        outs = [[torch.empty(numel, device=f"cuda:{r}") for _ in range(12)] for r in range(n)]
        for k in range(12):
            algo = [_lib.MB_AR_ALGO_ONESHOT, _lib.MB_AR_ALGO_TWOSHOT][k % 2]
            for r in range(n):
                with torch.cuda.device(r):
                    bufs[r].fill_(float((r + 1) * (k + 1)))
                    w.ctx[r].stage([bufs[r]], slot=k % 2)
                    w.ctx[r].allreduce_flat(outs[r][k], slot=k % 2, scale=False, algo=algo)
        w.sync()
        for k in range(12):
            exp = float(n * (n + 1) // 2 * (k + 1))
            for r in range(n):
                assert (outs[r][k] == exp).all().item(), (k, r)
    finally:
        w.close()


def test_barrier_timeout_is_reported():
    """A peer that never arrives: the kernel gives up after timeout_ms and the status says so (no hang)."""
    if NGPU < 2:
        pytest.skip("needs 2 GPUs")
    w = World(2, 1 << 12)
    try:
        with torch.cuda.device(0):
            d = torch.zeros(16, device="cuda:0")
            w.ctx[0].allreduce_flat(d, timeout_ms=200)
        torch.cuda.synchronize(0)
        _, st = w.ctx[0].result()
        assert st == _lib.MB_ETIMEOUT
    finally:
        w.close()
