"""HP-B parity on the GPU: the copy kernels, called through the C-ABI, against the CPU oracle and the golden
fixtures generated from the reference.  Bit-exact (byte movement)."""
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import batcher_trials, gen_input
from moolib_b200 import _lib

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _impls():
    return ["ldg", "tma"]


def _run_jobs(jobs_np, src_buf, dst_size, impl_env=None):
    """jobs_np: list of (src_off, dst_off, row_bytes, rows, src_pitch, dst_pitch) on flat uint8 buffers."""
    src_d = torch.from_numpy(src_buf).to(DEV)
    dst_d = torch.full((dst_size,), 0xA5, dtype=torch.uint8, device=DEV)
    jobs = [(src_d.data_ptr() + so, dst_d.data_ptr() + do, rb, rows, sp, dp) for so, do, rb, rows, sp, dp in jobs_np]
    _lib.copy2d_batch(jobs)
    torch.cuda.synchronize()
    exp = np.full(dst_size, 0xA5, dtype=np.uint8)
    for so, do, rb, rows, sp, dp in jobs_np:
        oracle.copy2d(src_buf, so, exp, do, rb, rows, sp, dp)
    return dst_d.cpu().numpy(), exp


@pytest.mark.parametrize("seed", range(6))
def test_random_pitched_jobs_match_oracle(seed):
    """Random job tables: every alignment class, tiny to multi-tile rows, pitches with gaps; untouched bytes stay."""
    rng = np.random.Generator(np.random.PCG64(seed))
    # 1..64: the 4 KiB parameter block; 65..512: the large parameter space; > 512: the multi-launch split
    njobs = [int(rng.integers(1, 64)), int(rng.integers(65, 200)), int(rng.integers(513, 700))][seed % 3]
    src_buf = rng.integers(0, 256, size=1 << 22, dtype=np.uint8)
    jobs, dst_cursor = [], 0
    for _ in range(njobs):
        kind = int(rng.integers(0, 5))
        if kind == 0:
            rb, rows = int(rng.integers(1, 64)), int(rng.integers(1, 40))
        elif kind == 1:
            rb, rows = int(rng.integers(1, 5000)), int(rng.integers(1, 12))
        elif kind == 2:
            rb, rows = int(rng.integers(16000, 70000)), int(rng.integers(1, 4))
        elif kind == 3:
            rb, rows = 16 * int(rng.integers(1, 3000)), int(rng.integers(1, 6))
        else:
            rb, rows = 28224, int(rng.integers(1, 8))
        align = [1, 2, 4, 8, 16][int(rng.integers(0, 5))] if kind != 3 else 16
        sp = rb + align * int(rng.integers(0, 9))
        dp = rb + align * int(rng.integers(0, 9))
        so = align * int(rng.integers(0, ((1 << 22) - sp * rows) // align))
        if dst_cursor > (200 << 20):
            break
        do = (dst_cursor + 15) // 16 * 16 + (0 if align == 16 else int(rng.integers(0, 16)))
        dst_cursor = do + dp * rows
        jobs.append((so, do, rb, rows, sp, dp))
    out, exp = _run_jobs(jobs, src_buf, dst_cursor + 64)
    assert out.tobytes() == exp.tobytes()


def test_empty_and_degenerate_jobs():
    assert _lib.copy2d_batch([]) == 0
    src = torch.arange(256, dtype=torch.uint8, device=DEV)
    dst = torch.zeros(256, dtype=torch.uint8, device=DEV)
    # zero rows / zero bytes are no-ops, a 1-byte job works
    n = _lib.copy2d_batch([(src.data_ptr(), dst.data_ptr(), 0, 5, 0, 0), (src.data_ptr(), dst.data_ptr(), 7, 0, 7, 7),
                           (src.data_ptr() + 3, dst.data_ptr() + 9, 1, 1, 1, 1)])
    torch.cuda.synchronize()
    assert n == 1
    exp = torch.zeros(256, dtype=torch.uint8)
    exp[9] = 3
    assert dst.cpu().equal(exp)


@pytest.mark.parametrize("impl", _impls())
def test_golden_batcher_trials_through_c_abi(golden_dir, impl, monkeypatch):
    """Replays the reference's Batcher trials with mb_stack_slot / mb_cat_narrow doing every copy_ (moolib.cc:676,
    745-751).  MB_COPY_IMPL is read once per process, so the two implementations run in separate subprocesses."""
    import subprocess
    import sys
    code = f"""
import sys, numpy as np, torch
sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r}); sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
from helpers import batcher_trials, gen_input
from moolib_b200 import _lib
g = np.load({golden_dir!r} + '/batcher_golden.npz')
DT = {{'u8': torch.uint8, 'f32': torch.float32, 'i64': torch.int64, 'bool': torch.bool}}
for ti, (mode, size, dim, shape, dt, n, seed) in enumerate(batcher_trials(g)):
    items = [torch.from_numpy(gen_input(seed * 100 + j, shape, dt)).cuda() for j in range(n)]
    nb, k = int(g[f't{{ti}}_nb']), 0
    if mode == 'stack':
        bshape = shape[:dim] + [size] + shape[dim:]
        for k in range(nb):
            dst = torch.empty(bshape, dtype=DT[dt], device='cuda')
            for s in range(size):
                _lib.stack_slot(dst, s, items[k * size + s], dim)
            torch.cuda.synchronize()
            assert dst.cpu().numpy().tobytes() == g[f't{{ti}}_b{{k}}'].tobytes(), (ti, k)
    else:
        bshape = list(shape); bshape[dim] = size
        dst = torch.empty(bshape, dtype=DT[dt], device='cuda'); fill = 0
        for it in items:
            off, m = 0, it.shape[dim]
            while off < m:
                take = min(m - off, size - fill)
                _lib.cat_narrow(dst, fill, it, off, take, dim)
                fill += take; off += take
                if fill == size:
                    torch.cuda.synchronize()
                    assert dst.cpu().numpy().tobytes() == g[f't{{ti}}_b{{k}}'].tobytes(), (ti, k)
                    k += 1; fill = 0
                    dst = torch.empty(bshape, dtype=DT[dt], device='cuda')
        assert k == nb
print('OK')
"""
    env = dict(os.environ, MB_COPY_IMPL=impl)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("impl", _impls())
def test_impala_time_stack_and_cat_full_size(impl):
    """BASELINE.json config 2 sizes (256 envs x 84x84x4, T=21 -> 32-wide learner batches) checked through
    size-independent properties: stack == torch.stack, cat == torch.cat, round trip stack->cat->unstack."""
    import subprocess
    import sys
    code = f"""
import sys, torch
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
from moolib_b200 import _lib
T, B, Bl = 21, 256, 32
g = torch.Generator(device='cuda'); g.manual_seed(1234)
steps = [torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8, device='cuda', generator=g) for _ in range(T)]
rew = [torch.randn(B, device='cuda', generator=g) for _ in range(T)]
tb = torch.empty((T, B, 4, 84, 84), dtype=torch.uint8, device='cuda'); tr = torch.empty((T, B), device='cuda')
for t in range(T):
    jobs = [(steps[t].data_ptr(), tb[t].data_ptr(), steps[t].numel(), 1, steps[t].numel(), steps[t].numel()),
            (rew[t].data_ptr(), tr[t].data_ptr(), B * 4, 1, B * 4, B * 4)]
    _lib.copy2d_batch(jobs)
torch.cuda.synchronize()
assert tb.equal(torch.stack(steps)) and tr.equal(torch.stack(rew))
for k in range(B // Bl):
    lb = torch.empty((T, Bl, 4, 84, 84), dtype=torch.uint8, device='cuda'); lr = torch.empty((T, Bl), device='cuda')
    _lib.cat_narrow(lb, 0, tb, k * Bl, Bl, 1)
    _lib.cat_narrow(lr, 0, tr, k * Bl, Bl, 1)
    torch.cuda.synchronize()
    assert lb.equal(tb[:, k * Bl:(k + 1) * Bl]) and lr.equal(tr[:, k * Bl:(k + 1) * Bl])
print('OK')
"""
    env = dict(os.environ, MB_COPY_IMPL=impl)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_gather_rows_pointer_array():
    """K-B1/K-B4: rows scattered in device memory gathered into a contiguous batch == torch.stack."""
    g = torch.Generator(device=DEV)
    g.manual_seed(7)
    for row_bytes, nrows in [(28224, 256), (28224, 64), (4, 33), (28229, 17), (100000, 5), (512, 4096)]:
        rows = [torch.randint(0, 256, (row_bytes,), dtype=torch.uint8, device=DEV, generator=g) for _ in range(nrows)]
        ptrs = torch.tensor([r.data_ptr() for r in rows], dtype=torch.int64, device=DEV)
        dst = torch.empty((nrows, row_bytes), dtype=torch.uint8, device=DEV)
        _lib.gather_rows(dst, ptrs, row_bytes, nrows)
        torch.cuda.synchronize()
        assert dst.equal(torch.stack(rows)), (row_bytes, nrows)


def test_host_mapped_source_rows():
    """Pinned host slab (the EnvPool shm layout: [maxEnvs, 28224] u8 + reward f32 + done u8) read by the kernel."""
    B = 64
    state = torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8).pin_memory()
    reward = torch.randn(B).pin_memory()
    done = (torch.rand(B) < 0.1).pin_memory()
    d_state = torch.empty_like(state, device=DEV)
    d_reward = torch.empty_like(reward, device=DEV)
    d_done = torch.empty_like(done, device=DEV)
    jobs = [(s.data_ptr(), d.data_ptr(), s.numel() * s.element_size(), 1, 0, 0)
            for s, d in ((state, d_state), (reward, d_reward), (done, d_done))]
    assert _lib.copy2d_batch(jobs) == 1  # one launch for all three leaves
    torch.cuda.synchronize()
    assert d_state.cpu().equal(state) and d_reward.cpu().equal(reward) and d_done.cpu().equal(done)


def test_scatter_actions_host_mailboxes():
    n = 300
    counters = torch.zeros(n * 2, dtype=torch.int32).pin_memory()
    counters[::2] = torch.arange(n, dtype=torch.int32) * 7
    exp = counters.numpy().view(np.uint32).copy()
    acts = torch.randint(0, 18, (n,), dtype=torch.int64, device=DEV)
    oracle.scatter_actions(exp, acts.cpu().numpy(), stride=2)
    L = _lib.load()
    import ctypes
    _lib.check(L.mb_scatter_actions(counters.data_ptr(), 2, acts.data_ptr(), n,
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert counters.numpy().view(np.uint32).tobytes() == exp.tobytes()
