"""moolib_b200.Batcher with device='cuda:0': the C++ host class drives the sm_100a copy kernels through the C-ABI.
Bit-exact against the reference's golden fixtures and torch.stack / torch.cat."""
import numpy as np
import pytest
import torch

import moolib_b200
from helpers import batcher_trials, gen_input
from moolib_b200 import _C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("src_kind", ["cuda", "pinned", "pageable"])
def test_golden_trials_on_device(golden_dir, src_kind):
    g = np.load(f"{golden_dir}/batcher_golden.npz")
    before = _C.kernel_launches()
    for ti, (mode, size, dim, shape, dt, n, seed) in enumerate(batcher_trials(g)):
        b = moolib_b200.Batcher(size=size, device=DEV, dim=dim)
        k = 0
        for j in range(n):
            x = torch.from_numpy(gen_input(seed * 100 + j, shape, dt))
            x = x.to(DEV) if src_kind == "cuda" else (x.pin_memory() if src_kind == "pinned" else x)
            getattr(b, mode)(x)
            while not b.empty():
                out = b.get()
                assert out.device.type == "cuda"
                assert out.cpu().numpy().tobytes() == g[f"t{ti}_b{k}"].tobytes(), (ti, k)
                k += 1
        assert k == int(g[f"t{ti}_nb"])
    if src_kind != "pageable":
        assert _C.kernel_launches() > before, "the CUDA path must run our kernels, not a torch fallback"


def test_impala_actor_learner_flow():
    """experiment.py:492-529: time-stack T=21 dict items (one launch per step for all 7 leaves), cat into 32-wide
    learner batches along dim 1; compared with torch.stack/cat on the same data."""
    T, B, Bl = 21, 128, 32
    g = torch.Generator(device=DEV)
    g.manual_seed(3)
    tb = moolib_b200.Batcher(size=T, device=DEV, dim=0)
    lb = moolib_b200.Batcher(size=Bl, device=DEV, dim=1)
    steps = []
    before = _C.kernel_launches()
    for t in range(T):
        item = {
            "env_outputs": {
                "state": torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8, device=DEV, generator=g),
                "reward": torch.randn(B, device=DEV, generator=g),
                "done": torch.rand(B, device=DEV, generator=g) < 0.1,
                "prev_action": torch.randint(0, 18, (B,), device=DEV, generator=g),
            },
            "actor_outputs": {
                "policy_logits": torch.randn(B, 18, device=DEV, generator=g),
                "baseline": torch.randn(B, device=DEV, generator=g),
                "action": torch.randint(0, 18, (B,), device=DEV, generator=g),
            },
        }
        steps.append(item)
        tb.stack(item)
    assert _C.kernel_launches() - before == T  # ONE launch per step, 7 leaves each (reference: 7 copy_ per step)
    data = tb.get()
    assert tb.empty()
    for grp in ("env_outputs", "actor_outputs"):
        for k in data[grp]:
            assert data[grp][k].equal(torch.stack([s[grp][k] for s in steps])), (grp, k)
    data["initial_core_state"] = ()
    before = _C.kernel_launches()
    lb.cat(data)
    assert lb.size() == B // Bl
    assert _C.kernel_launches() - before == 1  # all four 32-wide batches of the item in ONE launch (28 jobs)
    for i in range(B // Bl):
        mb = lb.get()
        assert mb["initial_core_state"] == ()
        for grp in ("env_outputs", "actor_outputs"):
            for k in data[grp]:
                assert mb[grp][k].equal(data[grp][k][:, i * Bl:(i + 1) * Bl]), (grp, k, i)


def test_cat_with_carry_on_device():
    b = moolib_b200.Batcher(size=5, device=DEV, dim=0)
    xs = [torch.arange(i * 21, i * 21 + 21, device=DEV, dtype=torch.int64).view(7, 3) for i in range(4)]
    outs = []
    for x in xs:
        b.cat(x)
        while not b.empty():
            outs.append(b.get())
    allx = torch.cat(xs)
    assert len(outs) == 28 // 5
    for i, o in enumerate(outs):
        assert o.equal(allx[i * 5:(i + 1) * 5])


def test_noncontiguous_and_dtype_converting_sources():
    b = moolib_b200.Batcher(size=3, device=DEV, dim=0)
    base = torch.randn(4, 6, device=DEV)
    items = [base.t(), base.t() * 2, base.t() * 3]  # non-contiguous views
    for it in items:
        b.stack(it)
    assert b.get().equal(torch.stack(items))


def test_stack_fields_on_device():
    items = tuple({"o": torch.randn(33, 7, device=DEV), "n": i} for i in range(6))
    before = _C.kernel_launches()
    st = _C.stack_fields(items, 0)
    assert _C.kernel_launches() == before + 1  # one gather launch for the leaf (reference: torch::stack)
    assert st["o"].equal(torch.stack([it["o"] for it in items])) and st["n"] == tuple(range(6))
    st1 = _C.stack_fields(items, 1)
    assert st1["o"].equal(torch.stack([it["o"] for it in items], dim=1))


def _impala_item(B, g):
    return {
        "env_outputs": {
            "state": torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8, device=DEV, generator=g),
            "reward": torch.randn(B, device=DEV, generator=g),
            "done": torch.rand(B, device=DEV, generator=g) < 0.1,
            "prev_action": torch.randint(0, 18, (B,), device=DEV, generator=g),
        },
        "actor_outputs": {
            "policy_logits": torch.randn(B, 18, device=DEV, generator=g),
            "baseline": torch.randn(B, device=DEV, generator=g),
            "action": torch.randint(0, 18, (B,), device=DEV, generator=g),
        },
    }


@pytest.mark.parametrize("B,Bl", [(256, 32), (96, 64)])
def test_unroll_batcher_full_size_one_launch(B, Bl):
    """BASELINE configs[1] shapes: T=21 steps of a 256-env slab gathered straight into 8 x [21,32,...] learner batches
    by ONE launch (21 x 7 x 8 = 1176 pitched copies in a device-resident table), bit-exact against torch.stack/cat.
    (96, 64): the carry across unrolls -- a learner batch is completed by the next unroll."""
    T, U = 21, 2
    g = torch.Generator(device=DEV)
    g.manual_seed(B)
    ub = moolib_b200.UnrollBatcher(T, Bl, DEV, cat_dim=1)
    got, stacked = [], []
    for u in range(U):
        steps = [_impala_item(B, g) for _ in range(T)]
        before = _C.kernel_launches()
        for t, it in enumerate(steps):
            if t == T - 1:
                ub.set_extra("initial_core_state", (torch.full((2, B, 8), float(u), device=DEV),))
            ub.stack(it)
        assert _C.kernel_launches() - before == 1, "one launch per unroll"
        stacked.append(steps)
        while not ub.empty():
            got.append(ub.get())
    assert len(got) == (U * B) // Bl
    cols = {}
    for grp in ("env_outputs", "actor_outputs"):
        for k in stacked[0][0][grp]:
            cols[(grp, k)] = torch.cat([torch.stack([s[grp][k] for s in steps]) for steps in stacked], dim=1)
    core = torch.cat([torch.full((2, B, 8), float(u), device=DEV) for u in range(U)], dim=1)
    for i, mb in enumerate(got):
        for (grp, k), full in cols.items():
            assert mb[grp][k].equal(full[:, i * Bl:(i + 1) * Bl]), (grp, k, i)
        assert mb["initial_core_state"][0].equal(core[:, i * Bl:(i + 1) * Bl])


def test_copy_table_through_the_c_abi_matches_oracle():
    """mb_copy2d_table: 1500 random pitched copies (every alignment class, bulk and tiny rows mixed) in one launch from a
    device-resident table, byte-exact against oracle_copy2d."""
    import oracle
    from moolib_b200 import _lib

    rng = np.random.Generator(np.random.PCG64(77))
    src = rng.integers(0, 256, size=128 << 20, dtype=np.uint8)
    dst_exp = np.zeros(128 << 20, dtype=np.uint8)
    s_dev = torch.from_numpy(src).to(DEV)
    d_dev = torch.zeros(128 << 20, dtype=torch.uint8, device=DEV)
    jobs, dpos, spos = [], 0, 0
    for i in range(1500):
        kind = i % 5
        if kind == 0:
            rb, rows = int(rng.integers(65536, 400000)) // 16 * 16, 1
        elif kind == 1:
            rb, rows = int(rng.integers(2048, 9000)) // 16 * 16, int(rng.integers(1, 6))
        elif kind == 2:
            rb, rows = int(rng.integers(1, 300)), int(rng.integers(1, 40))
        elif kind == 3:
            rb, rows = int(rng.integers(16, 600)) // 16 * 16, int(rng.integers(1, 30))
        else:
            rb, rows = int(rng.integers(1, 5000)), 1
        sp = rb + int(rng.integers(0, 3)) * 16
        dp = rb + int(rng.integers(0, 3)) * 16
        skew_s, skew_d = (int(rng.integers(0, 16)), int(rng.integers(0, 16))) if kind in (2, 4) else (0, 0)
        so, do = (spos + 255) // 256 * 256 + skew_s, (dpos + 255) // 256 * 256 + skew_d
        spos, dpos = so + sp * rows, do + dp * rows
        assert spos < src.size and dpos < src.size
        jobs.append((so, do, rb, rows, sp, dp))
        oracle.copy2d(src, so, dst_exp, do, rb, rows, sp, dp)
    ctx = _lib.CopyContext(0, max_jobs=4096)
    try:
        arr = _lib.make_jobs([(s_dev.data_ptr() + so, d_dev.data_ptr() + do, rb, rows, sp, dp)
                              for so, do, rb, rows, sp, dp in jobs])
        assert ctx.copy(arr, _lib.MB_SRC_DEVICE) == 1
        torch.cuda.synchronize()
        assert d_dev.cpu().numpy().tobytes() == dst_exp.tobytes()
        # and again with a table longer than the context's capacity (split into two launches), sources "unknown"
        d_dev.zero_()
        small = _lib.CopyContext(0, max_jobs=1000)
        assert small.copy(arr) == 2
        torch.cuda.synchronize()
        assert d_dev.cpu().numpy().tobytes() == dst_exp.tobytes()
        small.close()
    finally:
        ctx.close()


def test_to_device_reads_pinned_nest_in_one_launch():
    g = torch.Generator().manual_seed(5)
    host = {"state": torch.randint(0, 256, (64, 4, 84, 84), dtype=torch.uint8, generator=g).pin_memory(),
            "reward": torch.randn(64, generator=g).pin_memory(), "done": (torch.rand(64, generator=g) < 0.5).pin_memory(),
            "note": "x", "dev": torch.arange(5, device=DEV)}
    before = _C.kernel_launches()
    out = moolib_b200.to_device(host, DEV)
    assert _C.kernel_launches() - before == 1
    assert out["note"] == "x" and out["dev"].data_ptr() == host["dev"].data_ptr()
    for k in ("state", "reward", "done"):
        assert out[k].device.type == "cuda" and out[k].cpu().equal(host[k])
    # pageable and non-contiguous sources take ATen's path and still arrive intact
    odd = {"p": torch.randn(33, 7), "nc": torch.randn(16, 9).pin_memory().t()}
    out = moolib_b200.to_device(odd, DEV)
    assert out["p"].cpu().equal(odd["p"]) and out["nc"].cpu().equal(odd["nc"])


def test_non_contiguous_pinned_source_is_not_read_by_the_kernel():
    """A transposed view of a pinned tensor: making it contiguous would produce pageable memory, so the copy must go
    through ATen (round-1 advisor finding) -- and still be right."""
    x = torch.randn(12, 40).pin_memory().t()  # [40, 12], not contiguous
    b = moolib_b200.Batcher(size=3, device=DEV, dim=0)
    for _ in range(3):
        b.stack({"x": x, "y": torch.arange(7).pin_memory()})
    out = b.get()
    assert out["x"].cpu().equal(torch.stack([x, x, x])) and out["y"].cpu().equal(torch.stack([torch.arange(7)] * 3))
