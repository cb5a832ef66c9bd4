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
