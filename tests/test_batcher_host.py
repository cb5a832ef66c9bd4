"""Host logic of the Batcher class (C++ `moolib_b200._C.Batcher`) on CPU tensors: same results as the reference's
golden fixtures and the same control flow / error strings as src/moolib.cc:595-845.  (The byte movement of CUDA
batchers is covered by tests/test_batcher_gpu.py.)"""
import random

import numpy as np
import pytest
import torch

import moolib_b200
import oracle
from helpers import batcher_trials, gen_input


def _replay(make, trial):
    mode, size, dim, shape, dt, n, seed = trial
    b = make(size, dim)
    outs = []
    for j in range(n):
        x = torch.from_numpy(gen_input(seed * 100 + j, shape, dt))
        getattr(b, mode)(x)
        while not b.empty():
            outs.append(b.get())
    return outs


def test_golden_trials_cpu(golden_dir):
    g = np.load(f"{golden_dir}/batcher_golden.npz")
    for ti, trial in enumerate(batcher_trials(g)):
        outs = _replay(lambda size, dim: moolib_b200.Batcher(size=size, dim=dim), trial)
        assert len(outs) == int(g[f"t{ti}_nb"])
        for k, o in enumerate(outs):
            assert o.numpy().tobytes() == g[f"t{ti}_b{k}"].tobytes(), (trial, k)


def test_reference_unit_test_semantics():
    """The reference's own test (test/unit/test_batcher.py:13-52) restated against our class."""
    rnd = random.Random(7)
    for _ in range(64):
        size, dim = rnd.randint(1, 20), rnd.randint(0, 2)
        dims = rnd.randint(dim + 1, dim + 2)
        n = rnd.randint(20, 60)
        shape = [rnd.randint(1, 4) for _ in range(dims)]
        batcher = moolib_b200.Batcher(size=size, dim=dim)
        inputs = []
        for _ in range(n):
            x = torch.randn(shape)
            inputs.append(x.clone())
            batcher.stack(x)
            assert x.equal(inputs[-1])
            if not batcher.empty():
                assert batcher.get().equal(torch.stack(inputs, dim=dim))
                inputs = []
        batcher = moolib_b200.Batcher(size=size, dim=dim)
        inputs = []
        for _ in range(n):
            x = torch.randn(shape)
            inputs.append(x.clone())
            batcher.cat(x)
            assert x.equal(inputs[-1])
            while not batcher.empty():
                batched = batcher.get()
                catted = torch.cat(inputs, dim=dim)
                overflow = catted.narrow(dim, size, catted.size(dim) - size)
                assert batched.equal(catted.narrow(dim, 0, size))
                inputs = [overflow] if overflow.size(dim) > 0 else []


def test_nested_structures_and_passthrough(golden_dir):
    g = np.load(f"{golden_dir}/batcher_golden.npz")
    b = moolib_b200.Batcher(size=3, dim=0)
    for j in range(3):
        b.stack({
            "a": torch.from_numpy(gen_input(7000 + j, [2, 3], "f32")),
            "n": (torch.from_numpy(gen_input(7100 + j, [4], "u8")), [torch.from_numpy(gen_input(7200 + j, [1], "i64"))]),
            "tag": "first" if j == 0 else "later",
        })
    r = b.get()
    assert isinstance(r["n"], tuple) and isinstance(r["n"][1], list)
    assert r["a"].numpy().tobytes() == g["nested_a"].tobytes()
    assert r["n"][0].numpy().tobytes() == g["nested_n0"].tobytes()
    assert r["n"][1][0].numpy().tobytes() == g["nested_n1"].tobytes()
    assert r["tag"] == str(g["nested_tag"]) == "first"  # non-tensor leaves come from the first item (moolib.cc:689)


def test_error_behaviour_matches_reference():
    b = moolib_b200.Batcher(size=4, dim=0)
    b.stack(torch.zeros(2))
    with pytest.raises(RuntimeError, match="Previously called with stack; cannot mix cat/stack"):
        b.cat(torch.zeros(2))
    b = moolib_b200.Batcher(size=4, dim=0)
    b.cat(torch.zeros(2))
    with pytest.raises(RuntimeError, match="Previously called with cat; cannot mix cat/stack"):
        b.stack(torch.zeros(2))
    with pytest.raises(RuntimeError, match="Given input tensor with 1 dimensions, cannot cat in dimension 2"):
        moolib_b200.Batcher(size=4, dim=2).cat(torch.zeros(2))
    with pytest.raises(RuntimeError, match="Given input tensor with 0 dimensions, cannot stack in dimension 2"):
        moolib_b200.Batcher(size=4, dim=2).stack(torch.zeros(()))
    b = moolib_b200.Batcher(size=4, dim=0)
    b.stack({"a": torch.zeros(2), "b": torch.zeros(2)})
    with pytest.raises(RuntimeError, match="type mismatch in batch operation"):
        b.stack({"a": torch.zeros(2), "b": 3})
    b = moolib_b200.Batcher(size=8, dim=0)
    b.cat({"a": torch.zeros(2, 3), "b": torch.zeros(2)})
    with pytest.raises(RuntimeError, match="all tensors must have the same size in the batch dimension"):
        b.cat({"a": torch.zeros(2, 3), "b": torch.zeros(3)})


def test_matches_live_oracle_batcher_randomised():
    rnd = random.Random(11)
    for _ in range(40):
        size, dim = rnd.randint(1, 9), rnd.randint(0, 1)
        shape = [rnd.randint(1, 5) for _ in range(dim + rnd.randint(1, 2))]
        for mode in ("stack", "cat"):
            a, b = moolib_b200.Batcher(size=size, dim=dim), oracle.OracleBatcher(size, dim=dim)
            for j in range(rnd.randint(5, 30)):
                item = {"x": torch.randn(shape), "y": [torch.randint(0, 255, shape, dtype=torch.uint8)]}
                getattr(a, mode)(item)
                getattr(b, mode)({"x": item["x"].clone(), "y": [item["y"][0].clone()]})
                assert a.size() == len(b.queue)
                while not a.empty():
                    ra, rb = a.get(), b.get()
                    assert ra["x"].equal(rb["x"]) and ra["y"][0].equal(rb["y"][0])


def test_stack_and_unstack_fields_roundtrip():
    from moolib_b200 import _C
    items = tuple({"o": torch.randn(3, 2), "t": (torch.arange(4) + i, "s%d" % i)} for i in range(5))
    st = _C.stack_fields(items, 0)
    assert st["o"].equal(torch.stack([it["o"] for it in items]))
    assert st["t"][0].equal(torch.stack([it["t"][0] for it in items]))
    assert st["t"][1] == tuple("s%d" % i for i in range(5))  # non-tensor leaves become a tuple of N (batch_utils.cc:288)
    un = _C.unstack_fields(st, 5, 0)
    for i in range(5):
        assert un[i]["o"].equal(items[i]["o"]) and un[i]["t"][0].equal(items[i]["t"][0]) and un[i]["t"][1] == "s%d" % i
    # N == 1 fast path: unsqueeze / squeeze (batch_utils.cc:261-263, 318-320)
    one = _C.stack_fields((items[0],), 0)
    assert one["o"].shape == (1, 3, 2) and one["t"][1] == ("s0",)
    back = _C.unstack_fields(one, 1, 0)[0]
    assert back["o"].equal(items[0]["o"]) and back["t"][1] == "s0"


def _unroll_reference(items, extras, T, Bl, device="cpu"):
    """Batcher(T).stack x T followed by Batcher(Bl, dim=1).cat -- the two-pass composition UnrollBatcher fuses."""
    tb, lb = moolib_b200.Batcher(T, device), moolib_b200.Batcher(Bl, device, dim=1)
    outs, k = [], 0
    for it in items:
        tb.stack(it)
        if not tb.empty():
            data = tb.get()
            if extras is not None:
                data["extra"] = extras[k]
            k += 1
            lb.cat(data)
            while not lb.empty():
                outs.append(lb.get())
    return outs


def _same_nest(a, b):
    if isinstance(a, dict):
        return isinstance(b, dict) and list(a) == list(b) and all(_same_nest(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return type(a) is type(b) and len(a) == len(b) and all(_same_nest(x, y) for x, y in zip(a, b))
    if isinstance(a, torch.Tensor):
        return a.dtype == b.dtype and a.shape == b.shape and a.equal(b)
    return a == b


@pytest.mark.parametrize("B,Bl", [(8, 4), (6, 4), (5, 7), (4, 4), (3, 1)])
def test_unroll_batcher_equals_stack_then_cat(B, Bl):
    """Host logic of UnrollBatcher on CPU tensors (carry across unrolls when B is not a multiple of the learner batch,
    nested items, pass-through leaves, the extra nest that is concatenated only)."""
    T, U = 5, 4
    g = torch.Generator().manual_seed(B * 100 + Bl)
    items, extras = [], []
    for u in range(U):
        for t in range(T):
            items.append({"obs": {"state": torch.randint(0, 255, (B, 2, 3), dtype=torch.uint8, generator=g),
                                  "reward": torch.randn(B, generator=g)},
                          "act": [torch.randint(0, 9, (B,), generator=g), (torch.randn(B, 4, generator=g),)],
                          "tag": f"u{u}t{t}"})
        extras.append((torch.randn(2, B, 3, generator=g), torch.randn(1, B, generator=g)))
    ub = moolib_b200.UnrollBatcher(T, Bl, "cpu", cat_dim=1)
    got = []
    for i, it in enumerate(items):
        if i % T == T - 1:
            ub.set_extra("extra", extras[i // T])
        ub.stack(it)
        while not ub.empty():
            got.append(ub.get())
    exp = _unroll_reference(items, extras, T, Bl)
    assert len(got) == len(exp) == (U * B) // Bl
    for a, b in zip(got, exp):
        assert _same_nest(a, b)
        assert a["obs"]["state"].shape == (T, Bl, 2, 3) and a["extra"][0].shape == (2, Bl, 3)


def test_unroll_batcher_errors():
    with pytest.raises(RuntimeError, match="cat_dim must be >= 1"):
        moolib_b200.UnrollBatcher(3, 2, "cpu", cat_dim=0)
    ub = moolib_b200.UnrollBatcher(3, 2, "cpu")
    ub.stack({"a": torch.zeros(4, 2)})
    with pytest.raises(RuntimeError, match="same shapes and dtypes"):
        ub.stack({"a": torch.zeros(4, 3)})
    with pytest.raises(RuntimeError, match="type mismatch in batch operation"):
        ub.stack({"a": 3})
    ub = moolib_b200.UnrollBatcher(2, 2, "cpu")
    ub.stack({"a": torch.zeros(4), "b": torch.zeros(3)})
    with pytest.raises(RuntimeError, match="all tensors must have the same size in the batch dimension"):
        ub.stack({"a": torch.zeros(4), "b": torch.zeros(3)})


def test_to_device_cpu_is_identity_nest():
    nest = {"a": torch.arange(4), "b": (torch.ones(2), "s")}
    out = moolib_b200.to_device(nest, "cpu")
    assert _same_nest(out, nest) and out["b"][1] == "s"


def test_failed_calls_leave_no_stale_work_behind():
    """An item that fails half-way (ATen refuses a leaf after other leaves were already described) must not leave copies
    queued for the next call, and an UnrollBatcher that cannot finish an unroll starts the next one cleanly."""
    b = moolib_b200.Batcher(size=2, dim=0)
    b.stack({"a": torch.zeros(3), "b": torch.zeros(4)})
    with pytest.raises(RuntimeError):
        b.stack({"a": torch.ones(3), "b": torch.ones(5)})  # leaf b: shape mismatch inside copy_
    b.stack({"a": torch.full((3,), 2.0), "b": torch.full((4,), 2.0)})
    out = b.get()
    assert out["a"][1].eq(2).all() and out["b"][1].eq(2).all()
    ub = moolib_b200.UnrollBatcher(2, 2, "cpu")
    ub.set_extra("k", (torch.zeros(1, 4),))
    ub.stack([torch.zeros(4)])
    with pytest.raises(RuntimeError, match="set_extra needs dict items"):
        ub.stack([torch.ones(4)])
    assert ub.empty()
    for v in (3.0, 4.0):
        ub.stack([torch.full((4,), v)])
    got = [ub.get(), ub.get()]
    assert ub.empty() and all(g[0].shape == (2, 2) and g[0][0].eq(3).all() and g[0][1].eq(4).all() for g in got)
