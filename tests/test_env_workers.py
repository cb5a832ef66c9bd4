"""moolib_b200.EnvPool on CPU: the reference's own known-answer test (test/unit/test_envpool.py:39-88) restated, the
golden outputs recorded from the reference, and the argument errors.  Worker processes are real forked processes."""
import os
import random
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import moolib_b200 as moolib
from envs_for_tests import FrameEnv, ToyEnv


@pytest.mark.timeout(300)
def test_known_answer_toy_env():
    bs = 32
    envs = moolib.EnvPool(ToyEnv, batch_size=bs, num_batches=2, num_processes=4)
    with pytest.raises(RuntimeError):
        envs.step(0, torch.zeros(bs))
    with pytest.raises(RuntimeError):
        envs.step(0, torch.zeros(bs + 1).long())
    with pytest.raises(RuntimeError, match="out-of-range batch index"):
        envs.step(2, torch.zeros(bs).long())
    z = torch.zeros(bs).long()
    initial = torch.ones(4, 4)
    initial[0][0], initial[3][1], initial[1][2] = 4.0, 0.5, 0.25
    initial = initial.expand(bs, 4, 4)
    obs = envs.step(batch_index=0, action=z).result()
    assert obs["n"].equal(initial) and obs["n"].dtype == torch.float32
    assert obs["done"].dtype == torch.bool and obs["reward"].dtype == torch.float32
    fut0 = envs.step(batch_index=0, action=z + 1)
    fut1 = envs.step(batch_index=1, action=z)
    with pytest.raises(RuntimeError, match="twice concurrently"):
        envs.step(0, z)
    assert fut0.result()["n"].equal(initial * 2)
    assert fut1.result()["n"].equal(initial)
    assert envs.step(batch_index=0, action=z + 2).result()["n"].equal(initial)
    states = [initial.clone(), initial.clone()]
    rnd = random.Random(5)
    for _ in range(100):
        index = rnd.randint(0, 1)
        action = torch.randint(0, 3, [bs])
        s = states[index]
        obs = envs.step(index, action).result()
        for i in range(bs):
            if action[i] == 1:
                s[i] *= 2
            elif action[i] == 2:
                s[i] /= 2
            r = abs(s[i].sum() - 4)
            d = s[i].sum() < 1
            if d:
                s[i] = initial[i]
            assert d == obs["done"][i]
            assert s[i].equal(obs["n"][i])
            assert r == obs["reward"][i]


@pytest.mark.timeout(300)
def test_matches_reference_golden(golden_dir):
    """Seeded action sequence through FrameEnv (84x84x4 u8 observations): every step's state/reward/done equals what the
    reference's EnvPool produced (tests/golden/envpool_golden.npz)."""
    g = np.load(f"{golden_dir}/envpool_golden.npz")
    bs, steps = int(g["bs"]), int(g["steps"])
    envs = moolib.EnvPool(FrameEnv, batch_size=bs, num_batches=2, num_processes=3)
    rng = np.random.Generator(np.random.PCG64(77))
    for t in range(steps):
        index = t % 2
        action = torch.from_numpy(rng.integers(0, 18, size=bs, dtype=np.int64))
        obs = envs.step(index, action).result()
        assert set(obs) == {"state", "reward", "done"}
        assert obs["state"].numpy().tobytes() == g[f"state{t}"].tobytes(), t
        assert obs["reward"].numpy().tobytes() == g[f"reward{t}"].tobytes(), t
        assert obs["done"].numpy().tobytes() == g[f"done{t}"].tobytes(), t


@pytest.mark.timeout(120)
def test_env_exception_is_reported():
    envs = moolib.EnvPool(ToyEnv, batch_size=4, num_batches=1, num_processes=2)
    envs.step(0, torch.zeros(4).long()).result()
    with pytest.raises(RuntimeError, match="Error in env"):
        envs.step(0, torch.full((4,), 9).long()).result()  # ToyEnv raises on action 9


@pytest.mark.timeout(300)
def test_envpool_feeds_batcher_like_the_vtrace_loop():
    """EnvPool -> time Batcher -> learn Batcher on CPU: the data path of examples/vtrace/experiment.py:489-523 with real
    worker processes; the batches equal torch.stack / narrow of what the envs produced."""
    B, T, Bl = 6, 4, 3
    envs = moolib.EnvPool(FrameEnv, batch_size=B, num_batches=1, num_processes=2)
    tb = moolib.Batcher(T, "cpu")
    lb = moolib.Batcher(Bl, "cpu", dim=1)
    rng = np.random.Generator(np.random.PCG64(5))
    seen = []
    action = torch.zeros(B, dtype=torch.int64)
    for t in range(T):
        obs = envs.step(0, action).result()
        step = {k: v.clone() for k, v in obs.items()}  # result() aliases the slab until the next step()
        seen.append(step)
        tb.stack(step)
        action = torch.from_numpy(rng.integers(0, 18, size=B, dtype=np.int64))
    data = tb.get()
    for k in ("state", "reward", "done"):
        assert data[k].equal(torch.stack([s[k] for s in seen]))
    lb.cat(data)
    assert lb.size() == B // Bl
    for i in range(B // Bl):
        mb = lb.get()
        assert mb["state"].shape == (T, Bl, 4, 84, 84)
        assert mb["state"].equal(data["state"][:, i * Bl:(i + 1) * Bl])
