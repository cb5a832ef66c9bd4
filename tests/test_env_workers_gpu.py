"""EnvPool with a CUDA learner: actions are scattered into the worker mailboxes by the device (mb_scatter_actions on
host-mapped shared memory), observation slabs come back pinned, and the Batcher reads them in place."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.environ['MB_ROOT']); sys.path.insert(0, os.path.join(os.environ['MB_ROOT'], 'tests'))
import moolib_b200 as moolib
from moolib_b200 import _C
from envs_for_tests import FrameEnv
g = np.load(os.path.join(os.environ['MB_ROOT'], 'tests', 'golden', 'envpool_golden.npz'))
bs, steps = int(g['bs']), int(g['steps'])
envs = moolib.EnvPool(FrameEnv, batch_size=bs, num_batches=2, num_processes=3)   # forks before CUDA is touched
torch.cuda.init()
rng = np.random.Generator(np.random.PCG64(77))
tb = moolib.Batcher(steps // 2, 'cuda:0')
before = _C.kernel_launches()
stacked = []
for t in range(steps):
    action = torch.from_numpy(rng.integers(0, 18, size=bs, dtype=np.int64)).cuda()   # CUDA action tensor
    fut = envs.step(t % 2, action)
    obs = fut.result()
    if t >= 2 and t % 3 == 0:
        # EnvStepperFuture.result(device=...): every key on the device, read from the pinned slab by ONE launch
        l0 = _C.kernel_launches()
        dev_obs = fut.result(device='cuda:0')
        assert _C.kernel_launches() - l0 == 1
        for k in ('state', 'reward', 'done'):
            assert dev_obs[k].device.type == 'cuda' and dev_obs[k].cpu().numpy().tobytes() == g[f'{k}{t}'].tobytes(), (k, t)
    assert obs['state'].numpy().tobytes() == g[f'state{t}'].tobytes(), t
    assert obs['reward'].numpy().tobytes() == g[f'reward{t}'].tobytes(), t
    assert obs['done'].numpy().tobytes() == g[f'done{t}'].tobytes(), t
    if t >= 2:
        assert obs['state'].is_pinned(), 'slabs must be pinned once CUDA is up'
    if t % 2 == 0:
        tb.stack(obs)                       # pinned host slab -> device batch, ONE launch for state/reward/done
        stacked.append({k: v.clone() for k, v in obs.items()})
torch.cuda.synchronize()
out = tb.get()
for k in ('state', 'reward', 'done'):
    assert out[k].device.type == 'cuda' and out[k].cpu().equal(torch.stack([s[k] for s in stacked])), k
assert _C.kernel_launches() - before >= steps + steps // 2   # scatter per step + stack per even step
print('OK')
"""


def test_envpool_cuda_actions_and_pinned_slabs(tmp_path):
    script = tmp_path / "envpool_gpu.py"
    script.write_text(SCRIPT)
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, MB_ROOT=ROOT), capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
