"""Cross-PROCESS HP-A: one process per GPU, CUDA IPC handles exchanged over gloo (the torchrun layout of bench.py),
results checked against the CPU oracle in every rank."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0

WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ['MB_ROOT']); sys.path.insert(0, os.path.join(os.environ['MB_ROOT'], 'tests'))
import oracle
from helpers import gen_input
from moolib_b200 import _lib, peer
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('gloo')
numels = [992, 31, 5, 300001]
total = _lib.flat_numel(numels)
# a small context first: its 64 KiB sync block must not end up sub-allocated next to the next context's (CUDA IPC maps
# whole 2 MiB blocks -- regression test for handles that were opened at the wrong offset)
small = peer.make_context(4096, nslots=1)
ctx = peer.make_context(total * 4, nslots=2)
one = torch.full((8,), float(rank + 1), device='cuda')
small.stage([one]); out8 = torch.empty(8, device='cuda'); small.allreduce_flat(out8)
torch.cuda.synchronize()
assert (out8 == world * (world + 1) / 2).all().item() and small.result()[1] == 0
offs, _ = oracle.flat_layout(numels)
for rnd in range(6):
    algo = [_lib.MB_AR_ALGO_ONESHOT, _lib.MB_AR_ALGO_TWOSHOT][rnd % 2]
    ins = [[gen_input(1000 * rnd + 10 * r + i, [m], 'f32') for i, m in enumerate(numels)] for r in range(world)]
    mine = [torch.from_numpy(a.copy()).cuda() for a in ins[rank]]
    hdrs = [(r + 1, rnd, 32) for r in range(world)]
    ctx.stage(mine, slot=rnd % 2, zero_src=True)
    dst = [torch.empty(m, device='cuda') for m in numels]
    ctx.allreduce(dst, hdr=hdrs[rank] + (1,), slot=rnd % 2, algo=algo)
    torch.cuda.synchronize()
    flat_in = []
    for r in range(world):
        f = np.zeros(total, dtype=np.float32)
        for a, o, m in zip(ins[r], offs, numels):
            f[o:o + m] = a
        flat_in.append(f)
    exact, eh = oracle.allreduce_rankorder(flat_in, hdrs)
    got = np.zeros(total, dtype=np.float32)
    for t, o, m in zip(dst, offs, numels):
        got[o:o + m] = t.cpu().numpy()
    assert got.tobytes() == exact.tobytes(), f'rank {rank} round {rnd}'
    assert ctx.result(rnd % 2) == (eh, 0)
dist.barrier()
small.close()
ctx.close()
print(f'rank {rank} OK')
"""


@pytest.mark.parametrize("world", [n for n in (2, 4, 8) if n <= NGPU])
def test_cross_process_allreduce(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MB_ROOT=ROOT)
    port = 29600 + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == world
