"""world_size-2 gloo test (CPU) of the host-side N>1 plumbing: handle blobs travel intact and in rank order, the
import calls are made for every peer but self, shard ranges tile the work exactly."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ['MB_ROOT'])
from moolib_b200 import peer
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')

class FakeCtx:
    def __init__(self): self.imported = {}
    def export(self): return bytes([rank + 1]) * 192
    def import_peer(self, r, blob): self.imported[r] = blob

ctx = peer.connect_context(FakeCtx())
assert sorted(ctx.imported) == [r for r in range(world) if r != rank]
for r, b in ctx.imported.items():
    assert b == bytes([r + 1]) * 192
spans = [peer.shard_range(103, r, world) for r in range(world)]
assert spans[0][0] == 0 and spans[-1][1] == 103 and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
print('OK', rank)
"""


def test_handle_exchange_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, MB_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2
