"""moolib_b200.Accumulator / Group.all_reduce with CUDA tensors: the C++ host layer drives the stage + NVLink allreduce
kernels.  One GPU: single-member group (N=1 short-circuit, src/group.h:738-741, still through the kernels).
Several GPUs: one process per GPU under torchrun, CUDA IPC handles exchanged over the Group's own control plane."""
import os
import subprocess
import sys
import time

import numpy as np
import pytest
import torch

import moolib_b200 as moolib
import oracle
from helpers import gen_input
from moolib_b200 import _C

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_single_learner_accumulator_on_gpu():
    addr = "127.0.0.1:47301"
    broker = moolib.Broker()
    broker.listen(addr)
    m = torch.nn.Linear(32, 31).cuda()
    acc = moolib.Accumulator("acc", m.parameters(), m.buffers())
    acc.set_virtual_batch_size(20)
    acc.connect(addr)
    t0 = time.time()
    while not acc.connected():
        broker.update()
        acc.update()
        assert time.time() - t0 < 60
    assert acc.is_leader()
    before = _C.kernel_launches()
    gw1, gb1 = gen_input(1, [31, 32], "f32"), gen_input(2, [31], "f32")
    gw2, gb2 = gen_input(3, [31, 32], "f32"), gen_input(4, [31], "f32")
    for gw, gb in ((gw1, gb1), (gw2, gb2)):  # two local contributions before the gate (20) opens
        while not acc.wants_gradients():
            broker.update()
            acc.update()
        m.weight.grad = torch.from_numpy(gw.copy()).cuda()
        m.bias.grad = torch.from_numpy(gb.copy()).cuda()
        acc.reduce_gradients(10)
        assert not m.weight.grad.any().item()  # zeroed by the stage kernel
        for _ in range(20):
            broker.update()
            acc.update()
    t0 = time.time()
    while not acc.has_gradients():
        broker.update()
        acc.update()
        assert time.time() - t0 < 30
    assert _C.kernel_launches() - before == 3  # stage, stage(+=), allreduce
    st = np.zeros(oracle.flat_layout([992, 31])[1], dtype=np.float32)
    oracle.stage(st, [gw1.reshape(-1).copy(), gb1.copy()])
    oracle.stage(st, [gw2.reshape(-1).copy(), gb2.copy()], accumulate=True)
    exact, eh = oracle.allreduce_rankorder([st], [(2, 0, 20)])
    assert m.weight.grad.cpu().numpy().reshape(-1).tobytes() == exact[:992].tobytes()
    assert m.bias.grad.cpu().numpy().tobytes() == exact[992:992 + 31].tobytes()
    assert acc.get_gradient_stats() == {"num_gradients": 2, "num_skipped": 0, "batch_size": 20}
    assert acc.model_version() == 1
    acc.zero_gradients()
    assert not acc.has_gradients() and not m.weight.grad.any().item()

    # ---- zero-copy round: .grad now lives in the NVLink staging ring, a real backward() accumulates in place and
    # reduce_gradients() launches no stage kernel (one launch in total: K-A2) ----
    tm = acc.reduce_timings(clear=True)
    assert tm["device_gate"] and tm["stage_launches"] == 2 and tm["zero_copy_rounds"] == 0
    x = torch.from_numpy(gen_input(9, [20, 32], "f32")).cuda()
    ref = torch.nn.Linear(32, 31).cuda()
    ref.load_state_dict(m.state_dict())
    before = _C.kernel_launches()
    (m(x) ** 2).sum().backward()
    (ref(x) ** 2).sum().backward()
    gptr = m.weight.grad.data_ptr()
    acc.reduce_gradients(20)
    assert m.weight.grad.data_ptr() != gptr and not m.weight.grad.any().item()  # next ring buffer, zero-filled
    t0 = time.time()
    while not acc.has_gradients():
        broker.update()
        acc.update()
        assert time.time() - t0 < 30
    assert _C.kernel_launches() - before == 1
    tm = acc.reduce_timings()
    assert tm["zero_copy_rounds"] == 1 and tm["stage_launches"] == 0 and len(tm["reduce_us"]) == 1
    assert torch.equal(m.weight.grad, ref.weight.grad) and torch.equal(m.bias.grad, ref.bias.grad)  # 1 gradient: x 1.0f
    assert acc.model_version() == 2
    # clip + optimizer step work on the result views; zero_gradients() goes back to the (zeroed) staging ring
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    torch.optim.SGD(m.parameters(), lr=0.1).step()
    acc.zero_gradients()
    assert not m.weight.grad.any().item() and not m.bias.grad.any().item()


def test_parallel_gradients_on_gpu():
    """set_parallel_gradients(2) with CUDA parameters: the second slot is filled while the first is still in flight;
    results are applied one by one in order and never mix (round-1 advisor finding)."""
    addr = "127.0.0.1:47311"
    broker = moolib.Broker()
    broker.listen(addr)
    m = torch.nn.Linear(32, 31).cuda()
    acc = moolib.Accumulator("acc2", m.parameters(), m.buffers())
    acc.set_parallel_gradients(2)
    acc.set_virtual_batch_size(10)
    acc.connect(addr)
    t0 = time.time()
    while not (acc.connected() and acc.wants_gradients()):
        broker.update()
        acc.update()
        assert time.time() - t0 < 60
    gs = [(gen_input(50 + 2 * k, [31, 32], "f32"), gen_input(51 + 2 * k, [31], "f32")) for k in range(6)]
    fed = applied = 0
    t0 = time.time()
    while applied < 6:
        assert time.time() - t0 < 60
        broker.update()
        acc.update()
        if acc.has_gradients():
            gw, gb = gs[applied]
            assert m.weight.grad.cpu().numpy().tobytes() == gw.tobytes(), applied
            assert m.bias.grad.cpu().numpy().tobytes() == gb.tobytes(), applied
            acc.zero_gradients()
            applied += 1
        elif fed < 6 and acc.wants_gradients():
            with torch.no_grad():
                m.weight.grad.copy_(torch.from_numpy(gs[fed][0]))  # in place: into the slot's staging ring
                m.bias.grad.copy_(torch.from_numpy(gs[fed][1]))
            acc.reduce_gradients(10)
            fed += 1
    tm = acc.reduce_timings()
    # the very first contribution found an ordinary .grad tensor (created before the NVLink context existed): K-A1
    assert tm["zero_copy_rounds"] == 5 and tm["stage_launches"] == 1


LEGACY = r"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.environ['MB_ROOT']); sys.path.insert(0, os.path.join(os.environ['MB_ROOT'], 'tests'))
import oracle, moolib_b200 as moolib
from helpers import gen_input
addr = '127.0.0.1:47321'
broker = moolib.Broker(); broker.listen(addr)
m = torch.nn.Linear(32, 31).cuda()
acc = moolib.Accumulator('acc', m.parameters(), m.buffers())
acc.set_parallel_gradients(2)
acc.set_virtual_batch_size(10)
acc.connect(addr)
t0 = time.time()
while not (acc.connected() and acc.wants_gradients()):
    broker.update(); acc.update(); assert time.time() - t0 < 60
gs = [(gen_input(50 + 2 * k, [31, 32], 'f32'), gen_input(51 + 2 * k, [31], 'f32')) for k in range(4)]
fed = applied = 0
t0 = time.time()
while applied < 4:
    assert time.time() - t0 < 60
    broker.update(); acc.update()
    if acc.has_gradients():
        assert m.weight.grad.cpu().numpy().tobytes() == gs[applied][0].tobytes(), applied
        assert m.bias.grad.cpu().numpy().tobytes() == gs[applied][1].tobytes(), applied
        acc.zero_gradients(); applied += 1
    elif fed < 4 and acc.wants_gradients():
        m.weight.grad = torch.from_numpy(gs[fed][0].copy()).cuda(); m.bias.grad = torch.from_numpy(gs[fed][1].copy()).cuda()
        acc.reduce_gradients(10); fed += 1
tm = acc.reduce_timings()
assert not tm['device_gate'] and tm['stage_launches'] == 4 and tm['zero_copy_rounds'] == 0, tm
print('LEGACY OK')
os._exit(0)
"""


def test_host_counted_mode_still_works(tmp_path):
    """MOOLIB_B200_STRICT_COUNTING=0: the reference's accumulate-while-counting with the count on the control plane,
    K-A1 + un-gated K-A2 writing the .grad tensors; with set_parallel_gradients(2) no backward is admitted while a
    kernel is in flight, so results never mix."""
    script = tmp_path / "legacy.py"
    script.write_text(LEGACY)
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, MB_ROOT=ROOT, MOOLIB_B200_STRICT_COUNTING="0"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "LEGACY OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


WORKER = r"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.environ['MB_ROOT']); sys.path.insert(0, os.path.join(os.environ['MB_ROOT'], 'tests'))
import oracle, moolib_b200 as moolib
from helpers import gen_input
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
addr = '127.0.0.1:' + os.environ['MB_PORT']
broker = None
if rank == 0:
    broker = moolib.Broker(); broker.listen(addr)
rpc = moolib.Rpc(); rpc.set_name(f'peer{rank}'); rpc.set_timeout(30); rpc.connect(addr)
group = moolib.Group(rpc, 'g'); group.set_sort_order(rank)
torch.manual_seed(1000 + rank)   # every peer starts from DIFFERENT weights: the elected leader's must win
m = torch.nn.Linear(32, 31).cuda()
m.register_buffer('running', torch.full((5,), float(rank), device='cuda'))
m.register_buffer('steps', torch.tensor([rank], dtype=torch.int64, device='cuda'))   # non-float buffer: control plane
acc = moolib.Accumulator('acc', m.parameters(), m.buffers(), group=group)
acc.set_virtual_batch_size(10 * world)
def pump():
    if broker: broker.update()
    group.update(); acc.update()
    if acc.wants_state(): acc.set_state({'k': 1})
    if acc.has_new_state(): acc.state()
t0 = time.time()
while not (acc.connected() and len(group.members()) == world):
    pump(); time.sleep(0.001); assert time.time() - t0 < 90, group.members()
# late-joiner model sync (SURVEY 8(f)-3): parameters + float buffers came out of the leader's publish region over NVLink
import zlib
ds = acc.debug_state()
crc = zlib.crc32(torch.cat([p.detach().flatten() for p in m.parameters()] + [m.running]).cpu().numpy().tobytes())
print(f'PARAMCRC {rank} {crc:08x} {int(m.steps.item())}', flush=True)
was_leader = acc.is_leader()
if not was_leader:
    assert ds['nvlink_model_fetches'] >= 1, ds   # connected() means the model has arrived
# group.all_reduce on CUDA tensors (A8)
x = torch.from_numpy(gen_input(900 + rank, [64, 64], 'f32')).cuda()
f = group.all_reduce('t', x)
t0 = time.time()
while not f.done():
    pump(); assert time.time() - t0 < 60
r = f.result()
exact, _ = oracle.allreduce_rankorder([gen_input(900 + q, [64 * 64], 'f32') for q in range(world)], [(1, 0, 1)] * world, scale=False)
assert r.data_ptr() == x.data_ptr() and x.cpu().numpy().reshape(-1).tobytes() == exact.tobytes()
# two differently named operations in flight, started in OPPOSITE orders on odd and even ranks (round-1 advisor finding:
# they must never pair the wrong tensors) -- every name owns its context
xa = torch.full((5000,), float(rank + 1), device='cuda'); xb = torch.full((300,), float(10 * (rank + 1)), device='cuda')
order = [('opA', xa), ('opB', xb)] if rank % 2 == 0 else [('opB', xb), ('opA', xa)]
futs = [group.all_reduce(n_, t_) for n_, t_ in order]
t0 = time.time()
while not all(f_.done() for f_ in futs):
    pump(); assert time.time() - t0 < 60
for f_ in futs: f_.result()
assert (xa == world * (world + 1) / 2).all().item() and (xb == 10 * world * (world + 1) / 2).all().item()
# Accumulator rounds
numels = [992, 31]
offs, total = oracle.flat_layout(numels)
for rnd in range(5):
    t0 = time.time()
    while not acc.wants_gradients():
        pump(); assert time.time() - t0 < 60
    skip = (rnd == 3 and rank == world - 1)
    # the gate is evaluated on the device as part of reduce/skip_gradients(): set the virtual batch size first
    acc.set_virtual_batch_size(10 * (world - 1) if rnd == 3 else 10 * world)
    if skip:
        acc.skip_gradients()
    else:
        m.weight.grad = torch.from_numpy(gen_input(100 * rnd + 2 * rank, [31, 32], 'f32')).cuda()
        m.bias.grad = torch.from_numpy(gen_input(100 * rnd + 2 * rank + 1, [31], 'f32')).cuda()
        acc.reduce_gradients(10)
    t0 = time.time()
    while not acc.has_gradients():
        pump(); assert time.time() - t0 < 60, f'round {rnd}'
    ins, hdrs = [], []
    for q in range(world):
        if rnd == 3 and q == world - 1:
            ins.append(None); hdrs.append((0, 1, 0)); continue
        f_ = np.zeros(total, dtype=np.float32)
        f_[:992] = gen_input(100 * rnd + 2 * q, [992], 'f32'); f_[992:992 + 31] = gen_input(100 * rnd + 2 * q + 1, [31], 'f32')
        ins.append(f_); hdrs.append((1, 0, 10))
    exact, eh = oracle.allreduce_rankorder(ins, hdrs, numel=total)
    assert m.weight.grad.cpu().numpy().reshape(-1).tobytes() == exact[:992].tobytes(), f'rank {rank} round {rnd}'
    assert m.bias.grad.cpu().numpy().tobytes() == exact[992:1023].tobytes()
    s = acc.get_gradient_stats()
    assert (s['num_gradients'], s['num_skipped'], s['batch_size']) == eh[:3], (s, eh)
    acc.zero_gradients()
# zero-copy rounds: gradients are written IN PLACE into .grad (views of the NVLink staging ring); round 6 needs two
# contributions per rank before the device-side gate opens (the first attempt ends MB_AR_SHORT on every rank)
acc.set_virtual_batch_size(10 * world)
for rnd in range(5, 9):
    need = 2 if rnd == 6 else 1
    acc.set_virtual_batch_size(10 * world * need)
    for c in range(need):
        t0 = time.time()
        while not acc.wants_gradients():
            pump(); assert time.time() - t0 < 60
        with torch.no_grad():
            m.weight.grad.add_(torch.from_numpy(gen_input(100 * rnd + 2 * rank + 50 * c, [31, 32], 'f32')).cuda())
            m.bias.grad.add_(torch.from_numpy(gen_input(100 * rnd + 2 * rank + 1 + 50 * c, [31], 'f32')).cuda())
        acc.reduce_gradients(10)
    t0 = time.time()
    while not acc.has_gradients():
        pump(); assert time.time() - t0 < 60, f'round {rnd}'
    ins = []
    for q in range(world):
        f_ = np.zeros(total, dtype=np.float32)
        for c in range(need):
            g_ = np.zeros(total, dtype=np.float32)
            g_[:992] = gen_input(100 * rnd + 2 * q + 50 * c, [992], 'f32'); g_[992:1023] = gen_input(100 * rnd + 2 * q + 1 + 50 * c, [31], 'f32')
            oracle.stage(f_, [g_], accumulate=c > 0)
        ins.append(f_)
    exact, eh = oracle.allreduce_rankorder(ins, [(need, 0, 10 * need)] * world, numel=total)
    assert m.weight.grad.cpu().numpy().reshape(-1).tobytes() == exact[:992].tobytes(), f'rank {rank} round {rnd}'
    assert m.bias.grad.cpu().numpy().tobytes() == exact[992:1023].tobytes()
    s = acc.get_gradient_stats()
    assert (s['num_gradients'], s['num_skipped'], s['batch_size']) == eh[:3], (s, eh)
    acc.zero_gradients()
tm = acc.reduce_timings()
assert tm['device_gate'] and tm['zero_copy_rounds'] >= 4 and tm['short_rounds'] >= 1, tm
if was_leader:
    assert acc.debug_state()['nvlink_model_publishes'] >= 1, acc.debug_state()
for _ in range(200):
    pump(); time.sleep(0.001)
print(f'rank {rank} OK', flush=True)
os._exit(0)
"""


@pytest.mark.parametrize("world", [n for n in (2, 4, 8) if n <= NGPU])
def test_accumulator_across_processes(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MB_ROOT=ROOT, MB_PORT=str(47400 + world))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == world
    crcs = {tuple(ln.split()[2:]) for ln in r.stdout.splitlines() if ln.startswith("PARAMCRC")}
    assert len(crcs) == 1, f"peers hold different models after the NVLink model sync: {crcs}"
