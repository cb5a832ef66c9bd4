
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
# late-joiner model sync: parameters + float buffers came out of the leader's publish region over NVLink
import zlib
ds = acc.debug_state()
crc = zlib.crc32(torch.cat([p.detach().flatten() for p in m.parameters()] + [m.running]).cpu().numpy().tobytes())
print(f'PARAMCRC {rank} {crc:08x} {int(m.steps.item())}', flush=True)
if acc.is_leader():
    assert ds['nvlink_model_publishes'] >= 1, ds
else:
    assert ds['nvlink_model_fetches'] >= 1, ds
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
for _ in range(200):
    pump(); time.sleep(0.001)
print(f'rank {rank} OK', flush=True)
os._exit(0)
