
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
    if time.time() - t0 > 8: print('STUCK', fed, applied, acc.wants_gradients(), acc.has_gradients(), acc.debug_state(), flush=True); os._exit(1)
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
