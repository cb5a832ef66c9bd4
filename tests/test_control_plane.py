"""CPU tests of the moolib-API host layer: Rpc / Broker / Group / Accumulator with several peers in ONE process over
loopback, the way the reference's own script tests fake a cluster (test/test_reduce.py:97-104, test/test_group.py).
BASELINE.json configs[0]: 2-peer Accumulator, ~1k-param torch.nn.Linear, CPU (plumbing, no GPU)."""
import ast
import itertools
import time

import numpy as np
import pytest
import torch

import moolib_b200 as moolib
import oracle
from helpers import gen_input

_port = itertools.count(47100)


class Cluster:
    def __init__(self, n, group="g"):
        self.addr = f"127.0.0.1:{next(_port)}"
        self.broker_rpc = moolib.Rpc()
        self.broker_rpc.set_name("broker")
        self.broker = moolib.Broker(self.broker_rpc)
        self.broker_rpc.listen(self.addr)
        self.rpcs, self.groups = [], []
        for i in range(n):
            r = moolib.Rpc()
            r.set_name(f"peer{i}")
            r.set_timeout(20)
            r.connect(self.addr)
            g = moolib.Group(r, group)
            g.set_timeout(20)
            g.set_sort_order(i)
            self.rpcs.append(r)
            self.groups.append(g)

    def pump(self, extra=()):
        self.broker.update()
        for g in self.groups:
            g.update()
        for a in extra:
            a.update()

    def form(self, n=None, timeout=30):
        n = n if n is not None else len(self.groups)
        t0 = time.time()
        while True:
            self.pump()
            if all(g.active() and len(g.members()) == n for g in self.groups) and \
                    len({g.sync_id() for g in self.groups}) == 1:
                return
            assert time.time() - t0 < timeout, [g.members() for g in self.groups]
            time.sleep(0.005)


def test_group_membership_and_sort_order():
    c = Cluster(4)
    c.form()
    assert c.groups[0].members() == ["peer0", "peer1", "peer2", "peer3"]  # (sortOrder, creationOrder), broker.h:168
    assert c.groups[0].sync_id() != 0 and c.groups[0].name() == "g"
    # a member disappears -> regroup with a new sync id (test/test_group.py:57-86)
    old = c.groups[0].sync_id()
    gone = c.groups.pop(2)
    c.rpcs.pop(2)
    gone.set_timeout(0.3)
    del gone
    for g in c.groups:
        g.set_timeout(1.0)
    t0 = time.time()
    while not (all(len(g.members()) == 3 for g in c.groups) and c.groups[0].sync_id() != old):
        c.pump()
        time.sleep(0.01)
        assert time.time() - t0 < 30
    assert c.groups[0].members() == ["peer0", "peer1", "peer3"]


def test_all_reduce_cpu_tensor_and_python_op():
    n = 4
    c = Cluster(n)
    c.form()
    ins = [torch.from_numpy(gen_input(50 + r, [64, 64], "f32")) for r in range(n)]
    futs = [c.groups[r].all_reduce("test reduce", ins[r].clone()) for r in range(n)]
    res = [f.result(20) for f in futs]
    exact, _ = oracle.allreduce_rankorder([x.numpy().reshape(-1) for x in ins], [(1, 0, 1)] * n, scale=False)
    for r in range(n):
        assert res[r].numpy().reshape(-1).tobytes() == exact.tobytes()  # member order, all peers identical
    # the reference's own acceptance bound (test/test_reduce.py:66-81)
    assert abs(res[0].sum().item() - sum(x.sum().item() for x in ins)) < 0.01
    # python objects with an op (common/__init__.py:65-120 GlobalStatsAccumulator pattern)
    futs = [c.groups[r].all_reduce("stats", {"a": r, "b": [r]}, op=lambda x, y: {"a": x["a"] + y["a"], "b": x["b"] + y["b"]})
            for r in range(n)]
    for f in futs:
        assert f.result(20) == {"a": 6, "b": [0, 1, 2, 3]}
    with pytest.raises(RuntimeError, match="can only use the default operator on Tensor data"):
        c.groups[0].all_reduce("bad", {"x": 1})


def test_all_reduce_cancelled_on_group_change_and_concurrent_name():
    c = Cluster(2)
    c.form()
    f0 = c.groups[0].all_reduce("lonely", torch.ones(4))  # peer1 never joins
    with pytest.raises(RuntimeError, match="twice concurrently with the name 'lonely'"):
        c.groups[0].all_reduce("lonely", torch.ones(4))
    # a third peer joins -> sync id changes -> in-flight reductions are cancelled (src/group.h:453-461)
    r = moolib.Rpc()
    r.set_name("peer2")
    r.connect(c.addr)
    g = moolib.Group(r, "g")
    g.set_sort_order(2)
    c.rpcs.append(r)
    c.groups.append(g)
    c.form(3)
    assert f0.done()
    with pytest.raises(RuntimeError, match="cancelled due to a group change"):
        f0.result(1)


def test_rpc_define_async_sync():
    c = Cluster(2)
    c.rpcs[1].define("mul", lambda a, b=2: a * b)
    assert c.rpcs[0].sync("peer1", "mul", 21) == 42
    f = c.rpcs[0].async_("peer1", "mul", torch.ones(3), b=3)
    assert f.result(10).equal(torch.full((3,), 3.0))
    with pytest.raises(RuntimeError, match="does not exist"):
        c.rpcs[0].sync("peer1", "nope")


def _make_accumulators(c, n, vbs):
    models, accs = [], []
    torch.manual_seed(0)
    for i in range(n):
        m = torch.nn.Linear(32, 31)
        a = moolib.Accumulator("acc", m.parameters(), m.buffers(), group=c.groups[i])
        a.set_virtual_batch_size(vbs)
        models.append(m)
        accs.append(a)
    t0 = time.time()
    while not all(a.connected() for a in accs):
        c.pump(accs)
        for a in accs:
            if a.wants_state():
                a.set_state({"opt": 7})
            if a.has_new_state():
                assert a.state() == {"opt": 7}
        time.sleep(0.002)
        assert time.time() - t0 < 60, "accumulators did not connect"
    return models, accs


def test_accumulator_leader_and_model_sync():
    c = Cluster(3)
    c.form()
    models, accs = _make_accumulators(c, 3, 30)
    leaders = {a.get_leader() for a in accs}
    assert len(leaders) == 1 and sum(a.is_leader() for a in accs) == 1
    # every peer ends up with the leader's parameters (src/accumulator.cc:810-836)
    lead = [m for m, a in zip(models, accs) if a.is_leader()][0]
    for m in models:
        assert torch.equal(m.weight, lead.weight) and torch.equal(m.bias, lead.bias)


def test_accumulator_rounds_replay_reference_golden(golden_dir):
    """The controlled rounds recorded from the REFERENCE Accumulator (plain / local accumulation / skipping peer),
    driven through our Accumulator with the same calls.  N=2 is bit-identical (a+b == b+a); N=4 is within the stated
    1e-6 tolerance of the reference's tree order and bit-identical to the rank-order oracle."""
    g = np.load(f"{golden_dir}/accumulator_golden.npz")
    for n in (2, 4):
        c = Cluster(n, group=f"acc{n}")
        c.form()
        models, accs = _make_accumulators(c, n, 10 * n)
        for rec in g["rounds"]:
            tag, gn, plan, vbs, ngrad, nskip, bsz = ast.literal_eval(str(rec))
            if gn != n:
                continue
            for a in accs:
                a.set_virtual_batch_size(vbs)
            t0 = time.time()
            while not all(a.wants_gradients() for a in accs):
                c.pump(accs)
                assert time.time() - t0 < 30
            maxc = max(len(p) for p in plan)
            staged = [None] * n
            for k in range(maxc):
                for i, a in enumerate(accs):
                    t1 = time.time()
                    while not a.wants_gradients():
                        c.pump(accs)
                        assert time.time() - t1 < 30
                    if k < len(plan[i]):
                        gw, gb = gen_input(plan[i][k], [31, 32], "f32"), gen_input(plan[i][k] + 1, [31], "f32")
                        models[i].weight.grad = torch.from_numpy(gw.copy())
                        models[i].bias.grad = torch.from_numpy(gb.copy())
                        a.reduce_gradients(10)
                        assert not models[i].weight.grad.any()  # zeroed after staging (accumulator.cc:410-418)
                        flat = np.concatenate([gw.reshape(-1), gb])
                        staged[i] = flat if staged[i] is None else staged[i] + flat
                    else:
                        a.skip_gradients()
                for _ in range(10):
                    c.pump(accs)
                    time.sleep(0.002)
            t0 = time.time()
            while not all(a.has_gradients() for a in accs):
                c.pump(accs)
                time.sleep(0.001)
                assert time.time() - t0 < 30, tag
            stats = accs[0].get_gradient_stats()
            assert (stats["num_gradients"], stats["num_skipped"], stats["batch_size"]) == (ngrad, nskip, bsz), tag
            ref = np.concatenate([g[f"{tag}_w"].reshape(-1), g[f"{tag}_b"].reshape(-1)])
            hdrs = [(len(plan[i]), maxc - len(plan[i]), 10 * len(plan[i])) for i in range(n)]
            exact, _ = oracle.allreduce_rankorder(staged, hdrs, numel=ref.size)
            for i in range(n):
                got = np.concatenate([models[i].weight.grad.numpy().reshape(-1), models[i].bias.grad.numpy()])
                assert got.tobytes() == exact.tobytes(), (tag, i)
                assert (np.abs(got.astype(np.float64) - ref) <= 1e-6 * (np.abs(ref) + 1.0)).all(), tag
                if n == 2:
                    assert got.tobytes() == ref.tobytes(), tag
            for a in accs:
                a.zero_gradients()
                assert not a.has_gradients()
            c.pump(accs)


def test_accumulator_config0_two_peers_200_rounds():
    """BASELINE.json configs[0] / SURVEY.md section 8(d) row 1: peer i sets every .grad to i+1, 200 applied rounds.
    As in the reference, a peer may contribute several times before the count gate opens ("the
    first applied reduction can carry num_gradients >> peers"), so the average is (k0*1 + k1*2)/(k0+k1) with
    k0+k1 == num_gradients; both peers must hold identical bits."""
    c = Cluster(2, group="cfg0")
    c.form()
    models, accs = _make_accumulators(c, 2, 2)
    applied, t0 = 0, time.time()
    seen = [[], []]
    while applied < 200:
        c.pump(accs)
        for i, (m, a) in enumerate(zip(models, accs)):
            if a.has_gradients():
                v = m.weight.grad.flatten()[0].item()
                assert (m.weight.grad == v).all() and (m.bias.grad == v).all() and 1.0 <= v <= 2.0
                ng = a.get_gradient_stats()["num_gradients"]
                k1 = ng * (v - 1.0)
                assert abs(k1 - round(k1)) < 1e-3 * ng, (v, ng)
                seen[i].append(v)
                a.zero_gradients()
                applied += i == 0
            elif a.wants_gradients():
                m.weight.grad = torch.full_like(m.weight, float(i + 1))
                m.bias.grad = torch.full_like(m.bias, float(i + 1))
                a.reduce_gradients(1)
        assert time.time() - t0 < 120
    k = min(len(seen[0]), len(seen[1]))
    assert k >= 199 and seen[0][:k] == seen[1][:k]  # identical averaged gradients on both peers, every round
    assert abs(accs[0].model_version() - accs[1].model_version()) <= 1 and accs[0].model_version() >= 200


def test_reduce_without_wants_gradients_is_an_error():
    c = Cluster(1, group="solo")
    m = torch.nn.Linear(2, 2)
    a = moolib.Accumulator("acc", m.parameters(), m.buffers(), group=c.groups[0])
    assert not a.wants_gradients()
    with pytest.raises(RuntimeError, match="called while wantsGradients\\(\\) is false"):
        a.reduce_gradients(1)


@pytest.mark.timeout(120)
def test_recount_after_failed_count_does_not_deadlock():
    """A peer that contributes again while a count is in flight sets wantsMoreCounting; when that count comes back
    below the virtual batch size the next count is started from inside the result handling (src/accumulator.cc:
    1066-1071).  Regression test: this used to self-deadlock on the finished op's future mutex."""
    c = Cluster(2, group="recount")
    c.form()
    models, accs = _make_accumulators(c, 2, 50)
    for a in accs:
        a.set_virtual_batch_size(50)
    rounds = 0
    t0 = time.time()
    while rounds < 3:
        for i, (m, a) in enumerate(zip(models, accs)):
            # several contributions per pump: the 2nd..4th arrive while the first count is still in flight
            for _ in range(4):
                if a.wants_gradients():
                    m.weight.grad = torch.ones_like(m.weight)
                    m.bias.grad = torch.ones_like(m.bias)
                    a.reduce_gradients(5)
        c.pump(accs)
        for m, a in zip(models, accs):
            if a.has_gradients():
                assert (m.weight.grad == 1).all()
                a.zero_gradients()
                rounds += 1
        assert time.time() - t0 < 90


def _drive(c, models, accs, rounds, value_fn, max_s=90, extra_pump=()):
    """Standard loop: contribute when asked, apply when available; returns the list of applied gradients per peer."""
    seen = [[] for _ in accs]
    t0 = time.time()
    while min(len(s) for s in seen) < rounds:
        c.pump(list(accs) + list(extra_pump))
        for i, (m, a) in enumerate(zip(models, accs)):
            if a.wants_state():
                a.set_state({"k": i})
            if a.has_new_state():
                a.state()
            if a.has_gradients():
                seen[i].append((m.weight.grad.flatten()[0].item(), a.get_gradient_stats()["num_gradients"]))
                a.zero_gradients()
            elif a.wants_gradients():
                v = value_fn(i, len(seen[i]))
                m.weight.grad = torch.full_like(m.weight, v)
                m.bias.grad = torch.full_like(m.bias, v)
                a.reduce_gradients(1)
        assert time.time() - t0 < max_s, [len(s) for s in seen]
    return seen


@pytest.mark.timeout(180)
def test_parallel_gradients_ring():
    """set_parallel_gradients(2): two reduction slots used round-robin (src/accumulator.cc:889-903); every applied
    gradient is the average of one contribution per peer, identical on both peers, in order."""
    c = Cluster(2, group="ring")
    c.form()
    models, accs = _make_accumulators(c, 2, 2)
    for a in accs:
        a.set_parallel_gradients(2)
        a.set_virtual_batch_size(2)
    seen = _drive(c, models, accs, 12, lambda i, k: float(i + 1))
    for s in seen:
        assert all(abs(v - 1.5) < 1e-6 and ng == 2 for v, ng in s[:12]), s
    with pytest.raises(RuntimeError):
        accs[0].set_parallel_gradients(9)


@pytest.mark.timeout(240)
def test_training_survives_a_peer_joining_and_leaving():
    """Elasticity: a third peer joins mid-run (new syncId -> reductions reset, leader re-elected,
    model pushed to the joiner), later leaves (times out at the broker) and the remaining two carry on."""
    c = Cluster(2, group="elastic")
    c.form()
    models, accs = _make_accumulators(c, 2, 2)
    seen = _drive(c, models, accs, 5, lambda i, k: 1.0)
    assert all(abs(v - 1.0) < 1e-6 for s in seen for v, _ in s)
    # make the incumbents' parameters distinctive so that we can see the joiner receive them
    with torch.no_grad():
        for m in models:
            m.weight.fill_(0.25)
            m.bias.fill_(-0.5)
    for a in accs:
        a.set_model_version(100)
    r = moolib.Rpc()
    r.set_name("peer2")
    r.set_timeout(20)
    r.connect(c.addr)
    g = moolib.Group(r, "elastic")
    g.set_timeout(2.0)
    g.set_sort_order(2)
    m3 = torch.nn.Linear(32, 31)
    a3 = moolib.Accumulator("acc", m3.parameters(), m3.buffers(), group=g)
    a3.set_virtual_batch_size(3)
    for a in accs:
        a.set_virtual_batch_size(3)
    c.rpcs.append(r)
    c.groups.append(g)
    seen = _drive(c, models + [m3], accs + [a3], 5, lambda i, k: float(i + 1))
    assert torch.equal(m3.weight, models[0].weight) and torch.equal(m3.bias, models[0].bias)  # model sync to the joiner
    assert a3.model_version() >= 100
    assert all(abs(v - 2.0) < 1e-6 for v, ng in seen[2][-2:])  # (1+2+3)/3 once all three contribute
    # the joiner disappears
    c.groups.pop()
    c.rpcs.pop()
    del a3, g, r
    for a in accs:
        a.set_virtual_batch_size(2)
    t0 = time.time()
    while len(c.groups[0].members()) != 2:
        c.pump(accs)
        time.sleep(0.01)
        assert time.time() - t0 < 60
    seen = _drive(c, models, accs, 3, lambda i, k: 4.0)
    assert all(abs(v - 4.0) < 1e-6 for s in seen for v, _ in s[-2:])
