"""HP-A, device-gated rounds (K-A0 gate + barrier-free K-A2) through the C-ABI, against the CPU oracle.

The virtual-batch gate of moolib's Accumulator (count allreduce, src/accumulator.cc:1035-1078) is evaluated by the
gate kernel; gradients are produced IN the symmetric ring buffers (mb_ar_buffer) so no stage kernel runs.  All ranks
live in this process (peer access); the cross-process path is covered by tests/test_accumulator_gpu.py.
"""
import numpy as np
import pytest
import torch

import oracle
from helpers import gen_input
from moolib_b200 import _lib
from test_allreduce_gpu import ALGOS, World

pytestmark = pytest.mark.gpu

NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
WORLDS = [n for n in (1, 2, 4, 8) if n <= max(NGPU, 1)]


def launch_all(w, fn):
    for r in range(w.n):
        with torch.cuda.device(r):
            fn(r)
    w.sync()


@pytest.mark.parametrize("n", WORLDS)
@pytest.mark.parametrize("algo", list(ALGOS))
def test_zero_copy_ring_rounds(n, algo):
    """Seven rounds: gradients are written straight into the slot's current ring buffer, the gate opens at once
    (sum of batch sizes == min), the flat result is bit-exact, the ring advances; round 3 stays SHORT first (one more
    local contribution is folded in by K-A1 from the next ring buffer), round 5 has a skipping rank."""
    numel = 263_000  # ~1 MiB, not a multiple of the chunk size
    total = _lib.flat_numel([numel])
    w = World(n, total * 4)
    try:
        for rnd in range(7):
            ins = [gen_input(1000 * rnd + r, [numel], "f32") for r in range(n)]
            dst = [torch.full((total,), float("nan"), device=f"cuda:{r}") for r in range(n)]
            skipper = n - 1 if (rnd == 5 and n > 1) else -1

            def contribute(r, data, ahead=0):
                buf = w.ctx[r].buffer(total, ahead=ahead)
                buf.zero_()
                buf[:numel].copy_(torch.from_numpy(data))

            if rnd == 3:
                # first attempt: every rank has 5 of the 10 it must bring -> gate closed on every rank
                launch_all(w, lambda r: (contribute(r, ins[r]),
                                         w.ctx[r].reduce_gated(10 * n, flat_dst=dst[r], hdr=(1, 0, 5, 1), algo=ALGOS[algo])))
                for r in range(n):
                    h, st = w.ctx[r].result()
                    assert st == _lib.MB_AR_SHORT and h == (n, 0, 5 * n, n)
                    assert torch.isnan(dst[r]).all().item(), "a closed gate must not write the destination"
                # second local contribution: produced in the NEXT ring buffer, folded into the staging by K-A1
                more = [gen_input(5000 + r, [numel], "f32") for r in range(n)]

                def second(r):
                    contribute(r, more[r], ahead=1)
                    scratch = w.ctx[r].buffer(total, ahead=1)
                    w.ctx[r].stage([scratch], accumulate=True, zero_src=True)
                    w.ctx[r].reduce_gated(10 * n, flat_dst=dst[r], hdr=(2, 0, 10, 1), algo=ALGOS[algo])

                launch_all(w, second)
                flat_in = []
                for r in range(n):
                    st_ = np.zeros(total, dtype=np.float32)
                    oracle.stage(st_, [ins[r].copy()])
                    oracle.stage(st_, [more[r].copy()], accumulate=True)
                    flat_in.append(st_)
                hdrs = [(2, 0, 10)] * n
            else:
                def once(r):
                    if r == skipper:
                        w.ctx[r].reduce_gated(10 * (n - 1), flat_dst=dst[r], hdr=(0, 1, 0, 0), algo=ALGOS[algo])
                    else:
                        contribute(r, ins[r])
                        w.ctx[r].reduce_gated(10 * (n - 1) if skipper >= 0 else 10 * n, flat_dst=dst[r],
                                              hdr=(1, 0, 10, 1), algo=ALGOS[algo])

                launch_all(w, once)
                flat_in = [None if r == skipper else np.concatenate([ins[r], np.zeros(total - numel, np.float32)])
                           for r in range(n)]
                hdrs = [(0, 1, 0) if r == skipper else (1, 0, 10) for r in range(n)]
            exact, eh = oracle.allreduce_rankorder(flat_in, hdrs, numel=total)
            for r in range(n):
                assert w.ctx[r].result() == (eh, 0), (rnd, r, w.ctx[r].result())
                assert dst[r].cpu().numpy().tobytes() == exact.tobytes(), f"round {rnd} rank {r}"
                w.ctx[r].advance()
    finally:
        w.close()


@pytest.mark.parametrize("n", WORLDS)
def test_gated_tensor_list_destination(n):
    """The gated round can also scatter into a tensor list (atari gradient shapes, unaligned tail sizes)."""
    numels = [432, 16, 2304, 16, 991232, 256, 4950, 18, 275, 1, 3, 7]
    total = _lib.flat_numel(numels)
    offs, _ = oracle.flat_layout(numels)
    w = World(n, total * 4)
    try:
        ins = [[gen_input(31 * r + i, [m], "f32") for i, m in enumerate(numels)] for r in range(n)]
        dsts = [[torch.full((m,), float("nan"), device=f"cuda:{r}") for m in numels] for r in range(n)]

        def go(r):
            buf = w.ctx[r].buffer(total)
            buf.zero_()
            for a, o, m in zip(ins[r], offs, numels):
                buf[o:o + m].copy_(torch.from_numpy(a))
            w.ctx[r].reduce_gated(1, dst_tensors=dsts[r], hdr=(r + 1, r, 3, 1))

        launch_all(w, go)
        flat_in = []
        for r in range(n):
            f = np.zeros(total, dtype=np.float32)
            for a, o, m in zip(ins[r], offs, numels):
                f[o:o + m] = a
            flat_in.append(f)
        exact, eh = oracle.allreduce_rankorder(flat_in, [(r + 1, r, 3) for r in range(n)])
        for r in range(n):
            got = np.zeros(total, dtype=np.float32)
            for t, o, m in zip(dsts[r], offs, numels):
                got[o:o + m] = t.cpu().numpy()
            assert got.tobytes() == exact.tobytes()
            assert w.ctx[r].result() == (eh, 0)
    finally:
        w.close()


def test_gate_timeout_is_reported():
    """A peer that never reaches the gate: K-A0 gives up after timeout_ms, K-A2 returns without touching anything."""
    if NGPU < 2:
        pytest.skip("needs 2 GPUs")
    w = World(2, 1 << 12)
    try:
        with torch.cuda.device(0):
            d = torch.full((16,), 7.0, device="cuda:0")
            w.ctx[0].reduce_gated(1, flat_dst=d, timeout_ms=200)
        torch.cuda.synchronize(0)
        _, st = w.ctx[0].result()
        assert st == _lib.MB_ETIMEOUT
        assert (d == 7.0).all().item()
    finally:
        w.close()


def test_stage_kernel_many_small_tensors():
    """K-A1 with a table of 1500 tiny tensors (the per-thread cursor has to jump, not walk)."""
    rng = np.random.Generator(np.random.PCG64(5))
    numels = [int(x) for x in rng.integers(0, 40, size=1500)]
    total = _lib.flat_numel(numels)
    ctx = _lib.ArContext(0, 1, 0, total * 4)
    try:
        a = [gen_input(i, [m], "f32") for i, m in enumerate(numels)]
        ts = [torch.from_numpy(x.copy()).cuda() for x in a]
        ctx.stage(ts, zero_src=True)
        ctx.stage([torch.from_numpy(x.copy()).cuda() for x in a], accumulate=True)
        dst = [torch.empty(m, device="cuda") for m in numels]
        ctx.allreduce(dst, hdr=(2, 0, 1, 1))
        torch.cuda.synchronize()
        st = np.zeros(total, dtype=np.float32)
        oracle.stage(st, [x.copy() for x in a])
        oracle.stage(st, [x.copy() for x in a], accumulate=True)
        exact, _ = oracle.allreduce_rankorder([st], [(2, 0, 1)])
        offs, _ = oracle.flat_layout(numels)
        for t, o, m in zip(dst, offs, numels):
            assert t.cpu().numpy().tobytes() == exact[o:o + m].tobytes()
        assert all(not t.any().item() for t in ts)
    finally:
        ctx.close()


@pytest.mark.parametrize("n", [m for m in WORLDS if m >= 2])
def test_publish_region_roundtrip(n):
    """mb_ar_xfer_pack / mb_ar_xfer_unpack: rank n-1 publishes a ragged tensor list, every other rank pulls it into its
    own (differently aligned) tensors over NVLink; byte-exact, and the gradient ring is untouched."""
    numels = [992, 31, 5, 300001, 1, 0, 77]
    total = _lib.flat_numel(numels)
    w = World(n, total * 4)
    try:
        src_rank = n - 1
        vals = [gen_input(40 + i, [m], "f32") for i, m in enumerate(numels)]
        with torch.cuda.device(src_rank):
            ring = w.ctx[src_rank].buffer(total)
            ring.fill_(3.0)
            assert w.ctx[src_rank].xfer_pack([torch.from_numpy(v.copy()).cuda() for v in vals]) == 1
        torch.cuda.synchronize(src_rank)
        for r in range(n - 1):
            with torch.cuda.device(r):
                dst = []
                for m in numels:
                    buf = torch.full((m + 1,), float("nan"), device=f"cuda:{r}")
                    dst.append(buf[1:])  # 4-byte aligned only
                assert w.ctx[r].xfer_unpack(src_rank, dst) == 1
                torch.cuda.synchronize(r)
                for t, v in zip(dst, vals):
                    assert t.cpu().numpy().tobytes() == v.tobytes()
        assert (ring == 3.0).all().item()
    finally:
        w.close()


@pytest.mark.parametrize("n", [m for m in WORLDS if m >= 2])
def test_twoshot_in_place_result(n):
    """Two-shot with flat_dst == the rank's own staging buffer: the all-gather already left the averaged values there,
    the kernel skips its final copy; one-shot refuses an in-place destination (peers would still be reading it)."""
    numel = 700_001
    total = _lib.flat_numel([numel])
    w = World(n, total * 4)
    try:
        ins = [np.concatenate([gen_input(77 + r, [numel], "f32"), np.zeros(total - numel, np.float32)]) for r in range(n)]
        bufs = []
        for r in range(n):
            with torch.cuda.device(r):
                b = w.ctx[r].buffer(total)
                b.copy_(torch.from_numpy(ins[r]))
                bufs.append(b)
        launch_all(w, lambda r: w.ctx[r].reduce_gated(1, flat_dst=bufs[r], hdr=(1, 0, 4, 1), algo=_lib.MB_AR_ALGO_TWOSHOT))
        exact, eh = oracle.allreduce_rankorder(ins, [(1, 0, 4)] * n)
        for r in range(n):
            assert w.ctx[r].result() == (eh, 0)
            assert bufs[r].cpu().numpy().tobytes() == exact.tobytes(), r
        with torch.cuda.device(0):
            with pytest.raises(_lib.MoolibB200Error, match="in-place destination"):
                w.ctx[0].reduce_gated(1, flat_dst=bufs[0], algo=_lib.MB_AR_ALGO_ONESHOT)
        assert w.ctx[0].algo_for(4096) == _lib.MB_AR_ALGO_ONESHOT
        assert w.ctx[0].algo_for(64 << 20) == (_lib.MB_AR_ALGO_TWOSHOT if n > 2 else _lib.MB_AR_ALGO_ONESHOT)
    finally:
        w.close()
