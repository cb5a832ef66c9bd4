"""define_queue with dynamic batching (the consumer of stack_fields / unstack_fields in the reference,
src/moolib.cc:492-499) and the asyncio awaitables, after test/test_dynamic_batching_queue.py."""
import asyncio
import itertools

import pytest
import torch

import moolib_b200 as moolib

_port = itertools.count(47700)


async def _process(que, callback, n_batches):
    sizes = []
    for _ in range(n_batches):
        ret_cb, args, kwargs = await que
        sizes.append(args[0].shape[0] if args and torch.is_tensor(args[0]) and args[0].dim() == 2 else 1)
        ret_cb(callback(*args, **(kwargs or {})))
    return sizes


@pytest.mark.timeout(120)
def test_dynamic_batching_queue_and_await():
    addr = f"127.0.0.1:{next(_port)}"
    dim, n = 16, 24
    linear = torch.nn.Linear(dim, dim)

    async def main():
        server = moolib.Rpc()
        server.set_name("server")
        server.set_timeout(30)
        plain = server.define_queue("linear")
        batched = server.define_queue("batch_linear", batch_size=8, dynamic_batching=True)
        server.listen(addr)
        client = moolib.Rpc()
        client.set_name("client")
        client.set_timeout(30)
        client.connect(addr)

        def run_linear(x, info):
            with torch.no_grad():
                return linear(x), info

        xs = [torch.randn(dim) for _ in range(n)]
        ys = [linear(x) for x in xs]
        # un-batched queue: one call per await
        f = client.async_("server", "linear", xs[0], info={"index": [0, 1]})
        t = asyncio.ensure_future(_process(plain, run_linear, 1))
        y, info = await f
        assert torch.allclose(y, ys[0], atol=1e-6) and info == {"index": [0, 1]}
        assert await t == [1]
        # dynamic batching: calls parked while the server is busy are merged (stack_fields) and split (unstack_fields)
        futs = [client.async_("server", "batch_linear", x, info=dict(index=[i, i + 1])) for i, x in enumerate(xs)]
        while batched.size() < n:
            await asyncio.sleep(0.01)
        sizes = await _process(batched, run_linear, n // 8)
        assert sizes == [8, 8, 8]
        for i, fu in enumerate(futs):
            y, info = await fu
            assert torch.allclose(y, ys[i], atol=1e-6)
            assert info["index"] == (i, i + 1) or list(info["index"]) == [i, i + 1]

    asyncio.run(main())


@pytest.mark.timeout(60)
def test_batcher_is_awaitable():
    async def main():
        b = moolib.Batcher(size=3, dim=0)

        async def producer():
            for i in range(3):
                await asyncio.sleep(0.01)
                b.stack(torch.full((2,), float(i)))

        asyncio.ensure_future(producer())
        out = await b
        assert out.equal(torch.tensor([[0.0, 0.0], [1.0, 1.0], [2.0, 2.0]]))

    asyncio.run(main())
