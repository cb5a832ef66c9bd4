"""moolib_b200: the two data-parallel hot paths of moolib (gradient allreduce, observation batch gather) as
hand-written sm_100a kernels behind moolib's own Python API.

The names below are the ones `py/moolib/__init__.py` of the reference re-exports for these paths; they come from the
compiled host layer `moolib_b200._C` (C++/pybind11 over torch tensors) which calls the kernels through the C-ABI in
include/moolib_b200.h.  There is no Python or CPU fallback: importing fails loudly if the native pieces are missing.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
if not _os.path.exists(_os.path.join(_here, "lib", "libmoolib_b200.so")):
    raise ImportError("moolib_b200/lib/libmoolib_b200.so is missing; build it with `python moolib_b200/build.py`")

import torch as _torch  # noqa: E402,F401  (libtorch must be loaded before the extension)

try:
    from . import _C  # noqa: E402
except ImportError as e:  # pragma: no cover
    raise ImportError(
        "moolib_b200._C (the compiled host layer) is missing or failed to load; build it with "
        "`python moolib_b200/build.py`") from e

from ._C import Batcher, UnrollBatcher, to_device, u8_to_float, vtrace_from_importance_weights  # noqa: E402,F401

for _name in ("Accumulator", "Group", "Rpc", "Broker", "EnvPool", "EnvStepper", "EnvStepperFuture", "Future",
              "AllReduce", "create_uid", "set_log_level", "set_logging", "set_max_threads"):
    if hasattr(_C, _name):
        globals()[_name] = getattr(_C, _name)



# asyncio support (reference: FutureWrapper.__await__ / BatcherWrapper.__await__ / QueueWrapper.__await__,
# src/moolib.cc:316-393, 1440-1455, 553-576): the native objects are polled from the running event loop.
def _await_polling(ready, take):
    import asyncio

    async def _wait():
        delay = 0.0
        while not ready():
            await asyncio.sleep(delay)
            delay = min(0.002, delay + 0.0002)
        return take()

    return _wait().__await__()


def _future_await(self):
    return _await_polling(self.done, self.result)


def _batcher_await(self):
    return _await_polling(lambda: not self.empty(), self.get)


def _queue_await(self):
    box = []

    def ready():
        r = self.try_get()
        if r is not None:
            box.append(r)
        return bool(box)

    return _await_polling(ready, lambda: box.pop())


_C.Future.__await__ = _future_await
_C.Batcher.__await__ = _batcher_await
_C.UnrollBatcher.__await__ = _batcher_await
if hasattr(_C, "Queue"):
    _C.Queue.__await__ = _queue_await
    Queue = _C.Queue

__version__ = "0.1.0"
