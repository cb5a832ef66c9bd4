"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

* ctypes/numpy wrappers over oracle/_build/liboracle.so (moolib_oracle.c, the C restatement),
* `OracleBatcher`: a pure-Python restatement of the control flow of moolib's Batcher (src/moolib.cc:595-845) on top
  of torch.stack / torch.cat, which is what the reference's own test pins it to (test/unit/test_batcher.py:32,46-52),
* `load_reference()`: imports the compiled, unmodified reference from oracle/_ref (built by oracle/build_ref.sh).

Nothing here is used by the product path.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")

__all__ = [
    "build", "lib", "stack_slot", "cat_narrow", "copy2d", "fill_batch", "scatter_actions", "stage",
    "allreduce_tree", "allreduce_rankorder", "flat_layout", "OracleBatcher", "load_reference", "reference_available",
    "allreduce_tolerance", "vtrace", "u8_to_f32",
]


class _Hdr(ctypes.Structure):
    _fields_ = [("num_gradients", ctypes.c_uint64), ("num_skipped", ctypes.c_uint64), ("batch_size", ctypes.c_uint64),
                ("has_grads", ctypes.c_uint64)]


def build(force=False):
    src = os.path.join(_HERE, "moolib_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(src) > os.path.getmtime(_LIB):
        subprocess.run(["make", "-C", _HERE, "-s", "_build/liboracle.so"] + (["-B"] if force else []), check=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.oracle_allreduce_tree.restype = ctypes.c_int
        _lib.oracle_allreduce_rankorder.restype = ctypes.c_int
    return _lib


def _u8(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def stack_slot(dst: np.ndarray, slot: int, src: np.ndarray, dim: int = 0):
    """dst.select(dim, slot).copy_(src) (src/moolib.cc:676,751)."""
    outer = int(np.prod(dst.shape[:dim], dtype=np.int64))
    size = dst.shape[dim]
    inner = int(np.prod(dst.shape[dim + 1:], dtype=np.int64)) * dst.itemsize
    assert src.nbytes == outer * inner
    lib().oracle_stack_slot(_u8(dst), ctypes.c_size_t(outer), ctypes.c_size_t(size), ctypes.c_size_t(slot),
                            ctypes.c_size_t(inner), _u8(src))


def cat_narrow(dst: np.ndarray, dst_off: int, src: np.ndarray, src_off: int, n: int, dim: int = 0):
    """dst.narrow(dim, dst_off, n).copy_(src.narrow(dim, src_off, n)) (src/moolib.cc:665-668,745-748)."""
    outer = int(np.prod(dst.shape[:dim], dtype=np.int64))
    inner = int(np.prod(dst.shape[dim + 1:], dtype=np.int64)) * dst.itemsize
    lib().oracle_cat_narrow(_u8(dst), _u8(src), ctypes.c_size_t(outer), ctypes.c_size_t(dst.shape[dim]),
                            ctypes.c_size_t(dst_off), ctypes.c_size_t(src.shape[dim]), ctypes.c_size_t(src_off),
                            ctypes.c_size_t(n), ctypes.c_size_t(inner))


def copy2d(src_buf: np.ndarray, src_off: int, dst_buf: np.ndarray, dst_off: int, row_bytes, rows, src_pitch, dst_pitch):
    """One mb_copy_job on flat uint8 buffers."""
    assert src_buf.dtype == np.uint8 and dst_buf.dtype == np.uint8
    lib().oracle_copy2d(ctypes.c_void_p(src_buf.ctypes.data + src_off), ctypes.c_void_p(dst_buf.ctypes.data + dst_off),
                        ctypes.c_size_t(row_bytes), ctypes.c_size_t(rows), ctypes.c_ssize_t(src_pitch),
                        ctypes.c_ssize_t(dst_pitch))


def fill_batch(slab: np.ndarray, batch_index: int, src: np.ndarray):
    """src/env.h:248-263: slab is [maxEnvs, *shape]; copies one env's item into row batch_index."""
    lib().oracle_fill_batch(_u8(slab), ctypes.c_size_t(slab.itemsize), ctypes.c_size_t(src.size),
                            ctypes.c_size_t(batch_index), _u8(src))


def scatter_actions(counters: np.ndarray, actions: np.ndarray, stride: int = 1):
    assert counters.dtype == np.uint32 and actions.dtype == np.int64
    lib().oracle_scatter_actions(_u8(counters), ctypes.c_size_t(stride), _u8(actions), ctypes.c_size_t(actions.size))


def flat_layout(numels):
    """Offsets (in floats) of each tensor in the flat staging layout and the padded total."""
    offs, off = [], 0
    for n in numels:
        offs.append(off)
        off += (n + 3) & ~3
    return offs, off


def stage(staging: np.ndarray, grads, accumulate=False, zero_src=False):
    """src/accumulator.cc:941-980 (+ :410-418 when zero_src)."""
    assert staging.dtype == np.float32
    n = len(grads)
    ptrs = (ctypes.c_void_p * n)(*[g.ctypes.data for g in grads])
    numel = (ctypes.c_uint64 * n)(*[g.size for g in grads])
    lib().oracle_stage(_u8(staging), ptrs, numel, n, int(accumulate), int(zero_src))


def _ar_args(inputs, hdrs):
    n = len(inputs)
    numel = next(a.size for a in inputs if a is not None) if any(a is not None for a in inputs) else 0
    ptrs = (ctypes.c_void_p * n)(*[(a.ctypes.data if a is not None else None) for a in inputs])
    H = (_Hdr * n)()
    for i, h in enumerate(hdrs):
        H[i].num_gradients, H[i].num_skipped, H[i].batch_size = h[0], h[1], h[2]
        H[i].has_grads = 0 if inputs[i] is None else 1
    return n, numel, ptrs, H


def allreduce_tree(inputs, hdrs, order=0, scale=True, numel=None):
    """Reference order (binary tree, src/group.h:570-629,738-768).  inputs[r] = float32 array or None (skipped)."""
    n, ne, ptrs, H = _ar_args(inputs, hdrs)
    if numel is None:
        numel = ne
    out = np.empty(numel, dtype=np.float32)
    oh = _Hdr()
    rc = lib().oracle_allreduce_tree(ptrs, H, n, ctypes.c_size_t(numel), int(order), int(scale), _u8(out),
                                     ctypes.byref(oh))
    assert rc == 0
    return out, (oh.num_gradients, oh.num_skipped, oh.batch_size, oh.has_grads)


def allreduce_rankorder(inputs, hdrs, scale=True, numel=None):
    """The product's summation order (ascending rank); the bit-exact regression gate."""
    n, ne, ptrs, H = _ar_args(inputs, hdrs)
    if numel is None:
        numel = ne
    out = np.empty(numel, dtype=np.float32)
    oh = _Hdr()
    rc = lib().oracle_allreduce_rankorder(ptrs, H, n, ctypes.c_size_t(numel), int(scale), _u8(out), ctypes.byref(oh))
    assert rc == 0
    return out, (oh.num_gradients, oh.num_skipped, oh.batch_size, oh.has_grads)


def allreduce_tolerance(inputs, reference_out, scale_factor=1.0):
    """Per-element bound of SURVEY.md section 8(c): |ours - oracle| <= 1e-6 * max(|oracle|, sum_i|g_i| * scale)."""
    s = np.zeros_like(reference_out, dtype=np.float64)
    for a in inputs:
        if a is not None:
            s += np.abs(a.astype(np.float64))
    return 1e-6 * np.maximum(np.abs(reference_out.astype(np.float64)), s * scale_factor)


def vtrace(log_rhos, discounts, rewards, values, bootstrap_value, clip_rho=1.0, clip_pg_rho=1.0):
    """examples/common/vtrace.py:156-242 from_importance_weights on [T, B] float32 arrays -> (vs, pg_advantages)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (log_rhos, discounts, rewards, values, bootstrap_value)]
    T = arrs[0].shape[0]
    B = arrs[0].size // max(T, 1)
    vs, pg = np.empty_like(arrs[0]), np.empty_like(arrs[0])
    f = lib().oracle_vtrace
    f.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_float, ctypes.c_size_t,
                                          ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    f(*[_u8(a) for a in arrs], int(clip_rho is not None), float(clip_rho or 0.0), int(clip_pg_rho is not None),
      float(clip_pg_rho or 0.0), T, B, _u8(vs), _u8(pg))
    return vs, pg


def u8_to_f32(src, scale=np.float32(1.0) / np.float32(255.0)):
    """examples/atari/models.py:94 `x.float() / 255.0` as ATen evaluates it: x * fp32(1/255)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.empty(src.shape, dtype=np.float32)
    f = lib().oracle_u8_to_f32
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float]
    f(_u8(src), _u8(out), src.size, float(scale))
    return out


# ---------------------------------------------------------------------------------------------------------------
# Batcher restatement (host logic oracle)
# ---------------------------------------------------------------------------------------------------------------

class OracleBatcher:
    """Behavioural restatement of moolib.Batcher (src/moolib.cc:595-845,1411-1488) for flat or nested inputs.

    stack(): the k-th item goes to select(dim, k) of a [.., size, ..] batch; emitted when k == size (:813-845).
    cat(): items are concatenated along dim; a batch is emitted every `size` entries and the remainder of the
    item carries over into the next batch (:767-811).  Non-tensor leaves are taken from the first item (:689).
    """

    def __init__(self, size, device="cpu", dim=0):
        import torch
        self.torch = torch
        self.size, self.device, self.dim = size, device, dim
        self.items = []
        self.cat_parts = []
        self.cat_fill = 0
        self.mode = None
        self.queue = []

    def _map(self, f, *xs):
        x = xs[0]
        if isinstance(x, dict):
            return {k: self._map(f, *[y[k] for y in xs]) for k in x}
        if isinstance(x, list):
            return [self._map(f, *[y[i] for y in xs]) for i in range(len(x))]
        if isinstance(x, tuple):
            return tuple(self._map(f, *[y[i] for y in xs]) for i in range(len(x)))
        if isinstance(x, self.torch.Tensor):
            return f(*xs)
        return x

    def stack(self, item):
        if self.mode == "cat":
            raise RuntimeError("Batcher.stack: Previously called with cat; cannot mix cat/stack within the same batch")
        self.mode = "stack"
        self.items.append(item)
        if len(self.items) == self.size:
            out = self._map(lambda *ts: self.torch.stack(ts, dim=self.dim).to(self.device), *self.items)
            self.items, self.mode = [], None
            self.queue.append(out)

    def cat(self, item):
        if self.mode == "stack":
            raise RuntimeError("Batcher.cat: Previously called with stack; cannot mix cat/stack within the same batch")
        torch = self.torch
        sizes = []
        self._map(lambda t: sizes.append(t.size(self.dim)) or t, item)
        n = sizes[0] if sizes else 0
        off = 0
        while True:
            self.mode = "cat"
            left = self.size - self.cat_fill
            take = min(n - off, left)
            self.cat_parts.append(self._map(lambda t: t.narrow(self.dim, off, take), item))
            self.cat_fill += take
            off += take
            if self.cat_fill == self.size:
                out = self._map(lambda *ts: torch.cat(ts, dim=self.dim).to(self.device), *self.cat_parts)
                self.cat_parts, self.cat_fill, self.mode = [], 0, None
                self.queue.append(out)
                if off == n:
                    break
            else:
                break

    def empty(self):
        return not self.queue

    def get(self):
        return self.queue.pop(0)


# ---------------------------------------------------------------------------------------------------------------
# the compiled reference itself
# ---------------------------------------------------------------------------------------------------------------

def reference_available():
    d = os.path.join(REF_DIR, "moolib")
    return os.path.isdir(d) and any(f.startswith("_C") and f.endswith(".so") for f in os.listdir(d))


def load_reference():
    """import the unmodified reference build (oracle/_ref/moolib).  Raises ImportError when it was not built."""
    if not reference_available():
        raise ImportError("oracle/_ref is not built (run oracle/build_ref.sh where /root/reference is mounted)")
    if "moolib" in sys.modules and getattr(sys.modules["moolib"], "__file__", "").startswith(REF_DIR):
        return sys.modules["moolib"]
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    return importlib.import_module("moolib")
