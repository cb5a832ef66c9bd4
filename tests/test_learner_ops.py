"""V-trace targets and the uint8 observation normalisation.

CPU: the C oracle (oracle_vtrace / oracle_u8_to_f32) is pinned to fixtures produced by the reference's own Python
(tests/golden/make_golden.py imports examples/common/vtrace.py from the reference tree).  fp32 tolerance for V-trace:
|ours - ref| <= 1e-6 * max(|ref|, largest |ref| of the same batch column) -- the C library's expf and ATen's vectorised
exp may differ in the last bit, and the reverse scan carries that absolute error down the column; the normalisation
agrees to 1 ulp with ATen's CPU division and bit for bit with its CUDA evaluation.
GPU: the kernels are bit-exact against the PyTorch restatement run on the same device, within the same 1e-6 of the
golden fixtures, and called through the C-ABI.
"""
import ast
import ctypes

import numpy as np
import pytest
import torch

import oracle
from helpers import gen_input

TOL = 1e-6


def _cases(g):
    return [ast.literal_eval(str(c)) for c in g["cases"]]


def _inputs(ci, T, B):
    lr = gen_input(8000 + 10 * ci, [T, B], "f32") * 0.5
    disc = (gen_input(8001 + 10 * ci, [T, B], "bool") | gen_input(8002 + 10 * ci, [T, B], "bool")).astype(np.float32) * 0.99
    return lr, disc, gen_input(8003 + 10 * ci, [T, B], "f32"), gen_input(8004 + 10 * ci, [T, B], "f32"), \
        gen_input(8005 + 10 * ci, [B], "f32")


def _close(a, ref):
    ref = np.asarray(ref, dtype=np.float64)
    col = np.abs(ref).max(axis=0, keepdims=True)  # the scan runs down a column: errors are absolute within it
    return (np.abs(a.astype(np.float64) - ref) <= TOL * np.maximum(np.abs(ref), col)).all()


def test_oracle_vtrace_matches_reference_golden(golden_dir):
    g = np.load(f"{golden_dir}/vtrace_golden.npz")
    for ci, (T, B, clip, clip_pg) in enumerate(_cases(g)):
        vs, pg = oracle.vtrace(*_inputs(ci, T, B), clip_rho=clip, clip_pg_rho=clip_pg)
        assert _close(vs, g[f"c{ci}_vs"]) and _close(pg, g[f"c{ci}_pg"]), ci


def test_oracle_normalisation_matches_reference_golden(golden_dir):
    g = np.load(f"{golden_dir}/vtrace_golden.npz")
    x = gen_input(int(g["norm_in_seed"]), [3, 5, 4, 8, 8], "u8")
    ref = g["norm_out"]
    got = oracle.u8_to_f32(x)
    # ATen's CPU path divides, its CUDA path multiplies by the reciprocal: they agree to 1 ulp, the oracle restates CUDA
    assert np.abs(got.astype(np.float64) - ref).max() <= 2 ** -24
    assert oracle.u8_to_f32(np.arange(256, dtype=np.uint8))[255] == np.float32(255.0) * (np.float32(1.0) / np.float32(255.0))


def torch_vtrace(log_rhos, discounts, rewards, values, bootstrap_value, clip_rho, clip_pg_rho):
    """examples/common/vtrace.py:207-242 restated (the loop of the reference, plain PyTorch)."""
    rhos = torch.exp(log_rhos)
    clipped = torch.clamp(rhos, max=clip_rho) if clip_rho is not None else rhos
    cs = torch.clamp(rhos, max=1.0)
    v_tp1 = torch.cat([values[1:], bootstrap_value.unsqueeze(0)], dim=0)
    deltas = clipped * (rewards + discounts * v_tp1 - values)
    acc = torch.zeros_like(bootstrap_value)
    res = []
    for t in range(discounts.shape[0] - 1, -1, -1):
        acc = deltas[t] + discounts[t] * cs[t] * acc
        res.append(acc)
    res.reverse()
    vs = torch.add(torch.stack(res), values)
    vs_tp1 = torch.cat([vs[1:], (torch.ones_like(vs[0]) * bootstrap_value).unsqueeze(0)], dim=0)
    cpg = torch.clamp(rhos, max=clip_pg_rho) if clip_pg_rho is not None else rhos
    return vs, cpg * (rewards + discounts * vs_tp1 - values)


@pytest.mark.gpu
def test_vtrace_kernel_bit_exact_vs_torch_and_golden(golden_dir):
    import moolib_b200
    from moolib_b200 import _C
    g = np.load(f"{golden_dir}/vtrace_golden.npz")
    for ci, (T, B, clip, clip_pg) in enumerate(_cases(g)):
        ins = [torch.from_numpy(a).cuda() for a in _inputs(ci, T, B)]
        before = _C.kernel_launches()
        vs, pg = moolib_b200.vtrace_from_importance_weights(*ins, clip_rho_threshold=clip, clip_pg_rho_threshold=clip_pg)
        assert _C.kernel_launches() - before == 1
        evs, epg = torch_vtrace(*ins, clip, clip_pg)
        assert torch.equal(vs, evs) and torch.equal(pg, epg), f"case {ci}: kernel differs from the PyTorch restatement"
        assert _close(vs.cpu().numpy(), g[f"c{ci}_vs"]) and _close(pg.cpu().numpy(), g[f"c{ci}_pg"])
        ovs, opg = oracle.vtrace(*_inputs(ci, T, B), clip_rho=clip, clip_pg_rho=clip_pg)
        assert _close(vs.cpu().numpy(), ovs.astype(np.float64)) and _close(pg.cpu().numpy(), opg.astype(np.float64))
    # a long unroll (T = 700: the global-memory variant of the kernel), ragged column count
    ins = [torch.randn(700, 45, device="cuda") * 0.2, torch.full((700, 45), 0.97, device="cuda"),
           torch.randn(700, 45, device="cuda"), torch.randn(700, 45, device="cuda"), torch.randn(45, device="cuda")]
    vs, pg = moolib_b200.vtrace_from_importance_weights(*ins)
    evs, epg = torch_vtrace(*ins, 1.0, 1.0)
    assert torch.equal(vs, evs) and torch.equal(pg, epg)
    # the IMPALA learner's shape, extra trailing dimensions, NaN propagation through the clamps
    T, B = 20, 32
    ins = [torch.randn(T, B, 2, device="cuda") * 0.3, torch.full((T, B, 2), 0.99, device="cuda"),
           torch.randn(T, B, 2, device="cuda"), torch.randn(T, B, 2, device="cuda"), torch.randn(B, 2, device="cuda")]
    ins[0][3, 5, 1] = float("nan")
    vs, pg = moolib_b200.vtrace_from_importance_weights(*ins)
    evs, epg = torch_vtrace(*ins, 1.0, 1.0)
    assert torch.equal(torch.isnan(vs), torch.isnan(evs)) and torch.equal(vs.nan_to_num(7.0), evs.nan_to_num(7.0))
    assert torch.equal(pg.nan_to_num(7.0), epg.nan_to_num(7.0))


@pytest.mark.gpu
def test_u8_to_float_kernel_bit_exact(golden_dir):
    import moolib_b200
    from moolib_b200 import _lib
    g = np.load(f"{golden_dir}/vtrace_golden.npz")
    for shape in ([21 * 32, 4, 84, 84], [256, 4, 84, 84], [3, 5, 7], [1], [17]):
        x = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda")
        got = moolib_b200.u8_to_float(x)
        assert got.dtype == torch.float32 and torch.equal(got, x.float() / 255.0), shape
        assert got.cpu().numpy().tobytes() == oracle.u8_to_f32(x.cpu().numpy()).tobytes()
    x = torch.from_numpy(gen_input(int(g["norm_in_seed"]), [3, 5, 4, 8, 8], "u8")).cuda()
    assert np.abs(moolib_b200.u8_to_float(x).cpu().numpy().astype(np.float64) - g["norm_out"]).max() <= 2 ** -24
    # through the C-ABI, unaligned source / destination (no vector path)
    buf = torch.randint(0, 256, (1001,), dtype=torch.uint8, device="cuda")
    out = torch.zeros(1004, device="cuda")
    L = _lib.load()
    _lib.check(L.mb_u8_to_f32(buf.data_ptr() + 1, out.data_ptr() + 4, 1000, ctypes.c_float(1.0 / 255.0),
                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(out[1:1001], buf[1:].float() / 255.0) and out[0] == 0 and out[1001] == 0
