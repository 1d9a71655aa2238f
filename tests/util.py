"""Helpers shared by the GPU parity tests: NCHW<->NHWC conversion, C-ABI convolution calls, error metrics."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pytorch-kaldi-resnet_b200")
for p in (ROOT, PKG, os.path.join(PKG, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)

from svk import lib  # noqa: E402
from svk.lib import call  # noqa: E402

# the 14 conv shapes of ResNet-34 on 40x200 input + odd-size variants: (H, W, Cin, Cout, R, stride)
RESNET_SHAPES = [
    (40, 200, 32, 32, 3, 1), (40, 200, 32, 64, 3, 2), (20, 100, 64, 64, 3, 1), (20, 100, 64, 128, 3, 2),
    (10, 50, 128, 128, 3, 1), (10, 50, 128, 256, 3, 2), (5, 25, 256, 256, 3, 1),
    (40, 200, 32, 64, 1, 2), (20, 100, 64, 128, 1, 2), (10, 50, 128, 256, 1, 2),
]
ODD_SHAPES = [
    (30, 51, 32, 32, 3, 1), (30, 51, 32, 64, 3, 2), (15, 26, 64, 128, 3, 2), (8, 13, 128, 256, 3, 2),
    (4, 7, 256, 256, 3, 1), (15, 26, 64, 128, 1, 2), (5, 1, 256, 256, 3, 1), (7, 9, 64, 64, 3, 1),
]
WIDE_SHAPES = [(10, 38, 256, 512, 3, 2), (5, 19, 512, 512, 3, 1), (10, 38, 256, 512, 1, 2), (20, 75, 64, 64, 3, 1)]


def st():
    return torch.cuda.current_stream().cuda_stream


def tdtype(code):
    return torch.bfloat16 if code == lib.BF16 else torch.float32


def nhwc(t_nchw, code):
    """CPU/GPU NCHW fp32 -> CUDA NHWC tensor of the storage dtype."""
    return t_nchw.permute(0, 2, 3, 1).contiguous().to("cuda", tdtype(code))


def nchw(t_nhwc):
    return t_nhwc.float().cpu().permute(0, 3, 1, 2).contiguous()


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def pack(w_oihw, code):
    """OIHW fp32 (CPU) -> (w_fwd, w_dgrad) packed device tensors via svk_pack_conv_weight."""
    co, ci, r, _ = w_oihw.shape
    w = w_oihw.contiguous().float().cuda()
    wf = torch.empty(r * r * co * ci, dtype=tdtype(code), device="cuda")
    wd = torch.empty_like(wf)
    call.svk_pack_conv_weight(w.data_ptr(), wf.data_ptr(), wd.data_ptr(), co, ci, r, code, st())
    return wf, wd


def conv_fwd(x_nchw, w_oihw, stride, code, impl, stats=False, scale=None, shift=None, res_nchw=None, relu=0,
             valid_wo=None):
    N, ci, H, W = x_nchw.shape
    co, _, r, _ = w_oihw.shape
    d = lib.make_conv_desc(N, H, W, ci, co, r, stride, code, impl)
    x = nhwc(x_nchw, code)
    wf, _ = pack(w_oihw, code)
    y = torch.full((N, d.Ho, d.Wo, co), float("nan"), dtype=tdtype(code), device="cuda")
    sbuf = torch.zeros(2 * co, dtype=torch.float64, device="cuda") if stats else None
    sc = scale.float().cuda() if scale is not None else None
    sh = shift.float().cuda() if shift is not None else None
    res = nhwc(res_nchw, code) if res_nchw is not None else None
    vw = valid_wo.to("cuda", torch.int32) if valid_wo is not None else None
    call.svk_conv2d_fwd(d, x.data_ptr(), wf.data_ptr(), y.data_ptr(), 0 if sbuf is None else sbuf.data_ptr(),
                        0 if sc is None else sc.data_ptr(), 0 if sh is None else sh.data_ptr(),
                        0 if res is None else res.data_ptr(), relu, 0 if vw is None else vw.data_ptr(), st())
    torch.cuda.synchronize()
    return nchw(y), (sbuf.cpu() if stats else None)


def conv_dgrad(dy_nchw, w_oihw, H, W, stride, code, impl, res_nchw=None, resm_nchw=None, mask_nchw=None,
               accumulate_into=None):
    N, co, Ho, Wo = dy_nchw.shape
    _, ci, r, _ = w_oihw.shape
    d = lib.make_conv_desc(N, H, W, ci, co, r, stride, code, impl)
    assert (d.Ho, d.Wo) == (Ho, Wo)
    dy = nhwc(dy_nchw, code)
    _, wd = pack(w_oihw, code)
    if accumulate_into is not None:
        dx = nhwc(accumulate_into, code)
        res_ptr = dx.data_ptr()
    else:
        dx = torch.full((N, H, W, ci), float("nan"), dtype=tdtype(code), device="cuda")
        res = nhwc(res_nchw, code) if res_nchw is not None else None
        res_ptr = 0 if res is None else res.data_ptr()
    rm = nhwc(resm_nchw, code) if resm_nchw is not None else None
    mk = nhwc(mask_nchw, code) if mask_nchw is not None else None
    call.svk_conv2d_dgrad(d, dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), res_ptr, 0 if rm is None else rm.data_ptr(),
                          0 if mk is None else mk.data_ptr(), st())
    torch.cuda.synchronize()
    return nchw(dx)


def conv_wgrad(x_nchw, dy_nchw, r, stride, code, impl):
    N, ci, H, W = x_nchw.shape
    co = dy_nchw.shape[1]
    d = lib.make_conv_desc(N, H, W, ci, co, r, stride, code, impl)
    x = nhwc(x_nchw, code)
    dy = nhwc(dy_nchw, code)
    need = lib.load().svk_conv2d_wgrad_workspace_bytes(d)
    assert need > 0, lib.load().svk_last_error_string()
    ws = torch.full(((need + 3) // 4,), float("nan"), dtype=torch.float32, device="cuda")
    dw = torch.full((co, ci, r, r), float("nan"), dtype=torch.float32, device="cuda")
    call.svk_conv2d_wgrad(d, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel() * 4, st())
    torch.cuda.synchronize()
    return dw.cpu()


def ref_conv(x, w, stride):
    return F.conv2d(x, w, None, stride, w.shape[-1] // 2)


def ref_dgrad(dy, w, H, W, stride):
    return torch.nn.grad.conv2d_input((dy.shape[0], w.shape[1], H, W), w, dy, stride, w.shape[-1] // 2)


def ref_wgrad(x, dy, r, stride):
    return torch.nn.grad.conv2d_weight(x, (dy.shape[1], x.shape[1], r, r), dy, stride, r // 2)


def rel_err(a, b):
    """max |a-b| / max |b|  (the per-tensor relative error used for the north-star tolerances)."""
    a, b = a.double(), b.double()
    denom = float(b.abs().max())
    if denom == 0.0:
        return float((a - b).abs().max())
    return float((a - b).abs().max()) / denom


def make_case(shape, N, seed, quantize=True):
    H, W, ci, co, r, stride = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, ci, H, W, generator=g)
    w = torch.randn(co, ci, r, r, generator=g) * (2.0 / (ci * r * r)) ** 0.5
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    dy = torch.randn(N, co, Ho, Wo, generator=g)
    if quantize:       # values exactly representable in bf16, so fp32 and bf16 paths see identical inputs
        x, w, dy = bf16_round(x), bf16_round(w), bf16_round(dy)
    return x, w, dy


def sample_of(t, ns=96):
    """Same subsample as oracle/make_golden.py::sample (norm, mean, numel, strided values)."""
    v = t.detach().double().reshape(-1).cpu()
    n = v.numel()
    k = min(ns, n)
    idx = (torch.arange(k, dtype=torch.int64) * (n - 1)) // max(k - 1, 1)
    return np.concatenate([[float(v.norm()), float(v.mean()), float(n)], v[idx].numpy()])
