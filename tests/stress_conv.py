#!/usr/bin/env python3
"""Stress run of the resident-filter conv kernels under whatever SVK_* switches the environment selects: ITER launches of the
training forward (with statistics) and of the fused BatchNorm-backward data gradient, every result compared bit for bit with
the first launch's; prints a checksum so that runs under different switches can be compared with each other.

    python tests/stress_conv.py <channels> [iters]
"""
import ctypes
import sys

import torch

import util
from util import lib, call


def run(ch, iters):
    H, W = {32: (40, 200), 64: (20, 100)}[ch]
    N = 256
    st = util.st()
    torch.manual_seed(5)
    d = lib.make_conv_desc(N, H, W, ch, ch, 3, 1, lib.BF16, lib.IMPL_TCGEN05)
    x = torch.randn(N, H, W, ch, device="cuda").bfloat16()
    dy = torch.randn(N, H, W, ch, device="cuda").bfloat16()
    c = torch.randn(N, H, W, ch, device="cuda").bfloat16()
    mask = torch.randn(N, H, W, ch, device="cuda").bfloat16()
    w = torch.randn(ch, ch, 3, 3) * 0.05
    wf, wd = util.pack(w, lib.BF16)
    mean = torch.randn(ch, device="cuda") * 0.1
    rstd = torch.rand(ch, device="cuda") + 0.5
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    stats = torch.zeros(2 * ch, dtype=torch.float64, device="cuda")
    sums = torch.zeros(2 * ch, dtype=torch.float64, device="cuda")
    fuse = lib.BnBwdFuse(mask.data_ptr(), c.data_ptr(), mean.data_ptr(), rstd.data_ptr(), sums.data_ptr())
    ref = {}
    bad = 0
    for it in range(iters):
        stats.zero_(); sums.zero_()
        call.svk_conv2d_fwd(d, x.data_ptr(), wf.data_ptr(), y.data_ptr(), stats.data_ptr(), 0, 0, 0, 0, 0, st)
        call.svk_conv2d_dgrad_bn(d, dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), 0, 0, 0, ctypes.byref(fuse), st)
        if it == 0:
            torch.cuda.synchronize()
            ref = {"y": y.clone(), "dx": dx.clone(), "stats": stats.clone(), "sums": sums.clone()}
        elif it % 10 == 0 or it == iters - 1:
            ok = torch.equal(y, ref["y"]) and torch.equal(dx, ref["dx"])
            ok = ok and torch.allclose(stats, ref["stats"], rtol=1e-9, atol=1e-6) and torch.allclose(sums, ref["sums"], rtol=1e-9, atol=1e-6)
            bad += 0 if ok else 1
    torch.cuda.synchronize()
    print("ch %d iters %d mismatches %d | y %.6f dx %.6f stats %.6f sums %.6f" % (
        ch, iters, bad, ref["y"].double().abs().sum().item(), ref["dx"].double().abs().sum().item(),
        ref["stats"].abs().sum().item(), ref["sums"].abs().sum().item()))
    return bad


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 200)
