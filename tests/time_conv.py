#!/usr/bin/env python3
"""Back-to-back launch time of the tensor-core convolution kernels at the bench size (no profiling counters): 30 launches
between two CUDA events, per-launch microseconds.  Rows: training forward (with statistics), forward without statistics, fused
BatchNorm-backward data gradient with three distinct operand tensors, as in the training step.

    python tests/time_conv.py [N]
"""
import ctypes
import sys

import torch

import util
from util import lib, call

SHAPES = [(40, 200, 32, 32, 3, 1), (20, 100, 64, 64, 3, 1), (10, 50, 128, 128, 3, 1), (5, 25, 256, 256, 3, 1)]


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    st = util.st()
    for shape in SHAPES:
        H, W, ci, co, r, s = shape
        d = lib.make_conv_desc(N, H, W, ci, co, r, s, lib.BF16, lib.IMPL_TCGEN05)
        x = torch.randn(N, H, W, ci, device="cuda").bfloat16()
        y = torch.randn(N, d.Ho, d.Wo, co, device="cuda").bfloat16()
        c = torch.randn(N, H, W, ci, device="cuda").bfloat16()
        mask = torch.randn(N, H, W, ci, device="cuda").bfloat16()
        res = torch.randn(N, H, W, ci, device="cuda").bfloat16()
        dx = torch.empty_like(x)
        w = torch.randn(co, ci, r, r) * 0.05
        wf, wd = util.pack(w, lib.BF16)
        stats = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
        sums = torch.zeros(2 * ci, dtype=torch.float64, device="cuda")
        mean = torch.zeros(ci, device="cuda")
        rstd = torch.ones(ci, device="cuda")
        fuse = lib.BnBwdFuse(mask.data_ptr(), c.data_ptr(), mean.data_ptr(), rstd.data_ptr(), sums.data_ptr())
        ops = {
            "fwd+stats": lambda: call.svk_conv2d_fwd(d, x.data_ptr(), wf.data_ptr(), y.data_ptr(), stats.data_ptr(), 0, 0, 0, 0, 0, st),
            "fwd": lambda: call.svk_conv2d_fwd(d, x.data_ptr(), wf.data_ptr(), y.data_ptr(), 0, 0, 0, 0, 0, 0, st),
            "dgrad-bn": lambda: call.svk_conv2d_dgrad_bn(d, y.data_ptr(), wd.data_ptr(), dx.data_ptr(), 0, 0, 0, ctypes.byref(fuse), st),
            "dgrad-bn+res": lambda: call.svk_conv2d_dgrad_bn(d, y.data_ptr(), wd.data_ptr(), dx.data_ptr(), res.data_ptr(), 0, 0, ctypes.byref(fuse), st),
        }
        out = []
        for name, fn in ops.items():
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out.append("%s %6.1f us" % (name, e0.elapsed_time(e1) / 30 * 1e3))
        print("%-26s %s" % (shape, " | ".join(out)), flush=True)


if __name__ == "__main__":
    main()
