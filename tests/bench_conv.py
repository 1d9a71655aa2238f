#!/usr/bin/env python3
"""Per-shape timing of the tcgen05 convolution kernels at the bench batch (N=256), CUDA events, rotating buffers so
every launch reads from HBM/L2 like inside a training step.  Also the ncu target for profiles/:

    python tests/bench_conv.py [N] [iters] [stage-filter]
"""
import sys

import torch

import util
from util import lib, call

SHAPES = [(40, 200, 32, 32, 3, 1), (20, 100, 64, 64, 3, 1), (10, 50, 128, 128, 3, 1), (5, 25, 256, 256, 3, 1),
          (40, 200, 32, 64, 3, 2), (40, 200, 32, 64, 1, 2)]


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    st = util.st()
    nbuf = 3
    for shape in SHAPES:
        H, W, ci, co, r, s = shape
        if filt and filt != "%d" % ci:
            continue
        d = lib.make_conv_desc(N, H, W, ci, co, r, s, lib.BF16, lib.IMPL_TCGEN05)
        xs = [torch.randn(N, H, W, ci, device="cuda").bfloat16() for _ in range(nbuf)]
        ys = [torch.randn(N, d.Ho, d.Wo, co, device="cuda").bfloat16() for _ in range(nbuf)]
        w = torch.randn(co, ci, r, r) * 0.05
        wf, wd = util.pack(w, lib.BF16)
        stats = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
        need = lib.load().svk_conv2d_wgrad_workspace_bytes(d)
        ws = torch.empty((need + 3) // 4, dtype=torch.float32, device="cuda")
        dw = torch.empty(co, ci, r, r, device="cuda")
        flops = 2.0 * N * d.Ho * d.Wo * co * ci * r * r

        def fwd(i):
            call.svk_conv2d_fwd(d, xs[i % nbuf].data_ptr(), wf.data_ptr(), ys[i % nbuf].data_ptr(), stats.data_ptr(), 0, 0, 0, 0, 0, st)

        def dgrad(i):
            if r == 1:
                call.svk_conv2d_dgrad(d, ys[i % nbuf].data_ptr(), wd.data_ptr(), xs[i % nbuf].data_ptr(), xs[i % nbuf].data_ptr(), 0, 0, st)
            else:
                call.svk_conv2d_dgrad(d, ys[i % nbuf].data_ptr(), wd.data_ptr(), xs[i % nbuf].data_ptr(), 0, 0, 0, st)

        def wgrad(i):
            call.svk_conv2d_wgrad(d, xs[i % nbuf].data_ptr(), ys[i % nbuf].data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel() * 4, st)

        for name, fn in (("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)):
            fn(0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                fn(i + 1)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print("%-6s %-28s %8.4f ms  %7.1f TFLOP/s" % (name, shape, ms, flops / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
