#!/usr/bin/env python3
"""Per-role cycle breakdown of the wgrad9 / wgradr kernels (SVK_PROF=1 counters).

    SVK_PROF=1 python tests/prof_wgrad.py [N]
"""
import ctypes
import os
import sys

os.environ.setdefault("SVK_PROF", "1")
import torch

import util
from util import lib, call

SHAPES = [(40, 200, 32, 32, 3, 1), (20, 100, 64, 64, 3, 1), (10, 50, 128, 128, 3, 1), (5, 25, 256, 256, 3, 1)]


def read():
    buf = (ctypes.c_ulonglong * 16)()
    assert lib.load().svk_debug_prof_read(buf) == 0
    return list(buf)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    st = util.st()
    print("%-26s %8s | %9s %9s | %9s %9s | %9s %9s  (k-cycles per CTA)" % (
        "shape", "ms", "mma:loop", "wait-opnd", "tma:loop", "wait-slot", "epi:total", "wait-mma"))
    for shape in SHAPES:
        H, W, ci, co, r, s = shape
        d = lib.make_conv_desc(N, H, W, ci, co, r, s, lib.BF16, lib.IMPL_TCGEN05)
        x = torch.randn(N, H, W, ci, device="cuda").bfloat16()
        y = torch.randn(N, d.Ho, d.Wo, co, device="cuda").bfloat16()
        need = lib.load().svk_conv2d_wgrad_workspace_bytes(d)
        ws = torch.empty((need + 3) // 4, dtype=torch.float32, device="cuda")
        dw = torch.empty(co, ci, r, r, device="cuda")

        def fn():
            call.svk_conv2d_wgrad(d, x.data_ptr(), y.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel() * 4, st)

        fn()
        torch.cuda.synchronize()
        read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        p = read()
        n = max(p[0], 1)
        k = [v / n / 1e3 for v in p]
        span_us = (p[10] - ((~p[9]) & 0xFFFFFFFFFFFFFFFF)) / 1e3
        print("%-26s %8.4f | %9.1f %9.1f | %9.1f %9.1f | %9.1f %9.1f   CTAs=%d mma-loop span %.1f us, SM clock %.3f GHz" % (
            shape, e0.elapsed_time(e1), k[1], k[2], k[4], k[5], k[6], k[7], p[0], span_us, p[1] / max(p[8], 1)), flush=True)


if __name__ == "__main__":
    main()
