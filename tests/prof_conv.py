#!/usr/bin/env python3
"""Per-role cycle breakdown of the fprop / dgrad tensor-core kernels (SVK_PROF=1 counters): where a persistent CTA's
time goes — the MMA-issuing warp waiting for operands (TMA) or for a free accumulator (epilogue), the producer waiting
for a free stage, the first epilogue warp waiting for an accumulator.

    SVK_PROF=1 python tests/prof_conv.py [N] [stage-filter]
"""
import ctypes
import os
import sys

os.environ.setdefault("SVK_PROF", "1")
import torch

import util
from util import lib, call

SHAPES = [(40, 200, 32, 32, 3, 1), (20, 100, 64, 64, 3, 1), (10, 50, 128, 128, 3, 1), (5, 25, 256, 256, 3, 1),
          (40, 200, 32, 64, 3, 2)]


def read():
    buf = (ctypes.c_ulonglong * 16)()
    rc = lib.load().svk_debug_prof_read(buf)
    assert rc == 0, rc
    return list(buf)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    st = util.st()
    print("%-6s %-26s %8s | %9s %9s %9s | %9s %9s | %9s %9s   (k-cycles per CTA)" % (
        "op", "shape", "ms", "mma:loop", "wait-opnd", "wait-acc", "tma:loop", "wait-slot", "epi:loop", "wait-acc"))
    for shape in SHAPES:
        H, W, ci, co, r, s = shape
        if filt and filt != "%d" % ci:
            continue
        d = lib.make_conv_desc(N, H, W, ci, co, r, s, lib.BF16, lib.IMPL_TCGEN05)
        x = torch.randn(N, H, W, ci, device="cuda").bfloat16()
        y = torch.randn(N, d.Ho, d.Wo, co, device="cuda").bfloat16()
        c = torch.randn(N, H, W, ci, device="cuda").bfloat16()
        mask = torch.randn(N, H, W, ci, device="cuda").bfloat16()      # distinct operand streams, as in the training step
        w = torch.randn(co, ci, r, r) * 0.05
        wf, wd = util.pack(w, lib.BF16)
        stats = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
        sums = torch.zeros(2 * ci, dtype=torch.float64, device="cuda")
        mean = torch.zeros(ci, device="cuda")
        rstd = torch.ones(ci, device="cuda")
        fuse = lib.BnBwdFuse(mask.data_ptr(), c.data_ptr(), mean.data_ptr(), rstd.data_ptr(), sums.data_ptr())

        def fwd():
            call.svk_conv2d_fwd(d, x.data_ptr(), wf.data_ptr(), y.data_ptr(), stats.data_ptr(), 0, 0, 0, 0, 0, st)

        def fwd_nostats():
            call.svk_conv2d_fwd(d, x.data_ptr(), wf.data_ptr(), y.data_ptr(), 0, 0, 0, 0, 0, 0, st)

        def dgrad():
            call.svk_conv2d_dgrad(d, y.data_ptr(), wd.data_ptr(), x.data_ptr(), 0, 0, 0, st)

        def dgrad_bn():
            call.svk_conv2d_dgrad_bn(d, y.data_ptr(), wd.data_ptr(), x.data_ptr(), 0, 0, 0, ctypes.byref(fuse), st)

        for name, fn in (("fwd", fwd), ("fwd-ns", fwd_nostats), ("dgrad", dgrad), ("dgr-bn", dgrad_bn)):
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
            span_us = (p[10] - ((~p[9]) & 0xFFFFFFFFFFFFFFFF)) / 1e3       # first MMA-loop start .. last MMA-loop end (last launch)
            ghz = p[1] / max(p[8], 1)
            t_first = (~p[9]) & 0xFFFFFFFFFFFFFFFF
            start_spread = (p[11] - t_first) / 1e3
            end_spread = (p[10] - ((~p[12]) & 0xFFFFFFFFFFFFFFFF)) / 1e3
            print("%-6s %-26s %8.4f | %9.1f %9.1f %9.1f | %9.1f %9.1f | %9.1f %9.1f   CTAs=%d  mma-loop span %.1f us (starts within %.1f, ends within %.1f), SM clock %.3f GHz, peer producer waits slot %.1f" % (
                name, shape, e0.elapsed_time(e1), k[1], k[2], k[3], k[4], k[5], k[6], k[7], p[0], span_us, start_spread, end_spread, ghz, k[14]), flush=True)


if __name__ == "__main__":
    main()
