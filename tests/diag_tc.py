#!/usr/bin/env python3
"""Bring-up diagnostic for the tcgen05 convolution kernels: prints one line per (kind, shape) with the error of the
tcgen05 path and of the CUDA-core path against the CPU oracle (torch fp32 conv on bf16-exact inputs).  Each kind
runs in its own process under `timeout`, so a trapped kernel does not hide the other results:

    python tests/diag_tc.py fwd | dgrad | wgrad | simt
"""
import sys
import time

import torch

import util
from util import lib

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def main(kind):
    shapes = util.RESNET_SHAPES + util.ODD_SHAPES + util.WIDE_SHAPES
    N = 3
    print("device:", torch.cuda.get_device_name(0), flush=True)
    worst = 0.0
    for si, shape in enumerate(shapes):
        H, W, ci, co, r, stride = shape
        x, w, dy = util.make_case(shape, N, 100 + si)
        t0 = time.time()
        try:
            if kind == "simt":
                y, _ = util.conv_fwd(x, w, stride, lib.F32, lib.IMPL_SIMT)
                e1 = util.rel_err(y, util.ref_conv(x, w, stride))
                dx = util.conv_dgrad(dy, w, H, W, stride, lib.F32, lib.IMPL_SIMT)
                e2 = util.rel_err(dx, util.ref_dgrad(dy, w, H, W, stride))
                dw = util.conv_wgrad(x, dy, r, stride, lib.F32, lib.IMPL_SIMT)
                e3 = util.rel_err(dw, util.ref_wgrad(x, dy, r, stride))
                print("simt %-28s fwd %.2e dgrad %.2e wgrad %.2e  (%.2fs)" % (shape, e1, e2, e3, time.time() - t0), flush=True)
                worst = max(worst, e1, e2, e3)
                continue
            if kind == "fwd":
                ref = util.ref_conv(x, w, stride)
                y_tc, s_tc = util.conv_fwd(x, w, stride, lib.BF16, lib.IMPL_TCGEN05, stats=True)
                y_si, s_si = util.conv_fwd(x, w, stride, lib.BF16, lib.IMPL_SIMT, stats=True)
                es = util.rel_err(s_tc, s_si)
                e_tc, e_si, e_x = util.rel_err(y_tc, ref), util.rel_err(y_si, ref), util.rel_err(y_tc, y_si)
                print("fwd  %-28s tc-vs-ref %.2e simt-vs-ref %.2e tc-vs-simt %.2e stats %.2e nan=%d (%.2fs)" % (
                    shape, e_tc, e_si, e_x, es, int(torch.isnan(y_tc).sum()), time.time() - t0), flush=True)
                worst = max(worst, e_tc)
            elif kind == "dgrad":
                ref = util.ref_dgrad(dy, w, H, W, stride)
                if r == 1:
                    base = util.bf16_round(torch.randn(N, ci, H, W))
                    d_tc = util.conv_dgrad(dy, w, H, W, stride, lib.BF16, lib.IMPL_TCGEN05, accumulate_into=base)
                    d_si = util.conv_dgrad(dy, w, H, W, stride, lib.BF16, lib.IMPL_SIMT, accumulate_into=base)
                    ref = ref + base
                else:
                    d_tc = util.conv_dgrad(dy, w, H, W, stride, lib.BF16, lib.IMPL_TCGEN05)
                    d_si = util.conv_dgrad(dy, w, H, W, stride, lib.BF16, lib.IMPL_SIMT)
                e_tc, e_si, e_x = util.rel_err(d_tc, ref), util.rel_err(d_si, ref), util.rel_err(d_tc, d_si)
                print("dgrad %-28s tc-vs-ref %.2e simt-vs-ref %.2e tc-vs-simt %.2e nan=%d (%.2fs)" % (
                    shape, e_tc, e_si, e_x, int(torch.isnan(d_tc).sum()), time.time() - t0), flush=True)
                worst = max(worst, e_tc)
            elif kind == "wgrad":
                ref = util.ref_wgrad(x, dy, r, stride)
                g_tc = util.conv_wgrad(x, dy, r, stride, lib.BF16, lib.IMPL_TCGEN05)
                g_si = util.conv_wgrad(x, dy, r, stride, lib.BF16, lib.IMPL_SIMT)
                e_tc, e_si = util.rel_err(g_tc, ref), util.rel_err(g_si, ref)
                print("wgrad %-28s tc-vs-ref %.2e simt-vs-ref %.2e nan=%d (%.2fs)" % (
                    shape, e_tc, e_si, int(torch.isnan(g_tc).sum()), time.time() - t0), flush=True)
                worst = max(worst, e_tc)
        except Exception as ex:  # keep going: one unsupported shape must not hide the rest
            print("%s %-28s EXCEPTION %s" % (kind, shape, str(ex)[:200]), flush=True)
            if "CUDA" in str(ex) or "launch" in str(ex):
                break
    print("%s worst error vs oracle: %.3e" % (kind, worst), flush=True)


if __name__ == "__main__":
    main(sys.argv[1])
