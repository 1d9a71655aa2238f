#!/usr/bin/env python3
"""Per-stage timing of the HBM-bound BatchNorm kernels at the bench batch (N=256): CUDA events, rotating buffers.
    python tests/bench_bn.py [N] [iters] [C-filter]
Prints achieved GB/s on the algorithmic bytes of each launch."""
import sys

import torch

import util
from util import lib, call

STAGES = [(40, 200, 32), (20, 100, 64), (10, 50, 128), (5, 25, 256)]


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    filt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    st = util.st()
    nbuf = 4
    for H, W, C in STAGES:
        if filt and C != filt:
            continue
        M = N * H * W
        mk = lambda: [torch.randn(M, C, device="cuda").bfloat16() for _ in range(nbuf)]
        x, g, o, r = mk(), mk(), mk(), mk()
        stats = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
        call.svk_channel_stats(x[0].data_ptr(), M, C, lib.BF16, stats.data_ptr(), st)
        gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        coef = torch.zeros(4, C, device="cuda")
        sums = torch.zeros(3, C, dtype=torch.float64, device="cuda")
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        mu, rs = coef[2], coef[3]
        tsz = M * C * 2

        def act(i):
            call.svk_bn_train_act_fwd(x[i % nbuf].data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(),
                                      rv.data_ptr(), coef.data_ptr(), 0, 0, 0, 0, 0, 0, 0, C, 0.1, 1e-5, 1, o[i % nbuf].data_ptr(),
                                      M, C, lib.BF16, st)

        def act_res(i):
            call.svk_bn_train_act_fwd(x[i % nbuf].data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(),
                                      rv.data_ptr(), coef.data_ptr(), r[i % nbuf].data_ptr(), 0, 0, 0, 0, 0, 0, C, 0.1, 1e-5, 1,
                                      o[i % nbuf].data_ptr(), M, C, lib.BF16, st)

        def reduce(i):
            call.svk_bn_bwd_reduce(g[i % nbuf].data_ptr(), o[i % nbuf].data_ptr(), x[i % nbuf].data_ptr(), mu.data_ptr(),
                                   rs.data_ptr(), 0, 0, 0, sums.data_ptr(), M, C, lib.BF16, st)

        def apply_nomask(i):
            call.svk_bn_bwd_apply(g[i % nbuf].data_ptr(), 0, x[i % nbuf].data_ptr(), mu.data_ptr(), rs.data_ptr(),
                                  gamma.data_ptr(), r[i % nbuf].data_ptr(), 0, 0, 0, 0, 0, sums.data_ptr(), dg.data_ptr(),
                                  db.data_ptr(), 0, 0, M, C, lib.BF16, st)

        def apply_mask(i):
            call.svk_bn_bwd_apply(g[i % nbuf].data_ptr(), o[i % nbuf].data_ptr(), x[i % nbuf].data_ptr(), mu.data_ptr(),
                                  rs.data_ptr(), gamma.data_ptr(), r[i % nbuf].data_ptr(), 0, 0, 0, 0, 0, sums.data_ptr(),
                                  dg.data_ptr(), db.data_ptr(), 0, 0, M, C, lib.BF16, st)

        for name, fn, passes in (("bn_train_act", act, 2), ("bn_train_act+res", act_res, 3), ("bn_bwd_reduce", reduce, 3),
                                 ("bn_bwd_apply", apply_nomask, 3), ("bn_bwd_apply+mask", apply_mask, 4)):
            fn(0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                fn(i + 1)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print("%-18s M=%-8d C=%-4d %8.4f ms  %7.1f GB/s (%d passes of %.1f MB)" % (name, M, C, ms, passes * tsz / ms / 1e6,
                                                                                    passes, tsz / 1e6))


if __name__ == "__main__":
    main()
