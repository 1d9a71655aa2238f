#!/usr/bin/env python3
"""Cohort top-k statistics (compute_topk_mean_std.py:10-23): time per piece for several row-block sizes.
   python tests/bench_snorm.py [rows]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-kaldi-resnet_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from svk import lib, scoring  # noqa: E402
from svk.lib import call  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = "cuda"
X = scoring.l2_normalize_rows(torch.randn(n, 256, device=dev))
C = scoring.l2_normalize_rows(torch.randn(50000, 256, device=dev))
st = torch.cuda.current_stream().cuda_stream


def ev(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for br in (256, 512, 1024, 2048, 4096, 8192):
    S = torch.empty(br, 50000, device=dev)
    need = lib.load().svk_gemm_tf32_workspace_bytes(br, 50000, 256)
    ws = torch.empty(need // 4 + 4, device=dev)
    m, s = torch.empty(br, device=dev), torch.empty(br, device=dev)
    tg = ev(lambda: call.svk_gemm_tf32(X.data_ptr(), 256, 1, C.data_ptr(), 256, 1, S.data_ptr(), 50000, br, 50000, 256, 0, ws.data_ptr(), need, st))
    tk = ev(lambda: call.svk_topk_meanstd(S.data_ptr(), br, 50000, 300, m.data_ptr(), s.data_ptr(), st))
    tall = ev(lambda: scoring.cohort_topk_meanstd(X, C, topk=300, block_rows=br, tf32=True), 1)
    print("block %5d rows: gemm_tf32 %.3f ms (%.1f TFLOP/s)  topk %.3f ms (%.2f M rows/s)  | whole %d rows: %.2f ms = %.2f M rows/s" % (
        br, tg, 2.0 * br * 50000 * 256 / tg / 1e9, tk, br / tk / 1e3, n, tall, n / tall / 1e3))
