"""Helpers for the CPU-only tests (no libsvk compute calls, no CUDA)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pytorch-kaldi-resnet_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG, os.path.join(PKG, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)


def rel_err(a, b):
    a, b = a.double(), b.double()
    d = float(b.abs().max())
    return float((a - b).abs().max()) / (d if d > 0 else 1.0)


def sample_of(t, ns=96):
    """Same subsample as oracle/make_golden.py::sample (norm, mean, numel, strided values)."""
    v = t.detach().double().reshape(-1)
    n = v.numel()
    k = min(ns, n)
    idx = (torch.arange(k, dtype=torch.int64) * (n - 1)) // max(k - 1, 1)
    return np.concatenate([[float(v.norm()), float(v.mean()), float(n)], v[idx].numpy()])
