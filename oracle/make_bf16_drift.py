#!/usr/bin/env python3
"""How far does bf16 STORAGE move the benchmark network?  (test infrastructure; writes tests/golden/bf16_storage_drift.json)

The oracle (fp32 arithmetic on the CPU, pinned to the reference by make_golden.py) runs one training step of the cfg3 model
(5,994 speakers, 40 x 200 chunks, random init seed 1234) twice: as is, and with every tensor the CUDA engine stores in bf16
(conv weights, conv outputs, ReLU outputs, block outputs, and the gradients flowing back through them) rounded to bf16.
The difference is what ANY implementation with bf16 activation storage shows on this network — a random-init ResNet-34 with
training-mode BatchNorm amplifies a perturbation by ~1.1x per convolution — and is the yardstick
tests/test_bench_config_gpu.py holds the CUDA path to at batch 256.  Batch 32 here (the drift is batch-independent to ~3 %
once BatchNorm statistics are stable; measured 32 vs 256).

    python oracle/make_bf16_drift.py [batch]
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_model as O  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    C = 5994
    sd = O.init_state(C, 40, "mean+std", "AAM", seed=1234)
    names = O.param_names(sd)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(B, 40, 200, generator=g)
    y = torch.randint(0, C, (B,), generator=g)

    def run(rounding):
        s = {k: v.clone() for k, v in sd.items()}
        for n in names:
            s[n].requires_grad_(True)
        taps = {}
        if rounding:
            with O.storage_rounding(torch.bfloat16, grads=True):
                lg = O.model_forward(s, x, y, "mean+std", "AAM", 0.2, 30, True, {}, taps)
        else:
            lg = O.model_forward(s, x, y, "mean+std", "AAM", 0.2, 30, True, {}, taps)
        loss = O.cross_entropy(lg, y)
        loss.backward()
        return float(loss.detach()), {n: s[n].grad.clone() for n in names}, {k: v.detach() for k, v in taps.items()}
    l0, g0, t0 = run(False)
    l1, g1, t1 = run(True)
    out = {"batch": B, "loss_fp32": l0, "loss_bf16_storage": l1, "activations": {}, "gradients": {}}
    for k in t0:
        if k.startswith("pre/") or k in ("pool", "res.stem"):
            continue
        a, b = t1[k].double(), t0[k].double()
        out["activations"]["embedding" if k == "fc1" else k] = float(((a - b) ** 2).mean().sqrt() / (b ** 2).mean().sqrt())
    for n in names:
        a, b = g1[n].double().reshape(1, -1), g0[n].double().reshape(1, -1)
        out["gradients"][n] = float(F.cosine_similarity(a, b))
    path = os.path.join(ROOT, "tests", "golden", "bf16_storage_drift.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    acts = out["activations"]
    print("wrote", path, "| loss %.5f vs %.5f | layer4.2.conv2 rms-rel %.4f | layer1.0.conv1.weight cosine %.4f" % (
        l0, l1, acts["res.layer4.2.conv2"], out["gradients"]["res.layer1.0.conv1.weight"]))


if __name__ == "__main__":
    main()
