#!/usr/bin/env python3
"""Bring-up diagnostic: one training step of the CUDA model against the live CPU oracle, every activation gradient and
parameter gradient compared in full (not sampled).   python tests/diag_model.py <case> <precision> [B] [T]"""
import contextlib
import io
import sys

import numpy as np
import torch

import util
from oracle import ref_model as O

CASES = {
    "aam_f40": dict(spk_num=37, feat_dim=40, pooling="mean+std", loss="AAM"),
    "softmax_f30": dict(spk_num=11, feat_dim=30, pooling="mean", loss="softmax"),
    "aamv1_f40": dict(spk_num=19, feat_dim=40, pooling="mean+std", loss="AAM-v1"),
}


def main(case, precision, B=2, T=48, impl=None):
    from model import NeuralSpeakerModel
    from svk.loss import CrossEntropyLoss
    kw = CASES[case]
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(precision=precision, impl=impl, **kw)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, kw["feat_dim"], T, generator=g)
    y = torch.randint(0, kw["spk_num"], (B,), generator=g)
    # oracle
    names = O.param_names(sd)
    for n_ in names:
        sd[n_].requires_grad_(True)
    taps = {}
    ctx = O.storage_rounding(torch.bfloat16) if (precision == "bf16" and "fp32ref" not in sys.argv) else contextlib.nullcontext()
    with ctx:
        out = O.model_forward(sd, x, y, kw["pooling"], kw["loss"], 0.2, 30, True, {}, taps)
        loss = O.cross_entropy(out, y)
        loss.backward()
    # CUDA
    m.train()
    eng = m.engine
    eng.debug = {}
    logits = m(x.cuda(), y.cuda())
    l2 = CrossEntropyLoss()(logits, y.cuda())
    l2.backward()
    torch.cuda.synchronize()
    print("loss cuda %.7f oracle %.7f   logits rel err %.3e" % (float(l2), float(loss), util.rel_err(logits.detach().cpu(), out.detach())))
    ws = eng._train_ws[(B, kw["feat_dim"], T)][0]
    outs = {b.name: ws["o_%d" % bi] for bi, b in enumerate(eng.blocks)}
    rows = []
    for nm, t in eng.debug.items():
        ref = taps[nm].grad
        got = util.nchw(t)
        diff = (got - ref).abs()
        den = float(ref.abs().max())
        bad = diff > 1e-3 * den
        extra = ""
        if nm in outs and bad.any():
            o = util.nchw(outs[nm])
            extra = " | bad where out==0: %d of %d bad" % (int((bad & (o == 0)).sum()), int(bad.sum()))
            extra += " ; oracle out==0 there: %d" % int((bad & (taps[nm].detach() == 0)).sum())
        okm = ~torch.isnan(ref)
        l2 = float((got - ref)[okm].norm() / ref[okm].norm())
        cs = float(torch.nn.functional.cosine_similarity(got[okm].reshape(1, -1), ref[okm].reshape(1, -1)))
        rows.append((float(diff.max()) / den, nm, int(bad.sum()), ref.numel(), " l2rel %.3e cos %.6f" % (l2, cs) + extra))
    for e, nm, nb, n, extra in rows:
        print("dact %-28s rel %.3e  bad %6d / %d%s" % (nm, e, nb, n, extra))
    for nm in ["res.conv1"] + [b.name + s_ for b in eng.blocks for s_ in (".conv1", ".conv2", "")]:
        t = ws["c0"] if nm == "res.conv1" else None
        if t is None:
            bi = [b.name for b in eng.blocks].index(nm.replace(".conv1", "").replace(".conv2", ""))
            t = ws[("c1_%d" if nm.endswith("conv1") else "c2_%d" if nm.endswith("conv2") else "o_%d") % bi]
        ref = taps[nm].detach()
        got = util.nchw(t)
        print("act  %-28s rel %.3e l2rel %.3e" % (nm, util.rel_err(got, ref), float((got - ref).norm() / ref.norm())))
    prow = []
    for nm, p in m.named_parameters():
        ref = sd[nm].grad
        gg = p.grad.cpu()
        prow.append((util.rel_err(gg, ref), nm, float(ref.abs().max()), float((gg - ref).norm() / ref.norm())))
    for e, nm, mx, l2 in sorted(prow, reverse=True)[:12] + [r for r in prow if not r[1].startswith("res.")]:
        print("grad %-36s rel %.3e  l2rel %.3e (max |ref| %.3e)" % (nm, e, l2, mx))


if __name__ == "__main__":
    a = sys.argv
    main(a[1], a[2], int(a[3]) if len(a) > 3 else 2, int(a[4]) if len(a) > 4 else 48, a[5] if len(a) > 5 else None)
