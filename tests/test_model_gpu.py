"""End-to-end parity of the CUDA path against the reference-generated golden fixtures (tests/golden/*.npz, written by
oracle/make_golden.py from the UNMODIFIED reference) and against the live CPU oracle.

Tolerances (BASELINE.json north_star): per-layer activations and gradients <= 1e-4 relative in the fp32 validation
mode and <= 2e-2 in bf16; embeddings cosine >= 0.999; here "relative" = max |a-b| over the kept samples divided by
max |b| of the tensor's samples (per-tensor), plus a norm check."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu

CASES = {
    "aam_f40": dict(spk_num=37, feat_dim=40, pooling="mean+std", loss="AAM"),
    "softmax_f30": dict(spk_num=11, feat_dim=30, pooling="mean", loss="softmax"),
    "aamv1_f40": dict(spk_num=19, feat_dim=40, pooling="mean+std", loss="AAM-v1"),
}
TOL = {"fp32": dict(act=1e-4, grad=1e-4, emb_cos=0.999999, loss=1e-5, logits=1e-4),
       "bf16": dict(act=2e-2, grad=2e-2, emb_cos=0.999, loss=2e-2, logits=2e-2)}


def build(case, precision, impl=None):
    from model import NeuralSpeakerModel
    fx = np.load(os.path.join(util.ROOT, "tests", "golden", case + ".npz"))
    torch.manual_seed(int(fx["seed"]))
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(precision=precision, impl=impl, **CASES[case])
    for k, v in m.state_dict().items():          # same random-init weights as the reference (checksums from the fixture)
        got = np.array([float(v.double().abs().sum()), float(v.double().sum())])
        assert np.array_equal(got, fx["w/" + k]), "seeded init differs from the reference at " + k
    return m.cuda(), fx


def samp_err(got_sample, ref_sample):
    """(relative error over the strided samples, relative norm error) of two oracle-style samples."""
    g, r = np.asarray(got_sample), np.asarray(ref_sample)
    assert g[2] == r[2], "tensor sizes differ: %s vs %s" % (g[2], r[2])
    den = max(np.abs(r[3:]).max(), 1e-30)
    return float(np.abs(g[3:] - r[3:]).max() / den), float(abs(g[0] - r[0]) / max(r[0], 1e-30))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_extraction_matches_reference(case, precision):
    m, fx = build(case, precision)
    tol = TOL[precision]
    x = torch.from_numpy(fx["x"]).cuda()
    m.eval()
    e = m.predict(x).float().cpu()
    ref = torch.from_numpy(fx["embed_eval"])
    cos = torch.nn.functional.cosine_similarity(e, ref, dim=1)
    assert float(cos.min()) >= tol["emb_cos"], "embedding cosine %s" % cos
    if precision == "fp32":
        assert util.rel_err(e, ref) <= 1e-4
    # variable-length batching keeps batch-1 semantics: row 0 = utterance truncated to trunc_T, zero padded
    Tc = int(fx["trunc_T"])
    xb = x[:2].clone()
    xb[0, :, Tc:] = 0
    lengths = torch.tensor([Tc, x.shape[2]], dtype=torch.int32, device="cuda")
    eb = m.predict(xb, lengths=lengths).float().cpu()
    e1 = m.predict(x[:1, :, :Tc].contiguous()).float().cpu()
    ref1 = torch.from_numpy(fx["embed_eval_trunc"])
    assert float(torch.nn.functional.cosine_similarity(e1, ref1, dim=1).min()) >= tol["emb_cos"]
    assert util.rel_err(eb[0], e1[0]) <= 1e-5, "padded-batch row differs from its batch-1 result"
    assert util.rel_err(eb[1], e[1]) <= 1e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_train_step_matches_reference(case, precision):
    from svk.loss import CrossEntropyLoss, accuracy
    from svk.optim import SGD
    m, fx = build(case, precision)
    tol = TOL[precision]
    x = torch.from_numpy(fx["x"]).cuda()
    y = torch.from_numpy(fx["y"]).cuda()
    m.train()
    eng = m.engine
    eng.debug = {}
    crit = CrossEntropyLoss()
    opt = SGD(m.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
    logits = m(x, y)
    loss = crit(logits, y)
    acc1, acc5 = accuracy(logits, y, topk=(1, min(5, CASES[case]["spk_num"])))
    opt.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    report = []
    # ---- logits / loss / accuracy
    assert util.rel_err(logits.detach().cpu(), torch.from_numpy(fx["logits"])) <= tol["logits"]
    assert abs(float(loss) - float(fx["loss"])) <= tol["loss"] * max(1.0, abs(float(fx["loss"])))
    if precision == "fp32":
        assert [float(acc1), float(acc5)] == [float(fx["acc"][0]), float(fx["acc"][1])]
    # ---- per-layer activations (conv outputs, block outputs) from the engine workspace
    B, F, T = x.shape
    ws = eng._ws[("train", B, F, T)]
    names = {"res.conv1": ws["c0"]}
    for bi, b in enumerate(eng.blocks):
        names[b.name + ".conv1"] = ws["c1_%d" % bi]
        names[b.name + ".conv2"] = ws["c2_%d" % bi]
        names[b.name] = ws["o_%d" % bi]
        if b.convd is not None:
            names[b.name + ".downsample.0"] = ws["cd_%d" % bi]
    worst_act = 0.0
    for nm, t in names.items():
        e, en = samp_err(util.sample_of(util.nchw(t)), fx["act/" + nm])
        worst_act = max(worst_act, e)
        report.append(("act", nm, e, en))
    # ---- per-layer activation gradients captured by the engine's debug taps
    worst_dact = 0.0
    for nm, t in eng.debug.items():
        e, en = samp_err(util.sample_of(util.nchw(t)), fx["dact/" + nm])
        worst_dact = max(worst_dact, e)
        report.append(("dact", nm, e, en))
    # ---- parameter gradients
    worst_grad = 0.0
    for nm, p in m.named_parameters():
        e, en = samp_err(util.sample_of(p.grad), fx["grad/" + nm])
        worst_grad = max(worst_grad, e)
        report.append(("grad", nm, e, en))
    # ---- BatchNorm running statistics after one forward
    worst_buf = 0.0
    for nm, b in m.named_buffers():
        if "buf/" + nm in fx:
            e, _ = samp_err(util.sample_of(b), fx["buf/" + nm])
            worst_buf = max(worst_buf, e)
    bad = [r for r in report if r[2] > (tol["act"] if r[0] == "act" else tol["grad"])]
    msg = "worst act %.2e dact %.2e grad %.2e buf %.2e; offenders: %s" % (worst_act, worst_dact, worst_grad, worst_buf,
                                                                         bad[:8])
    print(case, precision, msg)
    assert not bad, msg
    assert worst_buf <= tol["act"], msg
    # ---- SGD step, then a second full step: loss must follow the reference trajectory
    opt.step()
    worst_step = max(samp_err(util.sample_of(p), fx["step/" + nm])[0] for nm, p in m.named_parameters())
    assert worst_step <= tol["act"], "parameters after SGD step off by %.2e" % worst_step
    eng.debug = None
    opt.zero_grad()
    loss2 = crit(m(x, y), y)
    loss2.backward()
    opt.step()
    assert abs(float(loss2) - float(fx["loss2"])) <= (5e-4 if precision == "fp32" else 0.15) * max(1.0, abs(float(fx["loss2"])))
    if precision == "fp32":
        worst2 = max(samp_err(util.sample_of(p), fx["step2/" + nm])[0] for nm, p in m.named_parameters())
        assert worst2 <= 1e-3, "parameters after two SGD steps off by %.2e" % worst2


def test_kat_seed0_full_size():
    """SURVEY.md §8c known answer at the real training shape (B=4, 40x200, 1211 speakers), bf16 tcgen05 path."""
    from model import NeuralSpeakerModel
    from svk.loss import CrossEntropyLoss
    fx = np.load(os.path.join(util.ROOT, "tests", "golden", "kat_seed0.npz"))
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(spk_num=1211, feat_dim=40, pooling="mean+std", loss="AAM", m=0.2, s=30).cuda()
    x, y = torch.from_numpy(fx["x"]).cuda(), torch.from_numpy(fx["y"]).cuda()
    m.train()
    out = m(x, y)
    loss = CrossEntropyLoss()(out, y)
    loss.backward()
    assert abs(float(loss) - float(fx["loss"])) <= 2e-2 * float(fx["loss"])
    assert abs(float(m.last.weight.grad.norm()) - float(fx["g_last"])) <= 2e-2 * float(fx["g_last"])
    assert abs(float(m.res.conv1.weight.grad.norm()) - float(fx["g_stem"])) <= 5e-2 * float(fx["g_stem"])
    m.eval()
    e = m.predict(x[:1]).float().cpu()
    cos = float(torch.nn.functional.cosine_similarity(e, torch.from_numpy(fx["embed"]), dim=1))
    assert cos >= 0.999, cos
