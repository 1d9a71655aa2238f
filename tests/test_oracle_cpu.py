"""CPU suite, part 1: the oracle (oracle/ref_model.py) against the reference-generated golden fixtures.
The fixtures were written by oracle/make_golden.py from the UNMODIFIED reference with exact-equality asserts against
the oracle; here the oracle is re-run on whatever host executes the tests (different vector units may reorder fp32
sums, hence small tolerances instead of bit equality)."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

import util_cpu as U
from oracle import ref_model as O

CASES = {
    "aam_f40": dict(spk_num=37, feat_dim=40, pooling="mean+std", loss="AAM"),
    "softmax_f30": dict(spk_num=11, feat_dim=30, pooling="mean", loss="softmax"),
    "aamv1_f40": dict(spk_num=19, feat_dim=40, pooling="mean+std", loss="AAM-v1"),
}


def seeded_state(case, fx):
    """The drop-in model's constructor draws the same random stream as the reference's: seed -> identical weights."""
    from model import NeuralSpeakerModel
    torch.manual_seed(int(fx["seed"]))
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(**CASES[case])
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for k, v in sd.items():
        got = np.array([float(v.double().abs().sum()), float(v.double().sum())])
        assert np.allclose(got, fx["w/" + k], rtol=1e-9, atol=1e-9), "seeded init differs from the reference at " + k
    return sd


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_reproduces_reference_fixture(case):
    fx = np.load(os.path.join(U.GOLDEN, case + ".npz"))
    cfg = CASES[case]
    sd = seeded_state(case, fx)
    x, y = torch.from_numpy(fx["x"]), torch.from_numpy(fx["y"])
    with torch.no_grad():
        e = O.embed(sd, x, cfg["pooling"], train=False)
    assert U.rel_err(e, torch.from_numpy(fx["embed_eval"])) <= 1e-5
    Tc = int(fx["trunc_T"])
    with torch.no_grad():
        e1 = O.embed(sd, x[:1, :, :Tc], cfg["pooling"], train=False)
    assert U.rel_err(e1, torch.from_numpy(fx["embed_eval_trunc"])) <= 1e-5
    names = O.param_names(sd)
    for n_ in names:
        sd[n_].requires_grad_(True)
    updates, taps = {}, {}
    logits = O.model_forward(sd, x, y, cfg["pooling"], cfg["loss"], float(fx["m"]), float(fx["s"]), True, updates, taps)
    loss = O.cross_entropy(logits, y)
    loss.backward()
    assert U.rel_err(logits.detach(), torch.from_numpy(fx["logits"])) <= 1e-5
    assert abs(float(loss.detach()) - float(fx["loss"])) <= 1e-5
    acc = O.accuracy(logits.detach(), y, (1, min(5, cfg["spk_num"])))
    assert [float(acc[0]), float(acc[1])] == [float(fx["acc"][0]), float(fx["acc"][1])]
    # every parameter gradient: norm within 1e-4 (robust to a ReLU tie flipping on another host), samples loosely
    for n_ in names:
        got, ref = U.sample_of(sd[n_].grad), fx["grad/" + n_]
        if n_ == "fc1.bias" and cfg["loss"] != "AAM":
            continue            # exactly-zero true gradient in front of BatchNorm1d: pure rounding noise
        assert abs(got[0] - ref[0]) <= 2e-3 * max(ref[0], 1e-12), n_
    for k, v in updates.items():
        got, ref = U.sample_of(v), fx["buf/" + k]
        assert np.abs(got[3:] - ref[3:]).max() <= 1e-5 * max(np.abs(ref[3:]).max(), 1e-12), k


def test_oracle_train_steps_follow_reference_losses():
    case = "aam_f40"
    fx = np.load(os.path.join(U.GOLDEN, case + ".npz"))
    cfg = CASES[case]
    sd = seeded_state(case, fx)
    names = O.param_names(sd)
    bufs = [None] * len(names)
    x, y = torch.from_numpy(fx["x"]), torch.from_numpy(fx["y"])
    l1, _ = O.train_step(sd, names, x, y, cfg["pooling"], cfg["loss"], 0.2, 30, bufs, 0.1, 0.9, 1e-4)
    l2, _ = O.train_step(sd, names, x, y, cfg["pooling"], cfg["loss"], 0.2, 30, bufs, 0.1, 0.9, 1e-4)
    assert abs(l1 - float(fx["loss"])) <= 1e-5
    assert abs(l2 - float(fx["loss2"])) <= 5e-3      # second step: chaotic sensitivity of the tiny-batch BN network


def test_oracle_known_answer_seed0():
    """loss 15.1372032, out.sum -78.62735, |grad last.weight| 24.729879, |grad stem| 114.912643."""
    from model import NeuralSpeakerModel
    fx = np.load(os.path.join(U.GOLDEN, "kat_seed0.npz"))
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(spk_num=1211, feat_dim=40, pooling="mean+std", loss="AAM", m=0.2, s=30)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x, y = torch.from_numpy(fx["x"]), torch.from_numpy(fx["y"])
    for n_ in O.param_names(sd):
        sd[n_].requires_grad_(True)
    updates = {}
    out = O.model_forward(sd, x, y, "mean+std", "AAM", 0.2, 30, True, updates)
    loss = O.cross_entropy(out, y)
    loss.backward()
    assert abs(float(loss.detach()) - 15.1372032) <= 2e-5
    assert abs(float(out.detach().sum()) - (-78.62735)) <= 2e-3
    assert abs(float(sd["last.weight"].grad.norm()) - 24.729879) <= 1e-3
    assert abs(float(sd["res.conv1.weight"].grad.norm()) - 114.912643) <= 2e-2
    with torch.no_grad():
        sde = {k: v.detach() for k, v in sd.items()}
        sde.update(updates)        # the reference's predict runs AFTER the training forward updated the running stats
        e = O.embed(sde, x[:1], "mean+std", train=False)
    assert np.allclose(e.numpy()[0, :4], [1.413416, -0.938731, 1.018672, -2.878340], atol=2e-5)
    assert abs(float(e.norm()) - 36.309711) <= 1e-3


def test_oracle_scoring_matches_reference_scripts():
    fx = np.load(os.path.join(U.GOLDEN, "scoring.npz"))
    emb, coh, mean = fx["emb"], fx["coh"], fx["mean"]
    # the reference scripts parse the text files back to float64 (kaldi_io.read_vec_flt_ark)
    emb64 = np.array([[float(t) for t in map(str, v)] for v in emb], dtype=np.float64)
    coh64 = np.array([[float(t) for t in map(str, v)] for v in coh], dtype=np.float64)
    mean64 = np.array([float(t) for t in map(str, mean)], dtype=np.float64)
    ie, it = fx["ie"], fx["it"]
    got = np.array([O.cosine_score(emb64[a], emb64[b], mean64) for a, b in zip(ie, it)])
    assert np.abs(got - fx["scores"]).max() <= 1e-6
    assert np.abs(O.cosine_scores_np(emb64, emb64, mean64, ie, it) - fx["scores"]).max() <= 1e-6
    m, s = O.topk_mean_std((emb64 - mean64).astype(np.float32), (coh64 - mean64).astype(np.float32), 300)
    assert np.abs(m - fx["topk_mean"]).max() <= 1e-6 and np.abs(s - fx["topk_std"]).max() <= 1e-6
    sn = np.array([O.adaptive_snorm(sc, fx["topk_mean"][a], fx["topk_std"][a], fx["topk_mean"][b], fx["topk_std"][b])
                   for sc, a, b in zip(fx["scores"], ie, it)])
    assert np.abs(sn - fx["snorm"]).max() <= 1e-9


def test_oracle_backend_matches_reference_scripts():
    """EER / minDCF / speaker means restated in numpy against the outputs of the reference's compute_eer.py,
    local/compute_min_dcf.py, compute_speaker_mean.py and compute_mean.py (fixture: oracle/make_golden.py backend)."""
    fx = np.load(os.path.join(U.GOLDEN, "backend.npz"))
    scores, labels = fx["scores"], fx["labels"]
    e, idx = O.eer(scores, labels)
    assert "{0:.2%}".format(e) == str(fx["eer_out"]) and e == float(fx["eer_value"])
    m1, _ = O.min_dcf(scores, labels, 0.01, 1.0, 1.0)
    m2, _ = O.min_dcf(scores, labels, 0.05, 2.0, 1.5)
    assert "{0:.4f}".format(m1) == str(fx["dcf001_out"]) and m1 == float(fx["dcf001_value"])
    assert "{0:.4f}".format(m2) == str(fx["dcf05_out"]) and m2 == float(fx["dcf05_value"])
    # ties: equal scores keep their file order (stable sort) — a permutation inside a tie group changes the curve
    fn, fp, thr = O.error_rates([0.5, 0.5, 0.1, 0.9], [1, 0, 0, 1])
    assert list(thr) == [0.1, 0.5, 0.5, 0.9] and list(fn) == [0.0, 0.5, 0.5, 1.0] and list(fp) == [0.5, 0.5, 0.0, 0.0]
    emb64 = np.array([[float(t) for t in map(str, v)] for v in fx["emb"]], dtype=np.float64)
    n_seg = int(fx["seg"].max()) + 1
    assert np.array_equal(O.speaker_means(emb64, fx["seg"], n_seg), fx["spk_mean"].astype(np.float32))
    assert np.abs(O.global_mean(emb64) - fx["mean"]).max() <= 1e-7


def test_storage_rounding_emulation_matches_measured_bf16_drift():
    """rounding every stored tensor to bf16 (fp32 math) makes layer4 drift by several percent on a
    random-init network — the reason the bf16 end-to-end test is layer-local."""
    fx = np.load(os.path.join(U.GOLDEN, "aam_f40.npz"))
    sd = seeded_state("aam_f40", fx)
    x, y = torch.from_numpy(fx["x"]), torch.from_numpy(fx["y"])
    t32, t16 = {}, {}
    with torch.no_grad():
        O.model_forward(sd, x, y, "mean+std", "AAM", 0.2, 30, True, {}, t32)
        with O.storage_rounding(torch.bfloat16):
            O.model_forward(sd, x, y, "mean+std", "AAM", 0.2, 30, True, {}, t16)
    first = float((t16["res.layer1.0"] - t32["res.layer1.0"]).norm() / t32["res.layer1.0"].norm())
    last = float((t16["res.layer4.2"] - t32["res.layer4.2"]).norm() / t32["res.layer4.2"].norm())
    assert 1e-3 < first < 2e-2 and 2e-2 < last < 0.3 and last > 3 * first
