#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/scripts) on CPU fp32, and pin the
oracle (oracle/ref_model.py) against it: every quantity must match the reference EXACTLY (same torch ops in the same
order) or this script aborts.  TEST INFRASTRUCTURE ONLY.  Run from the repo root in the build container:

    python oracle/make_golden.py

The reference ships no tests / golden vectors, so these fixtures are the parity pin.  Weights are NOT
stored (25 MB per model): the drop-in model's constructor consumes the torch RNG exactly like the reference's, so a
seed reproduces them; each fixture carries per-tensor float64 checksums of the reference's weights to prove it.
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/scripts"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import ref_model as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NS = 96   # samples kept per tensor


def sample(t):
    """Deterministic strided subsample + norm + mean of a tensor (what the fixtures keep of large tensors)."""
    v = t.detach().double().reshape(-1)
    n = v.numel()
    k = min(NS, n)
    idx = (torch.arange(k, dtype=torch.int64) * (n - 1)) // max(k - 1, 1)      # integer arithmetic: host independent
    return np.concatenate([[float(v.norm()), float(v.mean()), float(n)], v[idx].numpy()]).astype(np.float64)


def checksum(sd):
    return {k: np.array([float(v.double().abs().sum()), float(v.double().sum())]) for k, v in sd.items()}


def model_case(name, seed, spk_num, feat_dim, pooling, loss, B, T, m=0.2, s=30, tries=48):
    from model import NeuralSpeakerModel          # the reference's own class
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = NeuralSpeakerModel(spk_num=spk_num, feat_dim=feat_dim, pooling=pooling, loss=loss, m=m, s=s)
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    # Pick the input (out of `tries` seeded draws) whose smallest |ReLU pre-activation| is largest: two correct fp32
    # implementations differ by ~1e-7 in those values, and a sign flip there changes the backward mask of that element
    # (and, through the tiny-batch BatchNorm sums, perturbs the whole layer) — a property of the test, not of the code.
    best = None
    for t_ in range(tries):
        g = torch.Generator().manual_seed(1000 + seed + 7919 * t_)
        xt = torch.randn(B, feat_dim, T, generator=g)
        yt = torch.randint(0, spk_num, (B,), generator=g)
        taps_ = {}
        with torch.no_grad():
            O.model_forward(sd0, xt, yt, pooling, loss, m, s, True, {}, taps_)
        margin = min(float(v.abs().min()) for k, v in taps_.items() if k.startswith("pre/"))
        if best is None or margin > best[0]:
            best = (margin, xt, yt, t_)
    margin, x, y, t_best = best
    print("  input draw %d of %d: min |ReLU pre-activation| = %.3e" % (t_best, tries, margin))
    fx = {"seed": seed, "spk_num": spk_num, "feat_dim": feat_dim, "B": B, "T": T, "m": m, "s": s,
          "x": x.numpy(), "y": y.numpy(), "relu_margin": np.array(margin)}
    for k, v in checksum(sd0).items():
        fx["w/" + k] = v

    # ---- eval-mode embeddings on the initial weights (decode.py:198 path)
    ref.eval()
    with torch.no_grad():
        e_ref = ref.predict(x)
        e_orc = O.embed(sd0, x, pooling, train=False)
    assert torch.equal(e_ref, e_orc), "oracle embed != reference predict"
    fx["embed_eval"] = e_ref.numpy()
    # batch-1 predict of the first utterance truncated to an odd length (variable-length extraction)
    Tc = max(9, T - 7)
    with torch.no_grad():
        e1 = ref.predict(x[:1, :, :Tc])
        assert torch.equal(e1, O.embed(sd0, x[:1, :, :Tc], pooling, train=False))
    fx["embed_eval_trunc"] = e1.numpy()
    fx["trunc_T"] = Tc

    # ---- training forward/backward: reference with hooks for per-layer activations
    ref.train()
    acts = {}

    def hook(nm):
        def f(mod, inp, out):
            out.retain_grad()
            acts[nm] = out
        return f
    for nm, mod in ref.named_modules():
        if isinstance(mod, torch.nn.Conv2d) or nm.startswith("res.layer") and nm.count(".") == 2 and nm.split(".")[-1].isdigit():
            mod.register_forward_hook(hook(nm))
    ref.fc1.register_forward_hook(hook("fc1"))
    out_ref = ref(x, y)
    loss_ref = F.cross_entropy(out_ref, y)
    loss_ref.backward()

    sd = {k: v.clone() for k, v in sd0.items()}
    names = O.param_names(sd)
    for n_ in names:
        sd[n_].requires_grad_(True)
    updates, taps = {}, {}
    out_orc = O.model_forward(sd, x, y, pooling, loss, m, s, True, updates, taps)
    loss_orc = O.cross_entropy(out_orc, y)
    loss_orc.backward()
    assert torch.equal(out_ref, out_orc), "oracle logits != reference"
    assert torch.equal(loss_ref, loss_orc), "oracle loss != reference"
    ref_params = dict(ref.named_parameters())
    for n_ in names:
        assert torch.equal(ref_params[n_].grad, sd[n_].grad), "oracle grad != reference for " + n_
    for k, v in updates.items():
        assert torch.equal(dict(ref.named_buffers())[k], v), "oracle running stat != reference for " + k
    for nm, t in acts.items():
        if nm in taps:
            assert torch.equal(t, taps[nm]), "oracle activation != reference at " + nm
            assert torch.equal(t.grad, taps[nm].grad), "oracle activation grad != reference at " + nm
    fx["logits"] = out_ref.detach().numpy()
    fx["loss"] = np.array(float(loss_ref))
    a1, a5 = O.accuracy(out_ref.detach(), y, (1, min(5, spk_num)))
    fx["acc"] = np.array([float(a1), float(a5)])
    for nm, t in acts.items():
        fx["act/" + nm] = sample(t)
        fx["dact/" + nm] = sample(t.grad)
    for n_ in names:
        fx["grad/" + n_] = sample(ref_params[n_].grad)
    for k, v in updates.items():
        fx["buf/" + k] = sample(v)

    # ---- one SGD step (train_resnet.py:203-205,326-328) and the loss of the NEXT forward
    opt = torch.optim.SGD(ref.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
    opt.step()
    for n_ in names:
        fx["step/" + n_] = sample(ref_params[n_])
    opt.zero_grad()
    out2 = ref(x, y)
    l2 = F.cross_entropy(out2, y)
    l2.backward()
    opt.step()
    fx["loss2"] = np.array(float(l2))
    for n_ in names:
        fx["step2/" + n_] = sample(ref_params[n_])
    # oracle train_step must reproduce both steps
    sd = {k: v.clone() for k, v in sd0.items()}
    bufs = [None] * len(names)
    la, _ = O.train_step(sd, names, x, y, pooling, loss, m, s, bufs, 0.1, 0.9, 1e-4)
    lb, _ = O.train_step(sd, names, x, y, pooling, loss, m, s, bufs, 0.1, 0.9, 1e-4)
    assert la == float(loss_ref) and lb == float(l2), "oracle train_step losses != reference"
    for n_ in names:
        assert torch.equal(sd[n_], ref_params[n_].detach()), "oracle SGD param != reference for " + n_
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fx)
    print("wrote %s: loss %.7f -> %.7f, %d arrays" % (name, float(loss_ref), float(l2), len(fx)))


def kat_case():
    """SURVEY.md §8c smoke known-answer (seed 0, C=1211, B=4, 40x200)."""
    from model import NeuralSpeakerModel
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(spk_num=1211, feat_dim=40, pooling='mean+std', loss='AAM', m=0.2, s=30)
    x = torch.randn(4, 40, 200)
    y = torch.randint(0, 1211, (4,))
    fx = {"x": x.numpy(), "y": y.numpy()}
    for k, v in checksum(m.state_dict()).items():
        fx["w/" + k] = v
    m.train()
    out = m(x, y)
    loss = F.cross_entropy(out, y)
    loss.backward()
    fx["loss"] = np.array(float(loss))
    fx["out_sum"] = np.array(float(out.sum()))
    fx["g_last"] = np.array(float(m.last.weight.grad.norm()))
    fx["g_stem"] = np.array(float(m.res.conv1.weight.grad.norm()))
    m.eval()
    with torch.no_grad():
        e = m.predict(x[:1])
    fx["embed"] = e.numpy()
    assert abs(float(loss) - 15.1372032) < 1e-5, float(loss)
    np.savez_compressed(os.path.join(OUT, "kat_seed0.npz"), **fx)
    print("wrote kat_seed0: loss %.7f |e| %.6f" % (float(loss), float(e.norm())))


def scoring_case():
    """Run the reference's scoring scripts on generated Kaldi text files; keep inputs + outputs."""
    import kaldi_io as ref_kio                      # noqa: F401  (reference's, used by the scripts)
    import cosine_score as ref_cos
    import compute_topk_mean_std as ref_topk
    import adaptive_snorm as ref_snorm
    rs = np.random.RandomState(7)
    D, n_utt, n_coh, n_trials = 64, 40, 320, 200
    emb = rs.randn(n_utt, D).astype(np.float32) + 0.3
    coh = rs.randn(n_coh, D).astype(np.float32) + 0.3
    mean = emb.mean(0)
    utts = ["utt%03d" % i for i in range(n_utt)]
    ie = rs.randint(0, n_utt, n_trials)
    it = rs.randint(0, n_utt, n_trials)
    with tempfile.TemporaryDirectory() as d:
        def write_ark(path, keys, mat):
            with open(path, "w") as f:
                for k, v in zip(keys, mat):
                    f.write(k + " [ " + " ".join(map(str, v)) + " ]\n")      # decode.py:206 format
        write_ark(d + "/emb.iv", utts, emb)
        write_ark(d + "/coh.iv", ["spk%03d" % i for i in range(n_coh)], coh)
        with open(d + "/mean.vec", "w") as f:
            f.write(" [ " + " ".join(map(str, mean)) + " ]\n")               # compute_mean.py:28 format
        with open(d + "/trials", "w") as f:
            for a, b in zip(ie, it):
                f.write("%s %s %s\n" % (utts[a], utts[b], "target" if a == b else "nontarget"))
        argv = sys.argv
        with contextlib.redirect_stdout(io.StringIO()):
            sys.argv = ["cosine_score.py", "--mean", d + "/mean.vec", "--enroll", d + "/emb.iv", "--test", d + "/emb.iv",
                        "--trials", d + "/trials", "--score-file", d + "/scores"]
            ref_cos.main()
            sys.argv = ["compute_topk_mean_std.py", "--mean", d + "/mean.vec", "--ark-file", d + "/emb.iv",
                        "--cohort-file", d + "/coh.iv", "--mean-std-file", d + "/topk"]
            ref_topk.main()
            sys.argv = ["adaptive_snorm.py", "--enroll", d + "/topk", "--test", d + "/topk", "--score-in", d + "/scores",
                        "--score-out", d + "/snorm"]
            ref_snorm.main()
        sys.argv = argv
        scores = np.array([float(l.split()[2]) for l in open(d + "/scores")])
        topk = {l.split()[0]: (float(l.split()[1]), float(l.split()[2])) for l in open(d + "/topk")}
        snorm = np.array([float(l.split()[2]) for l in open(d + "/snorm")])
        files = {k: open(d + "/" + k).read() for k in ("emb.iv", "coh.iv", "mean.vec", "trials", "scores", "topk", "snorm")}
    tk_mean = np.array([topk[u][0] for u in utts])
    tk_std = np.array([topk[u][1] for u in utts])
    # pin the oracle's restatements on the same data (text round-trip: parse like kaldi_io.read_vec_flt_ark -> float64)
    emb64 = np.array([[float(t) for t in map(str, v)] for v in emb], dtype=np.float64)
    coh64 = np.array([[float(t) for t in map(str, v)] for v in coh], dtype=np.float64)
    mean64 = np.array([float(t) for t in map(str, mean)], dtype=np.float64)
    o_scores = np.array([O.cosine_score(emb64[a], emb64[b], mean64) for a, b in zip(ie, it)])
    assert np.array_equal(o_scores.astype(np.float32), scores.astype(np.float32)), "oracle cosine != reference script"
    o_mean, o_std = O.topk_mean_std((emb64 - mean64).astype(np.float32), (coh64 - mean64).astype(np.float32), 300)
    assert np.allclose(o_mean, tk_mean, rtol=0, atol=1e-7) and np.allclose(o_std, tk_std, rtol=0, atol=1e-7)
    o_sn = np.array([O.adaptive_snorm(s_, tk_mean[a], tk_std[a], tk_mean[b], tk_std[b]) for s_, a, b in zip(scores, ie, it)])
    assert np.allclose(o_sn, snorm, rtol=0, atol=1e-12), "oracle snorm != reference script"
    np.savez_compressed(os.path.join(OUT, "scoring.npz"), emb=emb, coh=coh, mean=mean, ie=ie, it=it, scores=scores,
                        topk_mean=tk_mean, topk_std=tk_std, snorm=snorm, utts=np.array(utts),
                        **{"file/" + k: np.array(v) for k, v in files.items()})
    print("wrote scoring: %d trials, score[0]=%.6f snorm[0]=%.6f" % (n_trials, scores[0], snorm[0]))


def backend_case():
    """Run the reference's compute_mean / compute_speaker_mean / compute_eer / local/compute_min_dcf on generated files."""
    import compute_mean as ref_mean
    import compute_speaker_mean as ref_spk
    import compute_eer as ref_eer
    sys.path.insert(0, os.path.join(os.path.dirname(REF), "local"))
    import compute_min_dcf as ref_dcf
    rs = np.random.RandomState(11)
    D, n_utt, n_spk, n_trials = 48, 150, 23, 3000
    emb = (rs.randn(n_utt, D) * 2 + 0.5).astype(np.float32)
    spk_of = rs.randint(0, n_spk, n_utt)
    utts = ["utt%04d" % i for i in range(n_utt)]
    ie, it = rs.randint(0, n_utt, n_trials), rs.randint(0, n_utt, n_trials)
    target = spk_of[ie] == spk_of[it]
    # scores with ties (rounded to 2 decimals) so that the stable order of equal scores matters
    sc = np.round(rs.randn(n_trials) + 1.5 * target, 2)
    with tempfile.TemporaryDirectory() as d:
        with open(d + "/emb.iv", "w") as f:
            for k, v in zip(utts, emb):
                f.write(k + " [ " + " ".join(map(str, v)) + " ]\n")
        with open(d + "/utt2spk", "w") as f:
            for k, s_ in zip(utts, spk_of):
                f.write("%s spk%02d\n" % (k, s_))
        seen = set()
        with open(d + "/trials", "w") as ft, open(d + "/scores", "w") as fs:
            keep = []
            for t in range(n_trials):
                key = (ie[t], it[t])
                if key in seen:
                    continue                      # the reference keys trials by the id pair: keep the pairs unique
                seen.add(key)
                keep.append(t)
                ft.write("%s %s %s\n" % (utts[ie[t]], utts[it[t]], "target" if target[t] else "nontarget"))
                fs.write("%s %s %s\n" % (utts[ie[t]], utts[it[t]], repr(float(sc[t]))))
        keep = np.asarray(keep)
        argv = sys.argv
        out = {}
        with contextlib.redirect_stdout(io.StringIO()):
            sys.argv = ["compute_mean.py", d + "/emb.iv", d + "/mean.vec"]
            ref_mean.main()
            sys.argv = ["compute_speaker_mean.py", d + "/emb.iv", d + "/utt2spk", d + "/spk_mean.iv"]
            ref_spk.main()
        for name, mod, extra in (("eer", ref_eer, []), ("dcf_001", ref_dcf, ["--p-target", "0.01"]),
                                 ("dcf_05", ref_dcf, ["--p-target", "0.05", "--c-miss", "2", "--c-fa", "1.5"])):
            buf, err = io.StringIO(), io.StringIO()
            with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(err):
                sys.argv = [name] + extra + [d + "/scores", d + "/trials"]
                mod.main()
            out[name] = buf.getvalue().strip()
            out[name + "_err"] = err.getvalue().strip().splitlines()[-1]
        sys.argv = argv
        files = {k: open(d + "/" + k).read() for k in ("emb.iv", "utt2spk", "trials", "scores", "mean.vec", "spk_mean.iv")}
    # pin the oracle restatements on the same data
    scores, labels = sc[keep], target[keep].astype(np.int64)
    e, _ = O.eer(scores, labels)
    assert "{0:.2%}".format(e) == out["eer"], (e, out["eer"])
    # exact float64 equality with the reference's own functions
    fn_r, fp_r, th_r = ref_eer.ComputeErrorRates(list(scores), list(labels))
    fn_o, fp_o, th_o = O.error_rates(scores, labels)
    assert np.array_equal(np.array(fn_r), fn_o) and np.array_equal(np.array(fp_r), fp_o) and np.array_equal(np.array(th_r), th_o)
    for p_t, cm, cf, key in ((0.01, 1.0, 1.0, "dcf_001"), (0.05, 2.0, 1.5, "dcf_05")):
        m_r, t_r = ref_dcf.ComputeMinDcf(fn_r, fp_r, th_r, p_t, cm, cf)
        m_o, t_o = O.min_dcf(scores, labels, p_t, cm, cf)
        assert m_r == m_o and t_r == t_o, (m_r, m_o, t_r, t_o)
        assert "{0:.4f}".format(m_o) == out[key]
    emb64 = np.array([[float(t) for t in map(str, v)] for v in emb], dtype=np.float64)
    spk_names = sorted(set(spk_of), key=lambda s_: list(spk_of).index(s_))          # first-appearance order
    remap = {s_: i for i, s_ in enumerate(spk_names)}
    seg = np.array([remap[s_] for s_ in spk_of])
    ref_spk_mean = np.array([[float(t) for t in l.split()[2:-1]] for l in files["spk_mean.iv"].splitlines()])
    ref_mean_vec = np.array([float(t) for t in files["mean.vec"].split()[1:-1]])
    assert np.array_equal(O.speaker_means(emb64, seg, len(spk_names)), ref_spk_mean.astype(np.float32)), "speaker means"   # str(float32) round-trips
    assert np.allclose(O.global_mean(emb64), ref_mean_vec, rtol=0, atol=1e-7), "global mean"
    np.savez_compressed(os.path.join(OUT, "backend.npz"), emb=emb, seg=seg, scores=scores, labels=labels,
                        eer_out=np.array(out["eer"]), dcf001_out=np.array(out["dcf_001"]), dcf05_out=np.array(out["dcf_05"]),
                        eer_value=np.float64(e), dcf001_value=np.float64(O.min_dcf(scores, labels, 0.01, 1.0, 1.0)[0]),
                        dcf05_value=np.float64(O.min_dcf(scores, labels, 0.05, 2.0, 1.5)[0]),
                        dcf05_err=np.array(out["dcf_05_err"]), spk_mean=ref_spk_mean, mean=ref_mean_vec,
                        **{"file/" + k: np.array(v) for k, v in files.items()})
    print("wrote backend: %d trials, eer %s, minDCF %s / %s" % (len(keep), out["eer"], out["dcf_001"], out["dcf_05"]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "backend":      # only the back-end fixture (the others stay as committed)
        os.makedirs(OUT, exist_ok=True)
        backend_case()
        sys.exit(0)
    if not os.path.isdir(REF):
        sys.exit("the reference is not mounted at %s: fixtures can only be generated in the build container" % REF)
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    model_case("aam_f40", seed=1, spk_num=37, feat_dim=40, pooling="mean+std", loss="AAM", B=4, T=48)
    model_case("softmax_f30", seed=2, spk_num=11, feat_dim=30, pooling="mean", loss="softmax", B=5, T=51)
    model_case("aamv1_f40", seed=3, spk_num=19, feat_dim=40, pooling="mean+std", loss="AAM-v1", B=6, T=40)
    kat_case()
    scoring_case()
    backend_case()
