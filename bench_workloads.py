"""Secondary workloads of bench.py (BASELINE.json configs 2 and 5) and the reference-side timings of every workload.

    cfg2  embedding extraction of variable-length utterances + cosine scoring of a 37,720-trial list   (1 GPU)
    cfg5  adaptive s-norm scoring: top-300 cohort statistics of 100,000 embeddings against a 50,000-row cohort, 1,000,000
          cosine trials, s-norm apply — embedding rows and trial blocks sharded over the ranks, one all-gather of the
          (mean, std) pairs ("no communication beyond the final gather")

Reference side: the UNMODIFIED reference scripts from baseline/_ref/scripts (a verbatim copy of /root/reference/scripts made
by __graft_entry__.build(); kind "reference"), else the oracle port (kind "port").  Everything here is measurement code:
the oracle / the reference are only ever the thing TIMED ON THE HOST as the baseline, never part of the product path.
"""
import contextlib
import importlib.util
import io
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
REF_SCRIPTS = os.path.join(ROOT, "baseline", "_ref", "scripts")
SPK, FEAT = 5994, 40
EXTRACT_FLOP_PER_FRAME = 22632960.0          # SURVEY.md §8d (conv MACs x 2 per frame) ; + 1.31 MFLOP per utterance (fc1)
EXTRACT_FLOP_PER_UTT = 1310720.0


def reference_module(name):
    """Import baseline/_ref/scripts/<name>.py (the unmodified reference) under a private module name, or None."""
    path = os.path.join(REF_SCRIPTS, name + ".py")
    if not os.path.exists(path):
        return None
    if REF_SCRIPTS not in sys.path:
        sys.path.append(REF_SCRIPTS)          # LAST: only names our drop-in scripts do not define resolve there (densenet)
    spec = importlib.util.spec_from_file_location("reference_unmodified_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------ cfg1 / cfg3: training on the CPU
def cpu_train(steps, warmup, batch=32, frames=200, widths=None):
    """Loop body of the reference's train_resnet.py:307-328 (forward, CrossEntropyLoss, accuracy, loss.item(), zero_grad,
    backward, SGD step) on the host cores, every core.  Unmodified reference model when baseline/_ref is present and the
    standard widths are asked for; the oracle port otherwise (the reference has no width knob, model.py:208-218)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, FEAT, frames, generator=g)
    y = torch.randint(0, SPK, (batch,), generator=g)
    ref = reference_module("model") if widths is None else None
    times = []
    if ref is not None:
        acc = reference_module("accuracy")
        torch.manual_seed(1234)
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref.NeuralSpeakerModel(spk_num=SPK, feat_dim=FEAT, pooling="mean+std", loss="AAM", m=0.2, s=30)
        crit = torch.nn.CrossEntropyLoss()
        opt = torch.optim.SGD(model.parameters(), 0.1, momentum=0.9, weight_decay=5e-4)
        model.train()
        # accuracy.py:12 does `correct[:k].view(-1)` on a non-contiguous slice, which raises on current torch for k = 5: the
        # unmodified function cannot run here, so the timed loop omits that (sub-millisecond) call and says so
        try:
            acc.accuracy(torch.zeros(2, SPK), torch.zeros(2, dtype=torch.long), topk=(1, 5))
            with_acc = True
        except RuntimeError:
            with_acc = False
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            output = model(x, y)                                   # train_resnet.py:316
            loss = crit(output, y)                                 # :317
            if with_acc:
                acc.accuracy(output, y, topk=(1, 5))               # :320
            loss.item()                                            # :321
            opt.zero_grad()                                        # :326
            loss.backward()                                        # :327
            opt.step()                                             # :328
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "reference"
        what = "unmodified reference scripts/model.py + loop body of train_resnet.py:307-328" + (
            "" if with_acc else " (without accuracy(): accuracy.py:12 raises on torch >= 2)")
    else:
        from oracle import ref_model as O
        kw = {} if widths is None else {"widths": widths}
        sd = O.init_state(SPK, FEAT, "mean+std", "AAM", seed=1234, **kw)
        names = O.param_names(sd)
        bufs = [None] * len(names)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.train_step(sd, names, x, y, "mean+std", "AAM", 0.2, 30, bufs, 0.1, 0.9, 5e-4)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "port"
        what = "oracle port of train_resnet.py:307-328"
    ms = 1e3 * statistics.median(times)
    return {"value": batch / (ms / 1e3), "unit": "chunks/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": "%d steps of %d chunks x %d frames (median) after %d warm-up; %s; fp32 oneDNN" % (steps, batch, frames, warmup, what),
            "ms_per_step": ms}


# ------------------------------------------------------------------------------------------------ cfg2: extraction
def cfg2_lengths(n=4708):
    import numpy as np
    rs = np.random.RandomState(1234)
    return np.clip(np.round(np.exp(rs.normal(np.log(650.0), 0.55, n))), 200, 6000).astype(int)


class MemoryUtterances(object):
    """The cfg2 utterance set held in host memory with the interface of datasets.EmbeddingDataset."""
    seq_len = -1

    def __init__(self, lengths, seed=1234):
        import numpy as np
        rs = np.random.RandomState(seed)
        self.utts = ["utt%05d" % i for i in range(len(lengths))]
        self.mats = [rs.randn(FEAT, int(t)).astype(np.float32) for t in lengths]

    def __len__(self):
        return len(self.mats)

    def num_frames(self, i):
        return self.mats[i].shape[1]

    def __getitem__(self, i):
        return self.mats[i], [self.utts[i]]


def cpu_extract(ds, n_utts):
    """decode_cpu.py:185-208 on the host cores: eval-mode model.predict per utterance (batch 1) + the text formatting of
    the embedding line, for the first `n_utts` utterances of the set."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    ref = reference_module("model")
    n = min(n_utts, len(ds))
    if ref is not None:
        torch.manual_seed(1234)
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref.NeuralSpeakerModel(spk_num=SPK, feat_dim=FEAT, pooling="mean+std", loss="AAM")
        model.eval()

        def predict(x):
            return model.predict(x)
        kind = "reference"
    else:
        from oracle import ref_model as O
        sd = O.init_state(SPK, FEAT, "mean+std", "AAM", seed=1234)

        def predict(x):
            return O.embed(sd, x, "mean+std", train=False)
        kind = "port"
    frames = 0
    with torch.no_grad():
        predict(torch.from_numpy(ds[0][0])[None])                                     # warm-up
        t0 = time.perf_counter()
        for i in range(n):
            mat, utt = ds[i]
            pred = predict(torch.from_numpy(mat)[None]).cpu().data.numpy()             # decode_cpu.py:198-199
            out = pred[0, :].flatten()
            line = utt[0] + " [ " + " ".join(map(str, out)) + " ]\n"                  # :206
            frames += mat.shape[1]
        dt = time.perf_counter() - t0
    assert line
    return {"value": n / dt, "unit": "utts/s", "frames_per_sec": frames / dt, "cores": torch.get_num_threads(), "kind": kind,
            "sample": "first %d of the cfg2 utterances (%d frames), batch 1 per utterance, decode_cpu.py:185-208 semantics" % (n, frames)}


def write_scoring_files(tmp, emb, ie, it, cohort=None):
    """Kaldi text files as the recipe produces them (decode.py:206, compute_mean.py:28, trials lists)."""
    import numpy as np
    keys = ["utt%06d" % i for i in range(len(emb))]
    p = {k: os.path.join(tmp, k) for k in ("emb", "mean", "trials", "cohort")}
    with open(p["emb"], "w") as f:
        for k, v in zip(keys, emb):
            f.write(k + " [ " + " ".join(map(str, v)) + " ]\n")
    with open(p["mean"], "w") as f:
        f.write(" [ " + " ".join(map(str, np.asarray(emb, dtype=np.float64).mean(0))) + " ]\n")
    with open(p["trials"], "w") as f:
        for j, (a, b) in enumerate(zip(ie, it)):
            f.write("%s %s %s\n" % (keys[a], keys[b], "target" if j % 2 else "nontarget"))
    if cohort is not None:
        with open(p["cohort"], "w") as f:
            for j, v in enumerate(cohort):
                f.write("spk%05d [ " % j + " ".join(map(str, v)) + " ]\n")
    return p


def run_script(scripts_dir, name, args):
    """Run a scoring script as the recipes do (its own directory first on sys.path) and return its wall time."""
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, os.path.join(scripts_dir, name)] + args, capture_output=True, text=True, cwd=scripts_dir)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("%s failed: %s" % (name, (r.stdout + r.stderr)[-500:]))
    return dt


def cpu_cosine_score(paths, n_trials):
    """The reference's cosine_score.py as a script on files (parse + per-trial loop :60-65 + write)."""
    scripts = REF_SCRIPTS if os.path.exists(os.path.join(REF_SCRIPTS, "cosine_score.py")) else None
    if scripts is None:
        return None
    out = paths["emb"] + ".ref_scores"
    dt = run_script(scripts, "cosine_score.py", ["--mean", paths["mean"], "--enroll", paths["emb"], "--test", paths["emb"],
                                                  "--trials", paths["trials"], "--score-file", out])
    return {"value": n_trials / dt, "unit": "trials/s", "cores": 1, "kind": "reference", "seconds": dt, "score_file": out,
            "sample": "reference cosine_score.py on files, %d trials (text parse + per-trial loop + write)" % n_trials}


# ------------------------------------------------------------------------------------------------ cfg5: s-norm scoring at scale
def cpu_snorm(n_emb=1000, n_trials=10000, n_cohort=50000, dim=256):
    """The reference's scoring path on a 1 % subsample of cfg5: compute_topk_mean_std.py:10-23 for `n_emb`
    embeddings against the FULL cohort, the per-trial cosine loop of cosine_score.py:60-65 and the s-norm arithmetic of
    adaptive_snorm.py:28-38 for `n_trials` trials — in memory (no text files), extrapolated linearly to the full job."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    torch.set_num_threads(os.cpu_count() or 1)
    rs = np.random.RandomState(1234)
    emb = rs.randn(n_emb, dim).astype(np.float32)
    cohort = torch.from_numpy(rs.randn(n_cohort, dim).astype(np.float32))
    ie, it = rs.randint(0, n_emb, n_trials), rs.randint(0, n_emb, n_trials)
    ref = reference_module("compute_topk_mean_std")
    utt2vec = {i: torch.from_numpy(emb[i]) for i in range(n_emb)}
    t0 = time.perf_counter()
    if ref is not None:
        with contextlib.redirect_stdout(io.StringIO()):
            mean, std = ref.compute_topk_mean_std(utt2vec, cohort)
        kind = "reference"
    else:
        from oracle import ref_model as O
        m_, s_ = O.topk_mean_std(emb, cohort.numpy(), 300)
        mean, std = dict(enumerate(m_)), dict(enumerate(s_))
        kind = "port"
    t_stats = time.perf_counter() - t0
    t0 = time.perf_counter()
    scores = []
    for a, b in zip(ie, it):                                           # cosine_score.py:60-65
        spkr_vec = torch.FloatTensor(emb[a])
        utt_vec = torch.FloatTensor(emb[b])
        scores.append(F.cosine_similarity(spkr_vec, utt_vec, dim=0).data.numpy())
    t_cos = time.perf_counter() - t0
    t0 = time.perf_counter()
    out = []
    for a, b, s in zip(ie, it, scores):                                # adaptive_snorm.py:28-38
        s = float(s)
        out.append((s - float(mean[a])) / max(float(std[a]), 1e-8) / 2 + (s - float(mean[b])) / max(float(std[b]), 1e-8) / 2)
    t_sn = time.perf_counter() - t0
    full = t_stats * (100000.0 / n_emb) + (t_cos + t_sn) * (1000000.0 / n_trials)
    return {"value": 1000000.0 / full, "unit": "trials/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": "1 %% subsample (%d embeddings x full %d-row cohort: %.2f s; %d trials cosine %.2f s + s-norm %.2f s), "
                      "EXTRAPOLATED linearly to 100,000 embeddings + 1,000,000 trials (%.0f s)" % (
                          n_emb, n_cohort, t_stats, n_trials, t_cos, t_sn, full),
            "stats_rows_per_sec": n_emb / t_stats, "cosine_trials_per_sec": n_trials / t_cos}
