"""CPU suite, part 2: the C-ABI library (loads, exports every symbol the header declares — no compute calls), the
host-side mirror of the reference interfaces (state-dict keys, loadParameters, kaldi_io formats, datasets, batching,
sharding) and the data-parallel plumbing under gloo with world size 2."""
import contextlib
import io
import os
import re
import tempfile

import numpy as np
import pytest
import torch

import util_cpu as U


# ------------------------------------------------------------------------------------------------ C-ABI
def header_functions():
    text = open(os.path.join(U.ROOT, "include", "svk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svk_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from svk import lib
    handle = lib.load()
    assert handle.svk_version() == 100
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(handle, n), "include/svk.h declares %s but libsvk.so does not export it" % n
    for n in lib.SIGNATURES:
        assert n in names, "svk.lib binds %s which include/svk.h does not declare" % n
    assert handle.svk_last_error_string() is not None
    assert lib.launch_count() == 0          # nothing was launched: there is no GPU here


def test_bad_arguments_are_rejected_without_touching_the_device():
    from svk import lib
    handle = lib.load()
    rc = handle.svk_sgd_step(None, None, None, 10, 0.1, 0.9, 0.0, 1.0, None)
    assert rc == -1 and b"sgd_step" in handle.svk_last_error_string()
    d = lib.make_conv_desc(1, 8, 8, 24, 32, 3, 1, lib.BF16, lib.IMPL_TCGEN05)      # Cin not a multiple of 16
    assert handle.svk_conv2d_fwd(d, 16, 16, 16, None, None, None, None, 0, None, None) == -2
    d = lib.make_conv_desc(1, 8, 8, 32, 32, 5, 1, lib.BF16, lib.IMPL_TCGEN05)      # 5x5 filters do not exist on this path
    assert handle.svk_conv2d_fwd(d, 16, 16, 16, None, None, None, None, 0, None, None) == -2


def test_product_path_fails_loudly_without_cuda():
    from model import NeuralSpeakerModel
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(spk_num=5, feat_dim=40, pooling="mean+std", loss="AAM")
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 40, 16), torch.zeros(2, dtype=torch.long))
    with pytest.raises(RuntimeError):
        m.predict(torch.zeros(1, 40, 16))


def test_torch_custom_operators_are_registered_with_accurate_schemas():
    """the C-ABI is wrapped as torch custom ops (torch.library.custom_op + register_autograd)."""
    from svk import ops
    for name in ops.REGISTERED:
        assert hasattr(torch.ops.svk, name), name
    s = str(torch.ops.svk.sgd_step.default._schema)
    assert "Tensor(a0!) param" in s and "Tensor(a2!) momentum_buf" in s and "Tensor grad" in s
    assert "Tensor(a1!) flat_grads" in str(torch.ops.svk.speaker_net_train_backward.default._schema)
    assert "!" not in str(torch.ops.svk.cross_entropy.default._schema)
    # CUDA-only kernels: a CPU tensor has no implementation to fall back to
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.svk.cross_entropy(torch.zeros(2, 3), torch.zeros(2, dtype=torch.long))
    # fake (meta) kernels propagate shapes without touching the GPU
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        z = torch.empty(4, 10, device="cuda")
        loss, lse = torch.ops.svk.cross_entropy(z, torch.empty(4, dtype=torch.long, device="cuda"))
        assert loss.shape == () and lse.shape == (4,)
        y = torch.ops.svk.conv2d(torch.empty(2, 10, 50, 128, dtype=torch.bfloat16, device="cuda"),
                                 torch.empty(256, 128, 3, 3, device="cuda"), 2)
        assert y.shape == (2, 5, 25, 256) and y.dtype == torch.bfloat16


def test_network_operator_autograd_plumbing_without_a_gpu():
    """The network-level op records ONE autograd node, hands the saved state of each forward to ITS backward (token), and
    drops the state of a forward whose graph is freed or that ran under no_grad.  (CPU kernels registered for this test
    only, around a stub engine: the real kernels are CUDA-only.)"""
    import gc
    from svk import ops

    class StubEngine(object):
        def __init__(self):
            self._params = [torch.nn.Parameter(torch.randn(3)) for _ in range(4)]
            self.flat_grads = torch.zeros(12)
            self.calls = []

        def ensure_device(self):
            pass

        def forward_train(self, x, y, with_head=True, save=True):
            self.calls.append(("fwd", save))
            return x.sum(2), {"id": len(self.calls)}

        def backward_train(self, d, sv):
            self.calls.append(("bwd", sv["id"]))
    ops._speaker_net_train.register_kernel("cpu")(ops._speaker_net_train._init_fn)
    ops._speaker_net_train_backward.register_kernel("cpu")(ops._speaker_net_train_backward._init_fn)
    e = StubEngine()
    x = torch.randn(2, 3, 4)
    a = ops.speaker_net_train(e, x, None)
    b = ops.speaker_net_train(e, x, None)
    assert a.requires_grad and sorted(e._pending) == [1, 2]
    b.sum().backward()
    a.sum().backward()
    assert e.calls == [("fwd", True), ("fwd", True), ("bwd", 2), ("bwd", 1)] and not e._pending
    c = ops.speaker_net_train(e, x, None)
    assert len(e._pending) == 1
    del c
    gc.collect()
    assert not e._pending
    with torch.no_grad():
        d = ops.speaker_net_train(e, x, None)
    assert not d.requires_grad and not e._pending and e.calls[-1] == ("fwd", False)


# ------------------------------------------------------------------------------------------------ model interface
def test_state_dict_keys_and_shapes_match_the_reference_layout():
    from model import NeuralSpeakerModel
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(spk_num=5994, feat_dim=40, pooling="mean+std", loss="AAM")
    sd = m.state_dict()
    assert len(sd) == 219
    assert sum(p.numel() for p in m.parameters()) == 7513440
    assert tuple(sd["res.conv1.weight"].shape) == (32, 1, 3, 3)
    assert tuple(sd["res.layer2.0.downsample.0.weight"].shape) == (64, 32, 1, 1)
    assert tuple(sd["res.layer4.2.conv2.weight"].shape) == (256, 256, 3, 3)
    assert tuple(sd["fc1.weight"].shape) == (256, 2560) and tuple(sd["last.weight"].shape) == (5994, 256)
    assert "res.layer3.5.bn2.num_batches_tracked" in sd
    with contextlib.redirect_stdout(io.StringIO()):
        s = NeuralSpeakerModel(spk_num=7, feat_dim=30, pooling="mean", loss="softmax")
    assert tuple(s.state_dict()["fc1.weight"].shape) == (256, 4 * 256) and "bn1.running_var" in s.state_dict()
    with pytest.raises(NotImplementedError):
        NeuralSpeakerModel(spk_num=7, loss="triplet")


def test_load_parameters_strips_module_prefix_and_skips_mismatches():
    from model import NeuralSpeakerModel
    with contextlib.redirect_stdout(io.StringIO()):
        a = NeuralSpeakerModel(spk_num=11, feat_dim=40, pooling="mean+std", loss="AAM")
        b = NeuralSpeakerModel(spk_num=13, feat_dim=40, pooling="mean+std", loss="AAM")
    ckpt = {"module." + k: v.clone() for k, v in a.state_dict().items()}
    ckpt["module.extra.weight"] = torch.zeros(3)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        b.loadParameters(ckpt)
    assert torch.equal(b.state_dict()["res.layer1.0.conv1.weight"], a.state_dict()["res.layer1.0.conv1.weight"])
    assert torch.equal(b.state_dict()["fc1.bias"], a.state_dict()["fc1.bias"])
    assert b.state_dict()["last.weight"].shape[0] == 13           # size mismatch: kept, and reported like the reference
    assert "Wrong parameter length: module.last.weight" in out.getvalue()
    assert "module.extra.weight is not in the model." in out.getvalue()


def test_wide_variant_constructor():
    from model import NeuralSpeakerModel
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(spk_num=10, feat_dim=40, pooling="mean+std", loss="AAM", widths=(64, 128, 256, 512))
    assert tuple(m.state_dict()["res.layer4.0.conv1.weight"].shape) == (512, 256, 3, 3)
    assert tuple(m.state_dict()["fc1.weight"].shape) == (256, 5 * 2 * 512)


# ------------------------------------------------------------------------------------------------ kaldi_io + datasets
def write_ark_scp(tmp, mats):
    import kaldi_io
    ark, scp = os.path.join(tmp, "feats.ark"), os.path.join(tmp, "feats.scp")
    with open(ark, "wb") as f, open(scp, "w") as s:
        for key, m in mats.items():
            f.write((key + " ").encode())
            s.write("%s %s:%d\n" % (key, ark, f.tell()))
            kaldi_io.write_mat(f, m)
    return ark, scp


def test_kaldi_io_roundtrip_and_reference_text_format():
    import kaldi_io
    rs = np.random.RandomState(0)
    mats = {"utt%d" % i: rs.randn(50 + 7 * i, 40).astype(np.float32) for i in range(4)}
    with tempfile.TemporaryDirectory() as tmp:
        ark, scp = write_ark_scp(tmp, mats)
        for key, m in kaldi_io.read_mat_scp(scp):
            assert m.dtype == np.float32 and np.array_equal(m, mats[key])
        got = dict(kaldi_io.read_mat_ark(ark))
        assert sorted(got) == sorted(mats) and all(np.array_equal(got[k], mats[k]) for k in mats)
        rx = open(scp).readline().split()[1]
        assert kaldi_io.mat_shape(rx) == (50, 40)
        assert np.array_equal(kaldi_io.read_mat_rows(rx, 13, 20), mats["utt0"][13:33])
        with pytest.raises(IndexError):
            kaldi_io.read_mat_rows(rx, 40, 20)
        # float64 matrices and binary vectors
        d = os.path.join(tmp, "d.mat")
        kaldi_io.write_mat(d, mats["utt1"].astype(np.float64))
        assert kaldi_io.read_mat(d).dtype == np.float64
        v = os.path.join(tmp, "v.ark")
        with open(v, "wb") as f:
            kaldi_io.write_vec_flt(f, np.arange(5, dtype=np.float32), key="a")
            kaldi_io.write_vec_flt(f, np.arange(3, dtype=np.float64), key="b")
        vecs = dict(kaldi_io.read_vec_flt_ark(v))
        assert vecs["a"].dtype == np.float32 and vecs["b"].dtype == np.float64 and vecs["b"][2] == 2.0
        with pytest.raises(kaldi_io.UnsupportedDataType):
            kaldi_io.write_mat(d, np.zeros((2, 2), dtype=np.int32))
    # files written by the reference's own scripts (decode.py:206 / compute_mean.py:28 text formats)
    fx = np.load(os.path.join(U.GOLDEN, "scoring.npz"))
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "emb.iv")
        open(p, "w").write(str(fx["file/emb.iv"]))
        rows = list(kaldi_io.read_vec_flt_ark(p))
        assert [k for k, _ in rows] == list(fx["utts"]) and rows[0][1].dtype == np.float64
        # shortest-repr text of a float32 is within half an ulp of it
        assert np.allclose(np.array([v for _, v in rows]), fx["emb"].astype(np.float64), rtol=1e-7, atol=1e-9)
        assert np.array_equal(np.array([v for _, v in rows]).astype(np.float32), fx["emb"])
        pm = os.path.join(tmp, "mean.vec")
        open(pm, "w").write(str(fx["file/mean.vec"]))
        assert np.allclose(kaldi_io.read_vec_flt(pm), fx["mean"], atol=1e-7)


def test_datasets_follow_the_reference_rules():
    import datasets
    rs = np.random.RandomState(1)
    mats = {"s%d_u%d" % (s, u): rs.randn(60 + u, 40).astype(np.float32) for s in range(3) for u in range(1 + 3 * s)}
    with tempfile.TemporaryDirectory() as tmp:
        _, scp = write_ark_scp(tmp, mats)
        u2s = os.path.join(tmp, "utt2spkid")
        with open(u2s, "w") as f:
            for k in mats:
                f.write("%s %d\n" % (k, int(k[1])))
        with contextlib.redirect_stdout(io.StringIO()):
            ds = datasets.SequenceDataset(scp, u2s, [32])
            ds2 = datasets.SequenceDataset2(scp, u2s, 32)
            ed = datasets.EmbeddingDataset(scp, -1)
        # balancing: cap = min(500, (max_count+1)//2) = 4; speaker 0 (1 utt) repeated 4x, speaker 2 (7 utts) once each
        labels = list(ds.labels)
        assert labels.count(0) == 4 and labels.count(1) == 4 and labels.count(2) == 7
        x, lab = ds[0]
        assert x.shape == (40, 32) and x.dtype == np.float32 and int(lab) == labels[0]
        assert len(ds2) == 3 * 4 and ds2[4][0].shape == (40, 32)
        full, utt = ed[2]
        key = list(mats)[2]
        assert utt == [key] and np.array_equal(full, mats[key].T) and ed.num_frames(2) == mats[key].shape[0]
        # a crop is a contiguous window of the utterance
        k0 = list(mats)[0]
        crop = ds[0][0].T
        full0 = mats[k0]
        assert any(np.array_equal(crop, full0[p:p + 32]) for p in range(full0.shape[0] - 31))


def test_decode_batch_planning_and_length_sharding():
    import decode
    from svk.parallel import shard_by_length, shard_range
    lengths = [200, 950, 400, 210, 6000, 640, 640, 205]
    batches = decode.plan_batches(lengths, list(range(8)), max_frames=2000)
    assert sorted(i for b in batches for i in b) == list(range(8))
    for b in batches:
        assert len(b) == 1 or max(lengths[i] for i in b) * len(b) <= 2000
    assert [4] in batches                                         # the 6000-frame utterance goes alone
    shards = [shard_by_length(lengths, r, 3) for r in range(3)]
    assert sorted(i for s in shards for i in s) == list(range(8))  # every utterance exactly once, no duplicates
    frames = [sum(lengths[i] for i in s) for s in shards]
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1 and min(frames) > 0
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


# ------------------------------------------------------------------------------------------------ data-parallel plumbing
def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    from svk.parallel import GradBucketer
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    flat = torch.randn(1000)
    mine = flat.clone()
    ranges = {0: (700, 1000), 1: (300, 700), 2: (0, 300)}
    gb = GradBucketer(flat, ranges)
    for idx in (0, 1, 2):                      # reverse-layer order, as the engine's backward issues them
        gb.reduce_bucket(idx)
    gb.finish()
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    expect = sum(gathered) / world
    q.put((rank, float((flat - expect).abs().max())))
    dist.destroy_process_group()


def test_gradient_bucketer_averages_across_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(err <= 1e-6 for _, err in res), res


def test_engine_gradient_buckets_cover_the_flat_buffer_in_reverse_layer_order():
    """Bucket plan: {head, fc1} first, then layer4, then the rest — contiguous, disjoint, complete.
    Computed from parameter names only (the flat buffer itself lives on the GPU)."""
    from model import NeuralSpeakerModel
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(spk_num=20, feat_dim=40, pooling="mean+std", loss="AAM")
    names = [n for n, _ in m.named_parameters()]
    first_l4 = next(i for i, n in enumerate(names) if n.startswith("res.layer4"))
    first_tail = next(i for i, n in enumerate(names) if not n.startswith("res."))
    assert names[first_tail:] == ["fc1.weight", "fc1.bias", "last.weight"]
    assert all(n.startswith("res.layer4") for n in names[first_l4:first_tail])
    assert m.engine.bucket_of_block.count(1) == 3 and m.engine.bucket_of_block[-1] == 1


def test_scored_trial_reader_matches_the_reference_parsing():
    """scripts/compute_eer.py::read_scored_trials (shared with compute_min_dcf.py): scores in score-file order as float64,
    labels from the trial list keyed by the id pair, and the reference's exception for a scored pair that is not a trial
    (compute_eer.py:81-93).  Checked on the reference's own files from the back-end fixture."""
    import tempfile
    import compute_eer
    fx = np.load(os.path.join(U.GOLDEN, "backend.npz"))
    with tempfile.TemporaryDirectory() as d:
        sp, tp = os.path.join(d, "scores"), os.path.join(d, "trials")
        open(sp, "w").write(str(fx["file/scores"]))
        open(tp, "w").write(str(fx["file/trials"]))
        scores, labels = compute_eer.read_scored_trials(sp, tp)
        assert scores.dtype == np.float64 and np.array_equal(scores, fx["scores"])
        assert np.array_equal(labels, fx["labels"])
        with open(sp, "a") as f:
            f.write("uttX uttY 0.5\n")
        with pytest.raises(Exception, match="Missing entry for uttX and uttY"):
            compute_eer.read_scored_trials(sp, tp)
