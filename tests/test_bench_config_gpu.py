"""Parity AT THE BENCHMARKED CONFIGURATION (BASELINE.json config 3: 256 chunks of 40 x 200 per GPU, 5,994 speakers).

The kernel tests elsewhere run 2-4 images; at 256 the tile pickers, split-K plans, persistent schedules and CTA-pair
paths take the branches the benchmark takes, and a weight gradient accumulates 2 M products.  Everything here compares
FULL tensors (no sampling) against torch's own fp32 / fp64 convolutions on the same device with TF32 disabled:

  * the ten convolution shapes of ResNet-34 at N = 256: forward (+ BatchNorm statistics), fused data gradient
    (svk_conv2d_dgrad_bn / svk_downsample_dgrad_bn with mask, residual and BatchNorm-backward sums) and weight gradient,
    each through the C-ABI, in whatever kernel variant the library selects by default;
  * one full training step of the benchmark model in bf16 against the SAME step in the fp32 validation mode (which is
    pinned to the reference at <= 4e-5, tests/test_model_gpu.py): loss, every stored activation, every parameter gradient;
  * extraction of 200 / 1,037 / 6,000-frame utterances against the CPU oracle (BASELINE.json config 2 lengths).
Tolerances are BASELINE.json's: 2e-2 relative for bf16 tensors (max |a - b| / max |b|), 1e-4 for fp32 accumulators,
embedding cosine >= 0.999.
"""
import contextlib
import io
import json
import os

import pytest
import torch
import torch.nn.functional as F

import util
from util import call, lib

pytestmark = pytest.mark.gpu

N = 256


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    torch.cuda.empty_cache()


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)     # bf16-exact values


def _nchw(t_nhwc, dtype=torch.float32):
    return t_nhwc.to(dtype).permute(0, 3, 1, 2)


def _rel(a, b):
    den = float(b.abs().max())
    return float((a.double() - b.double()).abs().max()) / max(den, 1e-30)


def _case(shape, seed):
    H, W, ci, co, r, stride = shape
    x = _rand((N, H, W, ci), seed)
    w = _rand((co, ci, r, r), seed + 1, (2.0 / (ci * r * r)) ** 0.5).float()
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    dy = _rand((N, Ho, Wo, co), seed + 2)
    d = lib.make_conv_desc(N, H, W, ci, co, r, stride, lib.BF16, lib.IMPL_TCGEN05)
    wf = torch.empty(r * r * co * ci, dtype=torch.bfloat16, device="cuda")
    wd = torch.empty_like(wf)
    call.svk_pack_conv_weight(w.data_ptr(), wf.data_ptr(), wd.data_ptr(), co, ci, r, lib.BF16, util.st())
    return x, w, dy, d, wf, wd


@pytest.mark.parametrize("shape", util.RESNET_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_forward_and_weight_gradient_at_batch_256(shape):
    H, W, ci, co, r, stride = shape
    x, w, dy, d, wf, wd = _case(shape, 100)
    y = torch.full((N, d.Ho, d.Wo, co), float("nan"), dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
    call.svk_conv2d_fwd(d, x.data_ptr(), wf.data_ptr(), y.data_ptr(), stats.data_ptr(), 0, 0, 0, 0, 0, util.st())
    ref = F.conv2d(_nchw(x), w, None, stride, r // 2)
    got = _nchw(y)
    assert not torch.isnan(got).any()
    assert _rel(got, ref) <= 2e-2
    # BatchNorm statistics = fp64 sums of the values AS STORED
    gd = got.double()
    s_ref = torch.cat([gd.sum((0, 2, 3)), (gd * gd).sum((0, 2, 3))])
    assert _rel(stats, s_ref) <= 1e-6
    del ref, got, gd
    # weight gradient: 256 * Ho * Wo products per entry, fp32 accumulation on the tensor cores, deterministic split-K
    need = lib.load().svk_conv2d_wgrad_workspace_bytes(d)
    ws = torch.full(((need + 3) // 4,), float("nan"), dtype=torch.float32, device="cuda")
    dw = torch.full((co, ci, r, r), float("nan"), dtype=torch.float32, device="cuda")
    call.svk_conv2d_wgrad(d, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel() * 4, util.st())
    ref_dw = torch.nn.grad.conv2d_weight(_nchw(x, torch.float64), (co, ci, r, r), _nchw(dy, torch.float64), stride, r // 2)
    assert not torch.isnan(dw).any()
    assert _rel(dw, ref_dw) <= 1e-4
    dw2 = torch.empty_like(dw)
    call.svk_conv2d_wgrad(d, x.data_ptr(), dy.data_ptr(), dw2.data_ptr(), ws.data_ptr(), ws.numel() * 4, util.st())
    assert torch.equal(dw, dw2), "split-K reduction is not deterministic"


STRIDE1 = [s for s in util.RESNET_SHAPES if s[4] == 3 and s[5] == 1]


@pytest.mark.parametrize("shape", STRIDE1, ids=lambda s: "x".join(map(str, s)))
def test_fused_data_gradient_at_batch_256(shape):
    """svk_conv2d_dgrad_bn = (dgrad + res) * (mask > 0) with the BatchNorm-backward sums of the stored values."""
    H, W, ci, co, r, stride = shape
    x, w, dy, d, wf, wd = _case(shape, 200)
    mask, c, res = _rand((N, H, W, ci), 7), _rand((N, H, W, ci), 8), _rand((N, H, W, ci), 9)
    g = torch.Generator(device="cuda").manual_seed(3)
    mean = torch.randn(ci, generator=g, device="cuda") * 0.3
    rstd = torch.rand(ci, generator=g, device="cuda") + 0.5
    ref0 = torch.nn.grad.conv2d_input((N, ci, H, W), w, _nchw(dy), stride, r // 2)
    for with_c in (True, False):
        dx = torch.full((N, H, W, ci), float("nan"), dtype=torch.bfloat16, device="cuda")
        sums = torch.zeros(2, ci, dtype=torch.float64, device="cuda")
        bn = lib.BnBwdFuse(mask.data_ptr(), c.data_ptr() if with_c else None, mean.data_ptr() if with_c else None,
                           rstd.data_ptr() if with_c else None, sums.data_ptr() if with_c else None)
        call.svk_conv2d_dgrad_bn(d, dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), res.data_ptr() if with_c else 0, 0, 0, bn,
                                 util.st())
        got = _nchw(dx)
        ref = (ref0 + (_nchw(res) if with_c else 0)) * (_nchw(mask) > 0)
        assert not torch.isnan(got).any()
        assert _rel(got, ref) <= 2e-2
        assert float(got[_nchw(mask) <= 0].abs().max()) == 0.0
        if with_c:
            gd = got.double()
            xhat = (_nchw(c, torch.float64) - mean.double().view(1, -1, 1, 1)) * rstd.double().view(1, -1, 1, 1)
            s_ref = torch.stack([gd.sum((0, 2, 3)), (gd * xhat).sum((0, 2, 3))])
            assert _rel(sums, s_ref) <= 1e-4
    # plain data gradient (no fusion): the stem-side and validation call
    dx = torch.full((N, H, W, ci), float("nan"), dtype=torch.bfloat16, device="cuda")
    call.svk_conv2d_dgrad(d, dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), 0, 0, 0, util.st())
    assert _rel(_nchw(dx), ref0) <= 2e-2


@pytest.mark.parametrize("hw", [(40, 200, 32), (20, 100, 64), (10, 50, 128)], ids=lambda s: "x".join(map(str, s)))
def test_downsample_block_data_gradient_at_batch_256(hw):
    """svk_downsample_dgrad_bn = dgrad(3x3/s2) + dgrad(1x1/s2), masked, + sums (first block of stages 2-4)."""
    H, W, ci = hw
    co = 2 * ci
    x, w3, dy3, d3, _, wd3 = _case((H, W, ci, co, 3, 2), 300)
    _, w1, dy1, d1, _, wd1 = _case((H, W, ci, co, 1, 2), 400)
    mask, c = _rand((N, H, W, ci), 17), _rand((N, H, W, ci), 18)
    g = torch.Generator(device="cuda").manual_seed(5)
    mean = torch.randn(ci, generator=g, device="cuda") * 0.3
    rstd = torch.rand(ci, generator=g, device="cuda") + 0.5
    dx = torch.full((N, H, W, ci), float("nan"), dtype=torch.bfloat16, device="cuda")
    sums = torch.zeros(2, ci, dtype=torch.float64, device="cuda")
    bn = lib.BnBwdFuse(mask.data_ptr(), c.data_ptr(), mean.data_ptr(), rstd.data_ptr(), sums.data_ptr())
    call.svk_downsample_dgrad_bn(d3, dy3.data_ptr(), wd3.data_ptr(), d1, dy1.data_ptr(), wd1.data_ptr(), dx.data_ptr(), bn,
                                 util.st())
    ref = (torch.nn.grad.conv2d_input((N, ci, H, W), w3, _nchw(dy3), 2, 1) +
           torch.nn.grad.conv2d_input((N, ci, H, W), w1, _nchw(dy1), 2, 0)) * (_nchw(mask) > 0)
    got = _nchw(dx)
    assert not torch.isnan(got).any()
    assert _rel(got, ref) <= 2e-2
    gd = got.double()
    xhat = (_nchw(c, torch.float64) - mean.double().view(1, -1, 1, 1)) * rstd.double().view(1, -1, 1, 1)
    s_ref = torch.stack([gd.sum((0, 2, 3)), (gd * xhat).sum((0, 2, 3))])
    assert _rel(sums, s_ref) <= 1e-4


# ------------------------------------------------------------------------------------------------ full benchmark step
def _bench_model(precision):
    from model import NeuralSpeakerModel
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        return NeuralSpeakerModel(spk_num=5994, feat_dim=40, pooling="mean+std", loss="AAM", m=0.2, s=30,
                                  precision=precision).cuda()


def _step(m, x, y):
    from svk.loss import CrossEntropyLoss
    m.train()
    logits = m(x, y)
    loss = CrossEntropyLoss()(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    eng = m.engine
    ws = eng._train_ws[tuple(x.shape)][0]
    acts = {"res.conv1": ws["c0"]}
    for bi, b in enumerate(eng.blocks):
        acts[b.name + ".conv1"] = ws["c1_%d" % bi]
        acts[b.name + ".conv2"] = ws["c2_%d" % bi]
        acts[b.name] = ws["o_%d" % bi]
    acts["embedding"] = ws["emb"]
    grads = {n: p.grad.detach().float().clone() for n, p in m.named_parameters()}
    return float(loss), logits.detach().float(), {k: v.float() for k, v in acts.items()}, grads


def test_benchmark_step_bf16_against_fp32_mode():
    """One cfg3 training step (B = 256, C = 5,994): the bf16 product path against the fp32 validation mode (pinned to the
    reference at <= 4e-5) on the same weights and batch, every stored activation and every parameter gradient.

    What bounds it.  At this batch BatchNorm statistics are stable (>= 32,000 samples per channel), so the difference is
    bf16 STORAGE noise only — and a random-init ResNet-34 in training mode amplifies a perturbation ~1.1x per convolution
    (2^-9 per stored tensor grows to ~10 % at layer4.2, and the gradients of the early layers inherit it).  That is a
    property of the network, not of the kernels: the CPU oracle with its stored tensors rounded to bf16 and ALL arithmetic
    in fp32 (oracle/make_bf16_drift.py -> tests/golden/bf16_storage_drift.json) shows the same growth to 2-3 digits.  The
    2e-2 per-layer bound of BASELINE.json is therefore enforced layer-locally (tests/test_model_gpu.py) and at batch 256 on
    every kernel (above); END TO END the CUDA path is held to the emulated bf16-storage drift: loss <= 2e-3 relative, every
    activation's RMS-relative error <= 1.25 x the emulation's (+1e-3), every parameter-gradient cosine >= the emulation's
    - 0.1.  The per-layer table goes to gpurun_out/cfg3_bf16_vs_fp32.json (copied to profiles/)."""
    with open(os.path.join(util.ROOT, "tests", "golden", "bf16_storage_drift.json")) as f:
        emul = json.load(f)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(N, 40, 200, generator=g).cuda()
    y = torch.randint(0, 5994, (N,), generator=g).cuda()
    m32 = _bench_model("fp32")
    loss32, logits32, acts32, grads32 = _step(m32, x, y)
    del m32
    torch.cuda.empty_cache()
    m16 = _bench_model("bf16")
    loss16, logits16, acts16, grads16 = _step(m16, x, y)
    rows = {"loss_fp32": loss32, "loss_bf16": loss16, "activations": {}, "gradients": {}}
    bad = []
    for k in acts32:
        a, b = acts16[k].double(), acts32[k].double()
        rms = float(((a - b) ** 2).mean().sqrt() / (b ** 2).mean().sqrt())
        rows["activations"][k] = {"rms_rel": rms, "max_rel": _rel(a, b), "emulated_bf16_storage_rms_rel": emul["activations"][k]}
        if rms > 1.25 * emul["activations"][k] + 1e-3:
            bad.append("activation %s: %.4g vs emulated %.4g" % (k, rms, emul["activations"][k]))
    for k in grads32:
        a, b = grads16[k].double().reshape(-1), grads32[k].double().reshape(-1)
        cos = float(F.cosine_similarity(a, b, dim=0))
        rows["gradients"][k] = {"cosine": cos, "emulated_bf16_storage_cosine": emul["gradients"][k], "numel": a.numel(),
                                "norm_rel": float((a.norm() - b.norm()).abs() / b.norm().clamp_min(1e-30))}
        if a.numel() >= 1000 and cos < emul["gradients"][k] - 0.1:
            bad.append("gradient %s: cosine %.4f vs emulated %.4f" % (k, cos, emul["gradients"][k]))
    rows["logits_max_rel"] = _rel(logits16, logits32)
    os.makedirs(os.path.join(util.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(util.ROOT, "gpurun_out", "cfg3_bf16_vs_fp32.json"), "w") as f:
        json.dump(rows, f, indent=1)
    assert loss16 == loss16 and abs(loss16 - loss32) <= 2e-3 * abs(loss32)
    assert rows["activations"]["res.conv1"]["max_rel"] <= 2e-2 and rows["activations"]["res.layer1.0.conv1"]["max_rel"] <= 2e-2
    assert not bad, "bf16 path drifts more than bf16 storage explains:\n" + "\n".join(bad[:10])


# ------------------------------------------------------------------------------------------------ long utterances
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_extraction_of_long_utterances(precision):
    """decode.py:198 semantics on cfg2-length utterances (200, 1,037 and 6,000 frames; the odd length exercises the
    ceil() of every stride-2 stage), alone and as one zero-padded batch with per-row lengths, against the CPU oracle."""
    from model import NeuralSpeakerModel
    from oracle import ref_model as O
    torch.manual_seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(spk_num=16, feat_dim=40, pooling="mean+std", loss="AAM", precision=precision)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():           # non-trivial eval-mode BatchNorm: random running statistics and affine parameters
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) * 0.5 + 0.75)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
    sd = {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}
    m.cuda().eval()
    lens = [200, 1037, 6000]
    xs = [torch.randn(1, 40, t, generator=g) for t in lens]
    refs = [O.embed(sd, x, "mean+std", train=False) for x in xs]
    min_cos = 0.999999 if precision == "fp32" else 0.999
    singles = []
    for x, ref in zip(xs, refs):
        e = m.predict(x.cuda()).float().cpu()
        singles.append(e)
        assert float(F.cosine_similarity(e, ref, dim=1)) >= min_cos
        if precision == "fp32":
            assert util.rel_err(e, ref) <= 1e-4
    xb = torch.zeros(3, 40, 6000)
    for i, x in enumerate(xs):
        xb[i, :, :x.shape[2]] = x[0]
    eb = m.predict(xb.cuda(), lengths=torch.tensor(lens, dtype=torch.int32, device="cuda")).float().cpu()
    for i in range(3):
        assert util.rel_err(eb[i], singles[i][0]) <= 1e-5, "padded-batch row %d differs from its batch-1 result" % i
