"""GPU parity of the individual kernels, called through the C-ABI (ctypes), against the CPU oracle / torch fp32
restatements on the same seeded inputs.  Tolerances: fp32 kernels <= 1e-4 relative, bf16 kernels <= 2e-2 relative
(per-tensor max error / max |reference|), scores <= 1e-3 absolute (BASELINE.json north_star)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import util
from oracle import ref_model as O
from util import call, lib

pytestmark = pytest.mark.gpu

ALL_SHAPES = util.RESNET_SHAPES + util.ODD_SHAPES + util.WIDE_SHAPES


# ------------------------------------------------------------------------------------------------ convolutions
@pytest.mark.parametrize("shape", ALL_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_fp32_validation_path(shape):
    H, W, ci, co, r, stride = shape
    x, w, dy = util.make_case(shape, 2, 7, quantize=False)
    y, stats = util.conv_fwd(x, w, stride, lib.F32, lib.IMPL_SIMT, stats=True)
    ref = util.ref_conv(x, w, stride)
    assert util.rel_err(y, ref) <= 1e-4
    s_ref = torch.cat([ref.double().sum((0, 2, 3)), (ref.double() ** 2).sum((0, 2, 3))])
    assert util.rel_err(stats, s_ref) <= 1e-5
    assert util.rel_err(util.conv_dgrad(dy, w, H, W, stride, lib.F32, lib.IMPL_SIMT), util.ref_dgrad(dy, w, H, W, stride)) <= 1e-4
    assert util.rel_err(util.conv_wgrad(x, dy, r, stride, lib.F32, lib.IMPL_SIMT), util.ref_wgrad(x, dy, r, stride)) <= 1e-4


@pytest.mark.parametrize("shape", ALL_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_tcgen05_path(shape):
    H, W, ci, co, r, stride = shape
    N = 3
    x, w, dy = util.make_case(shape, N, 11)            # bf16-exact inputs: the only error is accumulation + output rounding
    ref = util.ref_conv(x, w, stride)
    y, stats = util.conv_fwd(x, w, stride, lib.BF16, lib.IMPL_TCGEN05, stats=True)
    assert not torch.isnan(y).any()
    assert util.rel_err(y, ref) <= 2e-2
    s_ref = torch.cat([y.double().sum((0, 2, 3)), (y.double() ** 2).sum((0, 2, 3))])      # statistics of the STORED values
    assert util.rel_err(stats, s_ref) <= 1e-4
    if r == 1:   # 1x1/s2 data gradient accumulates into an existing block-input gradient
        base = util.bf16_round(torch.randn(N, ci, H, W))
        dx = util.conv_dgrad(dy, w, H, W, stride, lib.BF16, lib.IMPL_TCGEN05, accumulate_into=base)
        assert util.rel_err(dx, util.ref_dgrad(dy, w, H, W, stride) + base) <= 2e-2
    else:
        dx = util.conv_dgrad(dy, w, H, W, stride, lib.BF16, lib.IMPL_TCGEN05)
        assert util.rel_err(dx, util.ref_dgrad(dy, w, H, W, stride)) <= 2e-2
    dw = util.conv_wgrad(x, dy, r, stride, lib.BF16, lib.IMPL_TCGEN05)
    assert not torch.isnan(dw).any()
    assert util.rel_err(dw, util.ref_wgrad(x, dy, r, stride)) <= 1e-4      # fp32 accumulation and output


@pytest.mark.parametrize("impl", [lib.IMPL_SIMT, lib.IMPL_TCGEN05])
def test_conv_epilogues(impl):
    """Folded BN scale/shift, residual, ReLU and per-utterance valid widths (extraction path); masked residual (dgrad)."""
    shape = (10, 50, 128, 128, 3, 1)
    H, W, ci, co, r, stride = shape
    N = 3
    x, w, dy = util.make_case(shape, N, 5)
    g = torch.Generator().manual_seed(9)
    scale, shift = torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g)
    res = util.bf16_round(torch.randn(N, co, H, W, generator=g))
    valid = torch.tensor([50, 17, 33], dtype=torch.int32)
    y, _ = util.conv_fwd(x, w, stride, lib.BF16, impl, scale=scale, shift=shift, res_nchw=res, relu=1, valid_wo=valid)
    ref = F.relu(util.ref_conv(x, w, stride) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res)
    for n in range(N):
        ref[n, :, :, int(valid[n]):] = 0
    assert util.rel_err(y, ref) <= 2e-2
    assert float(y[1, :, :, 17:].abs().max()) == 0.0
    mask = util.bf16_round(torch.randn(N, ci, H, W, generator=g))
    resm = util.bf16_round(torch.randn(N, ci, H, W, generator=g))
    dx = util.conv_dgrad(dy, w, H, W, stride, lib.BF16, impl, resm_nchw=resm, mask_nchw=mask)
    assert util.rel_err(dx, util.ref_dgrad(dy, w, H, W, stride) + resm * (mask > 0)) <= 2e-2


BN_FUSE_SHAPES = [(40, 200, 32, 32, 3, 1), (20, 100, 64, 64, 3, 1), (10, 50, 128, 128, 3, 1), (5, 25, 256, 256, 3, 1),
                  (13, 27, 64, 32, 3, 1), (20, 100, 32, 64, 3, 2)]


@pytest.mark.parametrize("impl,code", [(lib.IMPL_TCGEN05, lib.BF16), (lib.IMPL_SIMT, lib.BF16), (lib.IMPL_SIMT, lib.F32)])
@pytest.mark.parametrize("shape", BN_FUSE_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_dgrad_bn_fusion(shape, impl, code):
    """svk_conv2d_dgrad_bn == svk_conv2d_dgrad -> ReLU mask -> svk_bn_bwd_reduce, on the values it stores."""
    H, W, ci, co, r, stride = shape
    N = 3
    x, w, dy = util.make_case(shape, N, 21, quantize=code == lib.BF16)
    g = torch.Generator().manual_seed(4)
    q = util.bf16_round if code == lib.BF16 else (lambda t: t)
    mask, c, res = [q(torch.randn(N, ci, H, W, generator=g)) for _ in range(3)]
    mean, rstd = torch.randn(ci, generator=g) * 0.3, torch.rand(ci, generator=g) + 0.5
    d = lib.make_conv_desc(N, H, W, ci, co, r, stride, code, impl)
    dyd, (_, wd) = util.nhwc(dy, code), util.pack(w, code)
    keep = [util.nhwc(t, code) for t in (mask, c, res)]
    md, rd = mean.cuda(), rstd.cuda()
    for with_c in (True, False):
        dx = torch.full((N, H, W, ci), float("nan"), dtype=util.tdtype(code), device="cuda")
        sums = torch.zeros(2, ci, dtype=torch.float64, device="cuda")
        bn = lib.BnBwdFuse(keep[0].data_ptr(), keep[1].data_ptr() if with_c else None, md.data_ptr() if with_c else None,
                           rd.data_ptr() if with_c else None, sums.data_ptr() if with_c else None)
        call.svk_conv2d_dgrad_bn(d, dyd.data_ptr(), wd.data_ptr(), dx.data_ptr(), keep[2].data_ptr() if with_c else 0, 0, 0,
                                 bn, util.st())
        torch.cuda.synchronize()
        got = util.nchw(dx)
        ref = (util.ref_dgrad(dy, w, H, W, stride) + (res if with_c else 0)) * (mask > 0)
        assert util.rel_err(got, ref) <= (2e-2 if code == lib.BF16 else 1e-4)
        assert float(got[mask <= 0].abs().max()) == 0.0
        if with_c:      # sums of the STORED gradient
            gd = got.double()
            xhat = (c.double() - mean.double().view(1, -1, 1, 1)) * rstd.double().view(1, -1, 1, 1)
            s_ref = torch.stack([gd.sum((0, 2, 3)), (gd * xhat).sum((0, 2, 3))])
            assert util.rel_err(sums.cpu(), s_ref) <= 1e-4
    # bad combinations are refused, not silently computed
    with pytest.raises(lib.SvkError):
        call.svk_conv2d_dgrad_bn(d, dyd.data_ptr(), wd.data_ptr(), dx.data_ptr(), 0, keep[2].data_ptr(), keep[0].data_ptr(),
                                 lib.BnBwdFuse(keep[0].data_ptr(), None, None, None, None), util.st())
    d1 = lib.make_conv_desc(N, H, W, ci, co, 1, 2, code, impl)
    with pytest.raises(lib.SvkError):
        call.svk_conv2d_dgrad_bn(d1, dyd.data_ptr(), wd.data_ptr(), dx.data_ptr(), 0, 0, 0,
                                 lib.BnBwdFuse(keep[0].data_ptr(), None, None, None, None), util.st())


@pytest.mark.parametrize("env", [{"SVK_DISABLE_EPI2": "1"}, {"SVK_DISABLE_WGRAD9": "1", "SVK_DISABLE_WGRADR": "1"},
                                 {"SVK_DISABLE_PAIR": "1"}, {"SVK_DISABLE_SINGLE_HALO": "1"},
                                 {"SVK_GATHER3_MMA": "1", "SVK_GATHER3_GROUPS": "2", "SVK_GATHER3_STAGES": "6"},
                                 {"SVK_GATHER3_MMA": "2", "SVK_GATHER3_GROUPS": "3", "SVK_GATHER3_STAGES": "12",
                                  "SVK_EPI2_MODE": "7", "SVK_EPI2_GROUPS": "3"}],
                         ids=["first-epilogue-only", "generic-wgrad", "single-cta-late-stages", "three-halo-loads",
                              "one-issuing-warp", "two-issuing-warps-everywhere"])
def test_conv_kernel_variants(env):
    """The library picks one kernel variant per shape (measured in the training step); the other variants stay selectable
    through the environment for A/B runs.  The switches are read once per process, so the convolution tests are re-run in a
    child process: every variant must meet the same parity bounds on every shape."""
    if os.environ.get("SVK_VARIANT_CHILD"):
        pytest.skip("child run")
    child = dict(os.environ, SVK_VARIANT_CHILD="1", **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                        "test_conv_tcgen05_path or (test_conv_dgrad_bn_fusion and 1-1)",
                        "-p", "no:cacheprovider"], env=child, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


@pytest.mark.parametrize("ch", [32, 64])
def test_resident_filter_conv_is_repeatable_at_full_size(ch):
    """Two MMA-issuing warps, three epilogue groups and six accumulator buffers share one CTA's barriers (conv_tc3.cu): 60
    back-to-back launches of the training forward and of the fused data gradient at the bench size must reproduce the first
    launch bit for bit (a barrier whose waiters drift two phases apart shows up here as corrupted tiles or a stalled kernel)."""
    import stress_conv
    assert stress_conv.run(ch, 60) == 0


@pytest.mark.parametrize("impl,code", [(lib.IMPL_TCGEN05, lib.BF16), (lib.IMPL_SIMT, lib.F32)])
@pytest.mark.parametrize("hw", [(40, 200, 32), (20, 100, 64), (10, 50, 128), (9, 27, 32)])
def test_downsample_block_dgrad(hw, impl, code):
    """svk_downsample_dgrad_bn == dgrad(3x3/s2) + dgrad(1x1/s2), masked, with the BatchNorm sums of the stored values."""
    H, W, ci = hw
    co, N = 2 * ci, 2
    x, w1, dy1 = util.make_case((H, W, ci, co, 3, 2), N, 31, quantize=code == lib.BF16)
    _, wd, dyd = util.make_case((H, W, ci, co, 1, 2), N, 32, quantize=code == lib.BF16)
    g = torch.Generator().manual_seed(6)
    q = util.bf16_round if code == lib.BF16 else (lambda t: t)
    mask, c = q(torch.randn(N, ci, H, W, generator=g)), q(torch.randn(N, ci, H, W, generator=g))
    mean, rstd = torch.randn(ci, generator=g) * 0.3, torch.rand(ci, generator=g) + 0.5
    d1 = lib.make_conv_desc(N, H, W, ci, co, 3, 2, code, impl)
    dd = lib.make_conv_desc(N, H, W, ci, co, 1, 2, code, impl)
    keep = [util.nhwc(dy1, code), util.pack(w1, code)[1], util.nhwc(dyd, code), util.pack(wd, code)[1],
            util.nhwc(mask, code), util.nhwc(c, code), mean.cuda(), rstd.cuda()]
    ref0 = util.ref_dgrad(dy1, w1, H, W, 2) + util.ref_dgrad(dyd, wd, H, W, 2)
    tol = 2e-2 if code == lib.BF16 else 1e-4
    for mode in ("plain", "mask", "sums"):
        dx = torch.full((N, H, W, ci), float("nan"), dtype=util.tdtype(code), device="cuda")
        sums = torch.zeros(2, ci, dtype=torch.float64, device="cuda")
        bn = None
        if mode == "mask":
            bn = lib.BnBwdFuse(keep[4].data_ptr(), None, None, None, None)
        elif mode == "sums":
            bn = lib.BnBwdFuse(keep[4].data_ptr(), keep[5].data_ptr(), keep[6].data_ptr(), keep[7].data_ptr(), sums.data_ptr())
        call.svk_downsample_dgrad_bn(d1, keep[0].data_ptr(), keep[1].data_ptr(), dd, keep[2].data_ptr(), keep[3].data_ptr(),
                                     dx.data_ptr(), bn, util.st())
        torch.cuda.synchronize()
        got = util.nchw(dx)
        ref = ref0 if mode == "plain" else ref0 * (mask > 0)
        assert not torch.isnan(got).any()
        assert util.rel_err(got, ref) <= tol, mode
        if mode == "sums":
            gd = got.double()
            xhat = (c.double() - mean.double().view(1, -1, 1, 1)) * rstd.double().view(1, -1, 1, 1)
            s_ref = torch.stack([gd.sum((0, 2, 3)), (gd * xhat).sum((0, 2, 3))])
            assert util.rel_err(sums.cpu(), s_ref) <= 1e-4


def test_stem_conv():
    g = torch.Generator().manual_seed(3)
    N, H, W, C = 3, 30, 51, 32
    x = torch.randn(N, H, W, generator=g)
    w = torch.randn(C, 1, 3, 3, generator=g)
    ref = F.conv2d(x.unsqueeze(1), w, None, 1, 1)
    xd, wd = x.cuda(), w.cuda()                      # keep the device copies alive across the asynchronous launches
    for code, tol in ((lib.F32, 1e-5), (lib.BF16, 2e-2)):
        y = torch.empty(N, H, W, C, dtype=util.tdtype(code), device="cuda")
        stats = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
        call.svk_stem_conv_fwd(xd.data_ptr(), wd.data_ptr(), y.data_ptr(), N, H, W, C, code, 0, 0, 0, 0, stats.data_ptr(), util.st())
        assert util.rel_err(util.nchw(y), ref) <= tol
        yd = util.nchw(y).double()           # statistics of the values as stored
        assert util.rel_err(stats.cpu(), torch.cat([yd.sum((0, 2, 3)), (yd * yd).sum((0, 2, 3))])) <= 1e-5
        # folded scale / shift / ReLU and per-utterance valid widths (extraction path)
        sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
        vw = torch.tensor([51, 7, 30], dtype=torch.int32)
        scd, shd, vwd = sc.cuda(), sh.cuda(), vw.cuda()
        call.svk_stem_conv_fwd(xd.data_ptr(), wd.data_ptr(), y.data_ptr(), N, H, W, C, code, scd.data_ptr(), shd.data_ptr(), 1,
                               vwd.data_ptr(), 0, util.st())
        ref2 = F.relu(ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
        for n in range(N):
            ref2[n, :, :, int(vw[n]):] = 0
        assert util.rel_err(util.nchw(y), ref2) <= tol
        dy = util.bf16_round(torch.randn(N, C, H, W, generator=g))
        dyd = util.nhwc(dy, code)
        dw = torch.empty(C, 9, device="cuda")
        call.svk_stem_conv_wgrad(xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), N, H, W, C, code, util.st())
        ref_dw = util.ref_wgrad(x.unsqueeze(1), dy, 3, 1).reshape(C, 9)
        assert util.rel_err(dw.cpu(), ref_dw) <= 1e-4


# ------------------------------------------------------------------------------------------------ BatchNorm
@pytest.mark.parametrize("code", [lib.F32, lib.BF16])
@pytest.mark.parametrize("C", [32, 256])
def test_batchnorm_train_forward_backward(code, C):
    g = torch.Generator().manual_seed(C)
    N, H, W = 4, 6, 9
    M = N * H * W
    tol = 1e-4 if code == lib.F32 else 2e-2
    q = (lambda t: t) if code == lib.F32 else util.bf16_round
    c = q(torch.randn(N, C, H, W, generator=g) * 2 + 0.5)
    cb = q(torch.randn(N, C, H, W, generator=g))
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    gamma_b, beta_b = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    dout = q(torch.randn(N, C, H, W, generator=g))
    # oracle: out = relu(bn(c) + bn_b(cb)); gradients w.r.t. c, cb, gamma, beta by autograd
    cr, cbr = c.clone().requires_grad_(True), cb.clone().requires_grad_(True)
    gr, br, gbr, bbr = [t.clone().requires_grad_(True) for t in (gamma, beta, gamma_b, beta_b)]
    rm, rv = torch.zeros(C), torch.ones(C)
    out_ref = F.relu(F.batch_norm(cr, rm, rv, gr, br, True, 0.1, 1e-5) + F.batch_norm(cbr, None, None, gbr, bbr, True, 0.1, 1e-5))
    out_ref.backward(dout)
    dev = lambda t: t.float().cuda()
    cd, cbd = util.nhwc(c, code), util.nhwc(cb, code)
    stats = torch.zeros(2, 2 * C, dtype=torch.float64, device="cuda")
    call.svk_channel_stats(cd.data_ptr(), M, C, code, stats[0].data_ptr(), util.st())
    call.svk_channel_stats(cbd.data_ptr(), M, C, code, stats[1].data_ptr(), util.st())
    coef = torch.zeros(2, 4, C, device="cuda")
    rmd, rvd = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    out = torch.empty_like(cd)
    gd, bd, gbd, bbd = dev(gamma), dev(beta), dev(gamma_b), dev(beta_b)
    call.svk_bn_train_act_fwd(cd.data_ptr(), stats[0].data_ptr(), gd.data_ptr(), bd.data_ptr(), rmd.data_ptr(), rvd.data_ptr(),
                              coef[0].data_ptr(), cbd.data_ptr(), stats[1].data_ptr(), gbd.data_ptr(), bbd.data_ptr(), 0, 0,
                              coef[1].data_ptr(), C, 0.1, 1e-5, 1, out.data_ptr(), M, C, code, util.st())
    assert util.rel_err(util.nchw(out), out_ref.detach()) <= tol
    assert util.rel_err(rmd.cpu(), rm) <= 1e-5 and util.rel_err(rvd.cpu(), rv) <= 1e-5      # running stats (unbiased var)
    # the separate finalise kernel publishes the same coefficients
    coef2 = torch.zeros(4, C, device="cuda")
    call.svk_bn_finalize(stats[0].data_ptr(), M, C, gd.data_ptr(), bd.data_ptr(), 0, 0, 0.1, 1e-5, coef2[0].data_ptr(),
                         coef2[1].data_ptr(), coef2[2].data_ptr(), coef2[3].data_ptr(), util.st())
    assert util.rel_err(coef2.cpu(), coef[0].cpu()) <= 1e-5
    # backward
    sums = torch.zeros(3 * C, dtype=torch.float64, device="cuda")
    dd = util.nhwc(dout, code)
    call.svk_bn_bwd_reduce(dd.data_ptr(), out.data_ptr(), cd.data_ptr(), coef[0][2].data_ptr(), coef[0][3].data_ptr(),
                           cbd.data_ptr(), coef[1][2].data_ptr(), coef[1][3].data_ptr(), sums.data_ptr(), M, C, code, util.st())
    dc, dcb = torch.empty_like(cd), torch.empty_like(cd)
    grads = torch.zeros(4, C, device="cuda")
    call.svk_bn_bwd_apply(dd.data_ptr(), out.data_ptr(), cd.data_ptr(), coef[0][2].data_ptr(), coef[0][3].data_ptr(),
                          gd.data_ptr(), dc.data_ptr(), cbd.data_ptr(), coef[1][2].data_ptr(), coef[1][3].data_ptr(),
                          gbd.data_ptr(), dcb.data_ptr(), sums.data_ptr(), grads[0].data_ptr(), grads[1].data_ptr(),
                          grads[2].data_ptr(), grads[3].data_ptr(), M, C, code, util.st())
    # in bf16 the CUDA mask comes from the rounded `out`; flips of |v| < 2^-9 are excluded by the norm-wise bound
    assert float((util.nchw(dc) - cr.grad).norm() / cr.grad.norm()) <= tol
    assert float((util.nchw(dcb) - cbr.grad).norm() / cbr.grad.norm()) <= tol
    for got, ref in zip(grads.cpu(), (gr.grad, br.grad, gbr.grad, bbr.grad)):
        assert util.rel_err(got, ref) <= tol


# ------------------------------------------------------------------------------------------------ pooling, FC, head, loss
@pytest.mark.parametrize("mode", [0, 1])
def test_stats_pooling(mode):
    g = torch.Generator().manual_seed(2)
    N, C, H, W = 3, 256, 5, 13
    x = F.relu(torch.randn(N, C, H, W, generator=g))
    x[0, 3, 2, :] = 0                                        # an all-zero row: sqrt(mean)=0, gradient defined as 0
    pooling = "mean+std" if mode else "mean"
    xr = x.clone().requires_grad_(True)
    ref = torch.flatten(O.stats_pooling(xr, pooling), 1, -1)
    dout = torch.randn(ref.shape, generator=g)
    xd = util.nhwc(x, lib.F32)
    out = torch.empty(N, ref.shape[1], device="cuda")
    call.svk_statspool_fwd(xd.data_ptr(), out.data_ptr(), N, H, W, C, mode, 0, lib.F32, util.st())
    assert util.rel_err(out.cpu(), ref.detach()) <= 1e-5
    dx = torch.empty_like(xd)
    doutd = dout.cuda()
    call.svk_statspool_bwd(xd.data_ptr(), doutd.data_ptr(), dx.data_ptr(), N, H, W, C, mode, 0, lib.F32, util.st())
    (ref * dout).sum().backward()
    gref = torch.nan_to_num(xr.grad, nan=0.0, posinf=0.0, neginf=0.0) * (x > 0)      # what survives the ReLU mask upstream
    assert util.rel_err(util.nchw(dx) * (x > 0), gref) <= 1e-5
    dxm = torch.empty_like(xd)                               # the same with the ReLU mask applied by the kernel
    call.svk_statspool_bwd(xd.data_ptr(), doutd.data_ptr(), dxm.data_ptr(), N, H, W, C, mode, 1, lib.F32, util.st())
    assert util.rel_err(util.nchw(dxm), gref) <= 1e-5
    assert float(util.nchw(dxm)[x <= 0].abs().max()) == 0.0
    # per-utterance valid widths
    lens = torch.tensor([13, 7, 10], dtype=torch.int32)
    lensd = lens.cuda()
    call.svk_statspool_fwd(xd.data_ptr(), out.data_ptr(), N, H, W, C, mode, lensd.data_ptr(), lib.F32, util.st())
    for n in range(N):
        r = torch.flatten(O.stats_pooling(x[n:n + 1, :, :, :int(lens[n])], pooling), 1, -1)
        assert util.rel_err(out[n:n + 1].cpu(), r) <= 1e-5


def test_sgemm_all_layouts():
    g = torch.Generator().manual_seed(4)
    M, N, K = 70, 130, 45
    A, B, bias = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g), torch.randn(N, generator=g)
    ref = A @ B + bias
    for at in (False, True):
        for bt in (False, True):
            a = (A.t().contiguous() if at else A).cuda()
            b = (B.t().contiguous() if bt else B).cuda()
            c = torch.empty(M, N, device="cuda")
            biasd = bias.cuda()
            call.svk_sgemm(a.data_ptr(), 1 if at else K, M if at else 1, b.data_ptr(), 1 if bt else N, K if bt else 1,
                           c.data_ptr(), N, M, N, K, 1.0, 0.0, biasd.data_ptr(), util.st())
            assert util.rel_err(c.cpu(), ref) <= 1e-5


@pytest.mark.parametrize("shape", [(256, 5994, 256), (256, 256, 5994), (5994, 256, 256), (256, 256, 2560), (37, 130, 45 * 4),
                                   (2, 256, 2560)])
def test_gemm_tf32_all_layouts(shape):
    """tcgen05 kind::tf32 GEMM (product-mode fc1 / AAM cosine / cohort GEMMs): all four operand layouts, split-K, ragged
    M/N/K, bias.  tf32 keeps 10 mantissa bits: 2e-3 relative to max |C| is the bound (the bf16-mode tolerance is 2e-2)."""
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    A, Bm, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    ref = A.double() @ Bm.double().t() + bias.double()
    pad4 = lambda n: (n + 3) // 4 * 4
    for a_k in (True, False):
        for b_k in (True, False):
            # stored layouts with 16-byte-aligned row pitches
            if a_k:
                a = torch.zeros(M, pad4(K)); a[:, :K] = A; lda = pad4(K)
            else:
                a = torch.zeros(K, pad4(M)); a[:, :M] = A.t(); lda = pad4(M)
            if b_k:
                b = torch.zeros(N, pad4(K)); b[:, :K] = Bm; ldb = pad4(K)
            else:
                b = torch.zeros(K, pad4(N)); b[:, :N] = Bm.t(); ldb = pad4(N)
            ad, bd, biasd = a.cuda(), b.cuda(), bias.cuda()
            c = torch.full((M, N), float("nan"), device="cuda")
            need = lib.load().svk_gemm_tf32_workspace_bytes(M, N, K)
            ws = torch.empty(need // 4 + 4, device="cuda")
            call.svk_gemm_tf32(ad.data_ptr(), lda, int(a_k), bd.data_ptr(), ldb, int(b_k), c.data_ptr(), N, M, N, K,
                               biasd.data_ptr(), ws.data_ptr(), need, util.st())
            torch.cuda.synchronize()
            assert not torch.isnan(c).any(), (a_k, b_k)
            assert util.rel_err(c.cpu(), ref) <= 2e-3, (a_k, b_k, util.rel_err(c.cpu(), ref))


def test_aam_head_and_cross_entropy():
    from svk.loss import CrossEntropyLoss, accuracy
    g = torch.Generator().manual_seed(6)
    B, E, C, m, s = 16, 256, 301, 0.2, 30.0
    x, w = torch.randn(B, E, generator=g), torch.randn(C, E, generator=g) * 0.1
    y = torch.randint(0, C, (B,), generator=g)
    w[y[0]] = x[0] * 3                                       # one very confident target: exercises cos > th and sine ~ 0
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    logits_ref = O.aam_logits(xr, wr, y, m, s)
    loss_ref = O.cross_entropy(logits_ref, y)
    loss_ref.backward()
    cos_m, sin_m, th, mm = O.aam_constants(m)
    st = util.st()
    xd, wd, yd = x.cuda(), w.cuda(), y.cuda()
    xh, wh = torch.empty_like(xd), torch.empty_like(wd)
    xinv, winv = torch.empty(B, device="cuda"), torch.empty(C, device="cuda")
    call.svk_l2norm_rows_fwd(xd.data_ptr(), xh.data_ptr(), xinv.data_ptr(), B, E, 1e-12, st)
    call.svk_l2norm_rows_fwd(wd.data_ptr(), wh.data_ptr(), winv.data_ptr(), C, E, 1e-12, st)
    logits = torch.empty(B, C, device="cuda")
    call.svk_sgemm(xh.data_ptr(), E, 1, wh.data_ptr(), 1, E, logits.data_ptr(), C, B, C, E, 1.0, 0.0, 0, st)
    cos_t = torch.empty(B, device="cuda")
    call.svk_aam_margin_fwd(logits.data_ptr(), yd.data_ptr(), cos_t.data_ptr(), B, C, cos_m, sin_m, th, mm, s, st)
    assert util.rel_err(logits.cpu(), logits_ref.detach()) <= 1e-5
    lg = logits.clone().requires_grad_(True)
    loss = CrossEntropyLoss()(lg, yd)
    assert abs(float(loss) - float(loss_ref.detach())) <= 1e-5
    loss.backward()
    a_ref = O.accuracy(logits_ref.detach(), y, (1, 5))
    a_got = accuracy(logits, yd, (1, 5))
    assert [float(a_got[0]), float(a_got[1])] == [float(a_ref[0]), float(a_ref[1])]
    d = lg.grad.clone()
    call.svk_aam_margin_bwd(d.data_ptr(), yd.data_ptr(), cos_t.data_ptr(), B, C, C, cos_m, sin_m, th, s, st)
    dxh, dwh = torch.empty(B, E, device="cuda"), torch.empty(C, E, device="cuda")
    call.svk_sgemm(d.data_ptr(), C, 1, wh.data_ptr(), E, 1, dxh.data_ptr(), E, B, E, C, 1.0, 0.0, 0, st)
    call.svk_sgemm(d.data_ptr(), 1, C, xh.data_ptr(), E, 1, dwh.data_ptr(), E, C, E, B, 1.0, 0.0, 0, st)
    dx, dw = torch.empty_like(xd), torch.empty_like(wd)
    call.svk_l2norm_rows_bwd(dxh.data_ptr(), xh.data_ptr(), xinv.data_ptr(), dx.data_ptr(), B, E, st)
    call.svk_l2norm_rows_bwd(dwh.data_ptr(), wh.data_ptr(), winv.data_ptr(), dw.data_ptr(), C, E, st)
    assert util.rel_err(dx.cpu(), xr.grad) <= 1e-4
    assert util.rel_err(dw.cpu(), wr.grad) <= 1e-4


def test_sgd_matches_torch_sgd_over_three_steps():
    from svk.optim import SGD
    g = torch.Generator().manual_seed(8)
    p0 = [torch.randn(37, 5, generator=g), torch.randn(11, generator=g)]
    grads = [[torch.randn(37, 5, generator=g), torch.randn(11, generator=g)] for _ in range(3)]
    ref = [t.clone().requires_grad_(True) for t in p0]
    mine = [t.clone().cuda().requires_grad_(True) for t in p0]
    o_ref = torch.optim.SGD(ref, 0.1, momentum=0.9, weight_decay=5e-4)
    o_mine = SGD(mine, 0.1, momentum=0.9, weight_decay=5e-4)
    for step in range(3):
        for t, m_, gr in zip(ref, mine, grads[step]):
            t.grad = gr.clone()
            m_.grad = gr.clone().cuda()
        o_ref.step()
        o_mine.step()
    for t, m_ in zip(ref, mine):
        assert util.rel_err(m_.detach().cpu(), t.detach()) <= 1e-6
    sd = o_mine.state_dict()                                  # torch's state-dict layout: loads into torch.optim.SGD
    o_ref.load_state_dict({"state": {k: {"momentum_buffer": v["momentum_buffer"].cpu()} for k, v in sd["state"].items()},
                           "param_groups": sd["param_groups"]})


# ------------------------------------------------------------------------------------------------ scoring
def test_scoring_kernels_match_reference_scripts():
    from svk import scoring
    fx = np.load(os.path.join(util.ROOT, "tests", "golden", "scoring.npz"))
    emb64 = np.array([[float(t) for t in map(str, v)] for v in fx["emb"]], dtype=np.float64)
    coh64 = np.array([[float(t) for t in map(str, v)] for v in fx["coh"]], dtype=np.float64)
    mean64 = np.array([float(t) for t in map(str, fx["mean"])], dtype=np.float64)
    e32 = (emb64 - mean64).astype(np.float32)
    c32 = (coh64 - mean64).astype(np.float32)
    s = scoring.cosine_scores(e32, e32, None, fx["ie"], fx["it"]).cpu().numpy()
    assert np.abs(s - fx["scores"]).max() <= 1e-6            # north-star bound: 1e-3 absolute
    m, sd = scoring.cohort_topk_meanstd(e32, c32, topk=300, block_rows=16)
    assert np.abs(m.cpu().numpy() - fx["topk_mean"]).max() <= 1e-6
    assert np.abs(sd.cpu().numpy() - fx["topk_std"]).max() <= 1e-6
    m2, sd2 = scoring.cohort_topk_meanstd(e32, c32, topk=300, tf32=True)          # tensor-core score matrix
    assert np.abs(m2.cpu().numpy() - fx["topk_mean"]).max() <= 1e-3 and np.abs(sd2.cpu().numpy() - fx["topk_std"]).max() <= 1e-3
    sn = scoring.snorm_apply(fx["scores"], fx["ie"], fx["it"], fx["topk_mean"], fx["topk_std"], fx["topk_mean"],
                             fx["topk_std"]).cpu().numpy()
    assert np.abs(sn - fx["snorm"]).max() <= 1e-4
    # edge cases: ties at the k-th value, k == cohort size, empty trial list
    ties = torch.zeros(2, 400)
    ties[0, :250] = 1.0
    ties[1] = torch.arange(400).float()
    mean = torch.empty(2, device="cuda")
    std = torch.empty(2, device="cuda")
    tiesd = ties.cuda()
    call.svk_topk_meanstd(tiesd.data_ptr(), 2, 400, 300, mean.data_ptr(), std.data_ptr(), util.st())
    ref0 = torch.cat([torch.ones(250), torch.zeros(50)])
    assert abs(float(mean[0]) - float(ref0.mean())) <= 1e-6 and abs(float(std[0]) - float(ref0.std())) <= 1e-6
    assert abs(float(mean[1]) - 249.5) <= 1e-3 and abs(float(std[1]) - float(torch.arange(100, 400).float().std())) <= 1e-3
    assert scoring.cosine_scores(e32, e32, None, np.zeros(0, np.int32), np.zeros(0, np.int32)).numel() == 0


def test_backend_kernels_match_reference_scripts():
    """Radix sort, detection metrics and mean kernels against the reference scripts' outputs."""
    from svk import scoring
    fx = np.load(os.path.join(util.ROOT, "tests", "golden", "backend.npz"))
    scores, labels = fx["scores"], fx["labels"]
    r = scoring.det_metrics(scores, labels, 0.01, 1.0, 1.0)
    assert r["eer"] == float(fx["eer_value"]) and "{0:.2%}".format(r["eer"]) == str(fx["eer_out"])          # bit-exact float64
    assert r["min_dcf"] == float(fx["dcf001_value"])
    r2 = scoring.det_metrics(scores, labels, 0.05, 2.0, 1.5)
    assert r2["min_dcf"] == float(fx["dcf05_value"]) and "{0:.4f}".format(r2["min_dcf"]) == str(fx["dcf05_out"])
    assert r2["min_dcf_threshold"] == O.min_dcf(scores, labels, 0.05, 2.0, 1.5)[1]
    assert r["n_target"] == int(labels.sum()) and r["n_nontarget"] == int((1 - labels).sum())
    # the sort itself: stable, float64 keys, negative / zero / denormal / tied values, sizes around the 2,048-key chunk
    g = np.random.RandomState(3)
    for n in (1, 2, 255, 2048, 2049, 100003):
        k = np.round(g.randn(n) * 3, 1 if n > 1000 else 3)
        k[g.randint(0, n, max(1, n // 50))] = 0.0
        k[g.randint(0, n, max(1, n // 50))] = -0.0
        k[g.randint(0, n, max(1, n // 100))] = 5e-324
        v = np.arange(n, dtype=np.int32)
        ks, vs = scoring.sort_pairs(k, v)
        order = np.argsort(k, kind="stable")
        assert np.array_equal(vs.cpu().numpy()[k[order] != 0], v[order][k[order] != 0])      # (+0.0 / -0.0 compare equal:
        assert np.array_equal(ks.cpu().numpy(), k[order])                                      #  their relative order is free)
    # properties at a size the Python reference would take minutes for: 1 M trials
    n = 1 << 20
    lab = (g.rand(n) < 0.3).astype(np.int32)
    sc = g.randn(n) + 2.0 * lab
    big = scoring.det_metrics(sc, lab)
    e_ref, _ = O.eer(sc, lab)
    m_ref, t_ref = O.min_dcf(sc, lab)
    assert big["eer"] == e_ref and big["min_dcf"] == m_ref and big["min_dcf_threshold"] == t_ref
    # means
    emb64 = np.array([[float(t) for t in map(str, v)] for v in fx["emb"]], dtype=np.float64)
    e32 = emb64.astype(np.float32)
    n_seg = int(fx["seg"].max()) + 1
    sm = scoring.speaker_means(e32, fx["seg"], n_seg).cpu().numpy()
    assert np.abs(sm - fx["spk_mean"]).max() <= 1e-6
    gm = scoring.global_mean(e32).cpu().numpy()
    assert np.abs(gm - fx["mean"]).max() <= 1e-6


def test_backend_scripts_end_to_end_on_reference_files():
    """compute_mean.py / compute_speaker_mean.py / compute_eer.py / compute_min_dcf.py drop-ins on the reference's files:
    same command lines, same output text (text embeddings and a binary-vector ark give the same means)."""
    import kaldi_io
    fx = np.load(os.path.join(util.ROOT, "tests", "golden", "backend.npz"))
    scripts = os.path.join(util.PKG, "scripts")
    with tempfile.TemporaryDirectory() as d:
        for k in ("emb.iv", "utt2spk", "trials", "scores"):
            open(os.path.join(d, k), "w").write(str(fx["file/" + k]))
        run = lambda *a: subprocess.run([sys.executable] + list(a), check=True, capture_output=True, text=True, cwd=d)
        assert run(os.path.join(scripts, "compute_eer.py"), "scores", "trials").stdout.strip() == str(fx["eer_out"])
        assert run(os.path.join(scripts, "compute_min_dcf.py"), "--p-target", "0.01", "scores", "trials").stdout.strip() == str(fx["dcf001_out"])
        r = run(os.path.join(scripts, "compute_min_dcf.py"), "--p-target", "0.05", "--c-miss", "2", "--c-fa", "1.5", "scores", "trials")
        assert r.stdout.strip() == str(fx["dcf05_out"])
        assert r.stderr.strip().splitlines()[-1] == str(fx["dcf05_err"])
        run(os.path.join(scripts, "compute_speaker_mean.py"), "emb.iv", "utt2spk", "spk_mean.iv")
        ref_lines = str(fx["file/spk_mean.iv"]).splitlines()
        got_lines = open(os.path.join(d, "spk_mean.iv")).read().splitlines()
        assert [l.split()[0] for l in got_lines] == [l.split()[0] for l in ref_lines]          # same speakers, same order
        got = np.array([[float(t) for t in l.split()[2:-1]] for l in got_lines])
        assert np.abs(got - fx["spk_mean"]).max() <= 1e-6
        # compute_mean_byspk.py: the same means from a spk2utt listing (the reference's own script gives the compute_speaker_mean
        # values within 4e-7 on this fixture: torch.mean of the float32 rows vs a float32 running sum)
        groups, order = {}, []
        for line in str(fx["file/utt2spk"]).splitlines():
            u, sp = line.split()
            if sp not in groups:
                groups[sp] = []
                order.append(sp)
            groups[sp].append(u)
        with open(os.path.join(d, "spk2utt"), "w") as f:
            f.writelines("%s %s\n" % (sp, " ".join(groups[sp])) for sp in reversed(order))
        r = run(os.path.join(scripts, "compute_mean_byspk.py"), "spk2utt", "emb.iv", "byspk.iv")
        assert "speakers: %d, feat-dim: %d" % (len(order), fx["spk_mean"].shape[1]) in r.stdout
        by_lines = open(os.path.join(d, "byspk.iv")).read().splitlines()
        assert [l.split()[0] for l in by_lines] == list(reversed(order))                        # spk2utt order
        ref_of = {l.split()[0]: row for l, row in zip(ref_lines, fx["spk_mean"])}
        for l in by_lines:
            assert np.abs(np.array([float(t) for t in l.split()[2:-1]]) - ref_of[l.split()[0]]).max() <= 1e-6
        run(os.path.join(scripts, "compute_mean.py"), "emb.iv", "mean.vec")
        txt = open(os.path.join(d, "mean.vec")).read()
        assert txt.startswith(" [ ") and txt.endswith(" ]\n")
        assert np.abs(np.array([float(t) for t in txt.split()[1:-1]]) - fx["mean"]).max() <= 1e-6
        # binary fast path: the same embeddings as a Kaldi binary float-vector ark
        keys = [l.split()[0] for l in str(fx["file/emb.iv"]).splitlines()]
        with open(os.path.join(d, "emb.ark"), "wb") as f:
            for k, v in zip(keys, fx["emb"]):
                kaldi_io.write_vec_flt(f, np.asarray(v, np.float32), key=k)
        run(os.path.join(scripts, "compute_mean.py"), "emb.ark", "mean_b.vec")
        tb = open(os.path.join(d, "mean_b.vec")).read()
        assert np.abs(np.array([float(t) for t in tb.split()[1:-1]]) - fx["mean"]).max() <= 1e-6


def test_scoring_scripts_end_to_end_on_reference_files():
    """The CLI drop-ins read the files the reference's own scripts read and reproduce their output files."""
    fx = np.load(os.path.join(util.ROOT, "tests", "golden", "scoring.npz"))
    scripts = os.path.join(util.PKG, "scripts")
    with tempfile.TemporaryDirectory() as d:
        for k in ("emb.iv", "coh.iv", "mean.vec", "trials"):
            open(os.path.join(d, k), "w").write(str(fx["file/" + k]))
        run = lambda *a: subprocess.run([sys.executable] + list(a), check=True, capture_output=True, cwd=d)
        run(os.path.join(scripts, "cosine_score.py"), "--mean", "mean.vec", "--enroll", "emb.iv", "--test", "emb.iv",
            "--trials", "trials", "--score-file", "scores")
        run(os.path.join(scripts, "compute_topk_mean_std.py"), "--mean", "mean.vec", "--ark-file", "emb.iv", "--cohort-file",
            "coh.iv", "--mean-std-file", "topk")
        run(os.path.join(scripts, "adaptive_snorm.py"), "--enroll", "topk", "--test", "topk", "--score-in", "scores",
            "--score-out", "snorm")
        ref_lines = str(fx["file/scores"]).splitlines()
        got_lines = open(os.path.join(d, "scores")).read().splitlines()
        assert [l.split()[:2] for l in got_lines] == [l.split()[:2] for l in ref_lines]
        assert np.abs(np.array([float(l.split()[2]) for l in got_lines]) - fx["scores"]).max() <= 1e-6
        got_sn = np.array([float(l.split()[2]) for l in open(os.path.join(d, "snorm"))])
        assert np.abs(got_sn - fx["snorm"]).max() <= 1e-3
        topk = [l.split() for l in open(os.path.join(d, "topk"))]
        assert [t[0] for t in topk] == list(fx["utts"])
        assert np.abs(np.array([float(t[1]) for t in topk]) - fx["topk_mean"]).max() <= 1e-6


# ------------------------------------------------------------------------------------------------ extraction CLI
def test_decode_script_matches_batch1_oracle():
    """decode.py on a synthetic ark of variable-length utterances: text vectors in the reference's format, equal (cosine
    >= 0.999, bf16) to the oracle's batch-1 eval embedding of every utterance."""
    import contextlib
    import io
    import kaldi_io
    from model import NeuralSpeakerModel
    rs = np.random.RandomState(5)
    lengths = [40, 41, 57, 57, 96, 133]
    mats = {"utt%02d" % i: rs.randn(t, 40).astype(np.float32) for i, t in enumerate(lengths)}
    torch.manual_seed(21)
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(spk_num=12, feat_dim=40, pooling="mean+std", loss="AAM")
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    with tempfile.TemporaryDirectory() as d:
        ark, scp = os.path.join(d, "feats.ark"), os.path.join(d, "feats.scp")
        with open(ark, "wb") as f, open(scp, "w") as s:
            for key, mat in mats.items():
                f.write((key + " ").encode())
                s.write("%s %s:%d\n" % (key, ark, f.tell()))
                kaldi_io.write_mat(f, mat)
        torch.save({"epoch": 1, "arch": "resnet34", "state_dict": {"module." + k: v for k, v in sd.items()},
                    "best_acc1": 0.0, "optimizer": {}}, os.path.join(d, "ckpt.pth.tar"))
        subprocess.run([sys.executable, os.path.join(util.PKG, "scripts", "decode.py"), "--spk_num", "12", "--input-dim", "40",
                        "--pooling", "mean+std", "--model-path", os.path.join(d, "ckpt.pth.tar"), "--decode-scp", scp,
                        "--out-path", os.path.join(d, "emb"), "--gpu", "0", "--max-batch-frames", "200"],
                       check=True, capture_output=True)
        got = dict(kaldi_io.read_vec_flt_ark(os.path.join(d, "emb", "0")))
        # binary fast path: same embeddings as Kaldi float-vector ark, bit-identical to the text ones
        subprocess.run([sys.executable, os.path.join(util.PKG, "scripts", "decode.py"), "--spk_num", "12", "--input-dim", "40",
                        "--pooling", "mean+std", "--model-path", os.path.join(d, "ckpt.pth.tar"), "--decode-scp", scp,
                        "--out-path", os.path.join(d, "emb"), "--gpu", "0", "--max-batch-frames", "200", "--embed-format", "ark"],
                       check=True, capture_output=True)
        got_b = dict(kaldi_io.read_vec_flt_ark(os.path.join(d, "emb", "0.ark")))
    assert sorted(got) == sorted(mats) and sorted(got_b) == sorted(mats)
    for key in mats:
        assert got_b[key].dtype == np.float32 and np.array_equal(got_b[key], got[key].astype(np.float32))
    for key, mat in mats.items():
        with torch.no_grad():
            ref = O.embed(sd, torch.from_numpy(mat.T.copy()).unsqueeze(0), "mean+std", train=False)[0]
        cos = float(F.cosine_similarity(torch.from_numpy(got[key]).float(), ref, dim=0))
        assert cos >= 0.999, (key, cos)


# ------------------------------------------------------------------------------------------------ training CLI
def test_train_script_end_to_end_then_decode():
    """The recipe's stage order on synthetic data: train_resnet.py (2 epochs, AAM, mean+std) writes the reference's
    checkpoint dictionary, decode.py loads that checkpoint and writes one embedding per utterance;
    --resume restores model/optimizer/epoch."""
    import kaldi_io
    rs = np.random.RandomState(11)
    n_spk, utts_per_spk, F_ = 4, 6, 40
    with tempfile.TemporaryDirectory() as d:
        ark, scp, u2s = os.path.join(d, "feats.ark"), os.path.join(d, "train.scp"), os.path.join(d, "utt2spkid")
        with open(ark, "wb") as f, open(scp, "w") as s, open(u2s, "w") as u:
            for spk in range(n_spk):
                proto = rs.randn(1, F_).astype(np.float32) * 2          # speaker-dependent mean: a learnable task
                for k in range(utts_per_spk):
                    key = "spk%d-utt%d" % (spk, k)
                    mat = (proto + rs.randn(70 + 5 * k, F_)).astype(np.float32)
                    f.write((key + " ").encode())
                    s.write("%s %s:%d\n" % (key, ark, f.tell()))
                    kaldi_io.write_mat(f, mat)
                    u.write("%s %d\n" % (key, spk))
        log = os.path.join(d, "exp")
        base = [sys.executable, os.path.join(util.PKG, "scripts", "train_resnet.py"), "--train-list", scp, "--cv-list", scp,
                "--utt2spkid", u2s, "--input-dim", "40", "--spk-num", str(n_spk), "--pooling", "mean+std", "--loss-type", "AAM",
                "--margin", "0.2", "--scale", "30", "--max-chunk-size", "64", "--log-dir", log, "-j", "0", "-b", "8",
                "--lr", "0.05", "--wd", "1e-4", "-p", "1", "--gpu", "0", "--seed", "3"]
        r = subprocess.run(base + ["--epochs", "2"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        losses = [float(l.split("Loss")[1].split()[0]) for l in r.stdout.splitlines() if l.startswith("Epoch: [")]
        assert len(losses) >= 4 and all(np.isfinite(losses))          # convergence parity: test_model_gpu.py::test_loss_trajectory
        assert " * Acc@1" in r.stdout
        ck = torch.load(os.path.join(log, "checkpoint_epoch1.pth.tar"), map_location="cpu", weights_only=False)
        assert sorted(ck) == ["arch", "best_acc1", "epoch", "optimizer", "state_dict"] and ck["epoch"] == 2
        assert len(ck["state_dict"]) == 219 and "res.layer4.2.bn2.running_var" in ck["state_dict"]
        assert os.path.exists(os.path.join(log, "model_best.pth.tar"))
        assert len(ck["optimizer"]["state"]) == 111                    # torch.optim.SGD layout: one momentum buffer per tensor
        r2 = subprocess.run(base + ["--epochs", "3", "--resume", os.path.join(log, "checkpoint_epoch1.pth.tar")],
                            capture_output=True, text=True)
        assert r2.returncode == 0, r2.stderr[-2000:]
        assert "=> loaded checkpoint" in r2.stdout and "Epoch: [2]" in r2.stdout and "Epoch: [1]" not in r2.stdout
        r3 = subprocess.run([sys.executable, os.path.join(util.PKG, "scripts", "decode.py"), "--spk_num", str(n_spk), "--input-dim",
                             "40", "--pooling", "mean+std", "--model-path", os.path.join(log, "model_best.pth.tar"),
                             "--decode-scp", scp, "--out-path", os.path.join(d, "emb"), "--gpu", "0"], capture_output=True, text=True)
        assert r3.returncode == 0, r3.stderr[-2000:]
        emb = dict(kaldi_io.read_vec_flt_ark(os.path.join(d, "emb", "0")))
        assert len(emb) == n_spk * utts_per_spk and all(v.shape == (256,) and np.isfinite(v).all() for v in emb.values())


# ------------------------------------------------------------------------------------------------ fused AAM-softmax head
@pytest.mark.parametrize("B,C", [(256, 5994), (32, 1211), (37, 1211), (5, 64), (130, 200)])
@pytest.mark.parametrize("exact", [1, 0])
def test_fused_aam_softmax_cross_entropy(B, C, exact):
    """svk_aam_ce_fwd / svk_aam_ce_bwd (4 launches) against the oracle's AAMLayer + cross-entropy and their autograd
    (model.py:483-501, train_resnet.py:317, accuracy.py:4-16): logits, loss, log-sum-exp, target rank, d h, d W.
    exact=1 (fp32 validation mode, 3xTF32 products): <= 1e-5; exact=0 (product mode, single tf32 products): <= 2e-3 on the
    logits (tf32 rounding of unit vectors), gradients <= 2e-2 (the bf16 path's bound)."""
    E, m, s = 256, 0.2, 30.0
    g = torch.Generator().manual_seed(B * 7 + C)
    h = torch.randn(B, E, generator=g) * 3.0
    W = torch.randn(C, E, generator=g) * 0.05
    y = torch.randint(0, C, (B,), generator=g)
    # rows that hit the "cos - th <= 0" branch of the margin (cos = -0.995 < th = -0.980) and a large positive cosine (0.9);
    # not +-1 exactly: sqrt(1 - cos^2) is ill-conditioned there and two fp32 implementations legitimately differ
    for r, (a, b) in ((0, (-0.995, 0.0999)), (1 % B, (0.9, 0.436))):
        w = W[y[r]] / W[y[r]].norm()
        u = torch.randn(E, generator=g)
        u = u - (u @ w) * w
        h[r] = 5.0 * (a * w + b * u / u.norm())
    hr, Wr = h.clone().double().requires_grad_(True), W.clone().double().requires_grad_(True)
    ref_logits = O.aam_logits(hr, Wr, y, m, s)
    ref_loss = O.cross_entropy(ref_logits, y)
    gout = 0.37                                   # an upstream gradient other than 1
    (ref_loss * gout).backward()
    cos_m, sin_m, th, mm = O.aam_constants(m)
    hd, Wd, yd = h.cuda(), W.cuda(), y.cuda()
    need = lib.load().svk_aam_ce_workspace_bytes(B, E, C)
    assert need > 0
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    logits = torch.full((B, C), float("nan"), device="cuda")
    cos_t, lse, rows = [torch.full((B,), float("nan"), device="cuda") for _ in range(3)]
    rank = torch.full((B,), -1, dtype=torch.int32, device="cuda")
    loss = torch.full((), float("nan"), device="cuda")
    call.svk_aam_ce_fwd(hd.data_ptr(), Wd.data_ptr(), yd.data_ptr(), logits.data_ptr(), cos_t.data_ptr(), lse.data_ptr(),
                        rows.data_ptr(), rank.data_ptr(), loss.data_ptr(), B, E, C, cos_m, sin_m, th, mm, s, exact, ws.data_ptr(),
                        ws.numel(), util.st())
    tol = 1e-5 if exact else 2e-3
    assert not torch.isnan(logits).any()
    assert util.rel_err(logits.cpu(), ref_logits.detach()) <= tol
    assert abs(float(loss) - float(ref_loss)) <= tol * abs(float(ref_loss))
    assert util.rel_err(lse.cpu(), torch.logsumexp(ref_logits.detach(), 1)) <= tol
    assert util.rel_err(rows.cpu(), (torch.logsumexp(ref_logits.detach(), 1) - ref_logits.detach().gather(1, y.view(-1, 1))[:, 0])) <= max(tol, 1e-5) * 10
    got_logits = logits.cpu().double()
    assert torch.equal(rank.cpu().long(), (got_logits > got_logits.gather(1, y.view(-1, 1))).sum(1)), "rank of the stored logits"
    if exact:
        assert torch.equal(rank.cpu().long(), (ref_logits.detach() > ref_logits.detach().gather(1, y.view(-1, 1))).sum(1))
    gd = torch.tensor(gout, device="cuda")
    dh = torch.full((B, E), float("nan"), device="cuda")
    dW = torch.full((C, E), float("nan"), device="cuda")
    call.svk_aam_ce_bwd(hd.data_ptr(), Wd.data_ptr(), yd.data_ptr(), logits.data_ptr(), lse.data_ptr(), cos_t.data_ptr(),
                        gd.data_ptr(), dh.data_ptr(), dW.data_ptr(), B, E, C, cos_m, sin_m, th, s, exact, ws.data_ptr(), ws.numel(),
                        util.st())
    # fp64 oracle: fp32 rounding of the margin derivative at cos = -0.995 (amplified ~100x by 1 / sqrt(1 - cos^2)) alone is ~1e-5
    gtol = 3e-5 if exact else 2e-2
    assert not torch.isnan(dh).any() and not torch.isnan(dW).any()
    assert util.rel_err(dh.cpu(), hr.grad) <= gtol
    assert util.rel_err(dW.cpu(), Wr.grad) <= gtol
    dW2 = torch.empty_like(dW)
    call.svk_aam_ce_bwd(hd.data_ptr(), Wd.data_ptr(), yd.data_ptr(), logits.data_ptr(), lse.data_ptr(), cos_t.data_ptr(),
                        gd.data_ptr(), dh.data_ptr(), dW2.data_ptr(), B, E, C, cos_m, sin_m, th, s, exact, ws.data_ptr(), ws.numel(),
                        util.st())
    assert torch.equal(dW, dW2), "fused head backward is not deterministic"


def test_topk_meanstd_candidate_path_and_radix_fallback():
    """The top-k select has a fast path (k <= 512 on long rows: thread-local top-2 -> threshold -> candidate list) and the
    radix path (any k; also taken when the candidate list overflows): both equal torch.topk + std_mean on clustered,
    heavy-tie and adversarial rows."""
    g = torch.Generator().manual_seed(3)
    rows = [torch.randn(50000, generator=g) * 0.06,                       # cosine-like cluster
            torch.round(torch.randn(50000, generator=g) * 4) / 4,         # heavy ties
            torch.cat([torch.full((3000,), 0.5), torch.randn(47000, generator=g) * 0.01]),   # > 2048 candidates tie at the top
            -torch.rand(50000, generator=g),                              # all negative
            torch.linspace(-1, 1, 50000)]
    S = torch.stack(rows).cuda()
    for k in (300, 2, 512, 600):
        mean = torch.empty(len(rows), device="cuda")
        std = torch.empty(len(rows), device="cuda")
        call.svk_topk_meanstd(S.data_ptr(), len(rows), 50000, k, mean.data_ptr(), std.data_ptr(), util.st())
        top = S.double().cpu().topk(k, dim=1).values
        assert util.rel_err(mean.cpu(), top.mean(1)) <= 1e-6, k
        assert float((std.cpu().double() - top.std(1)).abs().max()) <= 1e-6, k
