"""The torch custom operators of svk/ops.py (`torch.ops.svk.*`) and the re-entrancy of the network op."""
import contextlib
import io

import pytest
import torch
import torch.nn.functional as F

import util
from oracle import ref_model as O

pytestmark = pytest.mark.gpu


def _model(precision="fp32", spk=23):
    from model import NeuralSpeakerModel
    torch.manual_seed(21)
    with contextlib.redirect_stdout(io.StringIO()):
        return NeuralSpeakerModel(spk_num=spk, feat_dim=40, pooling="mean+std", loss="AAM", precision=precision).cuda()


def test_conv2d_operator_with_autograd_matches_torch():
    from svk import ops  # noqa: F401
    for (H, W, ci, co, r, stride) in [(10, 50, 128, 128, 3, 1), (20, 100, 64, 128, 3, 2), (20, 100, 64, 128, 1, 2)]:
        x, w, dy = util.make_case((H, W, ci, co, r, stride), 3, 31)
        xd = x.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16).requires_grad_(True)
        wd = w.cuda().requires_grad_(True)
        y = torch.ops.svk.conv2d(xd, wd, stride)
        y.backward(dy.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16))
        assert util.rel_err(util.nchw(y.detach()), util.ref_conv(x, w, stride)) <= 2e-2
        assert util.rel_err(util.nchw(xd.grad), util.ref_dgrad(dy, w, H, W, stride)) <= 2e-2
        assert util.rel_err(wd.grad.cpu(), util.ref_wgrad(x, dy, r, stride)) <= 1e-4
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.svk.conv2d(torch.zeros(1, 4, 4, 32), torch.zeros(32, 32, 3, 3), 1)      # CPU tensors: no kernel, no fallback


def test_cross_entropy_operator_opcheck_and_values():
    from svk import ops  # noqa: F401
    g = torch.Generator().manual_seed(2)
    z = (torch.randn(9, 1211, generator=g) * 4).cuda().requires_grad_(True)
    y = torch.randint(0, 1211, (9,), generator=g).cuda()
    loss, lse = torch.ops.svk.cross_entropy(z, y)
    loss.backward()
    zr = z.detach().cpu().double().requires_grad_(True)
    ref = F.cross_entropy(zr, y.cpu())
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert util.rel_err(z.grad.cpu(), zr.grad) <= 1e-5
    assert util.rel_err(lse.detach().cpu(), torch.logsumexp(zr.detach(), 1)) <= 1e-6
    rank = torch.ops.svk.target_rank(z.detach(), y).cpu()
    ref_rank = (zr.detach() > zr.detach().gather(1, y.cpu().view(-1, 1))).sum(1)
    assert torch.equal(rank.long(), ref_rank)
    torch.library.opcheck(torch.ops.svk.cross_entropy, (z.detach().requires_grad_(True), y),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    torch.library.opcheck(torch.ops.svk.target_rank, (z.detach(), y), test_utils=("test_schema", "test_faketensor"))


def test_out_of_range_label_traps_instead_of_reading_out_of_bounds():
    """A label >= C is a device-side assert (like torch's CrossEntropyLoss), not a silent out-of-bounds read.  The trap
    poisons the CUDA context, so it runs in a child process."""
    import subprocess
    import sys
    code = ("import sys; sys.path[:0] = %r\n"
            "import torch\nfrom svk.loss import CrossEntropyLoss\n"
            "z = torch.randn(4, 10).cuda(); y = torch.tensor([1, 2, 10, 3]).cuda()\n"
            "try:\n    CrossEntropyLoss()(z, y); torch.cuda.synchronize(); print('NO ERROR')\n"
            "except RuntimeError as e:\n    print('RAISED', str(e)[:80])\n") % ([util.PKG, util.ROOT],)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "RAISED" in r.stdout and "NO ERROR" not in r.stdout, r.stdout + r.stderr
    assert "label 10 of row 2 is outside [0, 10)" in r.stdout + r.stderr


def test_network_operator_is_reentrant_and_accumulates_like_torch():
    """Two training forwards before their backwards keep separate saved state (a single engine-global slot
    silently corrupted the first backward), predict() in training mode does not disturb a pending backward, and a second
    backward before optimizer.step() accumulates into .grad like torch."""
    from svk.loss import CrossEntropyLoss
    from svk.optim import SGD
    m = _model("fp32")
    crit = CrossEntropyLoss()
    opt = SGD(m.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(8)
    xa, xb = torch.randn(3, 40, 48, generator=g).cuda(), torch.randn(3, 40, 48, generator=g).cuda()
    ya, yb = torch.randint(0, 23, (3,), generator=g).cuda(), torch.randint(0, 23, (3,), generator=g).cuda()
    m.train()

    def grads_of(x, y):
        opt.zero_grad()
        crit(m(x, y), y).backward()
        torch.cuda.synchronize()
        return {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    ga, gb = grads_of(xa, ya), grads_of(xb, yb)
    # interleaved: forward a, forward b, predict (training mode), backward a, backward b
    opt.zero_grad()
    la = crit(m(xa, ya), ya)
    lb = crit(m(xb, yb), yb)
    m.predict(xa)
    la.backward()
    torch.cuda.synchronize()
    for n, p in m.named_parameters():
        assert util.rel_err(p.grad.cpu(), ga[n].cpu()) <= 1e-6, "first backward was corrupted by the second forward: " + n
    lb.backward()           # no zero_grad in between: accumulates
    torch.cuda.synchronize()
    for n, p in m.named_parameters():
        assert util.rel_err(p.grad.cpu(), (ga[n] + gb[n]).cpu()) <= 1e-5, "gradient accumulation differs from torch semantics: " + n
    # a forward whose graph is dropped releases its workspace
    for _ in range(6):
        out = m(xa, ya)
        del out
    with torch.no_grad():
        for _ in range(6):
            m(xa, ya)
    assert not getattr(m.engine, "_pending", {})
    opt.step()
    # double backward of one forward is refused
    out = crit(m(xa, ya), ya)
    out.backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        out.backward()


@pytest.mark.parametrize("loss_type", ["AAM", "AAM-v1", "softmax"])
def test_forward_loss_equals_model_plus_criterion(loss_type):
    """model.forward_loss (fused AAM-softmax-CE head, svk::speaker_net_train_loss: backward starts from d loss) gives the
    loss, logits, ranks and parameter gradients of model(x, y) + CrossEntropyLoss + backward (train_resnet.py:316-327)."""
    from model import NeuralSpeakerModel
    from svk.loss import CrossEntropyLoss, target_rank
    from svk.optim import SGD
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        m = NeuralSpeakerModel(spk_num=301, feat_dim=40, pooling="mean+std", loss=loss_type, precision="fp32").cuda()
    opt = SGD(m.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(6, 40, 56, generator=g).cuda()
    y = torch.randint(0, 301, (6,), generator=g).cuda()
    m.train()
    opt.zero_grad()
    logits_a = m(x, y)
    loss_a = CrossEntropyLoss()(logits_a, y)
    rank_a = target_rank(logits_a, y)
    (loss_a * 0.5).backward()
    torch.cuda.synchronize()
    ga = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    opt.zero_grad()
    loss_b, logits_b = m.forward_loss(x, y)
    (loss_b * 0.5).backward()
    torch.cuda.synchronize()
    assert abs(float(loss_a) - float(loss_b)) <= 1e-5 * abs(float(loss_a))
    assert util.rel_err(logits_b.cpu(), logits_a.detach().cpu()) <= 1e-5
    assert torch.equal(target_rank(logits_b, y).cpu(), rank_a.cpu())
    for n, p in m.named_parameters():
        if float(ga[n].abs().max()) < 1e-6:      # fc1.bias under a BatchNorm head: the gradient is identically zero, both sides are rounding noise
            assert float(p.grad.abs().max()) < 1e-6, n
        else:
            assert util.rel_err(p.grad.cpu(), ga[n].cpu()) <= 2e-5, n
    if loss_type != "softmax":
        assert m.engine.last_head is not None and hasattr(logits_b, "svk_rank")
        from svk import launch_count
        n0 = launch_count()
        m.engine._head_fwd(m.engine._train_ws[(6, 40, 56)][0]["emb"], y, m.engine._train_ws[(6, 40, 56)][0], None, True)
        assert launch_count() - n0 <= (2 if loss_type == "AAM" else 5), "fused head forward = 2 launches (+ 3 for the BatchNorm1d + ReLU of AAM-v1)"


def test_cuda_graph_training_step_equals_eager_steps():
    """svk.graph.GraphedTrainStep: replays of the captured step (forward, fused loss, backward with the side-stream
    weight gradients, SGD) leave the parameters, BatchNorm buffers and losses of four eager steps; a new learning rate
    captures a new graph."""
    from model import NeuralSpeakerModel
    from svk.graph import GraphedTrainStep
    from svk.optim import SGD

    def run(graphed):
        torch.manual_seed(9)
        with contextlib.redirect_stdout(io.StringIO()):
            m = NeuralSpeakerModel(spk_num=101, feat_dim=40, pooling="mean+std", loss="AAM", precision="fp32").cuda()
        opt = SGD(m.parameters(), 0.05, momentum=0.9, weight_decay=1e-4)
        m.train()
        g = torch.Generator().manual_seed(10)
        X = torch.randn(5, 8, 40, 64, generator=g).cuda()
        Y = torch.randint(0, 101, (5, 8), generator=g).cuda()
        step = GraphedTrainStep(m, opt, warmup=2) if graphed else None
        losses = []
        for i in range(4):
            if i == 2:
                opt.param_groups[0]["lr"] = 0.02
            if graphed:
                loss, logits = step(X[i], Y[i])
            else:
                loss, logits = m.forward_loss(X[i], Y[i])
                opt.zero_grad()
                loss.backward()
                opt.step()
            losses.append(float(loss))
            assert hasattr(logits, "svk_rank")
        torch.cuda.synchronize()
        if graphed:
            assert len(step._graphs) == 2
        return losses, {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}
    le, se = run(False)
    lg, sg = run(True)
    # fp32 validation mode: every kernel but the stem weight gradient (fp32 atomics across blocks, ~1e-7) is deterministic; a
    # random-init net with 8-chunk BatchNorm statistics amplifies that over the steps (measured 1.3e-3 after four), hence 1e-2
    assert max(abs(a - b) for a, b in zip(le, lg)) <= 1e-4 * max(abs(v) for v in le), (le, lg)
    for k in se:
        if k.endswith("num_batches_tracked"):
            assert int(se[k]) == int(sg[k]) == 4, k
        else:
            assert util.rel_err(sg[k], se[k]) <= 1e-2, k
