"""libsvk entry points registered as torch custom operators (`torch.ops.svk.*`).

SURVEY.md §8(b): the C-ABI of libsvk.so (include/svk.h) is called from Python through ctypes (svk/lib.py) and WRAPPED
as `torch.library.custom_op`s with `register_autograd`, so the drop-in modules (scripts/model.py, svk/loss.py,
svk/optim.py, svk/scoring.py) dispatch through the torch operator registry: the ops show up in profiler traces, FX graphs
and `torch.library.opcheck`, carry accurate mutation annotations, and have fake (meta) kernels for shape propagation.
CUDA only — there is no CPU kernel behind any of them (no fallback), a CPU tensor raises.

    op                                   replaces (reference)                       C-ABI behind it
    svk::speaker_net_train (+_backward)  NeuralSpeakerModel.forward, model.py:374   the whole plan of svk/engine.py
    svk::speaker_net_train_loss (+_bwd)  forward + CrossEntropyLoss (:316-317)      same, with the fused AAM-softmax-CE head
    svk::speaker_net_embed               NeuralSpeakerModel.predict, model.py:402   eval plan (BatchNorm folded)
    svk::cross_entropy (+_backward)      nn.CrossEntropyLoss, train_resnet.py:201   svk_ce_fwd / svk_ce_bwd
    svk::target_rank                     accuracy(), accuracy.py:4-16               svk_ce_fwd (rank output)
    svk::sgd_step                        torch.optim.SGD.step, train_resnet.py:328  svk_sgd_step
    svk::conv2d (+ autograd)             nn.Conv2d 3x3 / 1x1-s2, model.py:12-15     svk_conv2d_fwd / _dgrad / _wgrad
    svk::cosine_score_pairs              cosine_score.py:60-65                      svk_cosine_score_pairs
    svk::topk_meanstd                    compute_topk_mean_std.py:14-21             svk_topk_meanstd
    svk::snorm_apply                     adaptive_snorm.py:28-38                    svk_snorm_apply

The network-level op keeps the engine (a Python object owning workspaces and streams) out of the schema: it is addressed
by an integer handle, and the activations a forward saves for its backward by an integer token returned as a tensor.
"""
import weakref
from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.library import custom_op, register_autograd, register_fake

from . import lib
from .lib import call

_ENGINES = weakref.WeakValueDictionary()
_NEXT = [1]


def _st():
    return torch.cuda.current_stream().cuda_stream


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise lib.SvkError("svk ops run on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)


def engine_handle(engine):
    h = getattr(engine, "_handle", None)
    if h is None:
        h = engine._handle = _NEXT[0]
        _NEXT[0] += 1
        _ENGINES[h] = engine
    return h


def _engine(handle):
    eng = _ENGINES.get(handle)
    if eng is None:
        raise lib.SvkError("svk: engine handle %d is dead (its model was deleted)" % handle)
    return eng


# ------------------------------------------------------------------------------------------------ whole network
@custom_op("svk::speaker_net_train", mutates_args=(), device_types="cuda")
def _speaker_net_train(x: Tensor, y: Optional[Tensor], anchor: Tensor, handle: int, save: bool) -> Tuple[Tensor, Tensor]:
    """(B, F, T) fp32 [, (B,) int64] -> (logits (B, C) fp32, token).  Functional in its tensor arguments; the parameters
    (views of the engine's flat buffer) and the BatchNorm running statistics it updates are engine-owned state addressed by
    `handle`, not arguments — torch only lets functional operators carry an autograd formula.  `anchor` is a 1-element
    tensor that requires grad, created by the caller on the current stream: it is what makes autograd record the node (the
    backward writes parameter gradients in place and returns none).  Listing the parameters themselves would route the
    backward through their AccumulateGrad nodes, whose streams predate a CUDA-graph capture and invalidate it
    (tests/diag_graph.py).  save=False (no gradient will be asked for) keeps nothing for a backward."""
    eng = _engine(handle)
    logits, sv = eng.forward_train(x, y, save=save)
    tok = eng._next_token = getattr(eng, "_next_token", 0) + 1
    if save:
        eng._pending = getattr(eng, "_pending", {})
        eng._pending[tok] = sv
    return logits, torch.tensor(tok, dtype=torch.int64)


@register_fake("svk::speaker_net_train")
def _(x, y, anchor, handle, save):
    C = _engine(handle).model.last.weight.shape[0]
    return x.new_empty((x.shape[0], C), dtype=torch.float32), torch.empty((), dtype=torch.int64)


@custom_op("svk::speaker_net_train_backward", mutates_args={"flat_grads"}, device_types="cuda")
def _speaker_net_train_backward(dlogits: Tensor, flat_grads: Tensor, handle: int, token: int) -> None:
    """Writes every parameter gradient into `flat_grads` (the engine's flat buffer; each p.grad is a view of it)."""
    eng = _engine(handle)
    sv = getattr(eng, "_pending", {}).pop(token, None)
    eng.backward_train(dlogits, sv)


class _PendingGuard(object):
    """Drops the saved state of a forward whose autograd graph is freed without a backward."""
    __slots__ = ("handle", "token")

    def __init__(self, handle, token):
        self.handle, self.token = handle, token

    def __del__(self):
        eng = _ENGINES.get(self.handle)
        if eng is not None:
            getattr(eng, "_pending", {}).pop(self.token, None)


def _train_setup(ctx, inputs, output):
    x, y, anchor, handle, save = inputs
    ctx.handle = handle
    ctx.token = int(output[1])
    ctx.guard = _PendingGuard(handle, ctx.token)


def _train_backward(ctx, dlogits, dtoken):
    eng = _engine(ctx.handle)
    torch.ops.svk.speaker_net_train_backward(dlogits.contiguous(), eng.flat_grads, ctx.handle, ctx.token)
    # parameter gradients were written in place (p.grad = views of the flat gradient buffer): nothing flows through autograd
    return None, None, None, None, None


register_autograd("svk::speaker_net_train", _train_backward, setup_context=_train_setup)


def _anchor(x, save):
    return torch.empty(1, dtype=torch.float32, device=x.device, requires_grad=save)


def speaker_net_train(engine, x, y):
    engine.ensure_device()
    save = torch.is_grad_enabled() and any(p.requires_grad for p in engine._params)
    logits, _ = torch.ops.svk.speaker_net_train(x, y, _anchor(x, save), engine_handle(engine), save)
    return logits


# ---- the same network with the cross-entropy INSIDE the head (fused AAM-softmax-CE, csrc/aam_fused.cu): returns the mean loss,
# the margin logits and the target ranks; its backward starts from d loss (a scalar), so d_logits never exists in memory.
@custom_op("svk::speaker_net_train_loss", mutates_args=(), device_types="cuda")
def _speaker_net_train_loss(x: Tensor, y: Tensor, anchor: Tensor, handle: int, save: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(B, F, T) fp32, (B,) int64 -> (mean cross-entropy (), logits (B, C), rank (B,) int32, token).  AAM heads only."""
    eng = _engine(handle)
    logits, sv = eng.forward_train(x, y, save=save)
    head = eng.last_head
    if head is None:
        raise lib.SvkError("svk::speaker_net_train_loss needs the fused AAM head (loss 'AAM' / 'AAM-v1', 256-d embeddings)")
    tok = eng._next_token = getattr(eng, "_next_token", 0) + 1
    if save:
        eng._pending = getattr(eng, "_pending", {})
        eng._pending[tok] = sv
    return head["loss"], logits, head["rank"], torch.tensor(tok, dtype=torch.int64)


@register_fake("svk::speaker_net_train_loss")
def _(x, y, anchor, handle, save):
    C = _engine(handle).model.last.weight.shape[0]
    B = x.shape[0]
    return (x.new_empty((), dtype=torch.float32), x.new_empty((B, C), dtype=torch.float32),
            x.new_empty((B,), dtype=torch.int32), torch.empty((), dtype=torch.int64))


@custom_op("svk::speaker_net_train_loss_backward", mutates_args={"flat_grads"}, device_types="cuda")
def _speaker_net_train_loss_backward(dloss: Tensor, flat_grads: Tensor, handle: int, token: int) -> None:
    eng = _engine(handle)
    sv = getattr(eng, "_pending", {}).pop(token, None)
    eng.backward_train(None, sv, dloss=dloss)


def _train_loss_setup(ctx, inputs, output):
    x, y, anchor, handle, save = inputs
    ctx.handle = handle
    ctx.token = int(output[3])
    ctx.guard = _PendingGuard(handle, ctx.token)
    ctx.set_materialize_grads(False)


def _train_loss_backward(ctx, dloss, dlogits, drank, dtoken):
    if dlogits is not None:
        raise lib.SvkError("svk::speaker_net_train_loss: a gradient arrived through the logits; use model(x, y) + "
                           "CrossEntropyLoss when the logits feed another differentiable term")
    if dloss is not None:
        eng = _engine(ctx.handle)
        torch.ops.svk.speaker_net_train_loss_backward(dloss, eng.flat_grads, ctx.handle, ctx.token)
    return None, None, None, None, None


register_autograd("svk::speaker_net_train_loss", _train_loss_backward, setup_context=_train_loss_setup)


def speaker_net_train_loss(engine, x, y):
    """-> (loss, logits, rank); loss carries the autograd node."""
    engine.ensure_device()
    save = torch.is_grad_enabled() and any(p.requires_grad for p in engine._params)
    loss, logits, rank, _ = torch.ops.svk.speaker_net_train_loss(x, y, _anchor(x, save), engine_handle(engine), save)
    return loss, logits, rank


@custom_op("svk::speaker_net_embed", mutates_args=(), device_types="cuda")
def _speaker_net_embed(x: Tensor, lengths: Optional[Tensor], handle: int) -> Tensor:
    """(B, F, T) fp32 [, (B,) int32 valid frame counts] -> (B, 256) embeddings, eval-mode BatchNorm (model.py:402-409)."""
    return _engine(handle).forward_eval(x, lengths=lengths)


@register_fake("svk::speaker_net_embed")
def _(x, lengths, handle):
    return x.new_empty((x.shape[0], _engine(handle).model.fc1.weight.shape[0]), dtype=torch.float32)


def speaker_net_embed(engine, x, lengths=None):
    engine.ensure_device()
    return torch.ops.svk.speaker_net_embed(x, lengths, engine_handle(engine))


# ------------------------------------------------------------------------------------------------ loss
@custom_op("svk::cross_entropy", mutates_args=(), device_types="cuda")
def _cross_entropy(logits: Tensor, target: Tensor) -> Tuple[Tensor, Tensor]:
    """mean over rows of logsumexp(logits) - logits[target]; also returns the per-row logsumexp for the backward."""
    _cuda(logits, target)
    if logits.dtype != torch.float32 or logits.dim() != 2 or not logits.is_contiguous():
        raise lib.SvkError("svk::cross_entropy expects contiguous (B, C) fp32 logits")
    B, C = logits.shape
    loss_rows = torch.empty(B, dtype=torch.float32, device=logits.device)
    lse = torch.empty(B, dtype=torch.float32, device=logits.device)
    loss = torch.zeros((), dtype=torch.float32, device=logits.device)
    call.svk_ce_fwd(logits.data_ptr(), target.data_ptr(), loss_rows.data_ptr(), lse.data_ptr(), 0, loss.data_ptr(), B, C, _st())
    return loss, lse


@register_fake("svk::cross_entropy")
def _(logits, target):
    return logits.new_empty(()), logits.new_empty((logits.shape[0],))


@custom_op("svk::cross_entropy_backward", mutates_args=(), device_types="cuda")
def _cross_entropy_backward(logits: Tensor, target: Tensor, lse: Tensor, gout: Tensor) -> Tensor:
    B, C = logits.shape
    g = torch.empty_like(logits)
    call.svk_ce_bwd(logits.data_ptr(), target.data_ptr(), lse.data_ptr(), gout.data_ptr(), 1.0 / B, g.data_ptr(), B, C, _st())
    return g


@register_fake("svk::cross_entropy_backward")
def _(logits, target, lse, gout):
    return torch.empty_like(logits)


def _ce_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1], output[1])


def _ce_backward(ctx, gloss, glse):
    logits, target, lse = ctx.saved_tensors
    return torch.ops.svk.cross_entropy_backward(logits, target, lse, gloss.contiguous().float()), None


register_autograd("svk::cross_entropy", _ce_backward, setup_context=_ce_setup)


@custom_op("svk::target_rank", mutates_args=(), device_types="cuda")
def _target_rank(logits: Tensor, target: Tensor) -> Tensor:
    """rank[b] = number of classes scoring strictly above the target class (0 = top-1 hit)."""
    _cuda(logits, target)
    B, C = logits.shape
    loss_rows = torch.empty(B, dtype=torch.float32, device=logits.device)
    lse = torch.empty(B, dtype=torch.float32, device=logits.device)
    rank = torch.empty(B, dtype=torch.int32, device=logits.device)
    call.svk_ce_fwd(logits.data_ptr(), target.data_ptr(), loss_rows.data_ptr(), lse.data_ptr(), rank.data_ptr(), 0, B, C, _st())
    return rank


@register_fake("svk::target_rank")
def _(logits, target):
    return logits.new_empty((logits.shape[0],), dtype=torch.int32)


# ------------------------------------------------------------------------------------------------ optimizer
@custom_op("svk::sgd_step", mutates_args={"param", "momentum_buf"}, device_types="cuda")
def _sgd_step(param: Tensor, grad: Tensor, momentum_buf: Tensor, lr: float, momentum: float, weight_decay: float,
              grad_scale: float) -> None:
    """d = grad_scale * g + wd * p;  buf = momentum * buf + d;  p -= lr * buf   (torch.optim.SGD, no nesterov/dampening)."""
    _cuda(param, grad, momentum_buf)
    call.svk_sgd_step(param.data_ptr(), grad.data_ptr(), momentum_buf.data_ptr(), param.numel(), lr, momentum, weight_decay,
                      grad_scale, _st())


# ------------------------------------------------------------------------------------------------ convolution
def _conv_desc(x, w, stride):
    N, H, W, ci = x.shape
    co, ci2, r, r2 = w.shape
    if ci != ci2 or r != r2 or r not in (1, 3) or stride not in (1, 2):
        raise lib.SvkError("svk::conv2d: x (N,H,W,Cin) / w (Cout,Cin,R,R) mismatch or unsupported R/stride")
    code = lib.BF16 if x.dtype == torch.bfloat16 else lib.F32
    impl = lib.IMPL_TCGEN05 if code == lib.BF16 else lib.IMPL_SIMT
    return lib.make_conv_desc(N, H, W, ci, co, r, stride, code, impl), code


def _pack(w, code):
    co, ci, r, _ = w.shape
    dt = torch.bfloat16 if code == lib.BF16 else torch.float32
    wf = torch.empty(r * r * co * ci, dtype=dt, device=w.device)
    wd = torch.empty_like(wf)
    call.svk_pack_conv_weight(w.data_ptr(), wf.data_ptr(), wd.data_ptr(), co, ci, r, code, _st())
    return wf, wd


@custom_op("svk::conv2d", mutates_args=(), device_types="cuda")
def _conv2d(x: Tensor, weight: Tensor, stride: int) -> Tensor:
    """NHWC bf16 (tcgen05) or fp32 (validation) activations, OIHW fp32 master weights, padding R // 2, no bias."""
    _cuda(x, weight)
    x, weight = x.contiguous(), weight.contiguous().float()
    d, code = _conv_desc(x, weight, stride)
    wf, _ = _pack(weight, code)
    y = torch.empty((d.N, d.Ho, d.Wo, d.Cout), dtype=x.dtype, device=x.device)
    call.svk_conv2d_fwd(d, x.data_ptr(), wf.data_ptr(), y.data_ptr(), 0, 0, 0, 0, 0, 0, _st())
    return y


@register_fake("svk::conv2d")
def _(x, weight, stride):
    N, H, W, _ = x.shape
    return x.new_empty((N, (H - 1) // stride + 1, (W - 1) // stride + 1, weight.shape[0]))


@custom_op("svk::conv2d_backward", mutates_args=(), device_types="cuda")
def _conv2d_backward(dy: Tensor, x: Tensor, weight: Tensor, stride: int) -> Tuple[Tensor, Tensor]:
    dy, x, weight = dy.contiguous(), x.contiguous(), weight.contiguous().float()
    d, code = _conv_desc(x, weight, stride)
    _, wd = _pack(weight, code)
    dx = torch.zeros_like(x) if (d.R == 1 and stride == 2) else torch.empty_like(x)
    # the 1x1/s2 data gradient ACCUMULATES into an existing block-input gradient (res aliases dx)
    call.svk_conv2d_dgrad(d, dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), dx.data_ptr() if (d.R == 1 and stride == 2) else 0,
                          0, 0, _st())
    need = lib.load().svk_conv2d_wgrad_workspace_bytes(d)
    ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=x.device)
    dw = torch.empty_like(weight)
    call.svk_conv2d_wgrad(d, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel() * 4, _st())
    return dx, dw


@register_fake("svk::conv2d_backward")
def _(dy, x, weight, stride):
    return torch.empty_like(x), torch.empty_like(weight, dtype=torch.float32)


def _conv_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])
    ctx.stride = inputs[2]


def _conv_backward(ctx, dy):
    x, w = ctx.saved_tensors
    dx, dw = torch.ops.svk.conv2d_backward(dy, x, w, ctx.stride)
    return dx, dw, None


register_autograd("svk::conv2d", _conv_backward, setup_context=_conv_setup)


# ------------------------------------------------------------------------------------------------ scoring
@custom_op("svk::cosine_score_pairs", mutates_args=(), device_types="cuda")
def _cosine_score_pairs(enroll: Tensor, test: Tensor, mean: Optional[Tensor], idx_enroll: Tensor, idx_test: Tensor) -> Tensor:
    _cuda(enroll, test, mean, idx_enroll, idx_test)
    n, D = idx_enroll.numel(), enroll.shape[1]
    out = torch.empty(n, dtype=torch.float32, device=enroll.device)
    if n:
        call.svk_cosine_score_pairs(enroll.data_ptr(), test.data_ptr(), 0 if mean is None else mean.data_ptr(),
                                    idx_enroll.data_ptr(), idx_test.data_ptr(), out.data_ptr(), n, D, _st())
    return out


@register_fake("svk::cosine_score_pairs")
def _(enroll, test, mean, idx_enroll, idx_test):
    return enroll.new_empty((idx_enroll.numel(),))


@custom_op("svk::topk_meanstd", mutates_args=(), device_types="cuda")
def _topk_meanstd(scores: Tensor, topk: int) -> Tuple[Tensor, Tensor]:
    """Per row of `scores` (rows, n): mean and unbiased std of its `topk` largest entries (radix select, fp64 sums)."""
    _cuda(scores)
    rows, n = scores.shape
    mean = torch.empty(rows, dtype=torch.float32, device=scores.device)
    std = torch.empty(rows, dtype=torch.float32, device=scores.device)
    call.svk_topk_meanstd(scores.data_ptr(), rows, n, topk, mean.data_ptr(), std.data_ptr(), _st())
    return mean, std


@register_fake("svk::topk_meanstd")
def _(scores, topk):
    return scores.new_empty((scores.shape[0],)), scores.new_empty((scores.shape[0],))


@custom_op("svk::snorm_apply", mutates_args=(), device_types="cuda")
def _snorm_apply(scores: Tensor, idx_enroll: Tensor, idx_test: Tensor, mean_e: Tensor, std_e: Tensor, mean_t: Tensor,
                 std_t: Tensor) -> Tensor:
    _cuda(scores, idx_enroll, idx_test, mean_e, std_e, mean_t, std_t)
    out = torch.empty_like(scores)
    if scores.numel():
        call.svk_snorm_apply(scores.data_ptr(), idx_enroll.data_ptr(), idx_test.data_ptr(), mean_e.data_ptr(), std_e.data_ptr(),
                             mean_t.data_ptr(), std_t.data_ptr(), out.data_ptr(), scores.numel(), _st())
    return out


@register_fake("svk::snorm_apply")
def _(scores, idx_enroll, idx_test, mean_e, std_e, mean_t, std_t):
    return torch.empty_like(scores)


REGISTERED = ("speaker_net_train", "speaker_net_train_backward", "speaker_net_train_loss", "speaker_net_train_loss_backward",
              "speaker_net_embed", "cross_entropy", "cross_entropy_backward",
              "target_rank", "sgd_step", "conv2d", "conv2d_backward", "cosine_score_pairs", "topk_meanstd", "snorm_apply")
