"""Cross-entropy (mean reduction) and top-k accuracy on the logits, as libsvk kernels.
replaces: nn.CrossEntropyLoss (train_resnet.py:201, :317) and accuracy() (accuracy.py:4-16)."""
import torch
import torch.nn as nn

from .lib import call


def _st():
    return torch.cuda.current_stream().cuda_stream


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        if not logits.is_cuda or logits.dtype != torch.float32 or logits.dim() != 2:
            raise RuntimeError("svk CrossEntropyLoss expects (B, C) fp32 CUDA logits")
        logits = logits.contiguous()
        target = target.contiguous().long()
        B, C = logits.shape
        loss_rows = torch.empty(B, dtype=torch.float32, device=logits.device)
        lse = torch.empty(B, dtype=torch.float32, device=logits.device)
        loss = torch.zeros((), dtype=torch.float32, device=logits.device)
        call.svk_ce_fwd(logits.data_ptr(), target.data_ptr(), loss_rows.data_ptr(), lse.data_ptr(), 0,
                        loss.data_ptr(), B, C, _st())
        ctx.save_for_backward(logits, target, lse)
        return loss

    @staticmethod
    def backward(ctx, gout):
        logits, target, lse = ctx.saved_tensors
        B, C = logits.shape
        gout = gout.contiguous().float()
        g = torch.empty_like(logits)
        call.svk_ce_bwd(logits.data_ptr(), target.data_ptr(), lse.data_ptr(), gout.data_ptr(), 1.0 / B, g.data_ptr(),
                        B, C, _st())
        return g, None


class CrossEntropyLoss(nn.Module):
    def forward(self, logits, target):
        return _CEFn.apply(logits, target)


def target_rank(logits, target):
    """rank[b] = number of classes scoring strictly above the target class (0 = top-1 hit)."""
    logits = logits.detach().contiguous()
    target = target.contiguous().long()
    B, C = logits.shape
    loss_rows = torch.empty(B, dtype=torch.float32, device=logits.device)
    lse = torch.empty(B, dtype=torch.float32, device=logits.device)
    rank = torch.empty(B, dtype=torch.int32, device=logits.device)
    call.svk_ce_fwd(logits.data_ptr(), target.data_ptr(), loss_rows.data_ptr(), lse.data_ptr(), rank.data_ptr(), 0,
                    B, C, _st())
    return rank


def accuracy(output, target, topk=(1,)):
    """precision@k in percent, one 1-element tensor per k (same return shape as accuracy.py:4-16)."""
    rank = target_rank(output, target)
    B = target.size(0)
    res = []
    for k in topk:
        res.append((rank < k).float().sum(0, keepdim=True).mul_(100.0 / B))
    return res
