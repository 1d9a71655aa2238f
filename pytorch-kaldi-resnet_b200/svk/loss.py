"""Cross-entropy (mean reduction) and top-k accuracy on the logits, as libsvk kernels behind torch custom ops.
replaces: nn.CrossEntropyLoss (train_resnet.py:201, :317) and accuracy() (accuracy.py:4-16)."""
import torch
import torch.nn as nn

from . import ops  # noqa: F401  (registers torch.ops.svk.*)


def _check(logits):
    if not logits.is_cuda or logits.dtype != torch.float32 or logits.dim() != 2:
        raise RuntimeError("svk CrossEntropyLoss expects (B, C) fp32 CUDA logits")


class CrossEntropyLoss(nn.Module):
    def forward(self, logits, target):
        _check(logits)
        loss, _ = torch.ops.svk.cross_entropy(logits.contiguous(), target.contiguous().long())
        return loss


def target_rank(logits, target):
    """rank[b] = number of classes scoring strictly above the target class (0 = top-1 hit)."""
    rank = getattr(logits, "svk_rank", None)        # already computed by the fused AAM head (model.forward_loss)
    if rank is not None:
        return rank
    _check(logits)
    return torch.ops.svk.target_rank(logits.detach().contiguous(), target.contiguous().long())


def accuracy(output, target, topk=(1,)):
    """precision@k in percent, one 1-element tensor per k (same return shape as accuracy.py:4-16)."""
    rank = target_rank(output, target)
    B = target.size(0)
    res = []
    for k in topk:
        res.append((rank < k).float().sum(0, keepdim=True).mul_(100.0 / B))
    return res
