"""precision@k on the logits — same call and return value as the reference's accuracy.py:4-16, computed by the
svk_ce_fwd kernel (rank of the target logit) instead of torch.topk."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from svk.loss import accuracy  # noqa: E402,F401
