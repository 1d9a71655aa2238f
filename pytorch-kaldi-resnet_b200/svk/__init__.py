"""svk — host side of the B200-native speaker-embedding hot path.

    svk.lib       ctypes binding of libsvk.so (include/svk.h)
    svk.engine    forward/backward plan executor behind scripts/model.py
    svk.optim     SGD on the flat parameter buffer          (torch.optim.SGD semantics, train_resnet.py:203)
    svk.loss      CrossEntropyLoss + top-k accuracy         (train_resnet.py:201, accuracy.py)
    svk.parallel  data-parallel wrapper: bucketed NCCL all-reduce overlapped with backward (train_resnet.py:185)
    svk.scoring   cosine trial scoring, cohort top-k statistics, adaptive s-norm
"""
from . import lib  # noqa: F401
from .lib import SvkError, launch_count  # noqa: F401
