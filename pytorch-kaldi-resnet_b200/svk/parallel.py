"""Data parallelism for the training path and sharding helpers for extraction / scoring.

replaces: torch.nn.parallel.DistributedDataParallel as used at train_resnet.py:185 (one process per GPU, NCCL):
  C2  one flat broadcast of parameters (+ BatchNorm buffers) from rank 0 at wrap time;
  C3  gradient all-reduce-MEAN in three buckets, issued from inside the engine's backward as soon as a bucket's
      gradients are final ({AAM head, fc1} -> {layer4} -> {layer3 .. stem}), on a side stream so the transfer
      overlaps the remaining backward kernels; the flat gradient buffer makes every bucket one contiguous message;
  C4  the per-iteration BatchNorm-buffer broadcast is dropped (statistics stay per rank — no SyncBatchNorm, as in the
      reference — and rank 0's running statistics are the ones checkpointed, train_resnet.py:281-289).
Extraction and scoring shard independent units (utterances, trial blocks) with no collective.
"""
import torch
import torch.distributed as dist
import torch.nn as nn


class GradBucketer(object):
    """All-reduce-mean of contiguous ranges of one flat gradient tensor.  Device-agnostic (gloo on CPU in the tests,
    NCCL on GPU); on CUDA each bucket is reduced on `comm_stream` after an event recorded on the compute stream."""

    def __init__(self, flat_grads, ranges, process_group=None):
        self.flat = flat_grads
        self.ranges = dict(ranges)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.cuda = flat_grads.is_cuda
        self.comm_stream = torch.cuda.Stream(device=flat_grads.device) if self.cuda else None
        self.use_avg = self.cuda and dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        self.pending = 0

    def reduce_bucket(self, idx, also_wait=None):
        """`also_wait`: an event of another stream that wrote part of the bucket (the engine's weight-gradient stream)."""
        if self.world == 1:
            return
        lo, hi = self.ranges[idx]
        if hi <= lo:
            return
        view = self.flat[lo:hi]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                if also_wait is not None:
                    self.comm_stream.wait_event(also_wait)
                self._allreduce_mean(view)
            self.pending += 1
        else:
            self._allreduce_mean(view)

    def _allreduce_mean(self, view):
        if self.use_avg:
            dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
            view.div_(self.world)

    def finish(self):
        """Make the compute stream wait for every outstanding bucket."""
        if self.cuda and self.pending:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            self.pending = 0


class DistributedDataParallel(nn.Module):
    """Wrapper with the reference's interface: `.module`, forward(*args) -> module(*args), 'module.'-prefixed
    state-dict keys (what train_resnet.py:283-289 saves and model.loadParameters strips)."""

    def __init__(self, module, device_ids=None, process_group=None):
        super(DistributedDataParallel, self).__init__()
        self.module = module
        eng = module.engine
        eng.ensure_device()
        self._eng = eng
        if not dist.is_initialized():
            raise RuntimeError("init_process_group must be called before wrapping the model")
        with torch.no_grad():
            dist.broadcast(eng.flat_params, 0, group=process_group)
            bufs = [b for b in module.buffers() if b.dtype == torch.float32]
            if bufs:
                flat = torch.cat([b.reshape(-1) for b in bufs])
                dist.broadcast(flat, 0, group=process_group)
                o = 0
                for b in bufs:
                    b.copy_(flat[o:o + b.numel()].view_as(b))
                    o += b.numel()
        eng.invalidate()
        self.bucketer = GradBucketer(eng.flat_grads, eng.bucket_ranges, process_group)
        self._last_bucket = max(eng.bucket_ranges.keys())
        eng.grad_ready_cb = self._on_bucket

    def _on_bucket(self, idx):
        self.bucketer.reduce_bucket(idx, getattr(self._eng, "bucket_side_event", None))
        if idx == self._last_bucket:
            self.bucketer.finish()

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def forward_loss(self, x, y):
        return self.module.forward_loss(x, y)


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) slice of n independent units for this rank (trial blocks, embedding rows)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_length(lengths, rank, world):
    """Length-balanced static sharding of utterances: sort by length (descending) and deal round-robin in a
    serpentine order, so every rank gets the same number of utterances (+-1) and nearly the same number of frames.
    No duplicates and no shuffling, unlike DistributedSampler(shuffle=True) at decode.py:170."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    mine = []
    for pos, idx in enumerate(order):
        rnd, slot = divmod(pos, world)
        owner = slot if rnd % 2 == 0 else world - 1 - slot
        if owner == rank:
            mine.append(idx)
    return mine
