"""Host <-> device plumbing of the training loop (pinned-buffer batching, async H2D).

replaces: `audios.cuda(args.gpu, non_blocking=True)` / `losses.update(loss.item(), ...)` in train_resnet.py:310-320, whose
copies and reads sit on the compute stream (the copy of a batch delays its own step, `.item()` drains the launch queue
every step).  Nothing here touches the numbers: same tensors, same order.
"""
import torch


class DevicePrefetcher(object):
    """Iterate a loader of (x, y) host batches one batch ahead: while batch i trains, the host->device copies of batch
    i+1 run on a side stream (from pinned memory they are asynchronous DMA transfers) into one of two persistent device
    buffers per tensor shape — no allocator traffic in the loop.  A yielded batch stays valid until the batch after the
    next one is requested (its buffer is then overwritten, after the kernels that read it have finished)."""

    def __init__(self, loader, device, pin=True):
        self.loader = loader
        self.device = torch.device(device)
        self.pin = pin
        self.stream = torch.cuda.Stream(self.device)
        self._bufs = {}
        self._free = [None, None]        # per slot: event recorded on the compute stream when its last reader was queued
        self._k = 0

    def __len__(self):
        return len(self.loader)

    def _buf(self, slot, which, t):
        key = (slot, which, tuple(t.shape), t.dtype)
        b = self._bufs.get(key)
        if b is None:
            b = self._bufs[key] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        return b

    def _stage(self, it):
        try:
            x, y = next(it)
        except StopIteration:
            return None
        if self.pin:
            x = x if x.is_pinned() else x.pin_memory()
            y = y if y.is_pinned() else y.pin_memory()
        slot = self._k & 1
        self._k += 1
        xd, yd = self._buf(slot, 0, x), self._buf(slot, 1, y)
        if self._free[slot] is not None:
            self.stream.wait_event(self._free[slot])
        with torch.cuda.stream(self.stream):
            xd.copy_(x, non_blocking=True)
            yd.copy_(y, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return xd, yd, ev, slot, (x, y)    # the host tensors stay referenced until the copy has been consumed

    def __iter__(self):
        it = iter(self.loader)
        nxt = self._stage(it)
        while nxt is not None:
            xd, yd, ev, slot, _host = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            nxt = self._stage(it)          # next batch's copies are in flight before this batch's kernels are queued
            yield xd, yd
            done = torch.cuda.Event()      # everything the consumer queued on this batch
            done.record(torch.cuda.current_stream(self.device))
            self._free[slot] = done


class ScalarReader(object):
    """Read a device scalar (the loss) to the host EVERY step without stalling the launch queue: the value is copied to
    a pinned slot asynchronously and collected `depth` steps later (or at flush())."""

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.buf = torch.zeros(depth, dtype=torch.float32).pin_memory()
        self.events = [None] * depth
        self.values = []
        self.n = 0

    def push(self, scalar):
        slot = self.n % len(self.events)
        if self.events[slot] is not None:
            self.events[slot].synchronize()
            self.values.append(float(self.buf[slot]))
        self.buf[slot:slot + 1].copy_(scalar.detach().reshape(1).float(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.events[slot] = ev
        self.n += 1

    def flush(self):
        """-> every value pushed so far, in order (waits for the outstanding copies)."""
        depth = len(self.events)
        start = max(self.n - depth, 0)
        for k in range(start, self.n):
            slot = k % depth
            if self.events[slot] is not None:
                self.events[slot].synchronize()
                self.values.append(float(self.buf[slot]))
                self.events[slot] = None
        out, self.values = self.values, []
        return out
