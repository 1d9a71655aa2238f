"""One CUDA graph for the whole fixed-shape training step.

replaces: the Python loop body of train_resnet.py:316-328 (model(x, y), criterion, zero_grad, backward, optimizer.step) as
~250 kernel launches, ~40 event records / stream waits and a few dozen torch calls issued from Python every step.  The step
is GPU-bound, but its launch stream competes with everything else the host does per step (the loader, the host->device
prefetch, the loss read-back): replaying one captured graph takes the host out of the picture — `e2e` equals the
device-resident rate — and removes the inter-kernel gaps that depend on launch latency.

What is captured: forward (side-stream shortcut convolutions included), fused AAM-softmax-cross-entropy, backward (weight
gradients on the side stream, programmatic-dependent-launch edges kept), the bucketed NCCL all-reduce of a
DistributedDataParallel model, and the SGD kernel.  Scalars baked into kernel arguments (learning rate, momentum, weight
decay) are part of the cache key: a new value (CosineAnnealingLR steps once per epoch) captures a new graph.
"""
import torch


class GraphedTrainStep(object):
    """step = GraphedTrainStep(model, optimizer);  loss, logits = step(x, y)

    `model` is a NeuralSpeakerModel or its DistributedDataParallel wrapper in training mode; x (B, F, T) fp32 and y (B,)
    int64 are CUDA tensors (they are copied into the graph's static inputs).  The returned loss / logits are the graph's
    static outputs: valid until the next call."""

    def __init__(self, model, optimizer, warmup=3, max_graphs=4):
        self.model, self.opt = model, optimizer
        self.warmup, self.max_graphs = warmup, max_graphs
        self._graphs = {}
        self._pool = None
        self.launches_per_replay = 0
        self.replays = 0

    def _key(self, x, y):
        g = self.opt.param_groups[0]
        return (tuple(x.shape), tuple(y.shape), float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]),
                float(getattr(self.opt, "grad_scale", 1.0)))

    def _eager(self, x, y):
        loss, logits = self.model.forward_loss(x, y)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss, logits

    def _capture(self, x, y):
        if not self.model.training:
            raise RuntimeError("GraphedTrainStep captures the training step: call model.train() first")
        sx, sy = x.clone(), y.clone()
        net = self.model.module if hasattr(self.model, "module") else self.model
        eng = net.engine
        eng.ensure_device()
        # The warm-up iterations are real training steps (they create the workspaces, function attributes, NCCL
        # communicators and the optimizer's flat buffers OUTSIDE the capture): snapshot what they change — parameters,
        # momentum, BatchNorm buffers — and put it back, so that capturing is invisible to the training run.
        bufs = list(net.buffers())
        had_momentum = getattr(self.opt, "_flat_buf", None) is not None
        saved = (eng.flat_params.clone(), [b.clone() for b in bufs], self.opt._flat_buf.clone() if had_momentum else None)
        s = torch.cuda.Stream(device=x.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(max(1, self.warmup)):
                self._eager(sx, sy)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from . import lib
        g = torch.cuda.CUDAGraph()
        n0 = lib.launch_count()
        with torch.cuda.graph(g, pool=self._pool):
            loss, logits = self._eager(sx, sy)
        self.launches_per_replay = lib.launch_count() - n0      # libsvk kernels inside one replay (the host counter does not see replays)
        if self._pool is None:
            self._pool = g.pool()
        with torch.no_grad():
            eng.flat_params.copy_(saved[0])
            for b, v in zip(bufs, saved[1]):
                b.copy_(v)
            if getattr(self.opt, "_flat_buf", None) is not None:
                if had_momentum:
                    self.opt._flat_buf.copy_(saved[2])
                else:
                    self.opt._flat_buf.zero_()
        eng.invalidate()
        return {"graph": g, "x": sx, "y": sy, "loss": loss.detach(), "logits": logits.detach(),
                "rank": getattr(logits, "svk_rank", None)}

    def __call__(self, x, y):
        key = self._key(x, y)
        e = self._graphs.get(key)
        if e is None:
            if len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            e = self._graphs[key] = self._capture(x, y)
        e["x"].copy_(x, non_blocking=True)
        e["y"].copy_(y, non_blocking=True)
        e["graph"].replay()
        self.replays += 1
        logits = e["logits"]
        if e["rank"] is not None:
            logits.svk_rank = e["rank"]
        return e["loss"], logits
