"""SGD with momentum and weight decay, torch.optim.SGD semantics (train_resnet.py:203-205, :328):
    d = g + wd * p;  buf = momentum * buf + d  (buf starts at 0, so the first step gives buf = d);  p -= lr * buf.
When every parameter is a view of one SpeakerNetEngine flat buffer the whole update is ONE svk_sgd_step launch;
otherwise one launch per parameter.  state_dict()/load_state_dict() use torch's layout (per-parameter
'momentum_buffer'), so checkpoints interchange with the reference's torch.optim.SGD."""
import torch
from torch.optim import Optimizer

from . import ops  # noqa: F401  (registers torch.ops.svk.*)


class SGD(Optimizer):
    def __init__(self, params, lr=0.1, momentum=0.0, dampening=0, weight_decay=0.0, nesterov=False):
        if dampening != 0 or nesterov:
            raise NotImplementedError("the reference uses plain momentum SGD (no dampening, no nesterov)")
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super(SGD, self).__init__(params, defaults)
        self._flat_buf = None
        self._engine = None
        self.grad_scale = 1.0      # multiplied into the gradient inside the update kernel

    # ---- flat fast path --------------------------------------------------------------------------------------
    def _try_flatten(self):
        if self._flat_buf is not None or len(self.param_groups) != 1:
            return self._flat_buf is not None
        ps = self.param_groups[0]["params"]
        tags = [getattr(p, "_svk_flat", None) for p in ps]
        if not tags or any(t is None for t in tags):
            return False
        eng = tags[0][0]
        if any(t[0] is not eng for t in tags) or len(ps) != len(eng._params) or any(a is not b for a, b in zip(ps, eng._params)):
            return False
        self._engine = eng
        self._flat_buf = torch.zeros_like(eng.flat_params)
        for p, o in zip(ps, eng._offs):
            view = self._flat_buf[o:o + p.numel()].view(p.shape)
            old = self.state[p].get("momentum_buffer")
            if old is not None:
                view.copy_(old)
            self.state[p]["momentum_buffer"] = view
        return True

    def load_state_dict(self, state_dict):
        super(SGD, self).load_state_dict(state_dict)
        if self._flat_buf is not None:          # re-home loaded buffers into the flat buffer
            for p, o in zip(self._engine._params, self._engine._offs):
                view = self._flat_buf[o:o + p.numel()].view(p.shape)
                old = self.state[p].get("momentum_buffer")
                if old is not None and old.data_ptr() != view.data_ptr():
                    view.copy_(old)
                self.state[p]["momentum_buffer"] = view

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._try_flatten():
            eng = self._engine
            g = self.param_groups[0]
            grads_ok = all(p.grad is gv for p, gv in zip(eng._params, eng._grad_views))
            if grads_ok:
                torch.ops.svk.sgd_step(eng.flat_params, eng.flat_grads, self._flat_buf, float(g["lr"]), float(g["momentum"]),
                                       float(g["weight_decay"]), float(self.grad_scale))
                eng.invalidate()
                eng.grads_consumed()
                return loss
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("svk.optim.SGD updates fp32 CUDA parameters only")
                state = self.state[p]
                if "momentum_buffer" not in state:
                    state["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                buf = state["momentum_buffer"]
                grad = p.grad.contiguous()
                torch.ops.svk.sgd_step(p.data, grad, buf, float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]),
                                       float(self.grad_scale))
        if self._engine is not None:
            self._engine.invalidate()
            self._engine.grads_consumed()
        return loss

    def zero_grad(self, set_to_none=True):
        # The first backward after step() / zero_grad() OVERWRITES the flat gradient buffer (a second one before the next
        # step() accumulates into it, like torch): nothing needs clearing on the flat path.
        if self._flat_buf is not None or self._try_flatten():
            self._engine.grads_consumed()
            return
        super(SGD, self).zero_grad(set_to_none=set_to_none)
