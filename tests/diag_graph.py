"""Which part of the training step breaks CUDA-graph capture?  (diagnostic)"""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-kaldi-resnet_b200"), os.path.join(ROOT, "pytorch-kaldi-resnet_b200", "scripts")):
    sys.path.insert(0, p)
import torch
from model import NeuralSpeakerModel
from svk.optim import SGD
from svk.lib import call
from svk import lib

torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    m = NeuralSpeakerModel(spk_num=101, feat_dim=40, pooling="mean+std", loss="AAM").cuda()
opt = SGD(m.parameters(), 0.05, momentum=0.9, weight_decay=1e-4)
m.train()
x = torch.randn(8, 40, 64, device="cuda"); y = torch.randint(0, 101, (8,), device="cuda")


def full():
    loss, logits = m.forward_loss(x, y); opt.zero_grad(); loss.backward(); opt.step()


def fwd_only():
    with torch.no_grad():
        m.forward_loss(x, y)


def fwd_bwd():
    loss, logits = m.forward_loss(x, y); loss.backward()


def one_kernel():
    a = torch.zeros(1024, device="cuda"); b = torch.zeros(1024, device="cuda"); buf = torch.zeros(1024, device="cuda")
    torch.ops.svk.sgd_step(a, b, buf, 0.1, 0.9, 0.0, 1.0)


from svk import engine as _E
_orig_bwd = _E.SpeakerNetEngine.backward_train


def _traced_bwd(self, *a, **k):
    try:
        return _orig_bwd(self, *a, **k)
    except Exception:
        import traceback
        print("   exception inside backward_train:\n" + "".join(traceback.format_exc().splitlines(True)[-12:]), flush=True)
        raise


_E.SpeakerNetEngine.backward_train = _traced_bwd
state = {"bad": None, "n": 0}


def trace(name):
    if state["bad"] is None:
        state["n"] += 1
        try:
            ok = torch.cuda.is_current_stream_capturing()
            if not ok:
                state["bad"] = "%s (call #%d): stream no longer capturing" % (name, state["n"])
        except Exception as ex:
            state["bad"] = "%s (call #%d): %s" % (name, state["n"], str(ex).splitlines()[0][:120])


if os.environ.get("DIAG_NOPARAMS"):
    from svk import ops as _ops

    def _loss_noparams(engine, x_, y_):
        engine.ensure_device()
        anchor = torch.zeros(1, device=x_.device, requires_grad=True)
        loss, logits, rank, _ = torch.ops.svk.speaker_net_train_loss(x_, y_, [anchor], _ops.engine_handle(engine), True)
        return loss, logits, rank
    _ops.speaker_net_train_loss = _loss_noparams

for _ in range(3):
    full()
torch.cuda.synchronize()
for mode in ("global",):
    for name, fn in (("one_kernel", one_kernel), ("fwd_only", fwd_only), ("fwd_bwd", fwd_bwd), ("full", full)):
        g = torch.cuda.CUDAGraph()
        try:
            s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
            torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
            state["bad"], state["n"] = None, 0
            lib.TRACE = trace
            try:
                with torch.cuda.graph(g, capture_error_mode=mode):
                    fn()
            finally:
                lib.TRACE = None
                if state["bad"]:
                    print("   first invalidating libsvk call:", state["bad"], flush=True)
            g.replay(); torch.cuda.synchronize()
            print(mode, name, "OK", flush=True)
        except Exception as ex:
            print(mode, name, "FAILED:", str(ex).splitlines()[0][:150], flush=True)
            c = ex.__context__
            while c is not None:
                import traceback
                print("   caused by:", type(c).__name__, str(c).splitlines()[0][:300], flush=True)
                print("".join(traceback.format_tb(c.__traceback__)[-4:]))
                c = c.__context__
            try:
                torch.cuda.synchronize()
            except Exception as ex2:
                print("  sync:", str(ex2)[:100])
