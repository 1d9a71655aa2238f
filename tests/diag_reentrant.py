import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-kaldi-resnet_b200"), os.path.join(ROOT, "pytorch-kaldi-resnet_b200", "scripts"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import util
from model import NeuralSpeakerModel
from svk.loss import CrossEntropyLoss
from svk.optim import SGD
torch.manual_seed(21)
with contextlib.redirect_stdout(io.StringIO()):
    m = NeuralSpeakerModel(spk_num=23, feat_dim=40, pooling="mean+std", loss="AAM", precision=sys.argv[1] if len(sys.argv) > 1 else "fp32").cuda()
crit = CrossEntropyLoss(); opt = SGD(m.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
g = torch.Generator().manual_seed(8)
xa, xb = torch.randn(3, 40, 48, generator=g).cuda(), torch.randn(3, 40, 48, generator=g).cuda()
ya, yb = torch.randint(0, 23, (3,), generator=g).cuda(), torch.randint(0, 23, (3,), generator=g).cuda()
m.train()
def grads():
    torch.cuda.synchronize(); return {n: p.grad.detach().clone() for n, p in m.named_parameters()}
def cmp(tag, a, b):
    errs = sorted(((util.rel_err(a[n].cpu(), b[n].cpu()), n) for n in a), reverse=True)
    print(tag, "worst:", ["%.2e %s" % e for e in errs[:4]], "| #params > 1e-6:", sum(1 for e in errs if e[0] > 1e-6), "of", len(errs), flush=True)
opt.zero_grad(); crit(m(xa, ya), ya).backward(); g1 = grads()
opt.zero_grad(); crit(m(xa, ya), ya).backward(); g2 = grads()
cmp("same input twice            ", g2, g1)
opt.zero_grad(); la = crit(m(xa, ya), ya); lb = crit(m(xb, yb), yb); la.backward(); g3 = grads()
cmp("a-fwd, b-fwd, a-bwd         ", g3, g1)
lb.backward()
opt.zero_grad(); la = crit(m(xa, ya), ya); m.predict(xa); la.backward(); g4 = grads()
cmp("a-fwd, predict(train), a-bwd", g4, g1)
opt.zero_grad(); la = crit(m(xa, ya), ya); la.backward(); g5 = grads()
cmp("plain again                 ", g5, g1)
