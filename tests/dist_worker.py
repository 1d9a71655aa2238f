"""Worker of tests/test_multigpu_gpu.py, launched with `python -m torch.distributed.run --nproc-per-node N`.

Every rank builds the drop-in model (rank r seeds its own random init: the wrap-time broadcast must erase the
difference), wraps it in svk.parallel.DistributedDataParallel and runs K training steps on ITS OWN shard of a fixed
batch (train_resnet.py:185, 239-247: local BatchNorm statistics, gradient all-reduce-mean, identical SGD update).  It
dumps what the test compares: the gradients of step 1, the parameters and BatchNorm buffers after K steps."""
import contextlib
import io
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pytorch-kaldi-resnet_b200")
for p in (ROOT, PKG, os.path.join(PKG, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    out_dir, precision, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from model import NeuralSpeakerModel
    from svk.loss import CrossEntropyLoss
    from svk.optim import SGD
    from svk.parallel import DistributedDataParallel
    torch.manual_seed(100 + rank)                        # different init per rank: rank 0's must win
    with contextlib.redirect_stdout(io.StringIO()):
        net = NeuralSpeakerModel(spk_num=37, feat_dim=40, pooling="mean+std", loss="AAM", precision=precision).cuda(local)
    model = DistributedDataParallel(net, device_ids=[local])
    init = {k: v.detach().float().cpu().clone() for k, v in net.state_dict().items()}
    crit = CrossEntropyLoss()
    opt = SGD(model.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(4321)
    per = 4
    X = torch.randn(steps, world * per, 40, 64, generator=g)
    Y = torch.randint(0, 37, (steps, world * per), generator=g)
    model.train()
    dump = {"init": init}
    for s in range(steps):
        x = X[s, rank * per:(rank + 1) * per].cuda()
        y = Y[s, rank * per:(rank + 1) * per].cuda()
        loss = crit(model(x, y), y)
        opt.zero_grad()
        loss.backward()
        if s == 0:
            torch.cuda.synchronize()
            dump["grads_step1"] = {n: p.grad.detach().float().cpu().clone() for n, p in net.named_parameters()}
            dump["loss_step1"] = float(loss)
        opt.step()
    torch.cuda.synchronize()
    dump["final"] = {k: v.detach().float().cpu().clone() for k, v in net.state_dict().items()}
    dump["X"], dump["Y"], dump["per"] = X, Y, per
    torch.save(dump, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
