"""Drop-in for the reference's scripts/train_resnet.py (same flags :25-91, same process model — mp.spawn one process
per GPU :128, NCCL rendezvous :148 — same checkpoint dictionary :283-289, same log lines :412-427), with the hot loop
(:307-328) running on the libsvk kernels: NeuralSpeakerModel (model.py) -> svk.loss.CrossEntropyLoss ->
loss.backward() (one engine backward with bucketed NCCL all-reduce overlapped) -> svk.optim.SGD (one kernel).

Deliberate differences, none of which changes the printed numbers: metrics are accumulated on the device and read
back only every --print-freq steps (the reference syncs with loss.item() every step, :321); BatchNorm buffers are
not re-broadcast every iteration (rank 0's are the ones checkpointed).
"""
import argparse
import os
import random
import shutil
import sys
import time
import warnings

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.utils.data
import torch.utils.data.distributed

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
for _p in (_HERE, _PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from datasets import SequenceDataset, SequenceDataset2  # noqa: E402
from model import NeuralSpeakerModel  # noqa: E402
from svk.ckpt import load_checkpoint  # noqa: E402
from svk.data import DevicePrefetcher  # noqa: E402
from svk.loss import CrossEntropyLoss, target_rank  # noqa: E402
from svk.optim import SGD  # noqa: E402
from svk.parallel import DistributedDataParallel  # noqa: E402

# The reference's command line (train_resnet.py:25-91), kept flag for flag: (flags, options) rows, fed to argparse below.
_S, _I, _F = str, int, float
_FLAGS = (
    (("--train-list",), dict(type=_S, help="training scp")),
    (("--cv-list",), dict(type=_S, help="cv scp")),
    (("--utt2spkid",), dict(type=_S, help="utt2spkid")),
    (("--input-dim",), dict(type=_I, required=True, help="input feature dimension")),
    (("--spk-num",), dict(type=_I, required=True, help="number of speakers")),
    (("--pooling",), dict(type=_S, default="mean", help="mean or mean+std")),
    (("--loss-type",), dict(type=_S, default="softmax", help="softmax, AAM or AAM-v1")),
    (("--margin",), dict(type=_F, default=0.2, help="margin for AAM")),
    (("--scale",), dict(type=_F, default=30, help="scale for AAM")),
    (("--dataset",), dict(type=_S, default="v1", help="v1 or v2")),
    (("--min-chunk-size",), dict(type=_I, default=200, help="minimum feature map length (ignored, as in the reference)")),
    (("--max-chunk-size",), dict(type=_I, default=400, help="chunk length in frames")),
    (("--log-dir",), dict(type=_S, required=True, help="logging directory")),
    (("-a", "--arch"), dict(metavar="ARCH", default="resnet18", help="recorded in the checkpoint only")),
    (("-j", "--workers"), dict(type=_I, default=2, metavar="N", help="number of data loading workers")),
    (("--epochs",), dict(type=_I, default=10, metavar="N")),
    (("--start-epoch",), dict(type=_I, default=0, metavar="N")),
    (("-b", "--batch-size"), dict(type=_I, default=128, metavar="N", help="total batch size of all GPUs on the node")),
    (("--lr", "--learning-rate"), dict(type=_F, default=0.1, metavar="LR", dest="lr")),
    (("--lr-final", "--final-learning-rate"), dict(type=_F, default=0.0001, metavar="LR", dest="lr_final")),
    (("--momentum",), dict(type=_F, default=0.9, metavar="M")),
    (("--wd", "--weight-decay"), dict(type=_F, default=1e-4, metavar="W", dest="weight_decay")),
    (("-p", "--print-freq"), dict(type=_I, default=10, metavar="N")),
    (("--resume",), dict(type=_S, default="", metavar="PATH")),
    (("-e", "--evaluate"), dict(dest="evaluate", action="store_true")),
    (("--pretrained",), dict(dest="pretrained", type=_S, help="use pre-trained model")),
    (("--world-size",), dict(type=_I, default=-1, help="number of nodes")),
    (("--rank",), dict(type=_I, default=-1, help="node rank")),
    (("--dist-url",), dict(type=_S, default="tcp://224.66.41.62:23456")),
    (("--dist-backend",), dict(type=_S, default="nccl")),
    (("--seed",), dict(type=_I, default=None)),
    (("--gpu",), dict(type=_I, default=None, help="GPU id to use.")),
    (("--gpu-num",), dict(type=_I, default=-1, help="GPU nums to use.")),
    (("--multiprocessing-distributed",), dict(action="store_true")),
    # beyond the reference: numerics mode of the kernels
    (("--precision",), dict(default="bf16", choices=["bf16", "fp32"], help="activation storage (fp32 = validation mode)")),
    (("--cuda-graph",), dict(action="store_true", help="replay one captured CUDA graph per full-size training step")),
)
parser = argparse.ArgumentParser(description="B200-native ResNet speaker-embedding training")
for _names, _opts in _FLAGS:
    parser.add_argument(*_names, **_opts)

best_acc1 = 0


def main():
    args = parser.parse_args()
    if args.seed is not None:
        random.seed(args.seed)
        torch.manual_seed(args.seed)
        warnings.warn('You have chosen to seed training; data-loader crops are still drawn from numpy per worker.')
    if args.gpu is not None:
        warnings.warn('You have chosen a specific GPU. This will completely disable data parallelism.')
    if args.dist_url == "env://" and args.world_size == -1:
        args.world_size = int(os.environ["WORLD_SIZE"])
    args.distributed = args.world_size > 1 or args.multiprocessing_distributed
    ngpus_per_node = torch.cuda.device_count() if args.gpu_num == -1 else min(torch.cuda.device_count(), args.gpu_num)
    if ngpus_per_node == 0:
        raise RuntimeError("train_resnet.py needs at least one CUDA device (there is no CPU path)")
    if args.multiprocessing_distributed:
        args.world_size = ngpus_per_node * args.world_size
        mp.spawn(main_worker, nprocs=ngpus_per_node, args=(ngpus_per_node, args))
    else:
        main_worker(args.gpu, ngpus_per_node, args)


def main_worker(gpu, ngpus_per_node, args):
    global best_acc1
    args.gpu = gpu if gpu is not None else 0
    print("Use GPU: {} for training".format(args.gpu))
    torch.cuda.set_device(args.gpu)
    if args.distributed:
        if args.dist_url == "env://" and args.rank == -1:
            args.rank = int(os.environ["RANK"])
        if args.multiprocessing_distributed:
            args.rank = args.rank * ngpus_per_node + gpu
        dist.init_process_group(backend=args.dist_backend, init_method=args.dist_url, world_size=args.world_size,
                                rank=args.rank)
    print("=> creating model '{}'".format(args.arch))
    model = NeuralSpeakerModel(spk_num=args.spk_num, feat_dim=args.input_dim, pooling=args.pooling,
                               loss=args.loss_type, m=args.margin, s=args.scale, precision=args.precision)
    print('===> Model total parameter: {}'.format(sum(p.numel() for p in model.parameters() if p.requires_grad)))
    loc = 'cuda:{}'.format(args.gpu)
    if args.pretrained:
        if os.path.isfile(args.pretrained):
            print("=> using pre-trained model '{}'".format(args.pretrained))
            checkpoint = load_checkpoint(args.pretrained, map_location=loc)
            model.loadParameters(checkpoint['state_dict'])
        else:
            print("=> no pre-trained model found at '{}'".format(args.pretrained))
            return
    model.cuda(args.gpu)
    if args.distributed:
        args.batch_size = int(args.batch_size / ngpus_per_node)
        args.workers = int((args.workers + ngpus_per_node - 1) / ngpus_per_node)
        model = DistributedDataParallel(model, device_ids=[args.gpu])
    print("gpu: {}, batch size: {}, args.workers:{}, ngpus_per_node: {}".format(gpu, args.batch_size, args.workers,
                                                                               ngpus_per_node))
    # Without --gpu and without --multiprocessing-distributed the reference wraps the model in nn.DataParallel over all
    # visible GPUs (train_resnet.py:190-199) and therefore checkpoints 'module.'-prefixed keys.  This drop-in trains that
    # mode on ONE GPU (data parallelism here means one process per GPU: --multiprocessing-distributed), but keeps the key
    # prefix so the checkpoint still --resume's in the reference.
    key_prefix = ''
    if not args.distributed and gpu is None:
        key_prefix = 'module.'
        if ngpus_per_node > 1:
            warnings.warn('no --gpu / --multiprocessing-distributed: training on cuda:0 only (the reference would use '
                          'nn.DataParallel over %d GPUs); pass --multiprocessing-distributed for one process per GPU'
                          % ngpus_per_node)
    criterion = CrossEntropyLoss().cuda(args.gpu)
    optimizer = SGD(model.parameters(), args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, args.epochs, eta_min=args.lr_final, last_epoch=-1)

    if args.resume:
        if os.path.isfile(args.resume):
            print("=> loading checkpoint '{}'".format(args.resume))
            checkpoint = load_checkpoint(args.resume, map_location=loc)
            args.start_epoch = checkpoint['epoch']
            best_acc1 = checkpoint['best_acc1']
            state = checkpoint['state_dict']
            if args.distributed and not any(k.startswith('module.') for k in state):
                state = {'module.' + k: v for k, v in state.items()}
            if not args.distributed:
                state = {k[len('module.'):] if k.startswith('module.') else k: v for k, v in state.items()}
            model.load_state_dict(state)
            optimizer.load_state_dict(checkpoint['optimizer'])
            # the reference rebuilds the schedule with a hard-coded eta_min here (train_resnet.py:225)
            scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, args.epochs, eta_min=0.0001,
                                                                   last_epoch=args.start_epoch - 1)
            print("=> loaded checkpoint '{}' (epoch {})".format(args.resume, checkpoint['epoch']))
        else:
            print("=> no checkpoint found at '{}'".format(args.resume))

    if args.dataset == "v2":
        train_dataset = SequenceDataset2(scp_file=args.train_list, utt2spkid_file=args.utt2spkid, chunk_size=args.max_chunk_size)
    else:
        train_dataset = SequenceDataset(scp_file=args.train_list, utt2spkid_file=args.utt2spkid, chunk_size=[args.max_chunk_size])
    train_sampler = None
    if args.distributed:
        train_sampler = torch.utils.data.distributed.DistributedSampler(train_dataset, num_replicas=args.world_size,
                                                                        rank=args.rank, shuffle=True)
    train_loader = torch.utils.data.DataLoader(train_dataset, batch_size=args.batch_size, shuffle=(train_sampler is None),
                                               num_workers=args.workers, pin_memory=True, sampler=train_sampler)
    print("=> args.world_size: {}, args.rank: {}, args.batch_size: {}, train_loader samples: {}".format(
        args.world_size, args.rank, args.batch_size, len(train_loader)))
    if args.dataset == "v2":
        val = SequenceDataset2(scp_file=args.cv_list, utt2spkid_file=args.utt2spkid, chunk_size=args.max_chunk_size)
    else:
        val = SequenceDataset(scp_file=args.cv_list, utt2spkid_file=args.utt2spkid, chunk_size=[args.max_chunk_size])
    val_loader = torch.utils.data.DataLoader(val, batch_size=args.batch_size, shuffle=False, num_workers=args.workers,
                                             pin_memory=True)
    if args.evaluate:
        validate(val_loader, model, criterion, args)
        return
    os.makedirs(args.log_dir, exist_ok=True)
    for epoch in range(args.start_epoch, args.epochs):
        if args.distributed:
            train_sampler.set_epoch(epoch)
        train(train_loader, model, criterion, optimizer, epoch, args)
        acc1 = validate(val_loader, model, criterion, args)
        scheduler.step()
        is_best = acc1 > best_acc1
        best_acc1 = max(acc1, best_acc1)
        if not args.multiprocessing_distributed or (args.multiprocessing_distributed and args.rank % ngpus_per_node == 0):
            save_checkpoint({
                'epoch': epoch + 1,
                'arch': args.arch,
                'state_dict': {key_prefix + k: v.detach().clone() for k, v in model.state_dict().items()},
                'best_acc1': best_acc1,
                'optimizer': optimizer.state_dict(),
            }, is_best, os.path.join(args.log_dir, 'checkpoint_epoch{}.pth.tar'.format(epoch)))


class DeviceMeters(object):
    """loss / top-1 / top-5 sums kept on the device; one host read per display."""

    def __init__(self, device):
        self.acc = torch.zeros(4, dtype=torch.float64, device=device)   # sum loss*n, sum top1 hits, sum top5 hits, n
        self.last = None

    def update(self, loss, rank, n):
        hits1 = (rank < 1).sum()
        hits5 = (rank < 5).sum()
        cur = torch.stack([loss.detach().double() * n, hits1.double(), hits5.double(),
                           torch.tensor(float(n), dtype=torch.float64, device=loss.device)])
        self.acc += cur
        self.last = cur

    def read(self):
        tot, cur = self.acc.tolist(), self.last.tolist()
        n, cn = max(tot[3], 1.0), max(cur[3], 1.0)
        return {"loss": (cur[0] / cn, tot[0] / n), "acc1": (100.0 * cur[1] / cn, 100.0 * tot[1] / n),
                "acc5": (100.0 * cur[2] / cn, 100.0 * tot[2] / n)}


def _display(prefix, i, total, batch_time, data_time, stats):
    digits = len(str(total))
    entries = [prefix + ('[{:' + str(digits) + 'd}/{}]').format(i, total)]
    entries.append('Time {:6.3f} ({:6.3f})'.format(*batch_time))
    if data_time is not None:
        entries.append('Data {:6.3f} ({:6.3f})'.format(*data_time))
    entries.append('Loss {:.4e} ({:.4e})'.format(*stats["loss"]))
    entries.append('Acc@1 {:6.2f} ({:6.2f})'.format(*stats["acc1"]))
    entries.append('Acc@5 {:6.2f} ({:6.2f})'.format(*stats["acc5"]))
    print('\t'.join(entries))
    sys.stdout.flush()


def train(train_loader, model, criterion, optimizer, epoch, args):
    model.train()
    meters = DeviceMeters(torch.device('cuda', args.gpu))
    bt_sum = dt_sum = 0.0
    end = time.time()
    # batches arrive on the device one step ahead (svk.data.DevicePrefetcher; train_resnet.py:310-311 copied them on the
    # compute stream)
    graphed = None
    if getattr(args, "cuda_graph", False):
        from svk.graph import GraphedTrainStep
        graphed = GraphedTrainStep(model, optimizer)
    for i, (audios, target) in enumerate(DevicePrefetcher(train_loader, torch.device('cuda', args.gpu))):
        dt = time.time() - end
        if graphed is not None and audios.size(0) == args.batch_size:
            loss, output = graphed(audios, target)       # the whole step below as one graph replay
            meters.update(loss, target_rank(output, target), audios.size(0))
        else:
            # model(audios, target) + criterion(output, target) (train_resnet.py:316-317) as one call: AAM heads run the
            # fused AAM-softmax-cross-entropy kernels and hand back the target ranks accuracy() needs
            loss, output = model.forward_loss(audios, target)
            meters.update(loss, target_rank(output, target), audios.size(0))
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
        bt = time.time() - end
        end = time.time()
        bt_sum += bt
        dt_sum += dt
        if i % args.print_freq == 0:
            _display("Epoch: [{}]".format(epoch), i, len(train_loader), (bt, bt_sum / (i + 1)), (dt, dt_sum / (i + 1)),
                     meters.read())


def validate(val_loader, model, criterion, args):
    model.eval()
    meters = DeviceMeters(torch.device('cuda', args.gpu))
    bt_sum = 0.0
    with torch.no_grad():
        end = time.time()
        for i, (audios, target) in enumerate(val_loader):
            audios = audios.cuda(args.gpu, non_blocking=True)
            target = target.cuda(args.gpu, non_blocking=True)
            output = model(audios, target)          # the margin is applied in validation too (train_resnet.py:359)
            loss = criterion(output, target)
            meters.update(loss, target_rank(output, target), audios.size(0))
            bt = time.time() - end
            end = time.time()
            bt_sum += bt
            if i % args.print_freq == 0:
                _display('Test: ', i, len(val_loader), (bt, bt_sum / (i + 1)), None, meters.read())
        stats = meters.read() if meters.last is not None else {"acc1": (0, 0), "acc5": (0, 0)}
        print(' * Acc@1 {:.3f} Acc@5 {:.3f}'.format(stats["acc1"][1], stats["acc5"][1]))
    return stats["acc1"][1]


def save_checkpoint(state, is_best, filename='checkpoint.pth.tar'):
    torch.save(state, filename)
    if is_best:
        shutil.copyfile(filename, os.path.join(os.path.dirname(filename), 'model_best.pth.tar'))


if __name__ == '__main__':
    main()
