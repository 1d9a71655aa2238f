"""Drop-in for the reference's scripts/decode.py (same flags :26-61, same process model — one process per GPU, same
output: Kaldi text vectors `utt [ v0 ... v255 ]` in <out-path>/<gpu> or <out-path>/alone, :189-206).

What changes underneath (the embeddings are the same, decode.py:198 `model.predict` semantics = batch 1 per utterance):
  * utterances are sharded by LENGTH across ranks (svk.parallel.shard_by_length): every utterance exactly once, no
    sampler padding duplicates, no shuffle (the reference's DistributedSampler(shuffle=True) at :170 pads with
    duplicates that the recipe removes with awk afterwards, run_aam_v2.sh:134);
  * each rank sorts its utterances by length and runs them in zero-padded batches with per-row valid lengths; every
    layer re-zeroes the padding, so each row equals its batch-1 result exactly (tests/test_model_gpu.py) while the
    launch count per utterance drops by the batch factor (batch 1 at ~40 launches per utterance is launch-bound);
  * BatchNorm is folded into the conv epilogues (eval mode); the host -> device copy of batch i+1 overlaps batch i.
`-b/--batch-size` and `-j/--workers` are accepted for command-line compatibility and otherwise UNUSED: the batch is
whatever fits the frame budget `--max-batch-frames`, and the reader is the mmap crop reader of kaldi_io (no worker
processes).  `--chunk-size -1` (whole utterances) is what the recipes use.
"""
import argparse
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
for _p in (_HERE, _PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from datasets import EmbeddingDataset  # noqa: E402
from model import NeuralSpeakerModel  # noqa: E402
from svk.ckpt import load_checkpoint  # noqa: E402
from svk.parallel import shard_by_length  # noqa: E402

parser = argparse.ArgumentParser(description='B200-native speaker-embedding extraction')
parser.add_argument('--spk_num', type=int, help='number of speakers of the trained model (head is unused)')
parser.add_argument('-a', '--arch', metavar='ARCH', default='resnet34')
parser.add_argument('--input-dim', type=int, required=True, help='input feature dimension')
parser.add_argument('--pooling', type=str, default='mean', help='mean or mean+std')
parser.add_argument('--chunk-size', default=-1, type=int, help='-1: whole utterances')
parser.add_argument('--model-path', type=str, required=True, help='checkpoint written by train_resnet.py')
parser.add_argument('--world-size', default=-1, type=int)
parser.add_argument('--rank', default=-1, type=int)
parser.add_argument('-j', '--workers', default=2, type=int, help='accepted, unused (see module docstring)')
parser.add_argument('-b', '--batch-size', default=8, type=int, help='accepted, unused: see --max-batch-frames')
parser.add_argument('--dist-url', default='tcp://224.66.41.62:23456', type=str)
parser.add_argument('--dist-backend', default='nccl', type=str)
parser.add_argument('--seed', default=None, type=int)
parser.add_argument('--gpu', default=None, type=int)
parser.add_argument('--gpu-num', default=-1, type=int)
parser.add_argument('--decode-scp', type=str, required=True, help='feats.scp to embed')
parser.add_argument('--out-path', type=str, required=True)
parser.add_argument('--multiprocessing-distributed', action='store_true')
parser.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
parser.add_argument('--max-batch-frames', default=65536, type=int, help='frame budget of one padded batch')
parser.add_argument('--embed-format', default='text', choices=['text', 'ark'],
                    help="text: 'utt [ v0 ... ]' lines like the reference (decode.py:206); ark: Kaldi binary float vectors "
                         "(kaldi_io.write_vec_flt) in <out-path>/<gpu>.ark — formatting / re-parsing 256 floats per utterance "
                         "as text dominates the end-to-end scoring time at scale; every reader here accepts both")

MAX_BATCH_UTTS = 64
STAGE_SLOTS = int(os.environ.get("SVK_STAGE_SLOTS", "3"))   # pinned staging slots (= batches in flight on the host side)
_PINNED = {}             # (slot, kind) -> pinned tensor, grown on demand and kept across extract() calls


def _pinned(slot, kind, numel, dtype):
    t = _PINNED.get((slot, kind))
    if t is None or t.numel() < numel or t.dtype != dtype:
        t = torch.empty(numel, dtype=dtype).pin_memory()
        _PINNED[(slot, kind)] = t
    return t


def main():
    args = parser.parse_args()
    if args.seed is not None:
        random.seed(args.seed)
        torch.manual_seed(args.seed)
    if args.dist_url == "env://" and args.world_size == -1:
        args.world_size = int(os.environ["WORLD_SIZE"])
    args.distributed = args.world_size > 1 or args.multiprocessing_distributed
    ngpus = torch.cuda.device_count() if args.gpu_num == -1 else min(torch.cuda.device_count(), args.gpu_num)
    if ngpus == 0:
        raise RuntimeError("decode.py needs a CUDA device (there is no CPU path)")
    if args.multiprocessing_distributed:
        args.world_size = ngpus * args.world_size
        mp.spawn(main_worker, nprocs=ngpus, args=(ngpus, args))
    else:
        main_worker(args.gpu, ngpus, args)


def plan_batches(lengths, indices, max_frames, max_utts=MAX_BATCH_UTTS):
    """Group utterance indices (sorted by length) into batches whose padded size max_len * count <= max_frames."""
    order = sorted(indices, key=lambda i: (lengths[i], i))
    batches, cur = [], []
    for i in order:
        longest = lengths[i]            # ascending order: the newcomer is the longest
        if cur and ((len(cur) + 1) * longest > max_frames or len(cur) >= max_utts):
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def extract(model, dataset, indices, device, max_frames, on_result):
    """Run `indices` of `dataset` through model.predict in padded, length-sorted batches; on_result(utt, vector)."""
    lengths = [dataset.num_frames(i) if dataset.seq_len < 0 else dataset.seq_len for i in range(len(dataset))]
    batches = plan_batches(lengths, indices, max_frames)
    feat_dim = model.feat_dim
    if not batches:
        return
    copy_stream = torch.cuda.Stream(device=device)
    # Persistent pinned staging slots, sized for the largest padded batch (cudaHostAlloc per batch cost more than the batch's
    # forward pass), filled by a small thread pool a few batches ahead: reading / padding 64 utterances in Python takes
    # longer than their forward pass, and the numpy / torch copies release the GIL.
    from concurrent.futures import ThreadPoolExecutor
    nslot = min(STAGE_SLOTS, len(batches))
    cap = max(len(b) * max(lengths[i] for i in b) for b in batches) * feat_dim
    pinned = [_pinned(k, "x", cap, torch.float32) for k in range(nslot)]
    pinned_len = [_pinned(k, "len", MAX_BATCH_UTTS, torch.int32) for k in range(nslot)]

    pinned_np = [t.numpy() for t in pinned]                      # numpy views of the pinned slots: plain memcpy on the worker
    pinned_len_np = [t.numpy() for t in pinned_len]              # threads (torch CPU ops would fan out over the OpenMP pool
                                                                 # from every worker at once and make the timing erratic)

    def fill(slot, batch):                                       # host side only: no CUDA calls on the worker threads
        mats = [dataset[i][0] for i in batch]                    # (F, T_i) float32
        lens = [m.shape[1] for m in mats]
        tmax = max(lens)
        n = len(mats) * feat_dim * tmax
        host_np = pinned_np[slot][:n].reshape(len(mats), feat_dim, tmax)
        same = min(lens) == tmax
        if not same:
            host_np.fill(0.0)
        for r, m in enumerate(mats):
            host_np[r, :, :m.shape[1]] = m
        pinned_len_np[slot][:len(mats)] = lens
        return pinned[slot][:n].view(len(mats), feat_dim, tmax), pinned_len[slot][:len(mats)], same

    def collect(batch, emb):
        out = emb.float().cpu().numpy()                          # waits for that batch's kernels
        for r, i in enumerate(batch):
            on_result(dataset.utts[i], out[r])

    pool = ThreadPoolExecutor(max_workers=max(1, nslot - 1))
    try:
        futures = {j: pool.submit(fill, j % nslot, batches[j]) for j in range(nslot)}
        pending = None
        with torch.no_grad():
            for bi, batch in enumerate(batches):
                host, hl, same = futures.pop(bi).result()
                with torch.cuda.stream(copy_stream):
                    x = host.to(device, non_blocking=True)
                    ln = hl.to(device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                torch.cuda.current_stream().wait_event(ev)
                if pending is not None:
                    collect(*pending)                             # batch bi-1 is done: its result is read back and its
                    j = bi - 1 + nslot                            # pinned slot (copied out before its kernels ran) is refilled
                    if j < len(batches):
                        futures[j] = pool.submit(fill, j % nslot, batches[j])
                emb = model.predict(x, lengths=None if same else ln)   # asynchronous launches
                pending = (batch, emb)
            collect(*pending)
    finally:
        pool.shutdown(wait=True)


def main_worker(gpu, ngpus_per_node, args):
    args.gpu = gpu
    dev_index = gpu if gpu is not None else 0
    print("Use GPU: {} for training".format(dev_index))
    torch.cuda.set_device(dev_index)
    rank, world = 0, 1
    if args.distributed:
        if args.dist_url == "env://" and args.rank == -1:
            args.rank = int(os.environ["RANK"])
        if args.multiprocessing_distributed:
            args.rank = args.rank * ngpus_per_node + gpu
        dist.init_process_group(backend=args.dist_backend, init_method=args.dist_url, world_size=args.world_size,
                                rank=args.rank)
        rank, world = args.rank, args.world_size
    print("=> creating model '{}'".format(args.arch))
    model = NeuralSpeakerModel(spk_num=args.spk_num, feat_dim=args.input_dim, pooling=args.pooling,
                               precision=args.precision)
    if not os.path.isfile(args.model_path):
        print("=> no checkpoint found at '{}'".format(args.model_path))
        return
    print("=> loading checkpoint '{}'".format(args.model_path))
    checkpoint = load_checkpoint(args.model_path, map_location='cpu')
    model.loadParameters(checkpoint['state_dict'])
    print("=> loaded checkpoint '{}' (epoch {})".format(args.model_path, checkpoint.get('epoch')))
    model.cuda(dev_index)
    model.eval()
    dataset = EmbeddingDataset(scp_file=args.decode_scp, chunk_size=args.chunk_size)
    lengths = [dataset.num_frames(i) for i in range(len(dataset))]
    mine = shard_by_length(lengths, rank, world) if world > 1 else list(range(len(dataset)))
    print("=> args.world_size: {}, args.rank: {}, loaded embedding samples num: {}".format(world, rank, len(mine)))
    os.makedirs(args.out_path, exist_ok=True)
    name = str(args.gpu) if args.gpu is not None else 'alone'
    results = {}
    extract(model, dataset, mine, torch.device('cuda', dev_index), args.max_batch_frames,
            lambda utt, vec: results.__setitem__(utt, vec))
    if args.embed_format == 'ark':
        import numpy as np
        import kaldi_io
        with open(os.path.join(args.out_path, name + '.ark'), 'wb') as f:
            for i in mine:
                utt = dataset.utts[i]
                kaldi_io.write_vec_flt(f, np.asarray(results[utt], dtype=np.float32), key=utt)
    else:
        with open(os.path.join(args.out_path, name), 'w') as f:
            for i in mine:                                   # scp order within the shard
                utt = dataset.utts[i]
                f.write(utt + ' [ ' + ' '.join(map(str, results[utt])) + ' ]\n')   # decode.py:206 format
    if args.distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
