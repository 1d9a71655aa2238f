"""Drop-in for the reference's scripts/compute_topk_mean_std.py (flags :26-31; output 'utt mean std' per line, :56).
For every embedding: cosine scores against the L2-normalised cohort, top-300, mean and UNBIASED std (:10-23) — here a
normalise kernel, one GEMM per block of embeddings and a radix-select top-k kernel instead of a Python loop of
matmul/topk/std_mean.  Under torchrun embedding rows are sharded across ranks."""
import argparse
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
for _p in (_HERE, _PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import kaldi_io  # noqa: E402
from cosine_score import load_embeddings, merge_shards, shard  # noqa: E402

TOPK = 300   # compute_topk_mean_std.py:10


def compute_topk_mean_std(vecs, cohort_mat, topk=TOPK):
    """vecs (n, D) and cohort_mat (m, D): mean-subtracted float32 -> (mean, std) numpy float32 arrays of length n."""
    from svk import scoring
    print('cohort_mat.shape: {}, utt2vec: {}'.format(tuple(cohort_mat.shape), len(vecs)))
    mean, std = scoring.cohort_topk_meanstd(vecs, cohort_mat, topk=topk)
    return mean.cpu().numpy(), std.cpu().numpy()


def main():
    parser = argparse.ArgumentParser("Configuration for data preparation")
    parser.add_argument("--mean", type=str, help="mean vec file")
    parser.add_argument("--ark-file", type=str, help="test embeddings file")
    parser.add_argument("--cohort-file", type=str, help="cohort embeddings file")
    parser.add_argument("--mean-std-file", type=str, help="file to save mean and std")
    args = parser.parse_args()
    if args.mean and os.path.exists(args.mean):
        mean = np.asarray(kaldi_io.read_vec_flt(args.mean), dtype=np.float64)
        print("loaded mean from {}".format(args.mean))
    else:
        print("mean file missing")
        return
    import torch
    keys, mat = load_embeddings(args.ark_file, mean)
    _, cohort = load_embeddings(args.cohort_file, mean)
    rank, world, lo, hi = shard(len(keys))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    m, s = compute_topk_mean_std(mat[lo:hi], cohort)
    out_path = args.mean_std_file if world == 1 else "%s.%d" % (args.mean_std_file, rank)
    with open(out_path, 'w') as f:
        for key, mi, si in zip(keys[lo:hi], m, s):
            f.write('{} {} {}\n'.format(key, mi, si))
    merge_shards(args.mean_std_file, rank, world)
    print("saved speaker mean in {}".format(args.mean_std_file))


if __name__ == '__main__':
    main()
