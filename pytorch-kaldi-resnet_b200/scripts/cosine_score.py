"""Drop-in for the reference's scripts/cosine_score.py (flags :38-44, files: mean ' [ ... ]', enroll/test Kaldi text
arks, trials 'id1 id2 target|nontarget', output 'id1 id2 score' per line, :65-68).  The per-trial Python loop with two
tensor constructions and F.cosine_similarity (:60-65) becomes ONE svk_cosine_score_pairs launch over index arrays; the
arithmetic is the reference's: float64 parse, float64 mean subtraction, float32 cosine with eps 1e-8.
Under torchrun (WORLD_SIZE > 1) trial blocks are sharded across ranks, no communication beyond the final file merge.
"""
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


def load_embeddings(path, mean):
    """-> (list of keys, float32 matrix of (vec - mean)), subtraction in float64 like cosine_score.py:53-56."""
    keys, rows = [], []
    for key, vec in kaldi_io.read_vec_flt_ark(path):
        keys.append(key)
        rows.append(np.asarray(vec, dtype=np.float64) - mean)
    mat = np.asarray(rows, dtype=np.float64).astype(np.float32) if rows else np.zeros((0, len(mean)), np.float32)
    return keys, mat


def read_trials(path, enroll_index, test_index):
    pairs, ie, it = [], [], []
    with open(path) as f:
        for line in f:
            spkr, utt, _target = line.strip().split()
            pairs.append((spkr, utt))
            ie.append(enroll_index[spkr])          # KeyError on an unknown id, like the reference's dict lookup
            it.append(test_index[utt])
    return pairs, np.asarray(ie, dtype=np.int32), np.asarray(it, dtype=np.int32)


def shard(n):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    from svk.parallel import shard_range
    lo, hi = shard_range(n, rank, world)
    return rank, world, lo, hi


def merge_shards(path, rank, world):
    """Rank r wrote path.r; after a file-system barrier rank 0 concatenates them in order."""
    if world == 1:
        return
    import time
    open("%s.%d.done" % (path, rank), "w").close()
    if rank != 0:
        return
    for r in range(world):
        while not os.path.exists("%s.%d.done" % (path, r)):
            time.sleep(0.05)
    with open(path, "w") as out:
        for r in range(world):
            with open("%s.%d" % (path, r)) as f:
                out.write(f.read())
            os.remove("%s.%d" % (path, r))
            os.remove("%s.%d.done" % (path, r))


def main():
    parser = argparse.ArgumentParser("Configuration for data preparation")
    parser.add_argument("--mean", type=str, help="mean vec file")
    parser.add_argument("--enroll", type=str, help="enroll embeddings file")
    parser.add_argument("--test", type=str, help="test embeddings file")
    parser.add_argument("--trials", type=str, help="trials file")
    parser.add_argument("--score-file", type=str, help="score file")
    args = parser.parse_args()
    if args.mean and os.path.exists(args.mean):
        mean = np.asarray(kaldi_io.read_vec_flt(args.mean), dtype=np.float64)
        print("loaded mean from {}".format(args.mean))
    else:
        print("mean file missing")
        return
    import torch
    from svk import scoring
    ekeys, emat = load_embeddings(args.enroll, mean)
    if os.path.abspath(args.test) == os.path.abspath(args.enroll):
        tkeys, tmat = ekeys, emat
    else:
        tkeys, tmat = load_embeddings(args.test, mean)
    eidx = {k: i for i, k in enumerate(ekeys)}
    tidx = {k: i for i, k in enumerate(tkeys)}
    pairs, ie, it = read_trials(args.trials, eidx, tidx)
    rank, world, lo, hi = shard(len(pairs))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    scores = scoring.cosine_scores(emat, tmat, None, ie[lo:hi], it[lo:hi]).cpu().numpy()
    out_path = args.score_file if world == 1 else "%s.%d" % (args.score_file, rank)
    with open(out_path, "w") as f:
        for (spkr, utt), s in zip(pairs[lo:hi], scores):
            f.write('{} {} {}\n'.format(spkr, utt, s))
    merge_shards(args.score_file, rank, world)
    print("saved scores of {} in {}".format(args.trials, args.score_file))


if __name__ == '__main__':
    main()
