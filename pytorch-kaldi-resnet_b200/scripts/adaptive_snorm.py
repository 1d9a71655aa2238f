"""Drop-in for the reference's scripts/adaptive_snorm.py (flags :9-14): reads 'id mean std' tables for the enroll and
test sides and a score file 'id1 id2 score', writes 'id1 id2 snorm' with
    snorm = (s - mu_e) / max(sigma_e, 1e-8) / 2 + (s - mu_t) / max(sigma_t, 1e-8) / 2        (adaptive_snorm.py:33-34)
as one svk_snorm_apply launch (float32; the reference uses Python floats — results agree to ~1e-6)."""
import argparse
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
for _p in (_HERE, _PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def read_stats(path):
    index, mean, std = {}, [], []
    with open(path) as f:
        for line in f:
            key, m, s = line.strip().split()
            index[key] = len(mean)
            mean.append(float(m))
            std.append(float(s))
    return index, np.asarray(mean, np.float32), np.asarray(std, np.float32)


def main():
    parser = argparse.ArgumentParser("Configuration for data preparation")
    parser.add_argument("--enroll", type=str, help="enroll topk mean and std file")
    parser.add_argument("--test", type=str, help="test topk mean and std file")
    parser.add_argument("--score-in", type=str, help="score in file")
    parser.add_argument("--score-out", type=str, help="score out file")
    args = parser.parse_args()
    from svk import scoring
    eidx, emean, estd = read_stats(args.enroll)
    tidx, tmean, tstd = read_stats(args.test)
    pairs, ie, it, scores = [], [], [], []
    with open(args.score_in) as f:
        for line in f:
            spkr, utt, score = line.strip().split()
            pairs.append((spkr, utt))
            ie.append(eidx[spkr])
            it.append(tidx[utt])
            scores.append(float(score))
    out = scoring.snorm_apply(np.asarray(scores, np.float32), np.asarray(ie, np.int32), np.asarray(it, np.int32), emean,
                              estd, tmean, tstd).cpu().numpy()
    with open(args.score_out, 'w') as f:
        for (spkr, utt), s in zip(pairs, out):
            f.write('{} {} {}\n'.format(spkr, utt, s))
    print("saved adaptive S-norm scores in {}".format(args.score_out))


if __name__ == '__main__':
    main()
