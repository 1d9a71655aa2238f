"""Drop-in for the reference's scripts/compute_eer.py (positional args <scores-file> <trials-file>, :18-31; prints the EER as
'{0:.2%}' on stdout and 'eer is ...' on stderr, :102-103).  The O(N) Python list loops after a Python sort (:35-70) become a
device radix sort + prefix sums (svk_sort_pairs_f64, svk_det_metrics), float64 like the reference."""
import argparse
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
for _p in (_HERE, _PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def read_scored_trials(scores_filename, trials_filename):
    """-> (float64 scores, int32 labels) in score-file order; a scored pair missing from the trial list raises (:91-93)."""
    trials = {}
    for line in open(trials_filename, 'r'):
        utt1, utt2, target = line.rstrip().split()
        trials[utt1 + " " + utt2] = target
    scores, labels = [], []
    for line in open(scores_filename, 'r'):
        utt1, utt2, score = line.rstrip().split()
        trial = utt1 + " " + utt2
        if trial not in trials:
            raise Exception("Missing entry for " + utt1 + " and " + utt2 + " " + scores_filename)
        scores.append(float(score))
        labels.append(1 if trials[trial] == "target" else 0)
    return np.asarray(scores, dtype=np.float64), np.asarray(labels, dtype=np.int32)


def main():
    parser = argparse.ArgumentParser(description="Compute equal error rate "
                                     "Usage: scripts/compute_eer.py <scores-file> <trials-file> "
                                     "E.g., scripts/compute_eer.py exp/scores/trials data/test/trials",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("scores_filename", help="Input scores file, with columns of the form <utt1> <utt2> <score>")
    parser.add_argument("trials_filename", help="Input trials file, with columns of the form <utt1> <utt2> <target/nontarget>")
    sys.stderr.write(' '.join(sys.argv) + "\n")
    args = parser.parse_args()
    scores, labels = read_scored_trials(args.scores_filename, args.trials_filename)
    from svk import scoring
    eer = scoring.det_metrics(scores, labels)["eer"]
    sys.stdout.write("{0:.2%}\n".format(eer))
    sys.stderr.write("eer is {0:.2%}\n".format(eer))


if __name__ == "__main__":
    main()
