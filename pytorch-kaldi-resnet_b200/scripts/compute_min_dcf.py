"""Drop-in for the reference's local/compute_min_dcf.py (flags --p-target --c-miss --c-fa and positional <scores-file>
<trials-file>, :19-38; prints '{0:.4f}' of the minimum normalised detection cost on stdout, :118-119).  Same device path as
compute_eer.py: stable radix sort of the float64 scores, cumulative error rates, first minimum of the cost (:93-106)."""
import argparse
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
for _p in (_HERE, _PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from compute_eer import read_scored_trials  # noqa: E402


def main():
    parser = argparse.ArgumentParser(description="Compute the minimum of the detection cost function.  The comments "
                                     "refer to equations in Section 3 of the NIST 2016 Speaker Recognition Evaluation Plan.",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--p-target', type=float, dest="p_target", default=0.01,
                        help='The prior probability of the target speaker in a trial.')
    parser.add_argument('--c-miss', type=float, dest="c_miss", default=1, help='Cost of a missed detection.')
    parser.add_argument('--c-fa', type=float, dest="c_fa", default=1, help='Cost of a spurious detection.')
    parser.add_argument("scores_filename", help="Input scores file, with columns of the form <utt1> <utt2> <score>")
    parser.add_argument("trials_filename", help="Input trials file, with columns of the form <utt1> <utt2> <target/nontarget>")
    sys.stderr.write(' '.join(sys.argv) + "\n")
    args = parser.parse_args()
    if args.c_fa <= 0:                                       # CheckArgs, compute_min_dcf.py:43-50
        raise Exception("--c-fa must be greater than 0")
    if args.c_miss <= 0:
        raise Exception("--c-miss must be greater than 0")
    if args.p_target <= 0 or args.p_target >= 1:
        raise Exception("--p-target must be greater than 0 and less than 1")
    scores, labels = read_scored_trials(args.scores_filename, args.trials_filename)
    from svk import scoring
    r = scoring.det_metrics(scores, labels, args.p_target, args.c_miss, args.c_fa)
    sys.stdout.write("{0:.4f}\n".format(r["min_dcf"]))
    sys.stderr.write("minDCF is {0:.4f} at threshold {1:.4f} (p-target={2}, c-miss={3},c-fa={4})\n".format(
        r["min_dcf"], r["min_dcf_threshold"], args.p_target, args.c_miss, args.c_fa))


if __name__ == "__main__":
    main()
