"""Drop-in for the reference's scripts/compute_mean.py (argv: <ark-file> <mean-file>, :22-31): global mean of an embedding
ark (Kaldi text or binary vectors), written as ' [ v0 v1 ... ]' (:28).  The float32 table is averaged by svk_col_mean
(float64 accumulation; the reference's torch.mean over a FloatTensor agrees to ~1e-7 relative)."""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
for _p in (_HERE, _PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import kaldi_io  # noqa: E402


def compute_mean(ark_file):
    rows = [np.asarray(vec, dtype=np.float64) for _key, vec in kaldi_io.read_vec_flt_ark(ark_file)]
    mat = np.asarray(rows, dtype=np.float64).astype(np.float32)           # torch.FloatTensor(mat) (compute_mean.py:14)
    print("speakers: {}, feat-dim: {}".format(mat.shape[0], mat.shape[1]))
    from svk import scoring
    return scoring.global_mean(mat).cpu().numpy()


def main():
    ark_file = sys.argv[1]
    mean_file = sys.argv[2]
    mean = compute_mean(ark_file)
    with open(mean_file, 'w') as f:
        f.write(' [ ' + ' '.join(map(str, mean)) + ' ]\n')
    print("saved mean of {} in {}".format(ark_file, mean_file))


if __name__ == '__main__':
    main()
