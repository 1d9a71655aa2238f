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


def embedding_table(ark_file):
    """All vectors of the ark as one float32 matrix (text vectors parse to float64 first, like torch.FloatTensor(mat) at :14)."""
    rows = [np.asarray(vec, dtype=np.float64) for _utt, vec in kaldi_io.read_vec_flt_ark(ark_file)]
    return np.asarray(rows, dtype=np.float64).astype(np.float32)


def main(argv=None):
    ark_file, mean_file = (sys.argv[1:] if argv is None else argv)[:2]
    table = embedding_table(ark_file)
    print("speakers: {}, feat-dim: {}".format(*table.shape))
    from svk import scoring
    mean = scoring.global_mean(table).cpu().numpy()
    with open(mean_file, 'w') as out:
        out.write(" [ %s ]\n" % ' '.join(map(str, mean)))
    print("saved mean of {} in {}".format(ark_file, mean_file))


if __name__ == '__main__':
    main()
