"""Drop-in for the reference's scripts/compute_speaker_mean.py (argv: <ark-file> <utt2spk> <mean-file>, :32-44): per-speaker
mean embedding, one line 'spk [ v0 v1 ... ]' per speaker in order of first appearance (:40-41).  The per-utterance numpy
accumulation (:16-27) is one svk_segment_mean launch that adds every speaker's rows in file order."""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
for _p in (_HERE, _PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import kaldi_io  # noqa: E402


def compute_speaker_mean(ark_file, utt2spk_file):
    utt2spk = {}
    for line in open(utt2spk_file, 'r'):
        utt, spk = line.strip().split()
        utt2spk[utt] = spk
    spk_index, seg, rows = {}, [], []
    for utt, vec in kaldi_io.read_vec_flt_ark(ark_file):
        if utt not in utt2spk:
            raise Exception('{} not specified to any speaker'.format(utt))
        seg.append(spk_index.setdefault(utt2spk[utt], len(spk_index)))
        rows.append(np.asarray(vec, dtype=np.float64))
    mat = np.asarray(rows, dtype=np.float64).astype(np.float32)
    from svk import scoring
    means = scoring.speaker_means(mat, np.asarray(seg), len(spk_index)).cpu().numpy()
    print("speakers: {}, feat-dim: {}".format(len(spk_index), mat.shape[1]))
    return {spk: means[i] for spk, i in spk_index.items()}                # dicts keep insertion order, like the reference's


def main():
    ark_file = sys.argv[1]
    utt2spk_file = sys.argv[2]
    mean_file = sys.argv[3]
    speaker_mean = compute_speaker_mean(ark_file, utt2spk_file)
    with open(mean_file, 'w') as f:
        for spk in speaker_mean:
            f.write(spk + ' [ ' + ' '.join(map(str, speaker_mean[spk])) + ' ]\n')
    print("saved speaker mean in {}".format(mean_file))


if __name__ == '__main__':
    main()
