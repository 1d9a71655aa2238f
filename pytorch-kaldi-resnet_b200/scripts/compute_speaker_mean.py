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


def read_utt2spk(path):
    """'utt spk' per line -> dict (two whitespace-separated fields, like the reference's unpacking at :11-13)."""
    table = {}
    with open(path) as f:
        for line in f:
            utt, spk = line.strip().split()
            table[utt] = spk
    return table


def speaker_table(ark_file, utt2spk_file):
    """-> (speakers in order of first appearance, float32 matrix of their mean embeddings)."""
    speaker_of = read_utt2spk(utt2spk_file)
    speakers, row_speaker, rows = [], [], []
    slot = {}
    for utt, vec in kaldi_io.read_vec_flt_ark(ark_file):
        spk = speaker_of.get(utt)
        if spk is None:
            raise Exception('{} not specified to any speaker'.format(utt))      # same message as :19
        if spk not in slot:
            slot[spk] = len(speakers)
            speakers.append(spk)
        row_speaker.append(slot[spk])
        rows.append(np.asarray(vec, dtype=np.float64))
    table = np.asarray(rows, dtype=np.float64).astype(np.float32)
    from svk import scoring
    means = scoring.speaker_means(table, np.asarray(row_speaker), len(speakers)).cpu().numpy()
    print("speakers: {}, feat-dim: {}".format(len(speakers), table.shape[1]))
    return speakers, means


def main(argv=None):
    ark_file, utt2spk_file, mean_file = (sys.argv[1:] if argv is None else argv)[:3]
    speakers, means = speaker_table(ark_file, utt2spk_file)
    with open(mean_file, 'w') as out:
        out.writelines("%s [ %s ]\n" % (spk, ' '.join(map(str, row))) for spk, row in zip(speakers, means))
    print("saved speaker mean in {}".format(mean_file))


if __name__ == '__main__':
    main()
