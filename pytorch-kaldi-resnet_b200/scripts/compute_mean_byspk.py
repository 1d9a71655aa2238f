"""Drop-in for the reference's scripts/compute_mean_byspk.py (argv: <spk2utt> <ark-file> <mean-file>, :31-43): mean
embedding of every speaker LISTED IN spk2utt ('spk utt1 utt2 ...' per line, :14-24), one line 'spk [ v0 v1 ... ]' per speaker
in spk2utt order.  The reference builds a float32 tensor per speaker and calls torch.mean (:21-24); here all speakers are one
svk_segment_mean launch over the embedding table (rows of a speaker added in spk2utt order)."""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
for _p in (_HERE, _PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import kaldi_io  # noqa: E402


def speaker_means(spk2utt_file, ark_file):
    """-> (speakers in spk2utt order, float32 matrix of their mean embeddings)."""
    index, rows = {}, []
    for utt, vec in kaldi_io.read_vec_flt_ark(ark_file):
        index[utt] = len(rows)                                  # a repeated key keeps its last vector, like the dict at :12-13
        rows.append(np.asarray(vec, dtype=np.float64))
    table = np.asarray(rows, dtype=np.float64).astype(np.float32)   # torch.FloatTensor(mat), :21
    speakers, gather, seg = [], [], []
    with open(spk2utt_file) as f:
        for line in f:
            arr = line.strip().split()
            if not arr:
                continue
            if len(arr) == 1:
                raise ValueError("speaker %s has no utterances in %s" % (arr[0], spk2utt_file))
            for utt in arr[1:]:
                gather.append(index[utt])                       # KeyError on an unknown utterance, like utt2vec[key] at :19
                seg.append(len(speakers))
            speakers.append(arr[0])
    from svk import scoring
    picked = table[np.asarray(gather, dtype=np.int64)]
    means = scoring.speaker_means(picked, np.asarray(seg), len(speakers)).cpu().numpy()
    print("speakers: {}, feat-dim: {}".format(len(speakers), table.shape[1]))
    return speakers, means


def main(argv=None):
    spk2utt_file, ark_file, mean_file = (sys.argv[1:] if argv is None else argv)[:3]
    speakers, means = speaker_means(spk2utt_file, ark_file)
    with open(mean_file, 'w') as out:
        out.writelines("%s [ %s ]\n" % (spk, ' '.join(map(str, row))) for spk, row in zip(speakers, means))
    print("saved mean of {} by {} in {}".format(ark_file, spk2utt_file, mean_file))


if __name__ == '__main__':
    main()
