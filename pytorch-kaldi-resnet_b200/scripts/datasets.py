"""Datasets over uncompressed Kaldi feats.scp files, with the reference's class names, constructor arguments, sample
format ((F, T) float32 chunk, label) and balancing rules (scripts/datasets.py:7-72 SequenceDataset, :74-146
SequenceDataset2, :148-193 EmbeddingDataset).  Difference: a crop is read with
kaldi_io.read_mat_rows — mmap + copy of the 200 requested frames — instead of loading the whole utterance."""
import numpy as np
from torch.utils.data import Dataset

import kaldi_io


def _read_utt2spkid(path):
    table = {}
    with open(path) as f:
        for line in f:
            utt, label = line.rstrip().split()
            table[utt] = int(label)
    return table


def _crop(rxfile, seq_len):
    """Random `seq_len`-frame crop (all frames if seq_len < 0) of the matrix at rxfile, transposed to (F, T)."""
    try:
        rows, _ = kaldi_io.mat_shape(rxfile)
        if seq_len < 0:
            return np.ascontiguousarray(kaldi_io.read_mat_rows(rxfile, 0, -1).T)
        assert rows >= seq_len
        pin = np.random.randint(0, rows - seq_len + 1)
        return np.ascontiguousarray(kaldi_io.read_mat_rows(rxfile, pin, seq_len).T)
    except kaldi_io.UnsupportedDataType:
        full = kaldi_io.read_mat(rxfile)              # text / float64 matrices: whole-matrix path
        if seq_len < 0:
            return np.ascontiguousarray(full.T)
        assert len(full) >= seq_len
        pin = np.random.randint(0, len(full) - seq_len + 1)
        return np.ascontiguousarray(full[pin:pin + seq_len, :].T)


class SequenceDataset(Dataset):
    """One entry per (utterance x repetition); rare speakers are repeated up to min(500, (max_count+1)//2) samples
    per class (datasets.py:23-31)."""

    def __init__(self, scp_file, utt2spkid_file, chunk_size):
        self.utt2spkid = _read_utt2spkid(utt2spkid_file)
        counts = {}
        for label in self.utt2spkid.values():
            counts[label] = counts.get(label, 0) + 1
        cap = min(500, int((max(counts.values()) + 1) / 2))
        rxfiles, labels = [], []
        with open(scp_file) as f:
            for line in f:
                utt, rxfile = line.rstrip().split()
                label = self.utt2spkid[utt]
                rep = max(1, cap // counts[label])
                rxfiles.extend([rxfile] * rep)
                labels.extend([label] * rep)
        self.rxfiles = np.array(rxfiles)
        self.labels = np.array(labels, dtype=np.int64)
        n = len(self.labels)
        if isinstance(chunk_size, int):
            self.seq_len = np.full(n, chunk_size, dtype=np.int64)
        elif len(chunk_size) == 1:
            self.seq_len = np.full(n, chunk_size[0], dtype=np.int64)
        else:
            self.seq_len = np.random.randint(min(chunk_size), max(chunk_size) + 1, size=n)
        print("Totally " + str(n) + " samples with at most " + str(cap) + " samples for one class")

    def __len__(self):
        return len(self.labels)

    def set_chunk_size(self, seq_len):
        self.seq_len = seq_len

    def __getitem__(self, index):
        return _crop(self.rxfiles[index], int(self.seq_len[index])), np.array(self.labels[index])


class SequenceDataset2(Dataset):
    """Speaker-balanced: index i draws a random utterance of speaker i % num_spk; length num_spk * repetition with
    repetition = (max utterances per speaker + 1) // 2 (datasets.py:104-146)."""

    def __init__(self, scp_file, utt2spkid_file, chunk_size):
        utt2spkid = _read_utt2spkid(utt2spkid_file)
        self.rxfiles = {}
        with open(scp_file) as f:
            for line in f:
                utt, rxfile = line.rstrip().split()
                self.rxfiles.setdefault(utt2spkid[utt], []).append(rxfile)
        most = max(len(v) for v in self.rxfiles.values())
        self.repetition = int((most + 1) / 2)
        print("id_count: {}".format(most))
        self.labels = np.array(sorted(self.rxfiles))
        self.seq_len = chunk_size
        self.num_spk = len(self.rxfiles)
        print("Totally " + str(self.num_spk) + " speakers with at most " + str(self.repetition) + " samples for one class")

    def __len__(self):
        return len(self.labels) * self.repetition

    def set_chunk_size(self, seq_len):
        self.seq_len = seq_len

    def __getitem__(self, index):
        spkid = self.labels[index % self.num_spk]
        files = self.rxfiles[spkid]
        rxfile = files[np.random.randint(0, len(files))]
        return _crop(rxfile, int(self.seq_len)), np.array(spkid)


class EmbeddingDataset(Dataset):
    """Whole utterances (chunk_size = -1) or random crops for extraction; items are ((F, T) matrix, [utt])."""

    def __init__(self, scp_file, chunk_size=-1):
        self.rxfiles, self.utts = [], []
        with open(scp_file) as f:
            for line in f:
                utt, rxfile = line.rstrip().split()
                self.rxfiles.append(rxfile)
                self.utts.append(utt)
        self.rxfiles = np.array(self.rxfiles)
        self.seq_len = chunk_size
        print("Totally " + str(len(self.rxfiles)) + " samples")

    def __len__(self):
        return len(self.rxfiles)

    def set_chunk_size(self, seq_len):
        self.seq_len = seq_len

    def num_frames(self, index):
        """Frame count of utterance `index` without reading its data (length-balanced sharding / batching)."""
        try:
            return kaldi_io.mat_shape(self.rxfiles[index])[0]
        except kaldi_io.UnsupportedDataType:
            return len(kaldi_io.read_mat(self.rxfiles[index]))

    def __getitem__(self, index):
        return _crop(self.rxfiles[index], int(self.seq_len)), [self.utts[index]]
