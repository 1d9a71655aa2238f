"""Kaldi ark/scp I/O for the speaker-embedding path — written from the Kaldi file-format description, keeping the
call signatures and return dtypes of the reference's `scripts/kaldi_io.py` for the entry points this path uses:

    open_or_fd  (:41)   read_key (:110)   read_vec_flt (:256)   read_vec_flt_ark (:238)   read_vec_flt_scp (:217)
    write_vec_flt (:294)   read_mat (:376)   read_mat_ark (:354)   read_mat_scp (:333)   write_mat (:464)

Formats: binary objects start with b"\\0B" then a 3-byte type tag ("FM "/"DM " matrices, "FV "/"DV " vectors), each
dimension as b"\\x04" + little-endian int32, then raw little-endian samples.  Text vectors are "[ v0 v1 ... ]" on one
line and parse to float64; text matrices are " [" followed by rows, the last ending in "]", and parse to float32.
Readers accept a path, "ark:path", "path:offset", or an open binary file object.  Compressed ("CM") matrices, pipes
and gzip are outside this path (the recipe writes features with --compress false) and raise UnsupportedDataType.

Beyond the reference API: read_mat_rows() memory-maps an "FM " matrix and copies only the requested frame range
(the training loader crops 200 frames out of utterances several times longer — SURVEY.md §8f-1).
"""
import mmap
import os
import struct

import numpy as np


class UnsupportedDataType(Exception):
    pass


class UnknownVectorHeader(Exception):
    pass


class UnknownMatrixHeader(Exception):
    pass


class BadInputFormat(Exception):
    pass


_BIN = b"\0B"
_MAT_TAGS = {b"FM ": np.dtype("<f4"), b"DM ": np.dtype("<f8")}
_VEC_TAGS = {b"FV ": np.dtype("<f4"), b"DV ": np.dtype("<f8")}


def _split_rxspec(spec):
    """'ark:feats.ark:1234' -> ('feats.ark', 1234); 'feats.ark' -> ('feats.ark', None)."""
    for prefix in ("ark:", "scp:"):
        if spec.startswith(prefix):
            spec = spec[len(prefix):]
            break
    offset = None
    head, sep, tail = spec.rpartition(":")
    if sep and tail.isdigit() and head:
        spec, offset = head, int(tail)
    return spec, offset


def open_or_fd(file, mode="rb"):
    """Open `file` (path with optional 'ark:' prefix and ':offset' suffix) or pass an open descriptor through."""
    if not isinstance(file, str):
        return file
    path, offset = _split_rxspec(file)
    if path.endswith("|") or path.startswith("|") or path.endswith(".gz"):
        raise UnsupportedDataType("pipes and gzip are not supported on this path: %s" % file)
    fd = open(path, mode)
    if offset is not None:
        fd.seek(offset)
    return fd


def read_key(fd):
    """Next utterance key (bytes up to the first space) or None at end of stream."""
    chars = []
    while True:
        c = fd.read(1)
        if not c or c == b" ":
            break
        chars.append(c)
    key = b"".join(chars).decode("latin1").strip()
    if not key:
        return None
    if any(ch.isspace() for ch in key):
        raise BadInputFormat("whitespace inside key %r" % key)
    return key


def _read_dim(fd):
    raw = fd.read(5)
    if len(raw) != 5 or raw[0:1] != b"\x04":
        raise BadInputFormat("expected a 4-byte dimension marker")
    return struct.unpack("<i", raw[1:])[0]


# ------------------------------------------------------------------------------------------------ float vectors
def _vec_after_flag(fd, first2):
    if first2 == _BIN:
        tag = fd.read(3)
        if tag not in _VEC_TAGS:
            raise UnknownVectorHeader("The header contained '%s'" % tag.decode("latin1"))
        n = _read_dim(fd)
        dt = _VEC_TAGS[tag]
        if n == 0:
            return np.array([], dtype="float32")
        return np.frombuffer(fd.read(n * dt.itemsize), dtype=dt).astype(dt.newbyteorder("="), copy=False)
    toks = (first2 + fd.readline()).decode("latin1").split()
    toks = [t for t in toks if t not in ("[", "]")]
    return np.array(toks, dtype=float)


def read_vec_flt(file_or_fd):
    """One float vector, binary ('FV '/'DV ') or text ('[ ... ]' -> float64)."""
    fd = open_or_fd(file_or_fd)
    try:
        return _vec_after_flag(fd, fd.read(2))
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_vec_flt_ark(file_or_fd):
    """Generator of (key, vector) over an ark of float vectors (text or binary entries)."""
    fd = open_or_fd(file_or_fd)
    try:
        while True:
            key = read_key(fd)
            if key is None:
                break
            yield key, _vec_after_flag(fd, fd.read(2))
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_vec_flt_scp(file_or_fd):
    """Generator of (key, vector) following 'key rxfile' lines of an scp."""
    fd = open_or_fd(file_or_fd)
    try:
        for line in fd:
            key, rx = line.decode("latin1").split(None, 1)
            yield key, read_vec_flt(rx.strip())
    finally:
        if fd is not file_or_fd:
            fd.close()


def write_vec_flt(file_or_fd, v, key=""):
    """Binary Kaldi vector ('FV ' float32 / 'DV ' float64), optionally preceded by 'key '."""
    v = np.asarray(v)
    if v.dtype == np.float32:
        tag = b"FV "
    elif v.dtype == np.float64:
        tag = b"DV "
    else:
        raise UnsupportedDataType("'%s', please use 'float32' or 'float64'" % v.dtype)
    fd = open_or_fd(file_or_fd, mode="wb")
    try:
        if key:
            fd.write((key + " ").encode("latin1"))
        fd.write(_BIN + tag + b"\x04" + struct.pack("<i", v.shape[0]))
        fd.write(np.ascontiguousarray(v).astype(v.dtype.newbyteorder("<"), copy=False).tobytes())
    finally:
        if fd is not file_or_fd:
            fd.close()


# ------------------------------------------------------------------------------------------------ float matrices
def _mat_after_flag(fd, first2):
    if first2 == _BIN:
        tag = fd.read(3)
        if tag.startswith(b"CM"):
            raise UnsupportedDataType("compressed matrices are not used on this path (feats are written --compress false)")
        if tag not in _MAT_TAGS:
            raise UnknownMatrixHeader("The header contained '%s'" % tag.decode("latin1"))
        rows, cols = _read_dim(fd), _read_dim(fd)
        dt = _MAT_TAGS[tag]
        buf = fd.read(rows * cols * dt.itemsize)
        if len(buf) != rows * cols * dt.itemsize:
            raise BadInputFormat("truncated matrix: wanted %d x %d" % (rows, cols))
        return np.frombuffer(buf, dtype=dt).astype(dt.newbyteorder("="), copy=False).reshape(rows, cols)
    if first2 != b" [":
        raise BadInputFormat("neither a binary nor a text matrix")
    rows = []
    while True:
        line = fd.readline()
        if not line:
            raise BadInputFormat("end of file inside a text matrix")
        toks = line.decode("latin1").split()
        if not toks:
            continue
        last = toks[-1] == "]"
        if last:
            toks = toks[:-1]
        if toks:
            rows.append(np.array(toks, dtype="float32"))
        if last:
            return np.vstack(rows) if rows else np.zeros((0, 0), dtype="float32")


def read_mat(file_or_fd):
    """One matrix: binary 'FM '/'DM ' (float32/float64, shape (rows, cols)) or text (float32)."""
    fd = open_or_fd(file_or_fd)
    try:
        return _mat_after_flag(fd, fd.read(2))
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_mat_ark(file_or_fd):
    fd = open_or_fd(file_or_fd)
    try:
        while True:
            key = read_key(fd)
            if key is None:
                break
            yield key, _mat_after_flag(fd, fd.read(2))
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_mat_scp(file_or_fd):
    fd = open_or_fd(file_or_fd)
    try:
        for line in fd:
            key, rx = line.decode("latin1").split(None, 1)
            yield key, read_mat(rx.strip())
    finally:
        if fd is not file_or_fd:
            fd.close()


def write_mat(file_or_fd, m, key=""):
    """Binary Kaldi matrix; a path is opened 'wb' (scripts/kaldi_io.py:482 semantics), a descriptor is appended to.
    Returns nothing; use fd.tell() before the call to build scp offsets (offset points at the b'\\0B' flag)."""
    if not isinstance(m, np.ndarray) or m.ndim != 2:
        raise ValueError("'m' has to be a 2d numpy matrix")
    if m.dtype == np.float32:
        tag = b"FM "
    elif m.dtype == np.float64:
        tag = b"DM "
    else:
        raise UnsupportedDataType("'%s', please use 'float32' or 'float64'" % m.dtype)
    fd = open_or_fd(file_or_fd, mode="wb")
    try:
        if key:
            fd.write((key + " ").encode("latin1"))
        fd.write(_BIN + tag + b"\x04" + struct.pack("<i", m.shape[0]) + b"\x04" + struct.pack("<i", m.shape[1]))
        fd.write(np.ascontiguousarray(m).astype(m.dtype.newbyteorder("<"), copy=False).tobytes())
    finally:
        if fd is not file_or_fd:
            fd.close()


# ------------------------------------------------------------------------------------------------ fast crop reader
_MMAPS = {}


def _mapped(path):
    ent = _MMAPS.get(path)
    if ent is None:
        f = open(path, "rb")
        ent = (f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ))
        _MMAPS[path] = ent
    return ent[1]


def mat_shape(rxfile):
    """(rows, cols) of a binary float32 matrix at 'path:offset' without reading its data."""
    path, offset = _split_rxspec(rxfile)
    mm = _mapped(path)
    o = offset or 0
    if mm[o:o + 5] != _BIN + b"FM ":
        raise UnsupportedDataType("read_mat_rows/mat_shape need an uncompressed float32 'FM ' matrix")
    rows = struct.unpack_from("<i", mm, o + 6)[0]
    cols = struct.unpack_from("<i", mm, o + 11)[0]
    return rows, cols


def read_mat_rows(rxfile, start, count):
    """Rows [start, start+count) of the 'FM ' matrix at 'path:offset' — seeks offset + 15 + start*cols*4 and copies
    only the crop (count = -1: all rows)."""
    path, offset = _split_rxspec(rxfile)
    rows, cols = mat_shape(rxfile)
    if count < 0:
        start, count = 0, rows
    if start < 0 or start + count > rows:
        raise IndexError("rows [%d, %d) outside a %d-row matrix" % (start, start + count, rows))
    mm = _mapped(path)
    o = (offset or 0) + 15 + start * cols * 4
    return np.frombuffer(mm, dtype="<f4", count=count * cols, offset=o).reshape(count, cols).copy()
