"""Minimal FASTA I/O for the hot path (the reference uses Biopython's SeqIO, utils.py:301-309, 530-546)."""
from __future__ import annotations

import os
from typing import Iterator, List, Tuple

import numpy as np


def read_fasta(path: str) -> List[Tuple[str, str, np.ndarray]]:
    """[(id, full header text without '>', sequence as uint8 ASCII array)] for every record of a FASTA file."""
    with open(path, 'rb') as f:
        data = f.read()
    out: List[Tuple[str, str, np.ndarray]] = []
    if not data:
        return out
    buf = np.frombuffer(data, dtype=np.uint8)
    # record starts: '>' at file start or after a newline
    gt = np.flatnonzero(buf == ord('>'))
    starts = [int(p) for p in gt if p == 0 or buf[p - 1] == 10]
    for k, s in enumerate(starts):
        e = starts[k + 1] if k + 1 < len(starts) else len(buf)
        nl = data.find(b'\n', s, e)
        if nl < 0:
            nl = e
        header = data[s + 1:nl].decode('utf-8', 'replace').rstrip('\r')
        body = buf[nl + 1:e]
        seq = body[(body != 10) & (body != 13) & (body != 32)]
        rid = header.split()[0] if header.split() else ''
        out.append((rid, header, np.ascontiguousarray(seq)))
    return out


def write_fasta_record(path: str, header: str, seq: np.ndarray, width: int = 60) -> None:
    """One record, sequence wrapped at `width` columns (SeqIO.write's layout)."""
    n = len(seq)
    with open(path, 'wb') as f:
        f.write(b'>' + header.encode() + b'\n')
        if n:
            full = n // width
            if full:
                block = np.empty((full, width + 1), dtype=np.uint8)
                block[:, :width] = seq[:full * width].reshape(full, width)
                block[:, width] = 10
                f.write(block.tobytes())
            if n % width:
                f.write(seq[full * width:].tobytes() + b'\n')


def dir_records(seqdir: str) -> Iterator[Tuple[str, str, np.ndarray, str]]:
    """Every record of every file in a directory, files in sorted order: (id, header, seq, path)."""
    for fn in sorted(os.listdir(seqdir)):
        p = os.path.join(seqdir, fn)
        if os.path.isfile(p):
            for rid, header, seq in read_fasta(p):
                yield rid, header, seq, p
