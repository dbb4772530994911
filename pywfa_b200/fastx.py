"""FASTA / FASTQ streaming into the batched entry point.

The reference's own tests read their fixtures with ``pysam.FastxFile`` and align record by record
(``pywfa/tests/test.py:198-232``); here records are parsed with the standard library only (plain or
gzip, FASTA with wrapped lines, 4-line FASTQ) and handed to ``WavefrontAligner.align_arrays`` in
batches, so that reading batch ``i+1`` overlaps nothing on the GPU but at least never builds
Python strings per base.  ``FastxRecord`` mirrors the fields of pysam's proxy that the reference's
tests use (``name``, ``sequence``, ``comment``, ``quality``).
"""
from __future__ import annotations

import gzip
import io
from typing import Iterator, NamedTuple, Optional

import numpy as np


class FastxRecord(NamedTuple):
    name: str
    sequence: str
    comment: Optional[str] = None
    quality: Optional[str] = None


def _open(path):
    with open(path, "rb") as fh:
        magic = fh.read(2)
    if magic == b"\x1f\x8b":
        return io.TextIOWrapper(gzip.open(path, "rb"), encoding="ascii")
    return open(path, "r", encoding="ascii")


def read_fastx(path) -> Iterator[FastxRecord]:
    """Yield the records of a FASTA or FASTQ file (format detected from the first byte)."""
    with _open(path) as fh:
        line = fh.readline()
        while line and not line.strip():
            line = fh.readline()
        if not line:
            return
        if line[0] == ">":
            name, comment, chunks = None, None, []
            while line:
                line = line.rstrip("\r\n")
                if line.startswith(">"):
                    if name is not None:
                        yield FastxRecord(name, "".join(chunks), comment, None)
                    head = line[1:].split(None, 1)
                    name, comment, chunks = (head[0] if head else ""), (head[1] if len(head) > 1 else None), []
                elif line:
                    chunks.append(line.strip())
                line = fh.readline()
            if name is not None:
                yield FastxRecord(name, "".join(chunks), comment, None)
        elif line[0] == "@":
            while line:
                head = line.rstrip("\r\n")[1:].split(None, 1)
                seq = fh.readline().rstrip("\r\n")
                plus = fh.readline()
                qual = fh.readline().rstrip("\r\n")
                if not plus.startswith("+") or len(qual) != len(seq):
                    raise ValueError(f"{path}: malformed FASTQ record {head[0] if head else ''!r}")
                yield FastxRecord(head[0] if head else "", seq, head[1] if len(head) > 1 else None, qual)
                line = fh.readline()
                while line and not line.strip():
                    line = fh.readline()
        else:
            raise ValueError(f"{path}: neither FASTA ('>') nor FASTQ ('@')")


def _batch_arrays(patterns, texts):
    """[(bytes)...] x2 -> the (seq, p_off, p_len, t_off, t_len) layout of ``wfagpu_align_batch``."""
    p_len = np.fromiter((len(b) for b in patterns), np.int32, len(patterns))
    t_len = np.fromiter((len(b) for b in texts), np.int32, len(texts))
    rec = p_len.astype(np.int64) + t_len
    p_off = np.zeros(len(patterns), np.int64)
    np.cumsum(rec[:-1], out=p_off[1:])
    t_off = p_off + p_len
    seq = np.frombuffer(b"".join(x for pair in zip(patterns, texts) for x in pair) + b"\0", np.uint8)
    return seq, p_off, p_len, t_off, t_len


def align_fastx(aligner, texts_path, patterns_path=None, batch_size: int = 262144):
    """Align the records of ``texts_path`` against the records of ``patterns_path`` pairwise (or all
    against the aligner's cached pattern) and yield ``(names, BatchResult)`` per batch of
    ``batch_size`` pairs.  ``names[i]`` is ``(pattern_name, text_name)``."""
    texts = read_fastx(texts_path)
    patterns = read_fastx(patterns_path) if patterns_path is not None else None
    cached = None
    if patterns is None:
        if not aligner._pattern:
            raise ValueError("pattern is None")
        cached = aligner._pattern.upper().encode("ascii")
    while True:
        names, pb, tb = [], [], []
        for t in texts:
            if patterns is not None:
                p = next(patterns, None)
                if p is None:
                    raise ValueError("fewer pattern records than text records")
                names.append((p.name, t.name)); pb.append(p.sequence.upper().encode("ascii"))
            else:
                names.append((None, t.name)); pb.append(cached)
            tb.append(t.sequence.upper().encode("ascii"))
            if len(tb) >= batch_size:
                break
        if not tb:
            if patterns is not None and next(patterns, None) is not None:
                raise ValueError("fewer text records than pattern records")
            return
        yield names, aligner.align_arrays(*_batch_arrays(pb, tb))
        if len(tb) < batch_size:
            if patterns is not None and next(patterns, None) is not None:
                raise ValueError("fewer text records than pattern records")
            return
