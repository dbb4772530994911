"""FASTA / FASTQ / SAM streaming into the batched entry point.

The reference's own tests read their fixtures with ``pysam.FastxFile`` and align record by record
(``pywfa/tests/test.py:198-232``); here records are parsed with the standard library only (plain or
gzip, FASTA with wrapped lines, 4-line FASTQ) and handed to ``WavefrontAligner.align_arrays`` in
batches, so that reading batch ``i+1`` overlaps nothing on the GPU but at least never builds
Python strings per base.  ``FastxRecord`` mirrors the fields of pysam's proxy that the reference's
tests use (``name``, ``sequence``, ``comment``, ``quality``).  Text SAM is read as well (``read_sam``:
QNAME / SEQ / QUAL of every alignment line, reverse-strand records restored to the read's original
orientation on request), so that reads can be re-aligned straight from a mapper's output; ``read_seqs``
picks the parser from the first byte / the header.
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


_COMP = bytes.maketrans(b"ACGTNacgtnRYKMBVDHrykmbvdh", b"TGCANtgcanYRMKVBHDyrmkvbhd")


def read_sam(path, original_orientation: bool = False) -> Iterator[FastxRecord]:
    """Yield the reads of a text SAM file (plain or gzip): ``name`` = QNAME, ``sequence`` = SEQ, ``quality`` =
    QUAL, ``comment`` = ``"FLAG RNAME POS CIGAR"``.  Header lines (``@``) and records without a stored
    sequence (``*``) are skipped.  SAM stores reverse-strand alignments reverse-complemented;
    ``original_orientation=True`` undoes that (FLAG 0x10), which is what re-alignment against another
    reference wants."""
    with _open(path) as fh:
        for line in fh:
            if not line.strip() or line[0] == "@":
                continue
            f = line.rstrip("\r\n").split("\t")
            if len(f) < 11:
                raise ValueError(f"{path}: SAM line with {len(f)} fields")
            seq, qual = f[9], f[10]
            if seq == "*":
                continue
            if original_orientation and int(f[1]) & 0x10:
                seq = seq.encode("ascii").translate(_COMP)[::-1].decode("ascii")
                qual = qual[::-1] if qual != "*" else qual
            yield FastxRecord(f[0], seq, " ".join((f[1], f[2], f[3], f[5])), None if qual == "*" else qual)


def read_seqs(path, **kw) -> Iterator[FastxRecord]:
    """FASTA, FASTQ or SAM, told apart by content: ``>`` = FASTA; ``@`` followed by a two-letter SAM header
    tag and a tab (``@HD\t``, ``@SQ\t``, ...) or a line with >= 11 tab-separated fields = SAM; else FASTQ."""
    with _open(path) as fh:
        first = fh.readline()
        while first and not first.strip():
            first = fh.readline()
    is_sam = (len(first) > 3 and first[0] == "@" and first[3] == "\t") or (first[:1] not in (">", "@") and first.count("\t") >= 10)
    return read_sam(path, **kw) if is_sam else read_fastx(path)


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
    texts = read_seqs(texts_path)
    patterns = read_seqs(patterns_path) if patterns_path is not None else None
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
