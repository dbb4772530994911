"""Seeded synthetic read-pair generator.

pattern = L iid uniform ACGT; text = pattern where each base independently with probability
``div`` receives one edit drawn uniformly from {substitution by one of the three other bases,
keep the base and insert one uniform random base after it, deletion}.  Vectorised numpy so that
10M-pair batches are generated in seconds.  Output layout is the batch layout of
``wfagpu_align_batch``: one uint8 ASCII buffer plus offset/length arrays.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", np.uint8)


def generate_pairs(n: int, length: int, div: float, seed: int = 1234,
                   text_flank: int = 0):
    """Return ``(seq, p_off, p_len, t_off, t_len)``.

    ``text_flank`` > 0 adds that many random bases on both sides of every text
    (config 3b: ends-free with ``text_begin_free = text_end_free = flank``)."""
    rng = np.random.default_rng(seed)
    pat = rng.integers(0, 4, size=(n, length), dtype=np.uint8)
    u = rng.random((n, length), dtype=np.float32)
    edit = u < div
    kind = rng.integers(0, 3, size=(n, length), dtype=np.uint8)   # 0 sub, 1 ins, 2 del
    sub = edit & (kind == 0)
    ins = edit & (kind == 1)
    dele = edit & (kind == 2)
    base = pat.copy()
    shift = rng.integers(1, 4, size=(n, length), dtype=np.uint8)
    base[sub] = (pat[sub] + shift[sub]) & 3
    extra = rng.integers(0, 4, size=(n, length), dtype=np.uint8)
    # every pattern position emits 0 (deleted), 1, or 2 (insertion after it) text bases
    emit = np.ones((n, length), np.int32)
    emit[dele] = 0
    emit[ins] = 2
    t_core = emit.sum(axis=1).astype(np.int64)
    flank = int(text_flank)
    t_len = (t_core + 2 * flank).astype(np.int32)
    p_len = np.full(n, length, np.int32)
    # layout: [pattern_0 | text_0 | pattern_1 | text_1 | ...]
    rec = p_len.astype(np.int64) + t_len.astype(np.int64)
    rec_off = np.zeros(n + 1, np.int64)
    np.cumsum(rec, out=rec_off[1:])
    p_off = rec_off[:-1].copy()
    t_off = p_off + length
    seq = np.empty(int(rec_off[-1]), np.uint8)
    # patterns
    pidx = (p_off[:, None] + np.arange(length, dtype=np.int64)[None, :]).ravel()
    seq[pidx] = _ACGT[pat.ravel()]
    # texts: position of the first emitted base of every pattern position
    csum = np.cumsum(emit, axis=1, dtype=np.int64)
    first = csum - emit + (t_off + flank)[:, None]
    keep = emit > 0
    seq[first[keep]] = _ACGT[base[keep]]
    seq[first[ins] + 1] = _ACGT[extra[ins]]
    if flank:
        fl = rng.integers(0, 4, size=(n, 2 * flank), dtype=np.uint8)
        lidx = (t_off[:, None] + np.arange(flank, dtype=np.int64)[None, :]).ravel()
        ridx = ((t_off + flank + t_core)[:, None] + np.arange(flank, dtype=np.int64)[None, :]).ravel()
        seq[lidx] = _ACGT[fl[:, :flank].ravel()]
        seq[ridx] = _ACGT[fl[:, flank:].ravel()]
    return seq, p_off, p_len, t_off, t_len


def pairs_from_strings(pairs):
    """[(pattern, text), ...] -> batch layout (upper-cased like pywfa/align.pyx:431-435)."""
    chunks, p_off, p_len, t_off, t_len = [], [], [], [], []
    pos = 0
    for p, t in pairs:
        pb = p.upper().encode("ascii")
        tb = t.upper().encode("ascii")
        p_off.append(pos); p_len.append(len(pb)); pos += len(pb)
        t_off.append(pos); t_len.append(len(tb)); pos += len(tb)
        chunks.append(pb); chunks.append(tb)
    seq = np.frombuffer(b"".join(chunks) + b"\0", np.uint8).copy()
    return (seq, np.array(p_off, np.int64), np.array(p_len, np.int32),
            np.array(t_off, np.int64), np.array(t_len, np.int32))
