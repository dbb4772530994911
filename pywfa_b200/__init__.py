"""pywfa_b200 -- B200-native batched wavefront aligner behind pywfa's API.

``from pywfa_b200 import WavefrontAligner`` is a drop-in for ``from pywfa import
WavefrontAligner`` on the gap-affine / gap-affine-2p path (re-exports mirror
``pywfa/__init__.py:1-6``), with the batched entry points ``align_batch`` / ``align_arrays``.
"""
from .align import (  # noqa: F401
    AlignmentResult,
    BatchResult,
    WavefrontAligner,
    cigartuples_to_str,
    clip_cigartuples,
    elide_mismatches_from_cigar,
)

from .fastx import FastxRecord, align_fastx, read_fastx, read_sam, read_seqs  # noqa: F401,E402

__all__ = ["WavefrontAligner", "AlignmentResult", "BatchResult", "clip_cigartuples",
           "cigartuples_to_str", "elide_mismatches_from_cigar", "read_fastx", "read_sam", "read_seqs", "align_fastx", "FastxRecord"]
