#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py
It imports the reference's own Python module (oracle/_ref/pywfa, cythonized from
/root/reference/pywfa/align.pyx and linked against the reference WFA2-lib objects) and records
  * reference_kat.json : every known-answer case of the reference's test-suite
    (pywfa/tests/test.py) and README, incl. the FASTA fixtures, with the reference's outputs;
  * postprocess.json   : clip_cigartuples / elide_mismatches_from_cigar / cigartuples_to_str
    outputs of the reference for a set of inputs;
  * synthetic.json     : seeded synthetic batches per configuration (inputs are regenerated from
    the seed by pywfa_b200.synth.generate_pairs) with the reference's score/status/CIGAR/cells;
  * metrics.json       : the same three kinds for distance = linear / levenshtein / indel
    (`--metrics` regenerates this file only).
The GPU box has no /root/reference: tests only read the committed JSON.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
REF_TESTS = "/root/reference/pywfa/tests"

from pywfa import WavefrontAligner, cigartuples_to_str, clip_cigartuples, elide_mismatches_from_cigar  # noqa: E402
from pywfa.align import AlignmentResult  # noqa: E402

from oracle import oracle_py  # noqa: E402
from pywfa_b200.synth import generate_pairs  # noqa: E402


def read_fasta(path):
    recs, name, chunks = [], None, []
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith(">"):
            if name is not None:
                recs.append((name, "".join(chunks)))
            name, chunks = ln[1:].split()[0], []
        elif ln:
            chunks.append(ln)
    if name is not None:
        recs.append((name, "".join(chunks)))
    return recs


def run_case(name, ctor, pattern, text, call=None, source=""):
    call = call or {}
    a = WavefrontAligner(**ctor)
    res = a(text, pattern, **call)
    try:
        ap, at = res.aligned_pattern, res.aligned_text
    except Exception:                      # the reference's helper can raise on odd CIGARs
        ap = at = None
    return dict(name=name, source=source, ctor=ctor, pattern=pattern, text=text, call=call,
                expect=dict(score=int(res.score), status=int(res.status),
                            cigartuples=[[int(o), int(l)] for o, l in res.cigartuples],
                            cigarstring=res.cigarstring,
                            locations=[int(res.pattern_start), int(res.pattern_end), int(res.text_start), int(res.text_end)],
                            aligner_score=int(a.score), aligner_status=int(a.status),
                            aligner_cigarstring=a.cigarstring,
                            aligned_equal=(ap == at) if ap is not None else None))


def kat_cases():
    out = []
    P, T = "TCTTTACTCGCGCGTTGGAGAAATACAATAGT", "TCTATACTGCGCGTTTGGAGAAATAAAATAGT"
    out.append(run_case("readme_affine_default", {}, P, T, source="pywfa/tests/test.py:16-46, README.rst:34-42"))
    out.append(run_case("affine3", {}, "TCTATACTGCGCGTTTGGAGAAATAAAA", "TCTCCCCATACTGCGCGTTTGGAGAAATAAAA",
                        dict(clip_cigar=False), "pywfa/tests/test.py:47-51"))
    out.append(run_case("scope_score", dict(scope="score"), P, T, source="pywfa/tests/test.py:54-63"))
    out.append(run_case("supress_seqs_full", dict(scope="full"), P, T, dict(supress_sequences=True), "pywfa/tests/test.py:65-83"))
    P2, T2 = "AATTAATTTAAGTCTAGGCTACTTTCGGTACTTTGTTCTT", "AATTTAAGTCTAGGCTACTTTCGGTACTTTCTT"
    out.append(run_case("end_to_end", dict(span="end-to-end", mismatch=4, gap_opening=6, gap_extension=2), P2, T2,
                        source="pywfa/tests/test.py:94-102"))
    out.append(run_case("ends_free_clip_elide", dict(span="ends-free", mismatch=4, gap_opening=6, gap_extension=2), P2, T2,
                        dict(clip_cigar=True, elide_mismatches=True, min_aligned_bases_left=5, min_aligned_bases_right=5),
                        "pywfa/tests/test.py:104-113 (fails as shipped: SURVEY.md 0.2)"))
    ef = dict(span="ends-free", mismatch=4, gap_opening=6, gap_extension=2)
    for i, (p, t) in enumerate([
            ("AAAAACCTTTTTAAAAAA", "GGCCAAAAACCAAAAAA"), ("AAAAACCTTTTTAAAAAA", "GGCCAAAAACCGGGGGGG"),
            ("AAAAACCGGGG", "AAAAACC"), ("AAAAACC", "AAAAACCGGGG"), ("GGGGAAAAACC", "AAAAACCGGGG"),
            ("AAAAACCGGGG", "GGGGAAAAACC"), ("GGGGAAAAACC", "AAAAACC"), ("GGGGAAAAACC", "CCCCCAAAAACC"),
            ("GGGGAAAAACCGGGGG", "CCCCCAAAAACCTTTTT"), ("AAAAACC", "CCCCCAAAAACCTTTTT")]):
        out.append(run_case(f"ends_free2_{i}", ef, p, t, source="pywfa/tests/test.py:115-178"))
    for h in ("X-drop", "adaptive"):
        out.append(run_case(f"heuristic_{h}", dict(distance="affine", mismatch=4, gap_opening=6, gap_extension=2, heuristic=h),
                            "AAAAACCAAAAAA", "GGCCAAAAACCAAAAAA", source="pywfa/tests/test.py:180-194"))
        out.append(run_case(f"heuristic_swapped_{h}", dict(heuristic=h), "GGCCAAAAACCAAAAAA", "AAAAACCTTTTTAAAAAA",
                            source="SURVEY.md section 4 table"))
    reads = read_fasta(os.path.join(REF_TESTS, "long.fa"))
    refs = read_fasta(os.path.join(REF_TESTS, "long.reference.fa"))
    for (rn, rs), (_, fs) in zip(reads, refs):
        text, pattern = rs.upper(), fs.upper()
        lt, lp = len(text) // 2, len(pattern) // 2
        out.append(run_case(f"long_{rn}", dict(distance="affine", mismatch=4, gap_opening=6, gap_extension=2,
                                               pattern_begin_free=lp, pattern_end_free=lp, text_begin_free=lt, text_end_free=lt),
                            pattern, text, dict(clip_cigar=True), "pywfa/tests/test.py:196-212"))
    reads = read_fasta(os.path.join(REF_TESTS, "short.fa"))
    refs = read_fasta(os.path.join(REF_TESTS, "short.reference.fa"))
    for (rn, rs), (_, fs) in zip(reads, refs):
        text, pattern = rs.upper(), fs.upper()
        out.append(run_case(f"short_{rn}", dict(mismatch=5, gap_opening=6, gap_extension=2), pattern, text,
                            source="pywfa/tests/test.py:214-221"))
        out.append(run_case(f"short2p_{rn}", dict(distance="affine2p", mismatch=5, gap_opening=6, gap_extension=2), pattern, text,
                            dict(clip_cigar=True, elide_mismatches=True), "pywfa/tests/test.py:223-232"))
        out.append(run_case(f"short_e2e_{rn}", dict(span="end-to-end", mismatch=5, gap_opening=6, gap_extension=2), pattern, text,
                            source="derived from pywfa/tests/test.py:214-221"))
    out.append(run_case("acgt_vs_empty", dict(span="end-to-end"), "ACGT", "", source="SURVEY.md section 4"))
    out.append(run_case("empty_vs_acgt", dict(span="end-to-end"), "", "ACGT", source="derived"))
    out.append(run_case("max_steps_10", dict(span="end-to-end", max_steps=10), P2, "GGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGG", source="SURVEY.md section 4"))
    out.append(run_case("match_minus1", dict(span="end-to-end", match=-1), P, T, source="SURVEY.md section 4"))
    out.append(run_case("lowercase", {}, P.lower(), T.lower(), source="pywfa/align.pyx:431-435 upper()"))
    return out


def postprocess_cases():
    out = []
    rng = np.random.default_rng(11)
    cts = [[(0, 3), (8, 1), (0, 4), (2, 1), (0, 7), (1, 1), (0, 9), (8, 1), (0, 6)],
           [(1, 4), (0, 7), (2, 5), (0, 6)], [(2, 4)], [(0, 2), (8, 1), (0, 1), (1, 3), (0, 12), (2, 2), (0, 1)],
           [(8, 2), (0, 30), (8, 1)], []]
    for _ in range(40):
        n = int(rng.integers(1, 12))
        ct, last = [], -1
        for _ in range(n):
            op = int(rng.choice([0, 1, 2, 8]))
            if op == last:
                continue
            ct.append((op, int(rng.integers(1, 12)))); last = op
        cts.append(ct)
    for ct in cts:
        pl = sum(l for o, l in ct if o in (0, 2, 8)); tl = sum(l for o, l in ct if o in (0, 1, 8))
        for (ml, mr, ts) in ((5, 5, 0), (1, 1, 0), (3, 8, 2)):
            res = AlignmentResult(pl, tl, 0, pl, ts, tl, list(ct), -1, "", "", 0)
            res = clip_cigartuples(res, ml, mr)
            out.append(dict(kind="clip", cigartuples=[list(c) for c in ct], pattern_length=pl, text_length=tl,
                            text_start=ts, left=ml, right=mr,
                            expect=dict(cigartuples=[list(c) for c in res.cigartuples],
                                        locations=[res.pattern_start, res.pattern_end, res.text_start, res.text_end])))
        out.append(dict(kind="elide", cigartuples=[list(c) for c in ct],
                        expect=[list(c) for c in elide_mismatches_from_cigar(list(ct))]))
        out.append(dict(kind="str", cigartuples=[list(c) for c in ct], expect=cigartuples_to_str(list(ct))))
    return out


SYNTH = [
    ("cfg1", dict(span="end-to-end"), 256, 150, 0.05, 0),
    ("cfg1-endsfree0", dict(), 128, 150, 0.05, 0),
    ("cfg2", dict(span="end-to-end", scope="score"), 256, 250, 0.10, 0),
    ("cfg3a", dict(distance="affine2p"), 24, 1000, 0.10, 0),
    ("cfg3b", dict(distance="affine2p", text_begin_free=50, text_end_free=50), 24, 1000, 0.10, 50),
    ("cfg4-adaptive", dict(span="end-to-end", heuristic="adaptive"), 6, 10000, 0.15, 0),
    ("cfg4-xdrop-full", dict(span="end-to-end", heuristic="X-drop", xdrop=20), 8, 10000, 0.15, 0),
    ("cfg4-xdrop-score", dict(span="end-to-end", heuristic="X-drop", xdrop=20, scope="score"), 8, 10000, 0.15, 0),
    ("cfg4-none-2kbp", dict(span="end-to-end"), 6, 2000, 0.15, 0),
    ("cfg5-proxy-2kbp", dict(distance="affine2p", span="end-to-end"), 4, 2000, 0.20, 0),
    ("match-1", dict(span="end-to-end", match=-1), 128, 150, 0.1, 0),
    ("max-steps", dict(span="end-to-end", max_steps=10), 64, 150, 0.1, 0),
    ("endsfree-4", dict(pattern_begin_free=10, pattern_end_free=20, text_begin_free=5, text_end_free=7), 128, 150, 0.1, 4),
]


def synthetic_cases():
    out = []
    for i, (name, kw, n, length, div, flank) in enumerate(SYNTH):
        seed = 4000 + i
        batch = generate_pairs(n, length, div, seed, text_flank=flank)
        cfg = oracle_py.make_config(**kw)
        r = oracle_py.align_batch(cfg, *batch, kind="reference")
        cig = [oracle_py.runs_to_cigarstring(r["runs"][r["cig_off"][j]:r["cig_off"][j + 1]]) for j in range(n)]
        out.append(dict(name=name, config=kw, n=n, length=length, div=div, flank=flank, seed=seed,
                        input_checksum=int(np.frombuffer(batch[0].tobytes(), np.uint8).astype(np.uint64).sum()),
                        score=r["score"].tolist(), status=r["status"].tolist(), cigars=cig,
                        locations=r["locs"].tolist(), cells=r["cells"].tolist()))
    return out


METRIC_SYNTH = [
    ("linear-e2e", dict(distance="linear", span="end-to-end"), 128, 150, 0.08, 0),
    ("linear-endsfree", dict(distance="linear", pattern_begin_free=5, pattern_end_free=8, text_begin_free=6, text_end_free=9), 96, 150, 0.1, 6),
    ("linear-match-2", dict(distance="linear", span="end-to-end", match=-2, mismatch=4, gap_extension=3), 96, 150, 0.1, 0),
    ("linear-score", dict(distance="linear", span="end-to-end", scope="score", mismatch=2, gap_extension=5), 128, 250, 0.1, 0),
    ("linear-xdrop", dict(distance="linear", span="end-to-end", heuristic="X-drop", xdrop=30, steps_between_cutoffs=2), 64, 400, 0.15, 0),
    ("linear-1kbp", dict(distance="linear", span="end-to-end", mismatch=6, gap_extension=4), 12, 1000, 0.1, 0),
    ("edit-e2e", dict(distance="levenshtein", span="end-to-end"), 128, 150, 0.08, 0),
    ("edit-endsfree", dict(distance="levenshtein", text_begin_free=20, text_end_free=20), 96, 150, 0.1, 20),
    ("edit-score", dict(distance="levenshtein", span="end-to-end", scope="score"), 128, 250, 0.1, 0),
    ("edit-adaptive", dict(distance="levenshtein", span="end-to-end", heuristic="adaptive", min_wavefront_length=5,
                           max_distance_threshold=10, steps_between_cutoffs=2), 64, 400, 0.15, 0),
    ("edit-max-steps", dict(distance="levenshtein", span="end-to-end", max_steps=12), 64, 150, 0.1, 0),
    ("edit-2kbp", dict(distance="levenshtein", span="end-to-end"), 8, 2000, 0.15, 0),
    ("indel-e2e", dict(distance="indel", span="end-to-end"), 128, 150, 0.08, 0),
    ("indel-endsfree", dict(distance="indel", pattern_end_free=10, text_end_free=10), 96, 150, 0.1, 0),
    ("indel-score", dict(distance="indel", span="end-to-end", scope="score"), 128, 250, 0.1, 0),
    ("indel-1kbp", dict(distance="indel", span="end-to-end"), 12, 1000, 0.1, 0),
]


def metric_cases():
    """gap-linear / edit / indel: the reference's own Python object on its README / test pairs, seeded batches
    through the reference library, and the pairs on which levenshtein's exact pruning fires (sequences of very
    different lengths: wavefronts of >= 1000 diagonals, compute_edit.c:219-275)."""
    kat = []
    P, T = "TCTTTACTCGCGCGTTGGAGAAATACAATAGT", "TCTATACTGCGCGTTTGGAGAAATAAAATAGT"
    P2, T2 = "AATTAATTTAAGTCTAGGCTACTTTCGGTACTTTGTTCTT", "AATTTAAGTCTAGGCTACTTTCGGTACTTTCTT"
    for d in ("linear", "levenshtein", "indel"):
        kat.append(run_case(f"{d}_readme", dict(distance=d), P, T, source="pywfa/tests/test.py:16-46 with another metric"))
        kat.append(run_case(f"{d}_e2e", dict(distance=d, span="end-to-end"), P2, T2, source="pywfa/tests/test.py:94-102 with another metric"))
        kat.append(run_case(f"{d}_score", dict(distance=d, scope="score"), P, T, source="pywfa/tests/test.py:54-63 with another metric"))
        kat.append(run_case(f"{d}_endsfree", dict(distance=d, pattern_begin_free=4, pattern_end_free=4, text_begin_free=4, text_end_free=4),
                            "GGGGAAAAACCGGGGG", "CCCCCAAAAACCTTTTT", source="pywfa/tests/test.py:115-178 with another metric"))
        kat.append(run_case(f"{d}_vs_empty", dict(distance=d, span="end-to-end"), "ACGT", "", source="derived"))
    kat.append(run_case("linear_match_minus1", dict(distance="linear", span="end-to-end", match=-1, mismatch=3, gap_extension=2), P, T,
                        source="derived (penalties.c:78-82)"))
    # memory_mode="biwfa", scope="score": BiWFA is exact, same score / status as the other memory modes
    for nm, kw in (("affine", {}), ("affine2p", dict(distance="affine2p")), ("match_minus1", dict(match=-1)),
                   ("levenshtein", dict(distance="levenshtein")), ("linear", dict(distance="linear"))):
        kat.append(run_case(f"biwfa_score_{nm}", dict(memory_mode="biwfa", scope="score", span="end-to-end", **kw), P, T,
                            source="pywfa/align.pyx:386-387 (wavefront_memory_ultralow), score only"))
        kat.append(run_case(f"biwfa_score_{nm}_2", dict(memory_mode="biwfa", scope="score", span="end-to-end", **kw), P2, T2,
                            source="pywfa/align.pyx:386-387 (wavefront_memory_ultralow), score only"))
    syn = []
    for i, (name, kw, n, length, div, flank) in enumerate(METRIC_SYNTH):
        seed = 5000 + i
        batch = generate_pairs(n, length, div, seed, text_flank=flank)
        cfg = oracle_py.make_config(**kw)
        r = oracle_py.align_batch(cfg, *batch, kind="reference")
        cig = [oracle_py.runs_to_cigarstring(r["runs"][r["cig_off"][j]:r["cig_off"][j + 1]]) for j in range(n)]
        syn.append(dict(name=name, config=kw, n=n, length=length, div=div, flank=flank, seed=seed,
                        input_checksum=int(np.frombuffer(batch[0].tobytes(), np.uint8).astype(np.uint64).sum()),
                        score=r["score"].tolist(), status=r["status"].tolist(), cigars=cig,
                        locations=r["locs"].tolist(), cells=r["cells"].tolist()))
    # exact pruning: unrelated and repeat-derived pairs of very different lengths (inputs from the seed)
    import hashlib
    rng = np.random.default_rng(5)
    rs = lambda m: "".join("ACGT"[i] for i in rng.integers(0, 4, m))      # noqa: E731
    pairs = []
    for pl, tl in ((300, 3000), (3000, 300), (1500, 2500), (100, 2500)):
        p = rs(pl)
        pairs += [(p, rs(tl)), (p, (p * (tl // pl + 1))[:tl])]
    from pywfa_b200.synth import pairs_from_strings
    batch = pairs_from_strings(pairs)
    prune = []
    for kw in (dict(distance="levenshtein", span="end-to-end"), dict(distance="levenshtein", span="end-to-end", scope="score")):
        r = oracle_py.align_batch(oracle_py.make_config(**kw), *batch, kind="reference")
        prune.append(dict(config=kw, seed=5, score=r["score"].tolist(), status=r["status"].tolist(), cells=r["cells"].tolist(),
                          cigar_sha256=hashlib.sha256(r["runs"].tobytes()).hexdigest(), n_runs=int(len(r["runs"]))))
    return dict(kat=kat, synthetic=syn, prune=prune)


if __name__ == "__main__":
    if "--metrics" in sys.argv:
        json.dump(metric_cases(), open(os.path.join(HERE, "metrics.json"), "w"))
        print("metrics.json", os.path.getsize(os.path.join(HERE, "metrics.json")), "bytes")
        sys.exit(0)
    json.dump(kat_cases(), open(os.path.join(HERE, "reference_kat.json"), "w"), indent=1)
    json.dump(postprocess_cases(), open(os.path.join(HERE, "postprocess.json"), "w"))
    json.dump(synthetic_cases(), open(os.path.join(HERE, "synthetic.json"), "w"))
    json.dump(metric_cases(), open(os.path.join(HERE, "metrics.json"), "w"))
    for f in ("reference_kat.json", "postprocess.json", "synthetic.json", "metrics.json"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
