"""GPU: the batch staging of wfagpu_align_batch -- raw upload, device-side 2-bit packing, length
buckets, pinned / device-resident / scattered inputs, pinned result arrays, concurrent callers.
Every case is checked bit-exactly against the CPU checker through the C ABI."""
import threading

import numpy as np
import pytest

from conftest import assert_same
from pywfa_b200 import _ffi
from pywfa_b200.synth import generate_pairs, pairs_from_strings

pytestmark = pytest.mark.gpu


def _rnd(rng, m):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, m))


def _mutate(rng, s, rate):
    out = []
    for c in s:
        if rng.random() < rate:
            k = rng.integers(0, 3)
            if k == 0:
                out.append(_rnd(rng, 1))
            elif k == 1:
                out.append(c + _rnd(rng, 1))
        else:
            out.append(c)
    return "".join(out)


def test_device_packer_every_alignment_and_length(gpu_ctx, oracle):
    """The device packer reads 16 bases per thread with two aligned 16-byte loads and a funnel shift:
    every start alignment (0..15) x every tail length (0..40 bases), upper and lower case, with the
    pairs separated by junk bytes that must never be read as bases."""
    rng = np.random.default_rng(11)
    chunks, p_off, p_len, t_off, t_len = [], [], [], [], []
    pos = 0
    for length in range(0, 41):
        for shift in range(16):
            p = _rnd(rng, length)
            t = _mutate(rng, p, 0.15)
            if (length + shift) & 1:
                p, t = p.lower(), t.lower()
            pad = "#" * ((shift - pos) % 16)                 # junk; also fixes the start alignment
            chunks.append(pad); pos += len(pad)
            p_off.append(pos); p_len.append(len(p)); chunks.append(p); pos += len(p)
            gap = "?" * int(rng.integers(0, 5))
            chunks.append(gap); pos += len(gap)
            t_off.append(pos); t_len.append(len(t)); chunks.append(t); pos += len(t)
    seq = np.frombuffer(("".join(chunks) + "#" * 16).encode(), np.uint8)
    batch = (seq, np.array(p_off, np.int64), np.array(p_len, np.int32), np.array(t_off, np.int64), np.array(t_len, np.int32))
    # the checker wants clean upper-case input: same pairs, re-laid out
    clean = pairs_from_strings([(bytes(seq[a:a + l]).decode().upper(), bytes(seq[b:b + m]).decode().upper())
                                for a, l, b, m in zip(p_off, p_len, t_off, t_len)])
    for kw in (dict(span="end-to-end"), dict(distance="affine2p")):
        cfg = oracle.make_config(**kw)
        want = oracle.align_batch(cfg, *clean, kind=oracle.checker_kind())
        got = gpu_ctx.align_batch(cfg, *batch)
        assert_same(got, want, what=f"packer alignments {kw}")


def test_pinned_inputs_and_outputs(gpu_ctx, oracle, monkeypatch):
    """Pinned caller memory is read and written by the copy engine directly (no staging copies):
    same results as pageable memory, single call and chunked pipeline, full and score-only."""
    batch = generate_pairs(30000, 150, 0.06, seed=78)
    pinned = tuple(_ffi.pinned_copy(a) for a in batch)
    n = len(batch[1])
    for kw in (dict(span="end-to-end"), dict(span="end-to-end", scope="score")):
        cfg = oracle.make_config(**kw)
        want = oracle.align_batch(cfg, *batch, kind=oracle.checker_kind())
        for chunk in (None, "7000"):
            if chunk:
                monkeypatch.setenv("WFAGPU_CHUNK", chunk)
            out = dict(score=_ffi.pinned_empty(n, np.int32), status=_ffi.pinned_empty(n, np.int32),
                       locs=_ffi.pinned_empty((n, 4), np.int32), cig_off=_ffi.pinned_empty(n + 1, np.int64))
            got = gpu_ctx.align_batch(cfg, *pinned, out=out)
            assert_same(got, want, scope_full=kw.get("scope", "full") == "full", what=f"pinned chunk={chunk} {kw}")
            # mixed: pinned bases, pageable arrays and results
            got = gpu_ctx.align_batch(cfg, pinned[0], *batch[1:])
            assert_same(got, want, scope_full=kw.get("scope", "full") == "full", what=f"pinned bases only chunk={chunk} {kw}")
            if chunk:
                monkeypatch.delenv("WFAGPU_CHUNK")


def test_device_resident_bases(gpu_ctx, oracle):
    """`seq` may be memory of the context's device (here a torch tensor): packed in place, no upload."""
    torch = pytest.importorskip("torch")
    batch = generate_pairs(5000, 200, 0.08, seed=79)
    dev = torch.from_numpy(batch[0].copy()).to("cuda:0")
    cfg = oracle.make_config(span="end-to-end")
    want = oracle.align_batch(cfg, *batch, kind=oracle.checker_kind())
    got = gpu_ctx.align_batch(cfg, dev, *batch[1:])
    assert_same(got, want, what="device-resident bases")
    bt = gpu_ctx.prepare(cfg, dev, *batch[1:])
    bt.run()
    assert_same(bt.fetch(), want, what="device-resident bases, prepared batch")
    bt.free()


def test_scattered_pairs_are_gathered(gpu_ctx, oracle):
    """One pattern against texts spread over a large buffer: the byte range the pairs span is far
    larger than their bases, so the host gathers them back to back instead of uploading the range."""
    rng = np.random.default_rng(12)
    genome = np.frombuffer(_rnd(rng, 3_000_000).encode(), np.uint8).copy()
    pattern = _rnd(rng, 180)
    n = 400
    t_off = np.sort(rng.integers(0, len(genome) - 400, n)).astype(np.int64)
    t_len = rng.integers(150, 220, n).astype(np.int32)
    for o in t_off[::3]:                                   # plant mutated copies so that alignments are not all junk
        m = _mutate(rng, pattern, 0.08)[:150].encode()
        genome[o:o + len(m)] = np.frombuffer(m, np.uint8)
    seq = np.concatenate([genome, np.frombuffer(pattern.encode(), np.uint8)])
    p_off = np.full(n, len(genome), np.int64)
    p_len = np.full(n, len(pattern), np.int32)
    clean = pairs_from_strings([(pattern, bytes(seq[o:o + l]).decode()) for o, l in zip(t_off, t_len)])
    for kw in (dict(span="end-to-end"), dict(text_begin_free=20, text_end_free=20)):
        cfg = oracle.make_config(**kw)
        want = oracle.align_batch(cfg, *clean, kind=oracle.checker_kind())
        got = gpu_ctx.align_batch(cfg, seq, p_off, p_len, t_off, t_len)
        assert_same(got, want, what=f"scattered {kw}")
        got = gpu_ctx.align_batch(cfg, _ffi.pinned_copy(seq), p_off, p_len, t_off, t_len)
        assert_same(got, want, what=f"scattered, pinned {kw}")


def test_mixed_length_batch_is_bucketed(gpu_ctx, oracle, monkeypatch):
    """150 bp + 1 kbp + 10 kbp pairs in one batch, shuffled: each length class gets its own tier plan
    (the short reads stay on the register tier instead of inheriting the 10 kbp plan); results are
    positional and bit-exact, with and without bucketing, one call and chunked."""
    rng = np.random.default_rng(13)
    parts = [generate_pairs(6000, 150, 0.05, seed=31), generate_pairs(2500, 1000, 0.08, seed=32),
             generate_pairs(40, 10000, 0.05, seed=33), generate_pairs(3000, 260, 0.10, seed=34)]
    pairs = []
    for seq, po, pl, to, tl in parts:
        buf = seq.tobytes().decode()
        pairs += [(buf[a:a + l], buf[b:b + m]) for a, l, b, m in zip(po, pl, to, tl)]
    order = rng.permutation(len(pairs))
    batch = pairs_from_strings([pairs[i] for i in order])
    for kw in (dict(span="end-to-end"), dict(span="end-to-end", scope="score"), dict(span="end-to-end", heuristic="adaptive")):
        cfg = oracle.make_config(**kw)
        want = oracle.align_batch(cfg, *batch, kind=oracle.checker_kind())
        bt = gpu_ctx.prepare(cfg, *batch)
        bt.run()
        got = bt.fetch()
        st = bt.stats()
        bt.free()
        assert_same(got, want, scope_full=kw.get("scope", "full") == "full", what=f"bucketed {kw}")
        monkeypatch.setenv("WFAGPU_NO_BUCKETS", "1")
        b1 = gpu_ctx.prepare(cfg, *batch)
        b1.run()
        got1 = b1.fetch()
        st1 = b1.stats()
        b1.free()
        monkeypatch.delenv("WFAGPU_NO_BUCKETS")
        assert_same(got1, want, scope_full=kw.get("scope", "full") == "full", what=f"one bucket {kw}")
        assert st["cells"] == st1["cells"]
        monkeypatch.setenv("WFAGPU_CHUNK", "3000")
        got2 = gpu_ctx.align_batch(cfg, *batch)
        monkeypatch.delenv("WFAGPU_CHUNK")
        assert_same(got2, want, scope_full=kw.get("scope", "full") == "full", what=f"bucketed, chunked {kw}")


def test_two_threads_share_a_context(gpu_ctx, oracle):
    """pywfa aligners are independent objects, one per thread is fine there; here all aligners of a
    device share one context, so calls from several threads are serialised inside the library."""
    import pywfa_b200
    batches = [generate_pairs(4000, 150, 0.05 + 0.01 * i, seed=90 + i) for i in range(4)]
    cfg = oracle.make_config(span="end-to-end")
    wants = [oracle.align_batch(cfg, *b, kind=oracle.checker_kind()) for b in batches]
    errors = []

    def work(i):
        try:
            for _ in range(3):
                got = gpu_ctx.align_batch(cfg, *batches[i])
                assert_same(got, wants[i], what=f"thread {i}")
            a = pywfa_b200.WavefrontAligner("TCTTTACTCGCGCGTTGGAGAAATACAATAGT")
            for _ in range(20):
                assert a("TCTATACTGCGCGTTTGGAGAAATAAAATAGT").cigarstring == "3M1X4M1D7M1I9M1X6M"
        except Exception as e:          # surfaced after the join
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]


def test_negative_length_and_free_end_errors(gpu_ctx, oracle):
    batch = list(generate_pairs(100, 50, 0.1, seed=3))
    bad = batch[2].copy(); bad[17] = -1
    with pytest.raises(_ffi.WfaGpuError, match="negative length at pair 17"):
        gpu_ctx.align_batch(oracle.make_config(), batch[0], batch[1], bad, batch[3], batch[4], check=False)
    short = batch[4].copy(); short[40] = 3
    with pytest.raises(_ffi.WfaGpuError, match="pair 40"):
        gpu_ctx.align_batch(oracle.make_config(text_end_free=5), batch[0], batch[1], batch[2], batch[3], short)
    # the context stays usable after an error
    cfg = oracle.make_config(span="end-to-end")
    assert_same(gpu_ctx.align_batch(cfg, *batch), oracle.align_batch(cfg, *batch, kind=oracle.checker_kind()), what="after errors")


def test_single_pair_entry_point(gpu_ctx, oracle):
    """wfagpu_align_pair: the one-launch mailbox path (short gap-affine pairs without cut-offs) and its
    fall-backs to a batch of one (non-ACGT bytes, other distances / cut-offs, long or very divergent pairs)
    give what the checker gives for the same pairs."""
    rng = np.random.default_rng(14)
    pairs = [("", ""), ("ACGT", ""), ("", "ACGT"), ("A", "A"), ("A", "C"), ("ACGT" * 250, "ACGT" * 250)]
    for _ in range(120):
        lp = int(rng.integers(1, 400))
        p = _rnd(rng, lp)
        t = _mutate(rng, p, float(rng.choice([0.02, 0.1, 0.3]))) if rng.random() < 0.8 else _rnd(rng, int(rng.integers(1, 400)))
        pairs.append((p, t))
    pairs.append((_rnd(rng, 999), _rnd(rng, 1000)))                 # unrelated: wider than the 256-diagonal window
    pairs.append((_rnd(rng, 1500), _mutate(rng, _rnd(rng, 1500), 0.1)))   # longer than the mailbox takes
    pairs.append(("ACGTNACGTTTGA", "ACGTAACGTTTGA"))                # N: byte mode
    pairs.append(("acgtacgtacgt", "ACGTACCTACGT"))                  # lower case packs alike
    batch = pairs_from_strings(pairs)
    for kw in (dict(span="end-to-end"), dict(), dict(span="end-to-end", scope="score"),
               dict(pattern_begin_free=0, pattern_end_free=0, text_begin_free=0, text_end_free=0, match=-1, span="end-to-end"),
               dict(distance="affine2p"), dict(heuristic="adaptive", span="end-to-end"),
               dict(span="end-to-end", max_steps=30), dict(span="end-to-end", wildcard="N"), dict(span="end-to-end", wildcard="A")):
        cfg = oracle.make_config(**kw)
        want = oracle.align_batch(cfg, *batch, kind=oracle.checker_kind())
        full = kw.get("scope", "full") == "full"
        for i, (p, t) in enumerate(pairs):
            score, status, locs, runs = gpu_ctx.align_pair(cfg, p.encode(), t.encode())
            assert (score, status) == (int(want["score"][i]), int(want["status"][i])), (kw, i)
            if full:
                assert runs == want["runs"][want["cig_off"][i]:want["cig_off"][i + 1]].tolist(), (kw, i)
                assert locs == want["locs"][i].tolist(), (kw, i)
            else:
                assert runs == []
    # a wildcard that is not a base leaves clean pairs on the one-launch path; a pair holding it takes the batch path
    cfg = oracle.make_config(span="end-to-end", wildcard="N")
    assert gpu_ctx.align_pair(cfg, b"ACGTTACGTTTGA", b"ACGTAACGTTTGA")[0] == -4 and gpu_ctx.last_launches() == 1
    assert gpu_ctx.align_pair(cfg, b"ACGTNACGTTTGA", b"ACGTAACGTTTGA")[0] == 0 and gpu_ctx.last_launches() > 1
    # the pywfa surface rides on it
    import pywfa_b200
    a = pywfa_b200.WavefrontAligner("TCTTTACTCGCGCGTTGGAGAAATACAATAGT")
    for _ in range(3):
        r = a("TCTATACTGCGCGTTTGGAGAAATAAAATAGT")
        assert (r.score, r.status, r.cigarstring) == (-24, 0, "3M1X4M1D7M1I9M1X6M")
    assert gpu_ctx.last_launches() >= 0
