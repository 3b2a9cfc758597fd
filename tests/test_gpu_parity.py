"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle and the committed
reference outputs.  Everything here is bit-exact (integer / byte work)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from fuzz_util import fuzz_case
from golden_util import SUFS, args_to_kw, check_output, load_inputs, manifest, stdout_value
from oracle import oracle as O
from raft_b200 import api, synth
from sim_util import sim_dataset

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = manifest()
WHICH = {"coverage.txt": api.OUT_COVERAGE, "long_repeats.txt": api.OUT_LONG_REPEATS, "long_repeats.bed": api.OUT_BED,
         "reads.fasta": api.OUT_READS_FASTA}


def gpu_run(reads, paf, params, chunks=None):
    ctx = api.Context(params)
    ctx.set_reads(np.ascontiguousarray(reads.seq_off, np.int64), np.ascontiguousarray(reads.seq, np.uint8),
                  np.ascontiguousarray(reads.name_off, np.int64), np.ascontiguousarray(reads.names, np.uint8))
    if chunks is None:
        ctx.ingest_paf(np.frombuffer(paf, np.uint8) if paf else np.zeros(0, np.uint8), len(paf), last=True)
    else:
        pos = 0
        for k, c in enumerate(chunks):
            piece = np.frombuffer(paf[pos:pos + c], np.uint8)
            ctx.ingest_paf(piece, len(piece), last=(k == len(chunks) - 1))
            pos += c
        assert pos == len(paf)
    return ctx, ctx.run()


def compare_all(ctx, st, ref, real=True):
    assert st.n_records == ref.n_rec
    assert st.symmetric == ref.symmetric and st.high_cov == ref.high_cov and st.real_reads == ref.real_reads
    for tab, name in ((api.TAB_QID, "qid"), (api.TAB_TID, "tid"), (api.TAB_QS, "qs"), (api.TAB_QE, "qe"),
                      (api.TAB_TS, "ts"), (api.TAB_TE, "te"), (api.TAB_STRAND, "strand")):
        np.testing.assert_array_equal(ctx.table(tab), getattr(ref, name), err_msg=name)
    np.testing.assert_array_equal(ctx.table(api.TAB_BIN_OFF), ref.bin_off)
    np.testing.assert_array_equal(ctx.table(api.TAB_COV), ref.cov)
    np.testing.assert_array_equal(ctx.table(api.TAB_REP_OFF), ref.rep_off)
    np.testing.assert_array_equal(ctx.table(api.TAB_REP).reshape(-1, 2), np.stack([ref.rep_s, ref.rep_e], 1))
    np.testing.assert_array_equal(ctx.table(api.TAB_FRAG).reshape(-1, 3), np.stack([ref.frag_read, ref.frag_a, ref.frag_b], 1))
    assert (st.total_cov, st.total_windows, st.total_repeat_len, st.total_read_len) == (
        ref.total_cov, ref.total_windows, ref.total_repeat_len, ref.total_read_len)
    assert ctx.fetch(api.OUT_COVERAGE) == ref.cov_txt
    assert ctx.fetch(api.OUT_LONG_REPEATS) == ref.rep_txt
    assert ctx.fetch(api.OUT_READS_FASTA) == ref.fasta
    assert ctx.fetch(api.OUT_BED) == ref.bed_txt
    for which, data in ((api.OUT_COVERAGE, ref.cov_txt), (api.OUT_LONG_REPEATS, ref.rep_txt), (api.OUT_READS_FASTA, ref.fasta), (api.OUT_BED, ref.bed_txt)):
        assert ctx.digest(which) == O.digest(data)


@pytest.mark.parametrize("entry", CASES, ids=lambda e: e["name"])
def test_golden_through_c_abi(entry):
    """Outputs equal the reference binary's committed outputs (tests/golden)."""
    fa, paf = load_inputs(entry)
    reads = O.parse_fasta(fa)
    kw = args_to_kw(entry["args"])
    ctx, st = gpu_run(reads, paf, api.AlgoParams(**kw))
    for suf in SUFS:
        check_output(entry, suf, ctx.fetch(WHICH[suf]))
    assert stdout_value(entry, "Symmetric overlaps") == f"INFO, Symmetric overlaps {st.symmetric} "
    assert stdout_value(entry, "length of alignments") == f"INFO, length of alignments  {st.n_records}()"
    assert stdout_value(entry, "high_cov") == f"high_cov {st.high_cov}"
    compare_all(ctx, st, O.run(reads, paf, O.make_params(**kw)))
    ctx.close()


@pytest.mark.parametrize("seed", [5, 6])
def test_simulated_read_mode(seed):
    """Header variant and long_repeats.bed for simulated reads (chop.hpp:252-258,293-310; repeat.hpp:187-199)."""
    reads, paf = sim_dataset(seed)
    kw = dict(est_cov=30, repeat_length=4000, read_length=8000, flanking_length=300, overlap_length=200)
    ref = O.run(reads, paf, O.make_params(**kw))
    assert ref.status == 0 and ref.real_reads == 0 and len(ref.bed_txt) > 0 and ref.n_frag > reads.n
    ctx, st = gpu_run(reads, paf, api.AlgoParams(**kw))
    compare_all(ctx, st, ref)
    total = ctx.output_size(api.OUT_BED)
    assert ctx.fetch(api.OUT_BED, 3, min(1000, total - 3)) == ref.bed_txt[3:3 + min(1000, total - 3)]
    ctx.close()


@pytest.mark.parametrize("cfg,scale,sym", [("C1", 0.3, True), ("C1", 0.3, False), ("C2", 0.0004, True), ("C2", 0.0004, False),
                                           ("C4", 0.002, True), ("C4", 0.002, False), ("C5", 0.004, True), ("C5", 0.004, False)])
def test_configs_vs_oracle(cfg, scale, sym):
    ds = synth.make_dataset(cfg, scale, sym, seed=777)
    p = api.AlgoParams.from_args(ds.args)
    ctx, st = gpu_run(ds.reads, ds.paf, p)
    ref = O.run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    assert ref.status == 0
    compare_all(ctx, st, ref)
    ctx.close()


def test_chunked_ingest_and_windowed_fetch():
    ds = synth.make_dataset("C1", 0.1, False, seed=5)
    p = api.AlgoParams.from_args(ds.args)
    n = len(ds.paf)
    # chunk boundaries in the middle of lines, a 1-byte chunk, and an empty final chunk
    chunks = [1000, 1, 77777, n // 3, 0]
    chunks.append(n - sum(chunks))
    chunks.append(0)
    ctx, st = gpu_run(ds.reads, ds.paf, p, chunks=chunks)
    ref = O.run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    compare_all(ctx, st, ref)
    # windows at odd offsets / sizes reproduce the same bytes
    for which, data in ((api.OUT_COVERAGE, ref.cov_txt), (api.OUT_LONG_REPEATS, ref.rep_txt), (api.OUT_READS_FASTA, ref.fasta)):
        total = ctx.output_size(which)
        assert total == len(data)
        for off, ln in ((0, 1), (1, 17), (total // 2 + 3, 100001), (total - 5, 5), (min(12345, total // 3), 65536 + 7)):
            ln = min(ln, total - off)
            assert ctx.fetch(which, off, ln) == data[off:off + ln], (which, off, ln)
        with pytest.raises(api.RaftError):
            ctx.fetch(which, total - 1, 2)
    ctx.close()


def test_device_resident_inputs_and_outputs():
    torch = pytest.importorskip("torch")
    ds = synth.make_dataset("C2", 0.0003, True, seed=9)
    p = api.AlgoParams.from_args(ds.args)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    seq_off, seq, name_off, names = t(ds.reads.seq_off), t(ds.reads.seq), t(ds.reads.name_off), t(ds.reads.names)
    paf = t(np.frombuffer(ds.paf, np.uint8))
    torch.cuda.synchronize()
    ctx = api.Context(p)
    ctx.set_reads(seq_off, seq, name_off, names)
    ctx.ingest_paf(paf, paf.numel(), last=True)
    st = ctx.run()
    ref = O.run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    for which, data in ((api.OUT_COVERAGE, ref.cov_txt), (api.OUT_LONG_REPEATS, ref.rep_txt), (api.OUT_READS_FASTA, ref.fasta)):
        total = ctx.output_size(which)
        out = torch.empty(total + 3, dtype=torch.uint8, device=dev)
        ctx.fetch_into(which, 0, out[3:], total)  # deliberately misaligned destination
        torch.cuda.synchronize()
        assert out[3:].cpu().numpy().tobytes() == data
    assert st.n_records == ref.n_rec
    ctx.close()


def test_error_domain_matches_oracle():
    fa = b">a\nACGTACGTAC\n>b\nACGTACGTACGG\n"
    reads = O.parse_fasta(fa)
    line = lambda *f: b"\t".join(str(x).encode() for x in f) + b"\n"
    ok = line("a", 10, 0, 10, "+", "b", 12, 0, 10, 10, 10, 255)
    cases = [(line("zz", 10, 0, 10, "+", "b", 12, 0, 10, 10, 10, 255), dict(est_cov=1), -2),
             (ok + line("a", 10, 0, 10, "+", "nope", 12, 0, 10, 10, 10, 255), dict(est_cov=1), -2),
             (line("a", 10, 0, 500, "+", "b", 12, 0, 10, 10, 10, 255), dict(est_cov=1), -4),
             (ok, dict(est_cov=5, repeat_length=5, read_length=5, overlap_length=50), -5)]
    for paf, kw, want in cases:
        assert O.run(reads, paf, O.make_params(**kw)).status == want
        with pytest.raises(api.RaftError) as ei:
            gpu_run(reads, paf, api.AlgoParams(**kw))
        assert ei.value.status == want
    with pytest.raises(api.RaftError) as ei:
        api.Context(api.AlgoParams(est_cov=1, read_length=50, repeat_length=100))
    assert ei.value.status == -1
    dup = O.parse_fasta(b">a\nAC\n>a\nGT\n")
    with pytest.raises(api.RaftError) as ei:
        gpu_run(dup, b"", api.AlgoParams(est_cov=1))
    assert ei.value.status == -3


def test_cli_binary_against_golden():
    """The `raft` executable (C++ host over the C ABI): same files and stdout lines as the reference."""
    exe = os.path.join(ROOT, "raft_b200", "raft")
    assert os.path.exists(exe), "run make"
    for entry in CASES:
        fa, paf = load_inputs(entry)
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "r.fa"), "wb").write(fa)
            open(os.path.join(d, "o.paf"), "wb").write(paf)
            r = subprocess.run([exe] + entry["args"] + ["-o", os.path.join(d, "out"), os.path.join(d, "r.fa"), os.path.join(d, "o.paf")],
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
            assert r.returncode == 0, r.stdout.decode()
            for suf in SUFS:
                check_output(entry, suf, open(os.path.join(d, "out." + suf), "rb").read())
            got = [l for l in r.stdout.decode().splitlines() if not l.startswith("INFO, main(), program completed") and "CMD:" not in l]
            assert got == entry["stdout"], entry["name"]


def test_cli_two_paf_files_equal_cat():
    """f4(ii): several PAF files (plain, cut mid-line, + gzip) are ingested back to back like `cat a b` (README.md:35-36)."""
    import gzip
    exe = os.path.join(ROOT, "raft_b200", "raft")
    entry = max(CASES, key=lambda e: len(load_inputs(e)[1]))
    fa, paf = load_inputs(entry)
    cut = len(paf) // 2 + 7          # not on a line boundary: the tail of the first file joins the head of the second
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "r.fa"), "wb").write(fa)
        open(os.path.join(d, "a.paf"), "wb").write(paf[:cut])
        with gzip.open(os.path.join(d, "b.paf.gz"), "wb") as f:
            f.write(paf[cut:])
        multi = dict(os.environ, RAFT_B200_MULTI_PAF="1")   # the extension is opt-in: the reference ignores a third positional argument
        r = subprocess.run([exe] + entry["args"] + ["-o", os.path.join(d, "out"), os.path.join(d, "r.fa"), os.path.join(d, "a.paf"),
                                                    os.path.join(d, "b.paf.gz")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300, env=multi)
        assert r.returncode == 0, r.stdout.decode()
        for suf in SUFS:
            check_output(entry, suf, open(os.path.join(d, "out." + suf), "rb").read())
        got = [l for l in r.stdout.decode().splitlines() if not l.startswith("INFO, main(), program completed") and "CMD:" not in l]
        assert got == entry["stdout"], entry["name"]
        # python mirror, list form
        p = api.AlgoParams.from_args(entry["args"] + ["-o", os.path.join(d, "py")])
        api.break_long_reads(os.path.join(d, "r.fa"), [os.path.join(d, "a.paf"), os.path.join(d, "b.paf.gz")], p)
        for suf in SUFS:
            check_output(entry, suf, open(os.path.join(d, "py." + suf), "rb").read())
        # a missing second file is reported like a missing first one (chop.hpp:344-348)
        r = subprocess.run([exe] + entry["args"] + ["-o", os.path.join(d, "o2"), os.path.join(d, "r.fa"), os.path.join(d, "a.paf"),
                                                    os.path.join(d, "nope.paf")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300, env=multi)
        assert r.returncode == 1 and b"nope.paf input file either does not exist or is empty" in r.stdout
        # without the opt-in a third positional argument is ignored, as in the reference (main.cpp:75, chop.hpp:331)
        open(os.path.join(d, "whole.paf"), "wb").write(paf)
        r = subprocess.run([exe] + entry["args"] + ["-o", os.path.join(d, "o3"), os.path.join(d, "r.fa"), os.path.join(d, "whole.paf"),
                                                    os.path.join(d, "nope.paf")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
        assert r.returncode == 0, r.stdout.decode()
        for suf in SUFS:
            check_output(entry, suf, open(os.path.join(d, "o3." + suf), "rb").read())
        # a file without a final newline does not glue its last record to the first line of the next file
        lines = paf.split(b"\n")
        half = len(lines) // 2
        open(os.path.join(d, "c.paf"), "wb").write(b"\n".join(lines[:half]))          # no trailing newline
        open(os.path.join(d, "e.paf"), "wb").write(b"\n".join(lines[half:]))
        p2 = api.AlgoParams.from_args(entry["args"] + ["-o", os.path.join(d, "py2")])
        api.break_long_reads(os.path.join(d, "r.fa"), [os.path.join(d, "c.paf"), os.path.join(d, "e.paf")], p2)
        for suf in SUFS:
            check_output(entry, suf, open(os.path.join(d, "py2." + suf), "rb").read())


def test_async_fetch_lanes():
    """raftgpu_fetch_async queues emitters on the two emit lanes (text outputs / sequence outputs); raftgpu_sync joins them."""
    torch = pytest.importorskip("torch")
    ds = synth.make_dataset("C1", 0.2, True, seed=12)
    p = api.AlgoParams.from_args(ds.args)
    ctx, st = gpu_run(ds.reads, ds.paf, p)
    ref = O.run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    dev = torch.device("cuda:0")
    want = {api.OUT_COVERAGE: ref.cov_txt, api.OUT_LONG_REPEATS: ref.rep_txt, api.OUT_READS_FASTA: ref.fasta}
    bufs = {w: torch.zeros(len(d) + 16, dtype=torch.uint8, device=dev) for w, d in want.items()}
    torch.cuda.synchronize()
    # interleave windows of the three streams without waiting in between
    for lo_frac, hi_frac in ((0.0, 0.37), (0.37, 1.0)):
        for w, d in want.items():
            lo, hi = int(len(d) * lo_frac), int(len(d) * hi_frac)
            if hi > lo:
                ctx.fetch_async(w, lo, bufs[w][lo:], hi - lo)
    ctx.sync()
    for w, d in want.items():
        assert bufs[w][:len(d)].cpu().numpy().tobytes() == d, w
        assert int(bufs[w][len(d):].sum()) == 0
    ctx.close()


@pytest.mark.parametrize("block", range(3))
def test_fuzz_cuda_matches_oracle(block):
    """The 150 random small inputs of tests/fuzz_util.py (zero-length intervals, reads shorter than a bin, lengths that
    are multiples of -r / -p, mirrored first record, junk lines, CR LF, empty PAF): every table and every output byte."""
    for seed in range(block * 50, block * 50 + 50):
        fa, paf, args = fuzz_case(seed)
        reads = O.parse_fasta(fa)
        kw = args_to_kw(args)
        ref = O.run(reads, paf, O.make_params(**kw))
        assert ref.status == 0
        ctx, st = gpu_run(reads, paf, api.AlgoParams(**kw))
        try:
            compare_all(ctx, st, ref)
        except AssertionError as e:
            raise AssertionError(f"fuzz seed {seed} args {args}: {e}") from e
        finally:
            ctx.close()


def _stress_paf(ds, seed=0):
    """PAF text with everything the tokenizer's mask logic can trip on: lines longer than the staged
    overhang (kilobyte tags), lines crossing 16 KiB tile boundaries, blank-line runs, short lines,
    exactly-10-field lines, CRLF, numeric fields with spaces / signs / junk."""
    rng = np.random.default_rng(seed)
    lines = ds.paf.split(b"\n")[:-1]
    out = []
    for k, ln in enumerate(lines[:6000]):
        r = rng.integers(0, 20)
        f = ln.split(b"\t")
        if r == 0:
            out.append(b"")                                   # blank line
            out.append(ln)
        elif r == 1:
            out.append(b"\t".join(f[:9]))                    # 9 fields: not a record
        elif r == 2:
            out.append(b"\t".join(f[:10]))                   # 10 fields: a record
        elif r == 3:
            out.append(ln + b"\tcg:Z:" + b"5M1D" * int(rng.integers(100, 12000)))  # very long tag
        elif r == 4:
            out.append(ln + b"\r")                           # CRLF
        elif r == 5:
            f2 = list(f); f2[2] = b" " + f2[2]; f2[3] = b"+" + f2[3]; f2[8] = f2[8] + b"xyz"; f2[7] = b"0" + f2[7]
            out.append(b"\t".join(f2))
        elif r == 6:
            out.append(b"\n\n\n" + ln)                       # run of blank lines
        elif r == 7:
            out.append(b"garbage without tabs " * int(rng.integers(1, 200)))
        elif r == 8:
            out.append(b"x" * int(rng.integers(1000, 40000)) + b"\t" + ln)  # first field (name) longer than the overhang: unknown name? no: keep short-field line invalid
            out[-1] = out[-1].replace(b"\t", b" ")         # ... make it a tab-free garbage line
        else:
            out.append(ln)
    return b"\n".join(out) + (b"\n" if seed % 2 == 0 else b"")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tokenizer_stress(seed):
    ds = synth.make_dataset("C1", 0.05, seed % 2 == 0, seed=40 + seed)
    paf = _stress_paf(ds, seed)
    p = api.AlgoParams.from_args(ds.args)
    ref = O.run(ds.reads, paf, O.make_params(**args_to_kw(ds.args)))
    assert ref.status == 0 and ref.n_rec > 1000
    ctx, st = gpu_run(ds.reads, paf, p)
    compare_all(ctx, st, ref)
    ctx.close()
    # and the same text in awkward chunks
    cuts = sorted(set(int(x) for x in np.random.default_rng(seed).integers(0, len(paf), 7)))
    chunks = [b - a for a, b in zip([0] + cuts, cuts + [len(paf)])]
    ctx, st = gpu_run(ds.reads, paf, p, chunks=chunks)
    compare_all(ctx, st, ref)
    ctx.close()


def test_coverage_emitter_direct_path(monkeypatch):
    """Tiles whose text exceeds the shared-memory buffer are written byte-wise: force that path."""
    ds = synth.make_dataset("C5", 0.002, True, seed=12)
    p = api.AlgoParams.from_args(ds.args)
    ref = O.run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    monkeypatch.setenv("RAFT_B200_COV_CAP", "1000")
    ctx, st = gpu_run(ds.reads, ds.paf, p)
    assert ctx.fetch(api.OUT_COVERAGE) == ref.cov_txt
    total = len(ref.cov_txt)
    for off, ln in ((7, 333), (total // 2, 70001), (total - 9, 9)):
        assert ctx.fetch(api.OUT_COVERAGE, off, ln) == ref.cov_txt[off:off + ln]
    ctx.close()


def test_many_tiny_reads_and_fragments():
    """Thousands of records per 16 KiB output tile (multi-round gather) and zero-length reads."""
    rng = np.random.default_rng(3)
    n = 5000
    lens = rng.integers(0, 40, n)
    lens[::7] = 0
    names = [b"t%d" % i for i in range(n)]
    fa = b"".join(b">" + nm + b"\n" + bytes(rng.choice(list(b"ACGT"), int(L)).astype(np.uint8)) + b"\n" for nm, L in zip(names, lens))
    reads = O.parse_fasta(fa)
    assert reads.n == n
    lines = []
    for i in range(0, n - 1, 3):
        if lens[i] > 4 and lens[i + 1] > 4:
            lines.append(b"\t".join([names[i], b"%d" % lens[i], b"1", b"%d" % lens[i], b"+", names[i + 1], b"%d" % lens[i + 1], b"0",
                                     b"%d" % (lens[i + 1] - 1), b"3", b"3", b"60"]))
    paf = b"\n".join(lines) + b"\n"
    kw = dict(est_cov=1, reso=5, repeat_length=10, read_length=20, overlap_length=3, flanking_length=2)
    ref = O.run(reads, paf, O.make_params(**kw))
    assert ref.status == 0
    ctx, st = gpu_run(reads, paf, api.AlgoParams(**kw))
    compare_all(ctx, st, ref)
    ctx.close()


def test_reads_spanning_many_scan_tiles():
    """The coverage scan derives a tile's incoming running sum from the read that spans the tile boundary; reads
    longer than one 4096-slot scan tile take the chained path (tile waits for its predecessor).  Mix of reads of
    0.3-1.2 Mbp at -r 10 (30 k-120 k slots each, up to 30 tiles per read) with short ones."""
    rng = np.random.default_rng(17)
    lens = [300000, 700, 1200000, 41000, 90, 450000, 5000, 40960, 40950, 40970]
    names = [b"L%d" % i for i in range(len(lens))]
    fa = b"".join(b">" + nm + b"\n" + bytes(rng.choice(list(b"ACGT"), int(L)).astype(np.uint8)) + b"\n" for nm, L in zip(names, lens))
    reads = O.parse_fasta(fa)
    lines = []
    for _ in range(4000):
        q, t = rng.integers(0, len(lens), 2)
        if q == t:
            continue
        qs = int(rng.integers(0, lens[q])); qe = int(rng.integers(qs, lens[q] + 1))
        ts = int(rng.integers(0, lens[t])); te = int(rng.integers(ts, lens[t] + 1))
        lines.append(b"\t".join([names[q], b"%d" % lens[q], b"%d" % qs, b"%d" % qe, b"+", names[t], b"%d" % lens[t], b"%d" % ts, b"%d" % te,
                                 b"9", b"9", b"60"]))
    paf = b"\n".join(lines) + b"\n"
    kw = dict(est_cov=150, reso=10, repeat_length=5000, read_length=20000, overlap_length=500, flanking_length=500)
    ref = O.run(reads, paf, O.make_params(**kw))
    assert ref.status == 0 and ref.symmetric == 0
    ctx, st = gpu_run(reads, paf, api.AlgoParams(**kw))
    compare_all(ctx, st, ref)
    ctx.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_deferred_sequence_upload(pinned):
    """RAFTGPU_OPT_DEFER_SEQ_UPLOAD: the arena is uploaded in chunks behind the PAF; windows wait for their chunks."""
    torch = pytest.importorskip("torch")
    ds = synth.make_dataset("C2", 0.0006, True, seed=21)
    p = api.AlgoParams.from_args(ds.args)
    ref = O.run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    seq = np.ascontiguousarray(ds.reads.seq, np.uint8)
    if pinned:
        seq = torch.from_numpy(seq.copy()).pin_memory().numpy()
    ctx = api.Context(p)
    ctx.set_option(api.OPT_DEFER_SEQ_UPLOAD, 1)
    for rep in range(2):  # second pass: reset while nothing is pending, reuse of chunk events
        ctx.set_reads(np.ascontiguousarray(ds.reads.seq_off, np.int64), seq, np.ascontiguousarray(ds.reads.name_off, np.int64),
                      np.ascontiguousarray(ds.reads.names, np.uint8))
        ctx.ingest_paf(np.frombuffer(ds.paf, np.uint8), len(ds.paf), last=True)
        ctx.run()
        total = ctx.output_size(api.OUT_READS_FASTA)
        assert ctx.fetch(api.OUT_READS_FASTA, total - 1000, 1000) == ref.fasta[-1000:]   # a late window first
        assert ctx.fetch(api.OUT_READS_FASTA) == ref.fasta
        assert ctx.fetch(api.OUT_COVERAGE) == ref.cov_txt
        assert ctx.digest(api.OUT_READS_FASTA) == O.digest(ref.fasta)
    ctx.close()


def _fasta_variants():
    rng = np.random.default_rng(7)
    seqs = [bytes(rng.choice(list(b"ACGTN"), int(L)).astype(np.uint8)) for L in (1, 59, 60, 61, 0, 16384, 16385, 40000, 7, 100000, 3, 0, 250)]
    names = [b"r%d" % i for i in range(len(seqs))]
    def build(width, comment=b"", marker=b">", trailing=True, blank=False):
        out = []
        for i, (nm, sq) in enumerate(zip(names, seqs)):
            hdr = (b"@" if (marker == b"mix" and i % 3 == 0) else b">") + nm + ((b" " + comment + b"\t x=%d" % i) if comment else b"") + b"\n"
            body = b"\n".join(sq[k:k + width] for k in range(0, len(sq), width)) if width else sq
            out.append(hdr + body + (b"\n" if (body or not blank) else b"") + (b"\n\n" if blank and i % 2 else b""))
        txt = b"".join(out)
        return txt if trailing else txt.rstrip(b"\n")
    long_comment = b"c" * 20000  # header lines longer than a 16 KiB tile
    return {"wrap60": build(60), "wrap61c": build(61, b"some comment"), "unwrapped": build(0), "wrap80_noeol": build(80, trailing=False),
            "mixmarkers": build(70, marker=b"mix"), "blank_lines": build(75, blank=True), "huge_header": build(100, long_comment),
            "wrap1": build(1)}


@pytest.mark.parametrize("name", sorted(_fasta_variants()))
def test_device_fasta_ingest_matches_kseq_grammar(name):
    """Row f2: the device FASTA tokenizer yields the records of loadFASTA/kseq_read (oracle restatement, pinned by golden)."""
    text = _fasta_variants()[name]
    ref = O.parse_fasta(text)
    p = api.AlgoParams(est_cov=3)
    paf = b"".join(b"\t".join([b"r3", b"61", b"0", b"61", b"+", b"r%d" % j, b"1", b"0", b"1", b"1", b"1", b"9"]) + b"\n" for j in (5, 7))
    want = O.run(ref, paf, O.make_params(est_cov=3))
    assert want.status == 0
    for chunks in (None, [5, 1000, 16384, 1, len(text)]):
        ctx = api.Context(p)
        buf = np.frombuffer(text, np.uint8)
        if chunks is None:
            ctx.ingest_fasta(buf, len(text), last=True, total_hint=len(text))
        else:
            pos = 0
            for k, c in enumerate(chunks):
                piece = buf[pos:pos + c]
                ctx.ingest_fasta(np.ascontiguousarray(piece), len(piece), last=(k == len(chunks) - 1))
                pos += len(piece)
        ctx.ingest_paf(np.frombuffer(paf, np.uint8), len(paf), last=True)
        st = ctx.run()
        assert st.n_reads == ref.n
        np.testing.assert_array_equal(ctx.table(api.TAB_BIN_OFF), want.bin_off)
        assert ctx.fetch(api.OUT_READS_FASTA) == want.fasta     # names, lengths and every base in place
        assert ctx.fetch(api.OUT_COVERAGE) == want.cov_txt
        ctx.close()


def test_device_fasta_ingest_rejects_what_it_does_not_take():
    p = api.AlgoParams(est_cov=3)
    for text in (b">a\r\nACGT\r\n", b"junk\n>a\nAC\n", b">a\nAC\n>", b">a\nAC\n+\nII\n",
                 b"@a\nACGT\nAC\n+\nIIII\nII\n",            # FASTQ with wrapped bases / qualities
                 b"@a\nACGT\n+\nIII\n@b\nAC\n+\nII\n",      # quality line shorter than the bases (kseq would read on)
                 b"@a\nACGT\n+\nIIIII\n",                    # longer
                 b"@a\nACGT\n+\nIIII\n@b\nAC\n+\n",          # truncated last record
                 b"@a\r\nACGT\r\n+\r\nIIII\r\n"):
        ctx = api.Context(p)
        with pytest.raises(api.RaftError) as ei:
            ctx.ingest_fasta(np.frombuffer(text, np.uint8), len(text), last=True)
        assert ei.value.status == -12, text
        ctx.close()


def _fastq_variants():
    rng = np.random.default_rng(11)
    lens = (1, 59, 0, 16384, 16383, 40000, 7, 100000, 3, 0, 250, 33000)
    seqs = [bytes(rng.choice(list(b"ACGTN"), int(L)).astype(np.uint8)) for L in lens]
    names = [b"q%d" % i for i in range(len(seqs))]
    def build(comment=b"", plus_name=False, trailing=True, at_quals=False):
        out = []
        for i, (nm, sq) in enumerate(zip(names, seqs)):
            q = bytes(rng.integers(33, 74, len(sq)).astype(np.uint8))
            if at_quals and q:
                q = (b"@" if i % 2 else b"+") + q[1:]   # a quality line may start with '@' or '+': only its position says what it is
            out.append(b"@" + nm + ((b" " + comment) if comment else b"") + b"\n" + sq + b"\n+" + (nm if plus_name else b"") + b"\n" + q + b"\n")
        txt = b"".join(out)
        return txt if trailing else txt[:-1]
    return {"plain": build(), "comment_plusname": build(b"a comment\tx", True), "noeol": build(trailing=False), "at_in_quals": build(at_quals=True),
            "huge_header": build(b"c" * 20000)}


@pytest.mark.parametrize("name", sorted(_fastq_variants()))
def test_device_fastq_ingest_matches_kseq_grammar(name):
    """Row f2: strict four-line FASTQ is tokenised on the device (line role = line index mod 4) and yields kseq's records."""
    text = _fastq_variants()[name]
    ref = O.parse_fasta(text)
    assert ref.n == 12
    p = api.AlgoParams(est_cov=3)
    paf = b"".join(b"\t".join([b"q3", b"16384", b"0", b"61", b"+", b"q%d" % j, b"1", b"0", b"1", b"1", b"1", b"9"]) + b"\n" for j in (5, 7))
    want = O.run(ref, paf, O.make_params(est_cov=3))
    assert want.status == 0
    for chunks in (None, [5, 1000, 16384, 1, 70000, len(text)]):
        ctx = api.Context(p)
        buf = np.frombuffer(text, np.uint8)
        if chunks is None:
            ctx.ingest_fasta(buf, len(text), last=True, total_hint=len(text))
        else:
            pos = 0
            for k, c in enumerate(chunks):
                piece = buf[pos:pos + c]
                ctx.ingest_fasta(np.ascontiguousarray(piece), len(piece), last=(k == len(chunks) - 1))
                pos += len(piece)
        ctx.ingest_paf(np.frombuffer(paf, np.uint8), len(paf), last=True)
        st = ctx.run()
        assert st.n_reads == ref.n
        np.testing.assert_array_equal(ctx.table(api.TAB_BIN_OFF), want.bin_off)
        assert ctx.fetch(api.OUT_READS_FASTA) == want.fasta     # names, lengths and every base in place
        assert ctx.fetch(api.OUT_COVERAGE) == want.cov_txt
        ctx.close()


def test_device_fasta_ingest_full_pipeline():
    ds = synth.make_dataset("C1", 0.1, True, seed=77)
    p = api.AlgoParams.from_args(ds.args)
    text = synth.format_fasta(ds.reads, wrap=70)
    ref = O.run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    ctx = api.Context(p)
    ctx.ingest_fasta(np.frombuffer(text, np.uint8), len(text), last=True, total_hint=len(text))
    ctx.ingest_paf(np.frombuffer(ds.paf, np.uint8), len(ds.paf), last=True)
    st = ctx.run()
    compare_all(ctx, st, ref)
    ctx.close()


@pytest.mark.parametrize("sublen", [1, 13, 5000, 20000])
def test_split_naive_stream(sublen):
    """Row f4: split_naive.cpp through the same gather kernel; bytes equal the oracle's (pinned to the reference binary)."""
    ds = synth.make_dataset("C1", 0.03 if sublen > 1 else 0.002, True, seed=8)
    fa_bytes = synth.format_fasta(ds.reads) + b">empty\n\n>tail\nACGTAC\n"
    reads = O.parse_fasta(fa_bytes)
    want = O.split_naive(reads, sublen)
    ctx = api.Context(api.AlgoParams(est_cov=1))
    ctx.ingest_fasta(np.frombuffer(fa_bytes, np.uint8), len(fa_bytes), last=True)
    ctx.split_naive(sublen)
    assert ctx.output_size(api.OUT_SPLIT_NAIVE) == len(want)
    assert ctx.fetch(api.OUT_SPLIT_NAIVE) == want
    assert ctx.digest(api.OUT_SPLIT_NAIVE) == O.digest(want)
    off = len(want) // 3
    assert ctx.fetch(api.OUT_SPLIT_NAIVE, off, min(100000, len(want) - off)) == want[off:off + min(100000, len(want) - off)]
    ctx.close()


def test_split_naive_cli():
    exe = os.path.join(ROOT, "raft_b200", "split_naive")
    ds = synth.make_dataset("C1", 0.02, False, seed=4)
    fa_bytes = synth.format_fasta(ds.reads, wrap=70)
    with tempfile.TemporaryDirectory() as d:
        fa, out = os.path.join(d, "r.fa"), os.path.join(d, "o.fa")
        open(fa, "wb").write(fa_bytes)
        r = subprocess.run([exe, fa, out, "7000"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
        assert r.returncode == 0, r.stdout.decode()
        assert open(out, "rb").read() == O.split_naive(ds.reads, 7000)
