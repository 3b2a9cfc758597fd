"""Pins the CPU restatement (oracle/) to the unmodified reference binary's outputs (tests/golden/)."""
import os
import tempfile

import pytest

from golden_util import SUFS, args_to_kw, check_output, load_inputs, manifest, stdout_value
from oracle import oracle as O
from raft_b200 import synth
from fuzz_util import fuzz_case, fuzz_fastq_text
from sim_util import sim_dataset

CASES = manifest()


@pytest.mark.parametrize("entry", CASES, ids=[e["name"] for e in CASES])
def test_oracle_matches_reference_golden(entry):
    fa, paf = load_inputs(entry)
    reads = O.parse_fasta(fa)
    res = O.run(reads, paf, O.make_params(**args_to_kw(entry["args"])))
    assert res.status == 0
    for suf, data in zip(SUFS, (res.cov_txt, res.rep_txt, res.bed_txt, res.fasta)):
        check_output(entry, suf, data)
    assert stdout_value(entry, "Symmetric overlaps") == f"INFO, Symmetric overlaps {res.symmetric} "
    assert stdout_value(entry, "length of alignments") == f"INFO, length of alignments  {res.n_rec}()"
    assert stdout_value(entry, "high_cov") == f"high_cov {res.high_cov}"
    assert stdout_value(entry, "Real Reads") == f"Real Reads {res.real_reads} "
    cpw = res.total_cov / res.total_windows if res.total_windows else float("nan")
    assert stdout_value(entry, "coverage per window is") == "coverage per window is %f " % cpw
    assert stdout_value(entry, "fraction_of_repeat_length") == "fraction_of_repeat_length %f " % (
        res.total_repeat_len / res.total_read_len)


def _assert_digest_mode_equals_run(reads, paf, prm, threads):
    """orc_run_digest (the multi-threaded form used for outputs too large to hold) reports exactly the digests, lengths
    and counters of the bytes orc_run materialises."""
    res = O.run(reads, paf, prm)
    dg = O.run_digest(reads, paf, prm, threads=threads)
    assert dg.status == res.status
    if res.status:
        assert dg.bad_index == res.bad_index
        return
    for k, data in enumerate((res.cov_txt, res.rep_txt, res.bed_txt, res.fasta)):
        assert dg.bytes[k] == len(data), k
        assert dg.digest[k] == O.digest(data), k
    assert (dg.n_rec, dg.symmetric, dg.high_cov, dg.real_reads, dg.n_frag) == (res.n_rec, res.symmetric, res.high_cov, res.real_reads, res.n_frag)
    assert (dg.total_cov, dg.total_windows, dg.total_repeat_len, dg.total_read_len) == (res.total_cov, res.total_windows,
                                                                                        res.total_repeat_len, res.total_read_len)
    assert dg.n_rep == len(res.rep_s)


@pytest.mark.parametrize("entry", CASES, ids=[e["name"] for e in CASES])
def test_digest_mode_matches_run_on_golden(entry):
    fa, paf = load_inputs(entry)
    for threads in (1, 3):
        _assert_digest_mode_equals_run(O.parse_fasta(fa), paf, O.make_params(**args_to_kw(entry["args"])), threads)


def test_digest_mode_matches_run_on_fuzz_and_errors():
    for seed in range(150):
        fa, paf, args = fuzz_case(seed)
        _assert_digest_mode_equals_run(O.parse_fasta(fa), paf, O.make_params(**args_to_kw(args)), 1 + seed % 4)
    reads, paf = sim_dataset(5)
    _assert_digest_mode_equals_run(reads, paf, O.make_params(est_cov=30, repeat_length=4000, read_length=8000, flanking_length=300, overlap_length=200), 4)
    rd = O.parse_fasta(b">a\nACGTACGTAC\n>b\nACGTACGTACGG\n")
    line = lambda *f: b"\t".join(str(x).encode() for x in f) + b"\n"
    ok = line("a", 10, 0, 10, "+", "b", 12, 0, 10, 10, 10, 255)
    for bad in (line("zz", 10, 0, 10, "+", "b", 12, 0, 10, 10, 10, 255), line("a", 10, 0, 500, "+", "b", 12, 0, 10, 10, 10, 255)):
        _assert_digest_mode_equals_run(rd, ok * 3 + bad + ok, O.make_params(est_cov=1), 2)
    _assert_digest_mode_equals_run(rd, ok, O.make_params(est_cov=1, read_length=50, repeat_length=100), 2)
    _assert_digest_mode_equals_run(O.parse_fasta(b">a\nAC\n>a\nGT\n"), b"", O.make_params(est_cov=1), 2)


@pytest.mark.parametrize("cfg,scale,sym", [("C2", 0.0004, True), ("C5", 0.002, False), ("C4", 0.001, True)])
def test_digest_mode_matches_run_on_config_shapes(cfg, scale, sym):
    ds = synth.make_dataset(cfg, scale, sym, seed=77)
    _assert_digest_mode_equals_run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)), 8)


def test_oracle_defined_domain_errors():
    fa = b">a\nACGTACGTAC\n>b\nACGTACGTACGG\n"
    reads = O.parse_fasta(fa)
    line = lambda *f: b"\t".join(str(x).encode() for x in f) + b"\n"
    ok = line("a", 10, 0, 10, "+", "b", 12, 0, 10, 10, 10, 255)
    assert O.run(reads, ok, O.make_params(est_cov=1)).status == 0
    # unknown name: the reference indexes out of range (chop.hpp:162-168)
    assert O.run(reads, line("zz", 10, 0, 10, "+", "b", 12, 0, 10, 10, 10, 255), O.make_params(est_cov=1)).status == -2
    # interval past the last bin: the reference writes out of bounds (repeat.hpp:69-73)
    assert O.run(reads, line("a", 10, 0, 500, "+", "b", 12, 0, 10, 10, 10, 255), O.make_params(est_cov=1)).status == -4
    # l < p: div = 0 -> SIGFPE in the reference (chop.hpp:248,270)
    assert O.run(reads, ok, O.make_params(est_cov=1, read_length=50, repeat_length=100)).status == -1
    # duplicate FASTA names alias ids in the reference (chop.hpp:73-85)
    assert O.run(O.parse_fasta(b">a\nAC\n>a\nGT\n"), b"", O.make_params(est_cov=1)).status == -3


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/raft not built (reference sources not mounted)")
@pytest.mark.parametrize("cfg,scale,sym", [("C1", 0.05, True), ("C1", 0.05, False), ("C5", 0.001, False),
                                           ("C4", 0.0004, True), ("C2", 0.00006, False)])
def test_oracle_matches_reference_binary_live(cfg, scale, sym):
    ds = synth.make_dataset(cfg, scale, sym, seed=1234)
    with tempfile.TemporaryDirectory() as d:
        fa, pf = os.path.join(d, "r.fa"), os.path.join(d, "o.paf")
        fa_bytes = synth.format_fasta(ds.reads, wrap=61)
        open(fa, "wb").write(fa_bytes)
        open(pf, "wb").write(ds.paf)
        rc, out, files = O.run_ref(fa, pf, d, ds.args)
    assert rc == 0, out
    reads = O.parse_fasta(fa_bytes)
    assert (reads.seq_off == ds.reads.seq_off).all() and (reads.seq == ds.reads.seq).all()
    assert (reads.names == ds.reads.names).all()
    res = O.run(reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    assert res.status == 0
    for suf, data in zip(SUFS, (res.cov_txt, res.rep_txt, res.bed_txt, res.fasta)):
        assert files.get(suf, b"") == data, suf


def test_digest_is_window_additive():
    data = bytes(range(256)) * 7
    whole = O.digest(data)
    parts = (O.digest(data[:100], 0) + O.digest(data[100:1000], 100) + O.digest(data[1000:], 1000)) & (2**64 - 1)
    assert whole == parts


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/raft not built (reference sources not mounted)")
def test_oracle_sim_mode_matches_reference_binary_live():
    reads, paf = sim_dataset(5)
    with tempfile.TemporaryDirectory() as d:
        fa, pf = os.path.join(d, "r.fa"), os.path.join(d, "o.paf")
        open(fa, "wb").write(synth.format_fasta(reads, wrap=60))
        open(pf, "wb").write(paf)
        rc, out, files = O.run_ref(fa, pf, d, ["-e", "30", "-p", "4000", "-l", "8000", "-f", "300", "-v", "200"])
    assert rc == 0 and "Real Reads 0 " in out
    res = O.run(reads, paf, O.make_params(est_cov=30, repeat_length=4000, read_length=8000, flanking_length=300, overlap_length=200))
    assert res.status == 0 and res.real_reads == 0 and len(res.bed_txt) > 0
    for suf, data in zip(SUFS, (res.cov_txt, res.rep_txt, res.bed_txt, res.fasta)):
        assert files.get(suf, b"") == data, suf


@pytest.mark.skipif(not os.path.exists(O.REF_SPLIT_BIN), reason="oracle/_ref/split_naive not built")
@pytest.mark.parametrize("sublen", [1, 7, 20000, 1000000])
def test_oracle_split_naive_matches_reference_binary_live(sublen):
    import subprocess
    ds = synth.make_dataset("C1", 0.02, True, seed=3)
    fa_bytes = synth.format_fasta(ds.reads, wrap=60) + b">empty\n\n>tail\nACGTAC\n"
    reads = O.parse_fasta(fa_bytes)
    with tempfile.TemporaryDirectory() as d:
        fa, out = os.path.join(d, "r.fa"), os.path.join(d, "o.fa")
        open(fa, "wb").write(fa_bytes)
        subprocess.run([O.REF_SPLIT_BIN, fa, out, str(sublen)], check=True, timeout=120)
        assert open(out, "rb").read() == O.split_naive(reads, sublen)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/raft not built (reference sources not mounted)")
@pytest.mark.parametrize("block", range(6))
def test_oracle_fuzz_matches_reference_binary(block):
    """150 random small inputs of the shapes the closed form is sensitive to: the restatement and the unmodified reference
    binary write the same four files (or both leave the defined domain: v > star puts a fragment start below 0)."""
    compared = 0
    for seed in range(block * 25, block * 25 + 25):
        fa, paf, args = fuzz_case(seed)
        reads = O.parse_fasta(fa)
        res = O.run(reads, paf, O.make_params(**args_to_kw(args)))
        with tempfile.TemporaryDirectory() as d:
            fp, pp = os.path.join(d, "r.fa"), os.path.join(d, "o.paf")
            open(fp, "wb").write(fa)
            open(pp, "wb").write(paf)
            empty_input = len(fa) == 0 or len(paf) == 0
            rc, out, files = O.run_ref(fp, pp, d, args)
        if res.status != 0 or empty_input:
            continue  # outside the defined domain the reference crashes or exits early (covered by the error-domain tests)
        assert rc == 0, (seed, out)
        for suf, data in zip(SUFS, (res.cov_txt, res.rep_txt, res.bed_txt, res.fasta)):
            assert files.get(suf, b"") == data, (seed, suf, args)
        compared += 1
    assert compared >= 10


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/raft not built (reference sources not mounted)")
def test_oracle_fastq_grammar_fuzz_matches_reference_binary():
    """kseq's FASTQ reading (qualities read by length, wrapped lines, '@' / '+' opening a quality line, FASTA records in
    between): the oracle's parser and the reference binary see the same reads -- reads.fasta and coverage.txt of a run
    without overlaps list every read's name, length and bases."""
    for seed in range(60):
        text, lens = fuzz_fastq_text(seed)
        reads = O.parse_fasta(text)
        paf = b"x\n"                                   # no record: the reference refuses an empty PAF file
        res = O.run(reads, paf, O.make_params(est_cov=2, reso=10))
        assert res.status == 0
        with tempfile.TemporaryDirectory() as d:
            fp, pp = os.path.join(d, "r.fq"), os.path.join(d, "o.paf")
            open(fp, "wb").write(text)
            open(pp, "wb").write(paf)
            rc, out, files = O.run_ref(fp, pp, d, ["-e", "2", "-r", "10"])
        assert rc == 0, (seed, out)
        assert files.get("reads.fasta", b"") == res.fasta, seed
        assert files.get("coverage.txt", b"") == res.cov_txt, seed


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/raft not built (reference sources not mounted)")
@pytest.mark.parametrize("text", [b"@a\nACGT\n+\nIIII\n@b\nACGT\n+\nII\n",          # truncated quality at the end of the file: record dropped
                                  b"@a\nACGT\n+\nIIIII\n@b\nAC\n+\nII\n",          # quality longer than the bases: kseq stops there (kseq.h:296)
                                  b"@a\nACGT\n+\nIII\n@b\nAC\n+\nII\n",            # shorter: the next lines are read as quality
                                  b"@a\nAC\nGT\n+\nII\nII\n>b\nACG\n",             # wrapped FASTQ, then FASTA
                                  b"@a\n\n+\n\n@b\nA\n+\nI"])                      # empty record, no final newline
def test_oracle_fastq_error_cases_match_reference_binary(text):
    reads = O.parse_fasta(text)
    res = O.run(reads, b"x\n", O.make_params(est_cov=2, reso=10))
    assert res.status == 0
    with tempfile.TemporaryDirectory() as d:
        fp, pp = os.path.join(d, "r.fq"), os.path.join(d, "o.paf")
        open(fp, "wb").write(text)
        open(pp, "wb").write(b"x\n")
        rc, out, files = O.run_ref(fp, pp, d, ["-e", "2", "-r", "10"])
    assert rc == 0, out
    assert files.get("reads.fasta", b"") == res.fasta
    assert files.get("coverage.txt", b"") == res.cov_txt
