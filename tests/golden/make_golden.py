#!/usr/bin/env python
"""Generate tests/golden/ by running the UNMODIFIED reference binary (oracle/_ref/raft, built by
oracle/Makefile from /root/reference) on small fixtures.  Run in the build container only:

    make -C oracle && python tests/golden/make_golden.py

Outputs (committed):
  edge/      hand-made edge fixtures of SURVEY.md Appendix A (inputs + full reference outputs + stdout)
  synth/     small seeded synthetic datasets (inputs gz-compressed, reference outputs as sha256+length,
             plus full outputs for the smallest ones)
  manifest.json   one entry per case: inputs, CLI args, expected stdout markers, output digests
"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from raft_b200 import synth  # noqa: E402

import numpy as np  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SUFS = ("coverage.txt", "long_repeats.txt", "long_repeats.bed", "reads.fasta")


def acgt(n, seed):
    return bytes(synth._ACGT[(synth._h(seed, 7, np.arange(n)) & np.uint64(3)).astype(np.int64)])


def wrap(s, w=70):
    return b"\n".join(s[i:i + w] for i in range(0, len(s), w))


def edge_inputs():
    reads = [(b"a", 1234), (b"b", 999), (b"c", 50), (b"d", 2501), (b"e", 1)]
    seqs = {nm: acgt(L, 1000 + i) for i, (nm, L) in enumerate(reads)}
    fa = b"".join(b">" + nm + (b" some comment here" if nm == b"a" else b"") + b"\n" + wrap(seqs[nm]) + b"\n" for nm, _ in reads)
    fq = b"".join(b"@" + nm + b"\n" + seqs[nm] + b"\n+\n" + b"I" * len(seqs[nm]) + b"\n" for nm, _ in reads)
    T = b"\t"
    lines = [
        T.join([b"a", b"1234", b"100", b"600", b"+", b"b", b"999", b"0", b"500", b"500", b"500", b"255"]),
        T.join([b"a", b"1234", b"0", b"0", b"+", b"d", b"2501", b"0", b"0", b"0", b"0", b"255"]),
        T.join([b"a", b"1234", b"75", b"75", b"+", b"d", b"2501", b"130", b"130", b"0", b"0", b"255"]),
        T.join([b"a", b"1234", b"100", b"100", b"+", b"d", b"2501", b"150", b"150", b"0", b"0", b"255"]),
        T.join([b"d", b"2501", b"0", b"2501", b"+", b"d", b"2501", b"0", b"2501", b"2501", b"2501", b"255"]),
        b"garbage line with no tabs",
        b"",
        T.join([b"c", b"50", b"0", b"50", b"-", b"d", b"2501", b"2451", b"2501", b"50"]),
        T.join([b"c", b"50", b"0", b"50", b"-", b"d", b"2501", b"2451", b"2501"]),
        T.join([b"d", b"2501", b"49", b"51", b"+", b"a", b"1234", b"1199", b"1234", b"2", b"2", b"255"]),
        T.join([b"e", b"1", b"0", b"1", b"+", b"a", b"1234", b"1233", b"1234", b"1", b"1", b"255"]),
    ]
    asym = b"\n".join(lines) + b"\n"
    flip = asym + T.join([b"b", b"999", b"0", b"500", b"+", b"a", b"1234", b"100", b"600", b"500", b"500", b"255"]) + b"\n"
    # extra number-format edge cases: leading spaces / '+' sign / trailing junk in numeric fields, tags
    numfmt = b"\n".join([
        T.join([b"a", b"1234", b" 100", b"+600", b"+", b"b", b"999", b"0x", b"500abc", b"500", b"500", b"255", b"tp:A:P", b"cm:i:5"]),
        T.join([b"b", b"999", b"007", b"0999", b"-", b"d", b"2501", b"", b"12", b"1", b"1", b"0"]),
    ]) + b"\n"
    sim_reads = [(b"read=1,forward,position=1000-2234,length=1234,chr1", 1234),
                 (b"read=2,reverse,position=5000-5999,length=999,chr2", 999)]
    sim_fa = b"".join(b">" + nm + b"\n" + acgt(L, 2000 + i) + b"\n" for i, (nm, L) in enumerate(sim_reads))
    sim_paf = T.join([sim_reads[0][0], b"1234", b"100", b"600", b"+", sim_reads[1][0], b"999", b"0", b"500", b"500", b"500", b"255"]) + b"\n"
    empty_fa = b">x\n\n>y\nACGTACGTAC\n"
    empty_paf = T.join([b"y", b"10", b"0", b"10", b"+", b"y", b"10", b"0", b"10", b"10", b"10", b"255"]) + b"\n"
    return dict(fa=fa, fq=fq, asym=asym, flip=flip, numfmt=numfmt, sim_fa=sim_fa, sim_paf=sim_paf,
                empty_fa=empty_fa, empty_paf=empty_paf)


EDGE_ARGS = ["-e", "1", "-m", "1.0", "-r", "50", "-p", "100", "-l", "300", "-f", "10", "-v", "20"]


def run_case(name, fa_bytes, paf_bytes, args, outdir, keep_full, fa_name="r.fa", paf_name="o.paf", gz=False):
    with tempfile.TemporaryDirectory() as d:
        fa, pf = os.path.join(d, fa_name), os.path.join(d, paf_name)
        if gz:
            with gzip.open(fa, "wb") as fh:
                fh.write(fa_bytes)
            with gzip.open(pf, "wb") as fh:
                fh.write(paf_bytes)
        else:
            open(fa, "wb").write(fa_bytes)
            open(pf, "wb").write(paf_bytes)
        rc, stdout, outs = O.run_ref(fa, pf, d, args)
    assert rc == 0, (name, rc, stdout)
    keep = [l for l in stdout.splitlines() if not l.startswith("INFO, main(), program completed") and "CMD:" not in l]
    entry = dict(name=name, args=args, rc=rc, stdout=keep, outputs={})
    for suf in SUFS:
        data = outs.get(suf, b"")
        entry["outputs"][suf] = dict(len=len(data), sha256=hashlib.sha256(data).hexdigest())
        if keep_full:
            os.makedirs(outdir, exist_ok=True)
            open(os.path.join(outdir, f"{name}.{suf}"), "wb").write(data)
    return entry


def main():
    assert O.have_ref(), "build oracle/_ref/raft first: make -C oracle"
    for sub in ("edge", "synth"):
        shutil.rmtree(os.path.join(GOLD, sub), ignore_errors=True)
        os.makedirs(os.path.join(GOLD, sub))
    manifest = []
    E = edge_inputs()
    ed = os.path.join(GOLD, "edge")
    for k, v in E.items():
        open(os.path.join(ed, "in." + k), "wb").write(v)
    crlf = E["asym"].replace(b"\n", b"\r\n")
    cases = [
        ("asym", E["fa"], E["asym"], False), ("flip", E["fa"], E["flip"], False),
        ("asym_fastq", E["fq"], E["asym"], False), ("asym_crlf", E["fa"], crlf, False),
        ("asym_noeol", E["fa"], E["asym"][:-1], False), ("asym_gz", E["fa"], E["asym"], True),
        ("numfmt", E["fa"], E["numfmt"], False),
        ("sim", E["sim_fa"], E["sim_paf"], False), ("emptyseq", E["empty_fa"], E["empty_paf"], False),
    ]
    for name, fa, paf, gz in cases:
        e = run_case(name, fa, paf, EDGE_ARGS, ed, True, gz=gz)
        e["inputs"] = dict(kind="edge", fa={"asym_fastq": "fq", "sim": "sim_fa", "emptyseq": "empty_fa"}.get(name, "fa"),
                           paf={"flip": "flip", "numfmt": "numfmt", "sim": "sim_paf", "emptyseq": "empty_paf"}.get(name, "asym"),
                           transform={"asym_crlf": "crlf", "asym_noeol": "noeol", "asym_gz": "gz"}.get(name))
        manifest.append(e)
    # default-parameter run of the edge fixture (p=10000 > read lengths: whole reads)
    e = run_case("asym_defaults", E["fa"], E["asym"], ["-e", "1"], ed, True)
    e["inputs"] = dict(kind="edge", fa="fa", paf="asym", transform=None)
    manifest.append(e)

    sd = os.path.join(GOLD, "synth")
    synth_cases = [("C1", 0.02, True), ("C1", 0.02, False), ("C2", 0.00003, True), ("C2", 0.00003, False),
                   ("C4", 0.0002, True), ("C4", 0.0002, False), ("C5", 0.0005, True), ("C5", 0.0005, False)]
    for cfg, scale, sym in synth_cases:
        ds = synth.make_dataset(cfg, scale, sym)
        name = f"{cfg}_{'sym' if sym else 'asym'}"
        fa = synth.format_fasta(ds.reads, wrap=80)
        with gzip.GzipFile(os.path.join(sd, name + ".fa.gz"), "wb", mtime=0) as fh:
            fh.write(fa)
        with gzip.GzipFile(os.path.join(sd, name + ".paf.gz"), "wb", mtime=0) as fh:
            fh.write(ds.paf)
        e = run_case(name, fa, ds.paf, ds.args, sd, False)
        e["inputs"] = dict(kind="synth", config=cfg, scale=scale, symmetric=sym, n_reads=ds.reads.n, n_overlaps=ds.n_overlaps)
        manifest.append(e)
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1)
    print("wrote", len(manifest), "cases")
    os.system(f"du -sh {GOLD}")


if __name__ == "__main__":
    main()
