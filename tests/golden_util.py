"""Shared helpers: load tests/golden cases (inputs + reference outputs)."""
import gzip
import hashlib
import json
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUFS = ("coverage.txt", "long_repeats.txt", "long_repeats.bed", "reads.fasta")
ARGMAP = {"-e": "est_cov", "-r": "reso", "-p": "repeat_length", "-f": "flanking_length",
          "-v": "overlap_length", "-l": "read_length", "-m": "cov_mul"}


def manifest():
    return json.load(open(os.path.join(GOLD, "manifest.json")))


def args_to_kw(args):
    kw = {}
    for k in range(0, len(args), 2):
        name = ARGMAP[args[k]]
        kw[name] = float(args[k + 1]) if name == "cov_mul" else int(args[k + 1])
    return kw


def load_inputs(entry):
    """Returns (fasta_bytes, paf_bytes) — already inflated / transformed as the reference saw them
    (gz is transparent to the reference: gzread inflates)."""
    inp = entry["inputs"]
    if inp["kind"] == "edge":
        fa = open(os.path.join(GOLD, "edge", "in." + inp["fa"]), "rb").read()
        paf = open(os.path.join(GOLD, "edge", "in." + inp["paf"]), "rb").read()
        tr = inp.get("transform")
        if tr == "crlf":
            paf = paf.replace(b"\n", b"\r\n")
        elif tr == "noeol":
            paf = paf[:-1]
        return fa, paf
    name = entry["name"]
    fa = gzip.open(os.path.join(GOLD, "synth", name + ".fa.gz"), "rb").read()
    paf = gzip.open(os.path.join(GOLD, "synth", name + ".paf.gz"), "rb").read()
    return fa, paf


def expected_full(entry, suf):
    p = os.path.join(GOLD, "edge", f"{entry['name']}.{suf}")
    return open(p, "rb").read() if os.path.exists(p) else None


def check_output(entry, suf, data: bytes):
    exp = entry["outputs"][suf]
    full = expected_full(entry, suf)
    if full is not None:
        assert data == full, f"{entry['name']}.{suf} differs from the reference output"
    assert len(data) == exp["len"], f"{entry['name']}.{suf}: length {len(data)} != {exp['len']}"
    assert hashlib.sha256(data).hexdigest() == exp["sha256"], f"{entry['name']}.{suf}: sha256 mismatch"


def stdout_value(entry, marker):
    for l in entry["stdout"]:
        if marker in l:
            return l
    return None
