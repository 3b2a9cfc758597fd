"""The host FASTA/FASTQ(+gz) reader behind the C ABI (raftgpu_load_fasta, csrc/host_io.cpp: a streaming restatement of
kseq_read, kseq.h:240-298) against the oracle's parser, which is pinned to the reference binary.  No GPU involved."""
import gzip
import os
import tempfile

import numpy as np
import pytest

from fuzz_util import fuzz_case, fuzz_fastq_text
from oracle import oracle as O
from raft_b200 import api


def _same(text, path):
    ref = O.parse_fasta(text)
    seq_off, seq, name_off, names = api.load_fasta(path)
    np.testing.assert_array_equal(seq_off, ref.seq_off)
    np.testing.assert_array_equal(name_off, ref.name_off)
    assert bytes(seq) == bytes(ref.seq) and bytes(names) == bytes(ref.names)
    return ref.n


@pytest.mark.parametrize("block", range(3))
def test_host_reader_matches_oracle_on_fuzz_texts(block):
    n = 0
    with tempfile.TemporaryDirectory() as d:
        for seed in range(block * 20, block * 20 + 20):
            for k, text in enumerate((fuzz_fastq_text(seed)[0], fuzz_case(seed)[0])):
                p = os.path.join(d, f"t{seed}_{k}")
                open(p, "wb").write(text)
                n += _same(text, p)
                with gzip.open(p + ".gz", "wb") as f:       # gz is transparent (gzopen, chop.hpp:93)
                    f.write(text)
                _same(text, p + ".gz")
                crlf = text.replace(b"\n", b"\r\n")            # kseq strips a trailing CR of a line (kseq.h:189-190)
                open(p + ".crlf", "wb").write(crlf)
                _same(crlf, p + ".crlf")
    assert n > 50


def test_host_reader_edge_texts():
    cases = [b"", b"\n\n", b">a", b">a\n", b">a\nAC\n>", b"junk before\n>a\nACGT\n", b"@a\nACGT\n+\nIIII\n@b\nACGT\n+\nII\n",
             b">a b c\nAC\n\n\nGT\n>b\n\n>c\nA", b"@a\n\n+\n\n@b\nA\n+\nI"]
    with tempfile.TemporaryDirectory() as d:
        for k, text in enumerate(cases):
            p = os.path.join(d, f"e{k}")
            open(p, "wb").write(text)
            _same(text, p)
