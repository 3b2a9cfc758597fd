"""Synthetic reads named like seqrequester simulations (simulated-read mode of the reference)."""
import numpy as np

from raft_b200 import synth


def sim_dataset(seed, n_chr=3):
    """Reads named like seqrequester simulations (chop.hpp:99-106): read=N,forward|reverse,position=a-b,length=L,chr"""
    cfg = synth.CONFIGS["C1"]
    G = 150_000
    genome, fams = synth.make_genome(G, seed, [(20000, 3)])
    st, ln, strand = synth.make_read_layout(G, 30, seed, 9000, 0.4, 500, 40000)
    strs = [b"read=%d,%s,position=%d-%d,length=%d,chr%d" % (i + 1, b"forward" if strand[i] == 0 else b"reverse", st[i], st[i] + ln[i], ln[i], i % n_chr + 1)
            for i in range(len(st))]
    off = np.zeros(len(strs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in strs])
    names = (off, np.frombuffer(b"".join(strs), dtype=np.uint8).copy())
    reads = synth.build_reads(genome, st, ln, strand, names)
    ov = synth.make_overlaps(st, ln, strand, fams, 1000, seed % 2 == 0)
    return reads, synth.format_paf(reads, ov)
