"""Random small inputs inside the reference's defined domain (SURVEY.md §8.A preconditions): every PAF name is in
the FASTA, names unique, 0 <= s <= e <= L, p >= 1, l >= p, v <= p.  Shapes the closed form is sensitive to are drawn
on purpose: zero-length intervals, e == L, L < r, L a multiple of r or p, self overlaps, repeated records, a mirrored
first record (symmetric flag), junk / short lines, CR LF."""
import numpy as np


def fuzz_case(seed):
    rng = np.random.default_rng(seed)
    r = int(rng.choice([1, 3, 7, 10, 50]))
    p = int(rng.choice([r, 2 * r + 1, 40, 100, 250]))
    l = int(p * rng.integers(1, 4) + rng.integers(0, p))
    v = int(rng.integers(0, p + 1))
    f = int(rng.choice([0, 1, r, 3 * r + 2, 60]))
    e_cov = int(rng.integers(1, 6))
    mul = float(rng.choice([1.0, 1.5, 0.7, 2.25]))
    n = int(rng.integers(1, 12))
    lens = []
    for _ in range(n):
        kind = rng.integers(0, 6)
        if kind == 0:
            lens.append(int(rng.integers(0, r + 1)))                      # shorter than one bin (or empty)
        elif kind == 1:
            lens.append(int(r * rng.integers(1, 40)))                     # multiple of the resolution
        elif kind == 2:
            lens.append(int(p * rng.integers(1, 6)))                      # multiple of the star distance
        else:
            lens.append(int(rng.integers(1, 900)))
    names = [b"rd%d_%d" % (seed, i) for i in range(n)]
    wrap = int(rng.choice([0, 1, 7, 60]))
    fa = []
    for nm, L in zip(names, lens):
        seq = bytes(rng.choice(list(b"ACGTN"), L).astype(np.uint8))
        fa.append(b">" + nm + b" some comment\n")
        if wrap and L:
            fa.extend(seq[k:k + wrap] + b"\n" for k in range(0, L, wrap))
        else:
            fa.append(seq + b"\n")
    fa = b"".join(fa)
    lines = []
    n_rec = int(rng.integers(0, 60))
    usable = [i for i in range(n) if lens[i] > 0]
    for _ in range(n_rec if usable else 0):
        q, t = int(rng.choice(usable)), int(rng.choice(usable))

        def iv(L):
            k = rng.integers(0, 5)
            if k == 0:
                s = int(rng.integers(0, L + 1)); return s, s               # zero-length
            if k == 1:
                return int(rng.integers(0, L + 1)), L                      # ends at the read end
            if k == 2:
                return 0, int(rng.integers(0, L + 1))
            s = int(rng.integers(0, L + 1)); return s, int(rng.integers(s, L + 1))
        qs, qe = iv(lens[q]); ts, te = iv(lens[t])
        fields = [names[q], b"%d" % lens[q], b"%d" % qs, b"%d" % qe, b"+-"[int(rng.integers(0, 2))].to_bytes(1, "little"), names[t],
                  b"%d" % lens[t], b"%d" % ts, b"%d" % te, b"5", b"9", b"60"]
        if rng.integers(0, 6) == 0:
            fields.append(b"tp:A:P")
        lines.append(b"\t".join(fields))
    if lines and rng.integers(0, 3) == 0:                                   # mirror record 0 somewhere: symmetric flag
        f0 = lines[0].split(b"\t")
        m = [f0[5], f0[6], f0[7], f0[8], f0[4], f0[0], f0[1], f0[2], f0[3]] + f0[9:]
        lines.insert(int(rng.integers(1, len(lines) + 1)), b"\t".join(m))
    if lines and rng.integers(0, 3) == 0:                                   # non-records are skipped (paf.hpp:84,96-98)
        lines.insert(int(rng.integers(0, len(lines) + 1)), b"")
        lines.insert(int(rng.integers(0, len(lines) + 1)), b"short\tline\t3")
    eol = b"\r\n" if rng.integers(0, 5) == 0 else b"\n"
    paf = eol.join(lines) + (eol if lines and rng.integers(0, 4) else b"")
    args = ["-r", str(r), "-e", str(e_cov), "-m", repr(mul), "-l", str(l), "-p", str(p), "-f", str(f), "-v", str(v)]
    return fa, paf, args


def fuzz_fastq_text(seed):
    """Random FASTQ / mixed FASTA+FASTQ text inside kseq's grammar (kseq.h:240-298): wrapped bases and qualities,
    '+name' lines, quality lines that start with '@' or '+', empty records, missing final newline."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 9))
    out, lens = [], []
    for i in range(n):
        L = int(rng.choice([0, 1, 5, 60, 61, 200, int(rng.integers(1, 500))]))
        seq = bytes(rng.choice(list(b"ACGTN"), L).astype(np.uint8))
        name = b"fq%d_%d" % (seed, i)
        wrap = int(rng.choice([0, 0, 7, 60]))
        lines = lambda b: b"\n".join(b[k:k + wrap] for k in range(0, len(b), wrap)) if (wrap and b) else b
        if rng.integers(0, 4) == 0:                     # a FASTA record in between
            out.append(b">" + name + b" c\n" + lines(seq) + b"\n")
        else:
            q = bytes(rng.integers(33, 74, L).astype(np.uint8))
            if L and rng.integers(0, 3) == 0:
                q = (b"@" if rng.integers(0, 2) else b"+") + q[1:]
            out.append(b"@" + name + (b" comment" if rng.integers(0, 2) else b"") + b"\n" + lines(seq) + b"\n+" +
                       (name if rng.integers(0, 2) else b"") + b"\n" + lines(q) + b"\n")
        lens.append(L)
    text = b"".join(out)
    if rng.integers(0, 3) == 0:
        text = text[:-1]
    return text, lens
