"""Deterministic synthetic reads + all-vs-all PAF of the shapes BASELINE.json names.

Counter-hash based (no numpy RNG state), so the same (config, scale, seed) always yields the same
bytes.  Used by tests/ (small scales) and bench.py (host-side generation of the bench workload).
This module only builds INPUTS; it knows nothing about the fragmentation path.
"""
from dataclasses import dataclass, field

import numpy as np

_U = np.uint64
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[ord("A")], _COMP[ord("C")], _COMP[ord("G")], _COMP[ord("T")] = ord("T"), ord("G"), ord("C"), ord("A")


def mix64(x):
    x = np.asarray(x, dtype=_U).copy()
    with np.errstate(over="ignore"):
        x ^= x >> _U(33)
        x *= _U(0xFF51AFD7ED558CCD)
        x ^= x >> _U(33)
        x *= _U(0xC4CEB9FE1A85EC53)
        x ^= x >> _U(33)
    return x


def _h(seed, stream, idx):
    with np.errstate(over="ignore"):
        base = mix64(np.array([(seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF], dtype=_U))[0]
        return mix64(np.asarray(idx, dtype=_U) * _U(0x9E3779B97F4A7C15) + base)


def _unif(seed, stream, idx):
    return (_h(seed, stream, idx) >> _U(11)).astype(np.float64) * (1.0 / (1 << 53))


def _normal(seed, stream, idx):
    u1 = np.maximum(_unif(seed, stream, idx), 1e-300)
    u2 = _unif(seed, stream + 1, idx)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


@dataclass
class Reads:
    seq_off: np.ndarray
    seq: np.ndarray
    name_off: np.ndarray
    names: np.ndarray

    @property
    def n(self):
        return len(self.seq_off) - 1

    @property
    def lens(self):
        return np.diff(self.seq_off)

    def name(self, i):
        return bytes(self.names[self.name_off[i]:self.name_off[i + 1]])


@dataclass
class Dataset:
    reads: Reads
    paf: bytes
    args: list  # raft CLI flags for this config (without -o)
    n_overlaps: int
    meta: dict = field(default_factory=dict)


# ----------------------------------------------------------------------------- genome

def make_genome(G, seed, families=()):
    """families: iterable of (length, copies). Returns (uint8 bases, [(positions, length), ...])."""
    g = _ACGT[(_h(seed, 1, np.arange(G)) & _U(3)).astype(np.int64)]
    fams = []
    for fi, (length, copies) in enumerate(families):
        length = int(min(length, max(1, G // (2 * max(copies, 1)))))
        pos = np.sort((_unif(seed, 100 + fi, np.arange(copies)) * max(1, G - length)).astype(np.int64))
        # keep copies disjoint by pushing later copies right
        for k in range(1, copies):
            if pos[k] < pos[k - 1] + length:
                pos[k] = pos[k - 1] + length
        pos = pos[pos + length <= G]
        if len(pos) < 2:
            continue
        src = g[pos[0]:pos[0] + length].copy()
        for p_ in pos[1:]:
            g[p_:p_ + length] = src
        fams.append((pos, length))
    return g, fams


# ----------------------------------------------------------------------------- reads

def make_read_layout(G, coverage, seed, median, sigma, lo, hi, mixture=None):
    """Returns (start, length, strand) with read ids in random order w.r.t. genome position."""
    mean_len = median * np.exp(sigma * sigma / 2.0)
    if mixture is not None:
        w, median2, sigma2 = mixture
        mean_len = (1 - w) * mean_len + w * median2 * np.exp(sigma2 * sigma2 / 2.0)
    n = max(2, int(G * coverage / mean_len))
    idx = np.arange(n)
    z = _normal(seed, 10, idx)
    ln = median * np.exp(sigma * z)
    if mixture is not None:
        w, median2, sigma2 = mixture
        pick = _unif(seed, 12, idx) < w
        ln = np.where(pick, median2 * np.exp(sigma2 * z), ln)
    ln = np.clip(ln, lo, min(hi, G)).astype(np.int64)
    st = (_unif(seed, 13, idx) * (G - ln + 1)).astype(np.int64)
    strand = (_h(seed, 14, idx) & _U(1)).astype(np.int8)
    return st, ln, strand


def make_names(n, style, seed):
    if style == "uuid":
        hexd = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)
        a, b = _h(seed, 20, np.arange(n)), _h(seed, 21, np.arange(n))
        mat = np.full((n, 36), ord("-"), dtype=np.uint8)
        cols = [c for c in range(36) if c not in (8, 13, 18, 23)]
        for k, c in enumerate(cols):
            src = a if k < 16 else b
            mat[:, c] = hexd[((src >> _U(4 * (k % 16))) & _U(15)).astype(np.int64)]
        # uniqueness: overwrite the last 8 hex digits with the index
        for k in range(8):
            mat[:, 35 - k] = hexd[(np.arange(n) >> (4 * k)) & 15]
        return np.arange(n + 1, dtype=np.int64) * 36, mat.reshape(-1)
    if style == "ccs":
        strs = [b"m64011_190830_220126/%d/ccs" % i for i in range(n)]
    elif style == "short":
        strs = [b"r%d" % i for i in range(n)]
    else:
        raise ValueError(style)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in strs])
    return off, np.frombuffer(b"".join(strs), dtype=np.uint8).copy()


def build_reads(genome, st, ln, strand, names):
    n = len(st)
    seq_off = np.zeros(n + 1, dtype=np.int64)
    seq_off[1:] = np.cumsum(ln)
    seq = np.empty(int(seq_off[-1]), dtype=np.uint8)
    for i in range(n):
        s = genome[st[i]:st[i] + ln[i]]
        if strand[i]:
            s = _COMP[s[::-1]]
        seq[seq_off[i]:seq_off[i + 1]] = s
    return Reads(seq_off, seq, names[0], names[1])


# ----------------------------------------------------------------------------- overlaps

def _pairs_sorted(lo_a, hi_a, min_ovl):
    """All (i, j), i != j by index into arrays sorted by lo, with |[lo_i,hi_i) ∩ [lo_j,hi_j)| >= min_ovl,
    each unordered pair once with lo_i <= lo_j."""
    n = len(lo_a)
    # partners of i: j > i with lo_j <= hi_i - min_ovl  (then check hi_j)
    last = np.searchsorted(lo_a, hi_a - min_ovl, side="right")
    cnt = np.maximum(last - (np.arange(n) + 1), 0)
    tot = int(cnt.sum())
    if tot == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z
    i = np.repeat(np.arange(n), cnt)
    first = np.cumsum(cnt) - cnt
    j = np.arange(tot) - np.repeat(first, cnt) + i + 1
    ok = np.minimum(hi_a[i], hi_a[j]) - lo_a[j] >= min_ovl
    return i[ok], j[ok]


def make_overlaps(st, ln, strand, fams, min_ovl, symmetric, max_per_read=None, seed=0, contained_full=False):
    """Returns dict of int arrays q,t,qs,qe,ts,te,rev (PAF order: grouped by query id, then target id)."""
    n = len(st)
    en = st + ln
    order = np.argsort(st, kind="stable")
    i, j = _pairs_sorted(st[order], en[order], min_ovl)
    a, b = order[i], order[j]
    os_, oe_ = np.maximum(st[a], st[b]), np.minimum(en[a], en[b])
    A = [a]; Bb = [b]; AS = [os_ - st[a]]; AE = [oe_ - st[a]]; BS = [os_ - st[b]]; BE = [oe_ - st[b]]
    # repeat-induced overlaps: reads over copy x vs reads over copy y, mapped through the copy offset
    for pos, length in fams:
        for x in range(len(pos)):
            for y in range(len(pos)):
                if x == y:
                    continue
                # clip every read to copy x / copy y, express in family coordinates
                lx = np.clip(st, pos[x], pos[x] + length) - pos[x]; hx = np.clip(en, pos[x], pos[x] + length) - pos[x]
                ly = np.clip(st, pos[y], pos[y] + length) - pos[y]; hy = np.clip(en, pos[y], pos[y] + length) - pos[y]
                ix = np.nonzero(hx - lx >= min_ovl)[0]; iy = np.nonzero(hy - ly >= min_ovl)[0]
                if len(ix) == 0 or len(iy) == 0 or x > y:
                    continue  # unordered copy pairs once (x<y); direction doubling happens below
                # all pairs between ix and iy with enough intersection (sizes here are small: reads over one copy)
                lo = np.maximum(lx[ix][:, None], ly[iy][None, :]); hi = np.minimum(hx[ix][:, None], hy[iy][None, :])
                pi, pj = np.nonzero(hi - lo >= min_ovl)
                ra, rb = ix[pi], iy[pj]
                keep = ra != rb
                ra, rb, lo_, hi_ = ra[keep], rb[keep], lo[pi, pj][keep], hi[pi, pj][keep]
                A.append(ra); Bb.append(rb)
                AS.append(lo_ + pos[x] - st[ra]); AE.append(hi_ + pos[x] - st[ra])
                BS.append(lo_ + pos[y] - st[rb]); BE.append(hi_ + pos[y] - st[rb])
    a = np.concatenate(A); b = np.concatenate(Bb)
    as_, ae_, bs_, be_ = map(np.concatenate, (AS, AE, BS, BE))
    # strand flip: coordinates on a reverse-strand read are mirrored
    def flip(s, e, r):
        rv = strand[r] == 1
        return np.where(rv, ln[r] - e, s), np.where(rv, ln[r] - s, e)
    as_, ae_ = flip(as_, ae_, a)
    bs_, be_ = flip(bs_, be_, b)
    rev = (strand[a] != strand[b]).astype(np.int8)
    if contained_full:
        # contained reads report the whole read (qs=0, qe=ql): what hifiasm/minimap2 emit for containment
        pass  # already true for error-free simulated containment (intersection == whole read)
    if symmetric:
        q = np.concatenate([a, b]); t = np.concatenate([b, a])
        qs = np.concatenate([as_, bs_]); qe = np.concatenate([ae_, be_])
        ts = np.concatenate([bs_, as_]); te = np.concatenate([be_, ae_])
        rev = np.concatenate([rev, rev])
    else:
        # asymmetric: one direction per pair, query = smaller id
        sw = a > b
        q = np.where(sw, b, a); t = np.where(sw, a, b)
        qs = np.where(sw, bs_, as_); qe = np.where(sw, be_, ae_)
        ts = np.where(sw, as_, bs_); te = np.where(sw, ae_, be_)
    key = np.lexsort((ts, t, q))
    q, t, qs, qe, ts, te, rev = (x[key] for x in (q, t, qs, qe, ts, te, rev))
    if max_per_read is not None:
        # cap overlaps per query (hifiasm-like -N); keep the first max_per_read of each group
        first = np.searchsorted(q, q, side="left")
        keep = (np.arange(len(q)) - first) < max_per_read
        q, t, qs, qe, ts, te, rev = (x[keep] for x in (q, t, qs, qe, ts, te, rev))
    return dict(q=q, t=t, qs=qs, qe=qe, ts=ts, te=te, rev=rev)


# ----------------------------------------------------------------------------- text

def _dec_lens(v):
    v = np.asarray(v, dtype=np.int64)
    d = np.ones(len(v), dtype=np.int64)
    p = 10
    for _ in range(18):
        d += v >= p
        p *= 10
    return d


def _ragged_copy(dst, dst_off, src, src_off, lens):
    tot = int(lens.sum())
    if tot == 0:
        return
    rep_d = np.repeat(dst_off - (np.cumsum(lens) - lens), lens)
    rep_s = np.repeat(src_off - (np.cumsum(lens) - lens), lens)
    ar = np.arange(tot)
    dst[rep_d + ar] = src[rep_s + ar]


def _put_dec(dst, off, v, nd):
    v = np.asarray(v, dtype=np.int64).copy()
    for d in range(int(nd.max()) if len(nd) else 0):
        m = nd > d
        dst[(off + nd - 1 - d)[m]] = (ord("0") + v[m] % 10).astype(np.uint8)
        v //= 10


def format_paf(reads: Reads, ov, chunk=1 << 20, crlf=False) -> bytes:
    """12-column PAF: qn ql qs qe strand tn tl ts te nmatch blen 255."""
    out = []
    lens = reads.lens
    nlen = np.diff(reads.name_off)
    N = len(ov["q"])
    for c0 in range(0, N, chunk):
        sl = slice(c0, min(N, c0 + chunk))
        q, t = ov["q"][sl], ov["t"][sl]
        ql, tl = lens[q], lens[t]
        qs, qe, ts, te = ov["qs"][sl], ov["qe"][sl], ov["ts"][sl], ov["te"][sl]
        blen = np.maximum(qe - qs, te - ts)
        nums = [ql, qs, qe, tl, ts, te, blen, blen]
        nds = [_dec_lens(x) for x in nums]
        eol = 2 if crlf else 1
        line_len = nlen[q] + nlen[t] + sum(nds) + 2 + 3 + 11 + eol  # names, digits, strand+"255"... see below
        # layout: qn \t ql \t qs \t qe \t S \t tn \t tl \t ts \t te \t nm \t bl \t 255 EOL  -> 11 tabs, 1 strand, 3 for 255
        line_len = nlen[q] + nlen[t] + sum(nds) + 11 + 1 + 3 + eol
        off = np.cumsum(line_len) - line_len
        buf = np.full(int(line_len.sum()), ord("\t"), dtype=np.uint8)
        cur = off.copy()
        _ragged_copy(buf, cur, reads.names, reads.name_off[q], nlen[q]); cur = cur + nlen[q] + 1
        for k in range(3):
            _put_dec(buf, cur, nums[k], nds[k]); cur = cur + nds[k] + 1
        buf[cur] = np.where(ov["rev"][sl] == 1, ord("-"), ord("+")).astype(np.uint8); cur = cur + 2
        _ragged_copy(buf, cur, reads.names, reads.name_off[t], nlen[t]); cur = cur + nlen[t] + 1
        for k in range(3, 8):
            _put_dec(buf, cur, nums[k], nds[k]); cur = cur + nds[k] + 1
        buf[cur] = ord("2"); buf[cur + 1] = ord("5"); buf[cur + 2] = ord("5")
        if crlf:
            buf[cur + 3] = ord("\r"); buf[cur + 4] = ord("\n")
        else:
            buf[cur + 3] = ord("\n")
        out.append(buf.tobytes())
    return b"".join(out)


def format_fasta(reads: Reads, wrap=None, fastq=False, comment=None) -> bytes:
    parts = []
    for i in range(reads.n):
        s = bytes(reads.seq[reads.seq_off[i]:reads.seq_off[i + 1]])
        hdr = (b"@" if fastq else b">") + reads.name(i) + ((b" " + comment) if comment else b"") + b"\n"
        parts.append(hdr)
        if wrap and not fastq:
            parts.append(b"\n".join(s[k:k + wrap] for k in range(0, len(s), wrap)) + b"\n")
        else:
            parts.append(s + b"\n")
        if fastq:
            parts.append(b"+\n" + b"I" * len(s) + b"\n")
    return b"".join(parts)


# ----------------------------------------------------------------------------- configs

CONFIGS = {
    # name: genome, coverage, (median, sigma, lo, hi), mixture, families, names, args, seed, max_per_read
    "C1": dict(G=2_000_000, cov=42, dist=(15000, 0.3, 1000, 60000), mixture=None,
               families=[(30000, 6)], names="ccs", args=["-e", "42"], seed=11, cap=None),
    "C2": dict(G=3_100_000_000, cov=32, dist=(22000, 0.55, 1000, 300000), mixture=None,
               families="segdup0.5", names="uuid", args=["-e", "32"], seed=22, cap=None),
    "C4": dict(G=500_000_000, cov=30, dist=(8000, 0.5, 1000, 1_500_000), mixture=(0.4, 90000, 0.6),
               families=[], names="uuid", args=["-e", "30"], seed=44, cap=None),
    "C5": dict(G=200_000_000, cov=30, dist=(18000, 0.3, 1000, 100000), mixture=None,
               families="repeat-heavy", names="uuid",
               args=["-e", "30", "-r", "10", "-p", "7000", "-f", "500", "-v", "500", "-l", "15000"],
               seed=55, cap=2000),
}


def _families_for(cfg, G, seed):
    fam = cfg["families"]
    if fam == "segdup0.5":
        # ~0.5 % of the genome in 2-copy segmental duplications of 20-60 kb
        k = max(1, int(G * 0.005 / (2 * 40000)))
        return [(int(20000 + 40000 * u), 2) for u in _unif(seed, 200, np.arange(k))]
    if fam == "repeat-heavy":
        sc = G / 200_000_000
        segs = [(int(10000 + 190000 * u), int(2 + 8 * w)) for u, w in
                zip(_unif(seed, 201, np.arange(20)), _unif(seed, 202, np.arange(20)))]
        segs = [(max(8000, int(l * max(sc, 0.05))), c) for l, c in segs][:max(2, int(20 * min(1.0, sc * 8)))]
        return segs
    return list(fam)


def make_dataset(name, scale=1.0, symmetric=True, seed=None, min_ovl=2000, names=None) -> Dataset:
    """Build config `name` with its genome scaled by `scale` (reads/overlaps scale with it)."""
    cfg = CONFIGS[name]
    seed = cfg["seed"] if seed is None else seed
    G = max(50_000, int(cfg["G"] * scale))
    fams = _families_for(cfg, G, seed)
    genome, fam_pos = make_genome(G, seed, fams)
    med, sig, lo, hi = cfg["dist"]
    st, ln, strand = make_read_layout(G, cfg["cov"], seed, med, sig, lo, hi, cfg["mixture"])
    nm = make_names(len(st), names or cfg["names"], seed)
    reads = build_reads(genome, st, ln, strand, nm)
    ov = make_overlaps(st, ln, strand, fam_pos, min_ovl, symmetric, cfg["cap"], seed)
    paf = format_paf(reads, ov)
    return Dataset(reads, paf, list(cfg["args"]), len(ov["q"]),
                   dict(config=name, scale=scale, symmetric=symmetric, seed=seed, genome=G,
                        n_reads=reads.n, bases=int(reads.seq_off[-1]), paf_bytes=len(paf)))
