"""Device-side synthetic inputs for bench.py: the same shapes as raft_b200.synth but
generated with torch ops + the kernels of libraft_synth.so directly in HBM, so human-scale inputs
need no host generator.  INPUT GENERATION ONLY — nothing here is on the fragmentation path."""
import ctypes as C
import math
import os

import torch

from . import synth

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        p = os.path.join(HERE, "libraft_synth.so")
        if not os.path.exists(p):
            raise RuntimeError(f"{p} missing: run make")
        L = C.CDLL(p)
        vp, i64, u64 = C.c_void_p, C.c_int64, C.c_uint64
        L.synth_names.argtypes = [vp, i64, u64, vp]
        L.synth_seq.argtypes = [vp, vp, vp, vp, i64, i64, u64, vp]
        L.synth_fasta_text.argtypes = [vp, vp, vp, vp, i64, i64, i64, u64, vp]
        L.synth_paf_sizes.argtypes = [vp] * 8 + [i64, u64, vp, vp]
        L.synth_paf_write.argtypes = [vp] * 8 + [i64, u64, vp, vp, vp]
        _LIB = L
    return _LIB


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class GpuDataset:
    """Device tensors: seq_off(int64[n+1]), seq(uint8), name_off(int64[n+1]), names(uint8), paf(uint8)."""

    def __init__(self):
        self.meta = {}

    def to_host_reads(self):
        return synth.Reads(self.seq_off.cpu().numpy(), self.seq.cpu().numpy(), self.name_off.cpu().numpy(), self.names.cpu().numpy())


def _pairs(lo, hi, min_ovl):
    """pairs (i<j by sorted position) of intervals sorted by lo with intersection >= min_ovl"""
    n = lo.numel()
    last = torch.searchsorted(lo, hi - min_ovl, right=True)
    ar = torch.arange(n, device=lo.device)
    cnt = (last - (ar + 1)).clamp_(min=0)
    tot = int(cnt.sum())
    if tot == 0:
        z = torch.zeros(0, dtype=torch.int64, device=lo.device)
        return z, z
    i = torch.repeat_interleave(ar, cnt)
    first = torch.cumsum(cnt, 0) - cnt
    j = torch.arange(tot, device=lo.device) - first[i] + i + 1
    ok = torch.minimum(hi[i], hi[j]) - lo[j] >= min_ovl
    return i[ok], j[ok]


def make_dataset_gpu(name, scale=1.0, symmetric=True, seed=None, device="cuda:0", min_ovl=2000, with_seq=True,
                     read_slice=None, line_slice=None) -> GpuDataset:
    cfg = synth.CONFIGS[name]
    seed = cfg["seed"] if seed is None else seed
    dev = torch.device(device)
    G = max(50_000, int(cfg["G"] * scale))
    med, sig, lo, hi = cfg["dist"]
    mean_len = med * math.exp(sig * sig / 2.0)
    if cfg["mixture"] is not None:
        w, med2, sig2 = cfg["mixture"]
        mean_len = (1 - w) * mean_len + w * med2 * math.exp(sig2 * sig2 / 2.0)
    n = max(2, int(G * cfg["cov"] / mean_len))
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    z = torch.randn(n, generator=gen, device=dev, dtype=torch.float64)
    ln = med * torch.exp(sig * z)
    if cfg["mixture"] is not None:
        pick = torch.rand(n, generator=gen, device=dev, dtype=torch.float64) < w
        ln = torch.where(pick, med2 * torch.exp(sig2 * z), ln)
    ln = ln.clamp_(lo, min(hi, G)).to(torch.int64)
    st = (torch.rand(n, generator=gen, device=dev, dtype=torch.float64) * (G - ln + 1).to(torch.float64)).to(torch.int64)
    strand = torch.randint(0, 2, (n,), generator=gen, device=dev, dtype=torch.int8)
    en = st + ln

    # ---- overlaps from true genomic intersections
    order = torch.argsort(st, stable=True)
    i, j = _pairs(st[order], en[order], min_ovl)
    a, b = order[i], order[j]
    del i, j
    os_, oe_ = torch.maximum(st[a], st[b]), torch.minimum(en[a], en[b])
    A, B = [a], [b]
    AS, AE, BS, BE = [os_ - st[a]], [oe_ - st[a]], [os_ - st[b]], [oe_ - st[b]]
    # ---- repeat-induced overlaps between the copies of each family
    fams = synth._families_for(cfg, G, seed)
    if fams:
        import numpy as np
        for fi, (length, copies) in enumerate(fams):
            length = int(min(length, max(1, G // (2 * max(copies, 1)))))
            pos = np.sort((synth._unif(seed, 100 + fi, np.arange(copies)) * max(1, G - length)).astype(np.int64))
            for k in range(1, copies):
                if pos[k] < pos[k - 1] + length:
                    pos[k] = pos[k - 1] + length
            pos = pos[pos + length <= G]
            for x in range(len(pos)):
                for y in range(x + 1, len(pos)):
                    px, py = int(pos[x]), int(pos[y])
                    lx = st.clamp(px, px + length) - px; hx = en.clamp(px, px + length) - px
                    ly = st.clamp(py, py + length) - py; hy = en.clamp(py, py + length) - py
                    ix = torch.nonzero(hx - lx >= min_ovl).flatten(); iy = torch.nonzero(hy - ly >= min_ovl).flatten()
                    if ix.numel() == 0 or iy.numel() == 0:
                        continue
                    lo_ = torch.maximum(lx[ix][:, None], ly[iy][None, :]); hi_ = torch.minimum(hx[ix][:, None], hy[iy][None, :])
                    pi, pj = torch.nonzero(hi_ - lo_ >= min_ovl, as_tuple=True)
                    ra, rb = ix[pi], iy[pj]
                    keep = ra != rb
                    ra, rb, l2, h2 = ra[keep], rb[keep], lo_[pi, pj][keep], hi_[pi, pj][keep]
                    A.append(ra); B.append(rb)
                    AS.append(l2 + px - st[ra]); AE.append(h2 + px - st[ra]); BS.append(l2 + py - st[rb]); BE.append(h2 + py - st[rb])
    a, b = torch.cat(A), torch.cat(B)
    as_, ae_, bs_, be_ = torch.cat(AS), torch.cat(AE), torch.cat(BS), torch.cat(BE)
    del A, B, AS, AE, BS, BE, os_, oe_

    def flip(s, e, r):
        rv = strand[r] == 1
        return torch.where(rv, ln[r] - e, s), torch.where(rv, ln[r] - s, e)

    as_, ae_ = flip(as_, ae_, a)
    bs_, be_ = flip(bs_, be_, b)
    rev = (strand[a] != strand[b]).to(torch.int8)
    if symmetric:
        q, t = torch.cat([a, b]), torch.cat([b, a])
        qs, qe = torch.cat([as_, bs_]), torch.cat([ae_, be_])
        ts, te = torch.cat([bs_, as_]), torch.cat([be_, ae_])
        rev = torch.cat([rev, rev])
    else:
        sw = a > b
        q, t = torch.where(sw, b, a), torch.where(sw, a, b)
        qs, qe = torch.where(sw, bs_, as_), torch.where(sw, be_, ae_)
        ts, te = torch.where(sw, as_, bs_), torch.where(sw, ae_, be_)
    del a, b, as_, ae_, bs_, be_
    key = torch.argsort(q * n + t, stable=True)  # grouped by query, then target (hifiasm-like)
    q, t, qs, qe, ts, te, rev = (x[key] for x in (q, t, qs, qe, ts, te, rev))
    del key
    if cfg["cap"] is not None:
        first = torch.searchsorted(q, q, right=False)
        keep = (torch.arange(q.numel(), device=dev) - first) < cfg["cap"]
        q, t, qs, qe, ts, te, rev = (x[keep] for x in (q, t, qs, qe, ts, te, rev))
    if read_slice is not None:
        # PAF lines whose query falls in [lo, hi): the byte range of the PAF file one rank would read
        keep = (q >= read_slice[0]) & (q < read_slice[1])
        q, t, qs, qe, ts, te, rev = (x[keep] for x in (q, t, qs, qe, ts, te, rev))
    if line_slice is not None:
        # lines [N*r/P, N*(r+1)/P): the byte range of the PAF file that rank r of P would read
        r_, P_ = line_slice
        lo_l, hi_l = q.numel() * r_ // P_, q.numel() * (r_ + 1) // P_
        q, t, qs, qe, ts, te, rev = (x[lo_l:hi_l] for x in (q, t, qs, qe, ts, te, rev))
    N = q.numel()

    # ---- PAF text
    L = _lib()
    cols = [x.contiguous() for x in (q, t, qs, qe, ts, te, rev, ln)]
    sizes = torch.empty(N, dtype=torch.int32, device=dev)
    L.synth_paf_sizes(*[c.data_ptr() for c in cols], N, seed, sizes.data_ptr(), _stream())
    off = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    torch.cumsum(sizes, 0, out=off[1:])
    total = int(off[-1])
    paf = torch.empty(total, dtype=torch.uint8, device=dev)
    L.synth_paf_write(*[c.data_ptr() for c in cols], N, seed, off.data_ptr(), paf.data_ptr(), _stream())
    torch.cuda.synchronize()
    del cols, sizes, off, q, t, qs, qe, ts, te, rev

    ds = GpuDataset()
    ds.n, ds.n_overlaps, ds.paf = n, N, paf
    ds.lengths = ln
    ds.name_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * 36
    ds.names = torch.empty(n * 36, dtype=torch.uint8, device=dev)
    L.synth_names(ds.names.data_ptr(), n, seed, _stream())
    seq_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(ln, 0, out=seq_off[1:])
    ds.seq_off = seq_off
    ds.bases = int(seq_off[-1])
    ds.start, ds.strand, ds.seed = st, strand, seed
    ds.seq = None
    if with_seq:
        ds.seq = gen_seq(ds, 0, n)
    ds.args = list(cfg["args"])
    ds.meta = dict(config=name, scale=scale, symmetric=symmetric, seed=seed, genome=G, n_reads=n, bases=ds.bases,
                   paf_bytes=total, n_overlaps=N)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return ds


def gen_seq(ds, r0, r1):
    """sequence bytes of reads [r0, r1) (uint8 device tensor) and nothing else"""
    L = _lib()
    off = (ds.seq_off[r0:r1 + 1] - ds.seq_off[r0]).contiguous()
    total = int(off[-1])
    seq = torch.empty(total + 32, dtype=torch.uint8, device=ds.seq_off.device)[:total]
    L.synth_seq(seq.data_ptr(), off.data_ptr(), ds.start[r0:r1].contiguous().data_ptr(), ds.strand[r0:r1].contiguous().data_ptr(),
                r1 - r0, total, ds.seed, _stream())
    return seq


def gen_fasta_text(ds, r0, r1, out=None):
    """Unwrapped FASTA text (">" name "\\n" bases "\\n") of reads [r0, r1) as a uint8 device tensor."""
    L = _lib()
    total = int(ds.seq_off[r1] - ds.seq_off[r0]) + 39 * (r1 - r0)
    if out is None or out.numel() < total:
        out = torch.empty(total + 64, dtype=torch.uint8, device=ds.seq_off.device)
    L.synth_fasta_text(out.data_ptr(), ds.seq_off.data_ptr(), ds.start.data_ptr(), ds.strand.data_ptr(), r0, r1 - r0, total, ds.seed, _stream())
    return out, total
