"""World-size-2 gloo test of the multi-GPU protocol (raft_b200/sharded.py) on CPU.

The per-rank compute is a numpy/oracle stand-in with the same methods as raft_b200.api.Context, so
what is exercised here is the host logic: read partitioning, PAF byte-range splitting, record-0 /
symmetric-flag agreement, endpoint routing through all_to_all, and the global read= numbering.
The concatenation of the ranks' slices must equal the single-rank oracle result."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_util import args_to_kw
from oracle import oracle as O
from raft_b200 import sharded, synth


class Stats:
    pass


class NumpyEngine:
    def __init__(self, reads, kw, own_first, own_count):
        self.reads, self.kw, self.b0, self.b1 = reads, kw, own_first, own_first + own_count
        self.rec0, self.is_local, self.sym = None, False, 0
        self.ep = np.zeros((0, 3), np.int32)

    def _parse(self, text):
        r = O.run(self.reads, bytes(text), O.make_params(**self.kw), text=False)
        assert r.status in (0, -4), r.status  # range errors belong to coverage, not to parsing
        return r

    def peek_first_record(self, text, nbytes):
        r = self._parse(text[:nbytes])
        if r.status == 0 and r.n_rec > 0:
            return [int(x[0]) for x in (r.qid, r.tid, r.qs, r.qe, r.ts, r.te)], True
        return [0] * 6, False

    def set_first_record(self, rec, is_local):
        self.rec0, self.is_local = rec, is_local

    def ingest_paf(self, text, nbytes, last=True):
        self.r = self._parse(text[:nbytes])

    def get_symmetric(self):
        r, z = self.r, self.rec0
        if z is None or r.n_rec == 0:
            return 0
        m = (r.tid == z[0]) & (r.qid == z[1]) & (r.ts == z[2]) & (r.te == z[3]) & (r.qs == z[4]) & (r.qe == z[5])
        if self.is_local:
            m[0] = False  # record 0 is not compared with itself (chop.hpp:171)
        return int(m.any())

    def set_symmetric(self, f):
        self.sym = f

    def _all_endpoints(self):
        r = self.r
        e = [np.stack([r.qid, r.qs, r.qe], 1)]
        if not self.sym:
            k = r.tid != r.qid
            e.append(np.stack([r.tid[k], r.ts[k], r.te[k]], 1))
        return np.concatenate(e).astype(np.int32)

    def _endpoints(self):  # the ones that must travel
        e = self._all_endpoints()
        return e[(e[:, 0] < self.b0) | (e[:, 0] >= self.b1)]

    def accumulate_local(self):
        e = self._all_endpoints()
        self.local_ep = e[(e[:, 0] >= self.b0) & (e[:, 0] < self.b1)]

    def route_count(self, bounds):
        e = self._endpoints()
        owner = np.searchsorted(bounds, e[:, 0], side="right") - 1
        return np.bincount(owner, minlength=len(bounds) - 1).astype(np.int64)

    def route_pack(self, bounds, counts, sendbuf):
        e = self._endpoints()
        owner = np.searchsorted(bounds, e[:, 0], side="right") - 1
        e = e[np.argsort(owner, kind="stable")]
        sendbuf[:e.size] = torch.from_numpy(e.reshape(-1).copy())

    def accumulate_endpoints(self, recv, count):
        self.ep = np.concatenate([self.local_ep, recv[:3 * count].numpy().reshape(-1, 3)])

    def finalize(self):
        # coverage of the owned reads from the routed endpoints: feed them back to the oracle as self-overlaps
        assert ((self.ep[:, 0] >= self.b0) & (self.ep[:, 0] < self.b1)).all()
        rd = self.reads
        lens = np.diff(rd.seq_off)
        lines = []
        for i, s, e in self.ep:
            nm = bytes(rd.names[rd.name_off[i]:rd.name_off[i + 1]])
            lines.append(b"\t".join([nm, b"%d" % lens[i], b"%d" % s, b"%d" % e, b"+", nm, b"%d" % lens[i], b"%d" % s, b"%d" % e, b"0"]))
        res = O.run(rd, b"\n".join(lines) + (b"\n" if lines else b""), O.make_params(**self.kw))
        assert res.status == 0
        self.res = res
        own = (res.frag_read >= self.b0) & (res.frag_read < self.b1)
        self.frag = np.stack([res.frag_read[own], res.frag_a[own], res.frag_b[own]], 1)
        st = Stats()
        st.n_fragments, st.n_records = int(own.sum()), int(self.r.n_rec)
        return st

    def set_output_base(self, first_num):
        self.first_num = first_num

    def outputs(self):
        res = self.res
        cov_lines = res.cov_txt.split(b"\n")[:-1][self.b0:self.b1]
        rep_lines = res.rep_txt.split(b"\n")[:-1][self.b0:self.b1]
        return dict(cov=b"".join(l + b"\n" for l in cov_lines), rep=b"".join(l + b"\n" for l in rep_lines), frag=self.frag,
                    first_num=self.first_num, bin_cov=res.cov[res.bin_off[self.b0]:res.bin_off[self.b1]])


def _worker(rank, world, port, cfg, scale, sym, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = synth.make_dataset(cfg, scale, sym, seed=99)
    kw = args_to_kw(ds.args)
    bounds = sharded.partition_reads(ds.reads.lens, kw.get("reso", 50), world)
    paf = ds.paf
    ranges = sharded.split_text(len(paf), world, lambda p: paf.find(b"\n", p))
    lo, hi = ranges[rank]
    eng = NumpyEngine(ds.reads, kw, int(bounds[rank]), int(bounds[rank + 1] - bounds[rank]))
    comm = sharded.TorchComm(dist, torch.device("cpu"))
    st, info = sharded.run_rank(eng, comm, bounds, paf[lo:hi], hi - lo)
    out = eng.outputs()
    out["info"] = info
    torch.save(out, os.path.join(outdir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg,scale,sym", [("C1", 0.03, False), ("C1", 0.03, True), ("C5", 0.0008, False)])
def test_two_rank_protocol_matches_single_rank(cfg, scale, sym):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, cfg, scale, sym, d), nprocs=world, join=True)
        outs = [torch.load(os.path.join(d, f"r{r}.pt"), weights_only=False) for r in range(world)]
    ds = synth.make_dataset(cfg, scale, sym, seed=99)
    ref = O.run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    assert ref.status == 0
    assert b"".join(o["cov"] for o in outs) == ref.cov_txt
    assert b"".join(o["rep"] for o in outs) == ref.rep_txt
    np.testing.assert_array_equal(np.concatenate([o["frag"] for o in outs]), np.stack([ref.frag_read, ref.frag_a, ref.frag_b], 1))
    np.testing.assert_array_equal(np.concatenate([o["bin_cov"] for o in outs]), ref.cov)
    assert outs[0]["first_num"] == 1 and outs[1]["first_num"] == 1 + len(outs[0]["frag"])
    for o in outs:
        assert o["info"]["symmetric"] == ref.symmetric
        assert o["info"]["n_records_total"] == ref.n_rec and o["info"]["n_fragments_total"] == ref.n_frag
    if not sym:
        assert sum(o["info"]["sent_remote"] for o in outs) > 0  # the exchange really carried endpoints


def test_partition_and_split_helpers():
    lens = np.array([100, 0, 5000, 20, 20, 20, 70000, 1], dtype=np.int64)
    for world in (1, 2, 3, 8):
        b = sharded.partition_reads(lens, 50, world)
        assert b[0] == 0 and b[-1] == len(lens) and (np.diff(b) >= 0).all() and len(b) == world + 1
    text = b"aaa\nbbbbbb\ncc\n\nddddd"
    for world in (1, 2, 3, 5):
        r = sharded.split_text(len(text), world, lambda p: text.find(b"\n", p))
        assert r[0][0] == 0 and r[-1][1] == len(text)
        assert b"".join(text[a:b] for a, b in r) == text
        for a, b in r[1:]:
            assert a == len(text) or a == 0 or text[a - 1:a] == b"\n"
