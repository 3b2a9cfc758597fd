"""Multi-GPU fragmentation: one process per GPU, reads partitioned into contiguous id ranges.

The reference has no distributed mode; the path shards naturally because the
coverage of a read depends only on the records that name it (repeat.hpp:48-58) and everything after
coverage is per read.

The sharded run itself lives in the library (raftgpu_run_sharded, csrc/api.cu: NCCL collectives and
ncclSend/ncclRecv on the context's stream); `bench` below drives it under torchrun.  `run_rank` is the same
protocol spelled out over the caller-driven building blocks and torch.distributed, kept for the CPU test of the
protocol (gloo, numpy stand-in engine) and for boxes with a single GPU, where NCCL cannot put two ranks on one device:

  1. every rank tokenises its byte range of the PAF (split at newlines) against the full name table;
  2. record 0 of the whole file (chop.hpp:171-184 compares every later record with it) is taken from
     the first rank that has a record and broadcast; the symmetric flag is the max over ranks;
  3. intervals on reads the rank owns are scattered locally; every other contributing interval becomes
     a 12-byte endpoint (global read id, start, end) for the rank that owns the read: counts
     all-to-all, then ONE data all-to-all;
  4. each rank accumulates the endpoints it received, finalises its own reads, and the global
     `read=` numbering is fixed by an all-gather of fragment counts (chop.hpp:195,266,319).

Each rank's outputs are a contiguous slice of each output file, in read order.
`engine` is anything with the raft_b200.api.Context methods used below (the CPU tests plug in a
numpy stand-in); `comm` wraps torch.distributed.
"""
import numpy as np


def partition_reads(lengths, reso, world):
    """Read-id range boundaries (int64[world+1]) balanced by coverage slots (bins + 1 per read)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    slots = (lengths + reso - 1) // reso + 1
    csum = np.concatenate([[0], np.cumsum(slots)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(csum, total * r // world, side="left")))
    bounds.append(len(lengths))
    return np.maximum.accumulate(np.asarray(bounds, dtype=np.int64))


def split_text(nbytes, world, find_newline):
    """Byte ranges of a PAF for `world` ranks: rank r starts right after the first newline at or after
    nbytes*r/world (rank 0 at 0).  `find_newline(pos)` returns the offset of the first '\\n' at or after pos,
    or -1."""
    starts = [0]
    for r in range(1, world):
        p = nbytes * r // world
        nl = find_newline(p) if p > 0 else -1
        starts.append(nbytes if nl < 0 else nl + 1)
    starts.append(nbytes)
    starts = np.maximum.accumulate(np.asarray(starts, dtype=np.int64))
    return [(int(starts[r]), int(starts[r + 1])) for r in range(world)]


class TorchComm:
    """torch.distributed plumbing (nccl on GPUs, gloo in the CPU tests)."""

    def __init__(self, dist, device, buf_device=None):
        """device: where collectives run (cuda for nccl, cpu for gloo); buf_device: where endpoint buffers live
        (defaults to device; a CUDA buf_device with a cpu `device` stages the exchange through host memory)."""
        import torch
        self.dist, self.device, self.torch = dist, device, torch
        self.buf_device = device if buf_device is None else buf_device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._flat_gather = True

    def all_gather_i64(self, vals):
        """world x len(vals) int64 matrix; one flat output tensor and one device-to-host copy."""
        t = self.torch.tensor(list(vals), dtype=self.torch.int64, device=self.device)
        if self._flat_gather:
            try:
                out = self.torch.empty(self.world * t.numel(), dtype=self.torch.int64, device=self.device)
                self.dist.all_gather_into_tensor(out, t)
                return out.cpu().numpy().reshape(self.world, -1)
            except (RuntimeError, NotImplementedError):  # backend without the flat variant
                self._flat_gather = False
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])

    def all_to_all_counts(self, counts):
        """counts[d] = endpoints this rank sends to d; returns what every rank sends to this one.  An all-gather of the
        rows (the whole matrix is tiny) is one collective with a ring/tree schedule instead of world point-to-point pairs."""
        return self.all_gather_i64([int(c) for c in counts])[:, self.rank].astype(np.int64)

    def alloc_i32(self, n):
        return self.torch.empty(max(int(n), 1), dtype=self.torch.int32, device=self.buf_device)

    def all_to_all_v(self, recv, send, recv_counts, send_counts):
        n_s, n_r = int(sum(send_counts)), int(sum(recv_counts))
        osz, isz = [int(c) for c in recv_counts], [int(c) for c in send_counts]
        if self.buf_device == self.device:
            self.dist.all_to_all_single(recv[:n_r], send[:n_s], output_split_sizes=osz, input_split_sizes=isz)
            if recv.is_cuda:
                # the library launches on its own stream: the received endpoints must have landed before it reads them
                self.torch.cuda.current_stream(recv.device).synchronize()
        else:  # staged (tests: gloo collectives with CUDA buffers)
            r = self.torch.empty(max(n_r, 1), dtype=self.torch.int32, device=self.device)
            self.dist.all_to_all_single(r[:n_r], send[:n_s].to(self.device), output_split_sizes=osz, input_split_sizes=isz)
            recv[:n_r].copy_(r[:n_r])
            if recv.is_cuda:
                self.torch.cuda.current_stream(recv.device).synchronize()


def run_rank(engine, comm, bounds, text, nbytes):
    """Steps 1-4 for one rank.  Returns (stats, info dict)."""
    rank, world = comm.rank, comm.world
    import time
    tm, t_last = {}, time.perf_counter()

    def lap(name):  # host wall time of each protocol phase (every phase ends in a host-visible result)
        nonlocal t_last
        now = time.perf_counter()
        tm[name] = tm.get(name, 0.0) + (now - t_last) * 1e3
        t_last = now

    # -- record 0 of the whole file
    rec, found = engine.peek_first_record(text, nbytes)
    lap("peek")
    allrec = comm.all_gather_i64([1 if found else 0] + list(rec))
    lap("gather_rec0")
    holders = [r for r in range(world) if allrec[r][0]]
    if holders:
        h = holders[0]
        engine.set_first_record([int(x) for x in allrec[h][1:7]], is_local=(h == rank))
    else:
        engine.set_first_record(None, is_local=False)
    # -- tokenise the local byte range
    engine.ingest_paf(text, nbytes, last=True)
    lap("tokenize")
    # -- symmetric flag: OR over ranks
    sym = int(comm.all_gather_i64([engine.get_symmetric()]).max())
    engine.set_symmetric(sym)
    lap("gather_sym")
    # -- intervals on reads this rank owns are scattered directly; the rest is routed to the owners
    engine.accumulate_local()
    counts = engine.route_count(bounds)
    lap("route_count")
    recv_counts = comm.all_to_all_counts(counts)
    lap("a2a_counts")
    send = comm.alloc_i32(3 * int(counts.sum()))
    engine.route_pack(bounds, counts, send)
    recv = comm.alloc_i32(3 * int(recv_counts.sum()))
    comm.all_to_all_v(recv, send, 3 * recv_counts, 3 * counts)
    lap("a2a_data")
    engine.accumulate_endpoints(recv, int(recv_counts.sum()))
    # -- local coverage / repeats / cut points, then global numbering
    st = engine.finalize()
    lap("finalize")
    per_rank = comm.all_gather_i64([int(st.n_fragments), int(st.n_records)])
    first_num = 1 + int(per_rank[:rank, 0].sum())
    engine.set_output_base(first_num)
    lap("gather_frags")
    info = dict(symmetric=sym, sent=int(counts.sum()), received=int(recv_counts.sum()), first_read_num=first_num,
                n_records_total=int(per_rank[:, 1].sum()), n_fragments_total=int(per_rank[:, 0].sum()),
                sent_remote=int(counts.sum()), phase_ms=tm)
    return st, info


# ------------------------------------------------------------------------------------------- bench (N > 1)
def bench(a, rank, world, local, log):
    """Strong-scaling bench (BASELINE.json configs[2]): the SAME inputs as the 1-GPU run, reads sharded by id range and the
    PAF split by line range over `world` GPUs.  The sharded path runs inside the library (raftgpu_run_sharded: NCCL
    collectives + ncclSend/ncclRecv on the library stream); torch.distributed only carries the communicator id, the
    barriers around the timed region and the reduction of the reported numbers."""
    import json
    import os
    import torch
    import torch.distributed as dist
    from . import api, synth_gpu
    import bench as B

    dev = torch.device("cuda", local)
    comm = TorchComm(dist, dev)
    WINDOW = 1 << 30
    sym = not a.asymmetric
    ds = synth_gpu.make_dataset_gpu(a.config, a.scale, symmetric=sym, device=f"cuda:{local}", with_seq=False, line_slice=(rank, world))
    p = api.AlgoParams.from_args(ds.args)
    lengths = ds.lengths.cpu().numpy()
    bounds = api.partition_reads(lengths, p.reso, world)
    b0, b1 = int(bounds[rank]), int(bounds[rank + 1])
    own_seq = synth_gpu.gen_seq(ds, b0, b1)
    own_off = (ds.seq_off[b0:b1 + 1] - ds.seq_off[b0]).contiguous()
    torch.cuda.synchronize()
    log(f"[bench r{rank}] reads {b0}..{b1} of {ds.n}, {own_seq.numel()} bases, local PAF {ds.paf.numel()} bytes / {ds.n_overlaps} lines")
    ctx = api.Context(p, local)
    ids = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(world, rank, ids[0])
    win = torch.empty(WINDOW + 64, dtype=torch.uint8, device=dev)
    win2 = torch.empty(WINDOW + 64, dtype=torch.uint8, device=dev)
    OUTS = (api.OUT_COVERAGE, api.OUT_LONG_REPEATS, api.OUT_READS_FASTA)

    phase = {}

    def step(host=None):
        import time
        t0 = time.perf_counter()
        if host is None:
            ctx.set_reads_sharded(ds.n, ds.lengths, ds.name_off, ds.names, b0, b1 - b0, own_off, own_seq)
            st, info = ctx.run_sharded(bounds, ds.paf, ds.paf.numel())
        else:
            ctx.set_reads_sharded(ds.n, host["lengths"], host["name_off"], host["names"], b0, b1 - b0, host["own_off"], host["own_seq"])
            phase["set_reads_ms"] = (time.perf_counter() - t0) * 1e3
            st, info = ctx.run_sharded(bounds, host["paf"], ds.paf.numel())
            phase["tokenize_device_ms"] = st.ms_tokenize
        t1 = time.perf_counter()
        nout = 0
        if host is None:  # device windows: text streams and the gather overlap on the library's two emit streams
            for which in OUTS:
                n = ctx.output_size(which)
                for off in range(0, n, WINDOW):
                    ctx.fetch_async(which, off, win2 if which != api.OUT_READS_FASTA else win, min(WINDOW, n - off))
                nout += n
            ctx.sync()
        else:
            dst = host["out"].data_ptr()
            for which in OUTS:
                n = ctx.output_size(which)
                for off in range(0, n, WINDOW):
                    ctx.fetch_into(which, off, dst, min(WINDOW, n - off))
                nout += n
        t2 = time.perf_counter()
        phase["inputs+run_ms"] = (t1 - t0) * 1e3
        phase["outputs_ms"] = (t2 - t1) * 1e3
        return st, info, nout

    def timed(k, host=None):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            st, info, nout = step(host)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / k], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        return float(t), st, info, nout

    def all_digests(info):
        """file digests = sum over ranks of the slice digests taken at their file offsets (mod 2^64); sizes = the gathered totals"""
        mine = [ctx.digest(w, int(info.stream_base[w])) for w in OUTS]
        allv = comm.all_gather_i64([v - (1 << 64) if v >= (1 << 63) else v for v in mine])
        tot = [int(sum(int(x) for x in allv[:, k])) % (1 << 64) for k in range(3)]
        return {name: [int(info.stream_total[w]), tot[k]] for k, (w, name) in enumerate(zip(OUTS, B.STREAMS))}

    for _ in range(a.warmup):
        step()
    clocks = B.ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, st, info, nout = timed(a.steps)
    clk = clocks.stop() if rank == 0 else None
    s2 = ctx.stats()
    frag = ctx.table(api.TAB_FRAG).reshape(-1, 3)
    fasta_alg = int((frag[:, 2].astype(np.int64) - frag[:, 1]).sum()) + ctx.output_size(api.OUT_READS_FASTA)  # bases gathered + bytes written
    del frag
    launches = s2.kernel_launches * a.steps
    fasta_ms = s2.ms_emit[3]
    stage = {"set_reads": s2.ms_set_reads, "tokenize": s2.ms_tokenize, "exchange": info.ms_exchange, "scan": s2.ms_scan, "repeat_cut": s2.ms_repeat_cut,
             "layout": s2.ms_layout, "emit_cov": s2.ms_emit[0], "emit_rep": s2.ms_emit[1], "emit_fasta": s2.ms_emit[3]}
    tot = comm.all_gather_i64([int(info.endpoints_sent), nout, int(own_seq.numel()), int(ds.paf.numel()), launches])
    # ---- parity (not timed): the files the ranks' slices add up to, against the oracle-verified digests of the 1-GPU run
    checks = []
    dig = all_digests(info)
    key = B.digest_key(a)
    exp, exp_src = B.load_expected(key)
    if exp is not None:
        checks.append({"what": f"{world}-GPU run (NCCL inside the library) vs the recorded digests of the oracle-verified 1-GPU run of the same inputs",
                       "vs": "n1", "source": exp_src, "scale": a.scale, "identical": exp == dig})
    e2e = None
    if not a.no_e2e:
        def pinned(t):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h
        host = dict(lengths=pinned(ds.lengths).numpy(), name_off=pinned(ds.name_off).numpy(), names=pinned(ds.names).numpy(),
                    own_off=pinned(own_off).numpy(), own_seq=pinned(own_seq).numpy(), paf=pinned(ds.paf).numpy(),
                    out=torch.empty(WINDOW, dtype=torch.uint8, pin_memory=True))
        ctx.set_option(api.OPT_DEFER_SEQ_UPLOAD, 1)
        step(host)
        k = max(1, min(a.steps, 3))
        ms_e, _, info_e, nout_e = timed(k, host)
        h2d = sum(int(v.nbytes) for kk, v in host.items() if kk != "out")
        io = comm.all_gather_i64([h2d, nout_e])
        e2e = {"value": info.n_records_total / (ms_e / 1e3), "unit": "overlaps/s", "h2d_bytes_per_step": int(io[:, 0].sum()),
               "d2h_bytes_per_step": int(io[:, 1].sum()), "ms_per_step": ms_e, "steps": k,
               "host_wall_rank0_last_step": {kk: round(v, 1) for kk, v in phase.items() if kk.endswith("_ms")}}
        dig_e = all_digests(info_e)
        checks.append({"what": "e2e run (pinned host inputs through the C ABI on every rank) vs the device-resident run", "vs": "n-gpu device-resident",
                       "scale": a.scale, "identical": dig_e == dig})
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bytes_alg = int(tot[:, 1].sum() + tot[:, 2].sum() + tot[:, 3].sum()) + world * int(ds.names.numel())
        best = checks[0] if checks else None
        parity = {"identical": all(c["identical"] for c in checks) if checks else None, "scale": best["scale"] if best else None,
                  "vs": best["vs"] if best else None, "digests": dig, "checks": checks}
        out = {"metric": "PAF overlaps/sec end-to-end fragmentation", "value": info.n_records_total / (ms / 1e3), "unit": "overlaps/s",
               "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "int32/u8", "data": "synthetic", "config": B.config_block(a, ds.args),
               "workload_sizes": {"n_overlaps": int(info.n_records_total), "n_reads": ds.n, "bases": int(tot[:, 2].sum()), "paf_bytes": int(tot[:, 3].sum()),
                                  "out_bytes": {name: v[0] for name, v in dig.items()}, "symmetric": int(info.symmetric),
                                  "n_fragments": int(info.n_fragments_total)},
               "sharding": {"how": f"the same inputs as the 1-GPU run: reads sharded by id range (balanced by coverage slots) over {world} B200, PAF split by "
                                   f"line range, exchange inside the library (raftgpu_run_sharded)",
                            "collectives": "ncclAllGather(record 0), ncclAllReduce(symmetric), ncclAllGather(counts), grouped ncclSend/ncclRecv(12-byte endpoints), "
                                           "ncclAllGather(fragment counts), ncclAllGather(output sizes)",
                            "endpoints_sent_to_other_ranks": int(tot[:, 0].sum()), "exchange_bytes": 12 * int(tot[:, 0].sum())},
               "gbp_per_s": int(tot[:, 2].sum()) / (ms / 1e3) / 1e9, "stage_ms_rank0": stage,
               "roofline": {"kernel": "k_fasta_emit", "bound": "hbm", "peak": peak, "unit": "GB/s", "traffic": None,
                            "achieved": None if not fasta_ms else fasta_alg / (fasta_ms / 1e3) / 1e9,
                            "frac": None if not fasta_ms else fasta_alg / (fasta_ms / 1e3) / 1e9 / peak,
                            "kernel_ms_per_step": fasta_ms, "launches_per_step": s2.emit_launches[3],
                            "note": "rank 0's gather kernel over its own slice: bases gathered + reads.fasta bytes written, CUDA events on the library stream"},
               "path_roofline": {"bytes_alg": bytes_alg, "achieved_gbs_per_gpu": bytes_alg / world / (ms / 1e3) / 1e9,
                                 "frac": bytes_alg / world / (ms / 1e3) / 1e9 / peak},
               "parity": parity, "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(tot[:, 4].sum()), "clocks": clk}
        print(json.dumps(out), flush=True)
        if parity["identical"] is False:
            log("[bench] PARITY CHECK FAILED: " + json.dumps(checks))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
