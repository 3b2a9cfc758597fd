"""Rank body of the multi-rank CUDA parity test (launched by test_gpu_sharded.py or by torchrun).
Runs the sharded protocol with the real CUDA engine and writes each rank's output slices."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from raft_b200 import api, sharded, synth  # noqa: E402


def main():
    cfg, scale, sym, outdir, backend = sys.argv[1], float(sys.argv[2]), sys.argv[3] == "1", sys.argv[4], sys.argv[5]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ngpu = torch.cuda.device_count()
    local = int(os.environ.get("LOCAL_RANK", "0")) % ngpu
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if backend == "lib":      # the library's own NCCL communicator; torch.distributed only hands the id around
        dist.init_process_group("gloo")
        comm = None
    elif backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
        comm = sharded.TorchComm(dist, dev)
    else:
        dist.init_process_group("gloo")
        comm = sharded.TorchComm(dist, torch.device("cpu"), buf_device=dev)
    ds = synth.make_dataset(cfg, scale, sym, seed=99)
    p = api.AlgoParams.from_args(ds.args)
    lens = np.ascontiguousarray(ds.reads.lens, np.int64)
    bounds = sharded.partition_reads(lens, p.reso, world)
    b0, b1 = int(bounds[rank]), int(bounds[rank + 1])
    paf = ds.paf
    lo, hi = sharded.split_text(len(paf), world, lambda q: paf.find(b"\n", q))[rank]
    own_off = np.ascontiguousarray(ds.reads.seq_off[b0:b1 + 1] - ds.reads.seq_off[b0], np.int64)
    own_seq = np.ascontiguousarray(ds.reads.seq[ds.reads.seq_off[b0]:ds.reads.seq_off[b1]], np.uint8)
    ctx = api.Context(p, local)
    ctx.set_reads_sharded(ds.reads.n, lens, np.ascontiguousarray(ds.reads.name_off, np.int64), np.ascontiguousarray(ds.reads.names, np.uint8),
                          b0, b1 - b0, own_off, own_seq)
    text = np.frombuffer(paf[lo:hi], np.uint8)
    if backend == "lib":
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(world, rank, ids[0])
        assert np.array_equal(bounds, api.partition_reads(lens, p.reso, world))  # the C and the Python partition agree
        st, sh = ctx.run_sharded(bounds, text, hi - lo)
        info = dict(symmetric=sh.symmetric, n_records_total=sh.n_records_total, sent_remote=sh.endpoints_sent, received=sh.endpoints_received,
                    first_read_num=sh.first_read_num, n_fragments_total=sh.n_fragments_total, stream_base=list(sh.stream_base),
                    stream_total=list(sh.stream_total), backend="nccl-lib", peek_retries=sh.peek_retries,
                    digest=[ctx.digest(w, sh.stream_base[w]) for w in (api.OUT_COVERAGE, api.OUT_LONG_REPEATS, api.OUT_READS_FASTA)])
    else:
        st, info = sharded.run_rank(ctx, comm, bounds, text, hi - lo)
        info["backend"] = backend
    out = dict(info=info, cov=ctx.fetch(api.OUT_COVERAGE), rep=ctx.fetch(api.OUT_LONG_REPEATS), fasta=ctx.fetch(api.OUT_READS_FASTA),
               frag=ctx.table(api.TAB_FRAG).reshape(-1, 3), bin_cov=ctx.table(api.TAB_COV))
    torch.save(out, os.path.join(outdir, f"r{rank}.pt"))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
