"""Multi-rank CUDA parity: the ranks' output slices, concatenated, are byte-identical to the single-rank
oracle result.  With enough GPUs the run is the library's own sharded path (raftgpu_run_sharded: NCCL collectives and
ncclSend/ncclRecv over NVLink on the library stream) -- a box with >= `world` GPUs MUST take that path, there is no quiet
fallback; on a 1-GPU box the ranks share cuda:0 (NCCL refuses two ranks on one device) and the caller-driven building
blocks are exercised instead with the exchange staged through gloo (same kernels)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from golden_util import args_to_kw
from oracle import oracle as O
from raft_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("cfg,scale,sym,world,route_cap", [("C1", 0.2, False, 2, None), ("C1", 0.2, True, 2, None), ("C5", 0.003, False, 3, None),
                                                           ("C4", 0.002, False, 2, None), ("C1", 0.2, False, 2, 64)])
def test_sharded_cuda_matches_oracle(cfg, scale, sym, world, route_cap):
    """route_cap: RAFT_B200_ROUTE_CAP forces the two-pass packing (endpoint list overflow) instead of the collected list."""
    ngpu = torch.cuda.device_count()
    backend = "lib" if ngpu >= world else "gloo"
    port = 29600 + (os.getpid() % 2000)
    with tempfile.TemporaryDirectory() as d:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(HERE, "mgpu_worker.py"), cfg, str(scale), "1" if sym else "0", d, backend]
        env = dict(os.environ)
        if route_cap is not None:
            env["RAFT_B200_ROUTE_CAP"] = str(route_cap)
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900, env=env)
        assert r.returncode == 0, r.stdout.decode()[-4000:]
        outs = [torch.load(os.path.join(d, f"r{k}.pt"), weights_only=False) for k in range(world)]
    ds = synth.make_dataset(cfg, scale, sym, seed=99)
    ref = O.run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    assert ref.status == 0
    assert b"".join(o["cov"] for o in outs) == ref.cov_txt
    assert b"".join(o["rep"] for o in outs) == ref.rep_txt
    assert b"".join(o["fasta"] for o in outs) == ref.fasta
    np.testing.assert_array_equal(np.concatenate([o["frag"] for o in outs]), np.stack([ref.frag_read, ref.frag_a, ref.frag_b], 1))
    np.testing.assert_array_equal(np.concatenate([o["bin_cov"] for o in outs]), ref.cov)
    for o in outs:
        assert o["info"]["symmetric"] == ref.symmetric and o["info"]["n_records_total"] == ref.n_rec
    if not sym:
        assert sum(o["info"]["sent_remote"] for o in outs) > 0
    if ngpu >= world:  # enough GPUs: the NCCL path inside the library is the one that ran, and its bookkeeping is right
        assert all(o["info"]["backend"] == "nccl-lib" for o in outs)
        for k, data in enumerate((ref.cov_txt, ref.rep_txt, ref.fasta)):
            w = (0, 1, 3)[k]
            assert [o["info"]["stream_base"][w] for o in outs] == list(np.cumsum([0] + [len(o[("cov", "rep", "fasta")[k]]) for o in outs[:-1]]))
            assert all(o["info"]["stream_total"][w] == len(data) for o in outs)
            assert sum(o["info"]["digest"][k] for o in outs) % 2**64 == O.digest(data)   # per-rank digests at their file offsets add up
        assert sum(o["info"]["sent_remote"] for o in outs) == sum(o["info"]["received"] for o in outs)


def _oracle_case(cfg="C1", scale=0.2, sym=False, seed=99):
    ds = synth.make_dataset(cfg, scale, sym, seed=seed)
    ref = O.run(ds.reads, ds.paf, O.make_params(**args_to_kw(ds.args)))
    assert ref.status == 0
    return ds, ref


@pytest.mark.parametrize("chunked", [False, True])
def test_library_sharded_path_with_one_rank(chunked):
    """raftgpu_comm_init / raftgpu_run_sharded (and the chunked begin / ingest / finish form) with a one-rank NCCL
    communicator: every collective of the sharded path runs, the result is the single-GPU result."""
    from raft_b200 import api
    ds, ref = _oracle_case()
    p = api.AlgoParams.from_args(ds.args)
    lens = np.ascontiguousarray(ds.reads.lens, np.int64)
    bounds = api.partition_reads(lens, p.reso, 1)
    assert list(bounds) == [0, ds.reads.n]
    with api.Context(p, 0) as ctx:
        ctx.comm_init(1, 0, api.comm_unique_id())
        for rep in range(2):  # the communicator is reused across runs
            ctx.set_reads_sharded(ds.reads.n, lens, np.ascontiguousarray(ds.reads.name_off, np.int64), np.ascontiguousarray(ds.reads.names, np.uint8),
                                  0, ds.reads.n, np.ascontiguousarray(ds.reads.seq_off, np.int64), np.ascontiguousarray(ds.reads.seq, np.uint8))
            text = np.frombuffer(ds.paf, np.uint8)
            if not chunked:
                st, sh = ctx.run_sharded(bounds, text, len(ds.paf))
            else:
                import ctypes as C
                from raft_b200 import _lib
                cut = [0, 1000, len(ds.paf) // 3 + 5, len(ds.paf)]
                ctx._ck(ctx.L.raftgpu_sharded_begin(ctx._h, bounds.ctypes.data, text.ctypes.data, cut[1], 0))
                for a, b in zip(cut[:-1], cut[1:]):
                    ctx.ingest_paf(text[a:b], b - a, last=(b == len(ds.paf)))
                st, sh = _lib.Stats(), _lib.ShardInfo()
                ctx._ck(ctx.L.raftgpu_sharded_finish(ctx._h, C.byref(st), C.byref(sh)))
            assert (sh.n_records_total, sh.symmetric, sh.first_read_num, sh.n_fragments_total) == (ref.n_rec, ref.symmetric, 1, ref.n_frag)
            assert (sh.total_cov, sh.total_repeat_len, sh.total_read_len) == (ref.total_cov, ref.total_repeat_len, ref.total_read_len)
            assert list(sh.stream_base) == [0, 0, 0, 0] and sh.stream_total[3] == len(ref.fasta)
            assert ctx.fetch(api.OUT_COVERAGE) == ref.cov_txt and ctx.fetch(api.OUT_LONG_REPEATS) == ref.rep_txt
            assert ctx.fetch(api.OUT_READS_FASTA) == ref.fasta


def test_library_sharded_peek_retry_and_errors():
    """A PAF whose head holds no record: the record-0 protocol notices and redoes the step with whole-text peeks.
    Data errors come back as the rank's own status (no hang: one rank, but the same code path the ranks agree through)."""
    from raft_b200 import api
    ds, ref = _oracle_case(sym=True)
    p = api.AlgoParams.from_args(ds.args)
    lens = np.ascontiguousarray(ds.reads.lens, np.int64)
    junk = b"no tabs here\n" * 400
    paf = junk + ds.paf
    ref2 = O.run(ds.reads, paf, O.make_params(**args_to_kw(ds.args)))
    os.environ["RAFT_B200_PEEK_BYTES"] = "1024"
    try:
        with api.Context(p, 0) as ctx:
            ctx.comm_init(1, 0, api.comm_unique_id())
            args = (ds.reads.n, lens, np.ascontiguousarray(ds.reads.name_off, np.int64), np.ascontiguousarray(ds.reads.names, np.uint8),
                    0, ds.reads.n, np.ascontiguousarray(ds.reads.seq_off, np.int64), np.ascontiguousarray(ds.reads.seq, np.uint8))
            ctx.set_reads_sharded(*args)
            st, sh = ctx.run_sharded(np.array([0, ds.reads.n], np.int64), np.frombuffer(paf, np.uint8), len(paf))
            assert sh.peek_retries == 1 and sh.symmetric == ref2.symmetric == 1 and sh.n_records_total == ref2.n_rec
            assert ctx.fetch(api.OUT_COVERAGE) == ref2.cov_txt and ctx.fetch(api.OUT_READS_FASTA) == ref2.fasta
            bad = ds.paf + b"nobody\t10\t0\t5\t+\tnobody2\t10\t0\t5\t5\t5\t255\n"
            ctx.set_reads_sharded(*args)
            with pytest.raises(api.RaftError) as ei:
                ctx.run_sharded(np.array([0, ds.reads.n], np.int64), np.frombuffer(bad, np.uint8), len(bad))
            assert ei.value.status == -2
            ctx.set_reads_sharded(*args)   # the context and its communicator survive a failed run
            st, sh = ctx.run_sharded(np.array([0, ds.reads.n], np.int64), np.frombuffer(ds.paf, np.uint8), len(ds.paf))
            assert ctx.fetch(api.OUT_READS_FASTA) == ref.fasta
    finally:
        del os.environ["RAFT_B200_PEEK_BYTES"]


@pytest.mark.parametrize("ndev,fastq", [(2, False), (2, True), (3, False)])
def test_cli_on_several_gpus(ndev, fastq):
    """RAFT_B200_DEVICES=0,1[,2]: the `raft` CLI shards the run over the GPUs of the box inside one process (one thread and
    context per device, raftgpu_break_long_reads_mgpu) and writes the same four files and stdout lines as on one GPU.
    FASTQ reads take the host-reader route, FASTA is cut by byte range and tokenised on every GPU."""
    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    ds, ref = _oracle_case("C5", 0.004, False, seed=5)
    exe = os.path.join(os.path.dirname(HERE), "raft_b200", "raft")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "r.fa"), "wb").write(synth.format_fasta(ds.reads, wrap=None if fastq else 70, fastq=fastq))
        half = ds.paf.rfind(b"\n", 0, len(ds.paf) // 2) + 1
        open(os.path.join(d, "a.paf"), "wb").write(ds.paf[:half])
        if fastq:   # the second file gzip-compressed: it cannot be cut into byte ranges and goes to one rank as a whole
            import gzip
            with gzip.open(os.path.join(d, "b.paf"), "wb") as f:
                f.write(ds.paf[half:])
        else:
            open(os.path.join(d, "b.paf"), "wb").write(ds.paf[half:])
        outs = {}
        for tag, env in (("one", dict(os.environ, RAFT_B200_MULTI_PAF="1")),
                         ("many", dict(os.environ, RAFT_B200_MULTI_PAF="1", RAFT_B200_DEVICES=",".join(str(k) for k in range(ndev))))):
            r = subprocess.run([exe] + ds.args + ["-o", os.path.join(d, tag), os.path.join(d, "r.fa"), os.path.join(d, "a.paf"), os.path.join(d, "b.paf")],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, env=env)
            assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
            # stdout alone is compared: NCCL's version banner and debug lines must have gone to stderr
            outs[tag] = [l for l in r.stdout.decode().splitlines() if "program completed" not in l and "CMD:" not in l]
            for suf, data in zip(("coverage.txt", "long_repeats.txt", "long_repeats.bed", "reads.fasta"), (ref.cov_txt, ref.rep_txt, ref.bed_txt, ref.fasta)):
                assert open(os.path.join(d, f"{tag}.{suf}"), "rb").read() == data, (tag, suf)
        assert outs["one"] == outs["many"]


def test_cli_on_two_gpus_simulated_reads():
    """Simulated-read mode (seqrequester names: genome coordinates in the headers, long_repeats.bed) sharded over two GPUs:
    the per-rank header tables, the BED slices and their file offsets add up to the single-rank oracle files."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from sim_util import sim_dataset
    reads, paf = sim_dataset(5)
    args = ["-e", "30", "-p", "4000", "-l", "8000", "-f", "300", "-v", "200"]
    ref = O.run(reads, paf, O.make_params(est_cov=30, repeat_length=4000, read_length=8000, flanking_length=300, overlap_length=200))
    assert ref.status == 0 and ref.real_reads == 0 and len(ref.bed_txt) > 0
    exe = os.path.join(os.path.dirname(HERE), "raft_b200", "raft")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "r.fa"), "wb").write(synth.format_fasta(reads, wrap=60))
        open(os.path.join(d, "o.paf"), "wb").write(paf)
        r = subprocess.run([exe] + args + ["-o", os.path.join(d, "out"), os.path.join(d, "r.fa"), os.path.join(d, "o.paf")],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, env=dict(os.environ, RAFT_B200_DEVICES="0,1"))
        assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
        assert "Real Reads 0 " in r.stdout.decode()
        for suf, data in zip(("coverage.txt", "long_repeats.txt", "long_repeats.bed", "reads.fasta"), (ref.cov_txt, ref.rep_txt, ref.bed_txt, ref.fasta)):
            assert open(os.path.join(d, f"out.{suf}"), "rb").read() == data, suf


def test_cli_on_two_gpus_data_error_reaches_every_rank():
    """A PAF name that is not a read name, in the part of the file one rank reads: that rank reports it, the other one
    learns about it through the collectives (RAFTGPU_E_PEER) and the process ends with the reference-crash exit code instead
    of hanging in NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ds, ref = _oracle_case("C1", 0.2, False, seed=99)
    lines = ds.paf.split(b"\n")
    bad = lines[(3 * len(lines)) // 4].split(b"\t")
    bad[5] = b"not_a_read"
    lines[(3 * len(lines)) // 4] = b"\t".join(bad)
    exe = os.path.join(os.path.dirname(HERE), "raft_b200", "raft")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "r.fa"), "wb").write(synth.format_fasta(ds.reads, wrap=80))
        open(os.path.join(d, "o.paf"), "wb").write(b"\n".join(lines))
        r = subprocess.run([exe] + ds.args + ["-o", os.path.join(d, "out"), os.path.join(d, "r.fa"), os.path.join(d, "o.paf")],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, env=dict(os.environ, RAFT_B200_DEVICES="0,1"))
        assert r.returncode == 2, r.stdout.decode() + r.stderr.decode()
        assert b"not in the reads file" in r.stderr
