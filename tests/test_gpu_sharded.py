"""Multi-rank CUDA parity: the ranks' output slices, concatenated, are byte-identical to the single-rank
oracle result.  With >= 2 GPUs the exchange is NCCL all_to_all over NVLink; on a 1-GPU box both ranks share
cuda:0 and the exchange is staged through gloo (same protocol, same kernels)."""
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
    backend = "nccl" if ngpu >= world else "gloo"
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
