"""The C ABI boundary without a GPU: libraft_b200.so loads, exports every function include/raft_b200.h declares (and the
ctypes binding declares exactly those), its host-only entry points work, and the compute entry points fail loudly
instead of falling back to anything when no sm_100 device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from raft_b200 import _lib, api, sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "raft_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(raftgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 45
    L = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/raft_b200.h but not exported by libraft_b200.so"
    assert sorted(_lib.SYMBOLS) == names, "raft_b200/_lib.py and include/raft_b200.h list different entry points"
    _lib.lib()  # resolves every symbol with its signature


def test_struct_layouts_match_the_header():
    # sizes the C compiler gives the same declarations (a mismatch would corrupt stats / shard info silently)
    src = '#include "raft_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(raftgpu_params), sizeof(raftgpu_stats), sizeof(raftgpu_shard_info));return 0;}\n'
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(_lib.Params), C.sizeof(_lib.Stats), C.sizeof(_lib.ShardInfo)]


def test_host_only_entry_points():
    L = _lib.lib()
    p = _lib.Params()
    L.raftgpu_default_params(C.byref(p))
    assert (p.reso, p.est_cov, p.cov_mul, p.repeat_length, p.interval_length, p.read_length, p.overlap_length, p.flanking_length) == \
        (50, 0, 1.5, 10000, 10000, 20000, 500, 1000)                      # param.hpp:18-31
    assert L.raftgpu_strerror(0) == b"ok" and b"another rank" in L.raftgpu_strerror(-14)
    rng = np.random.default_rng(7)
    for n, reso, world in ((1, 50, 4), (1000, 50, 8), (977, 10, 3), (5, 1, 64), (0, 50, 2)):
        lens = rng.integers(0, 300000, n).astype(np.int64)
        np.testing.assert_array_equal(api.partition_reads(lens, reso, world), sharded.partition_reads(lens, reso, world))


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(api.RaftError) as ei:
        api.Context(api.AlgoParams(est_cov=30))
    assert ei.value.status == -8                                          # RAFTGPU_E_CUDA: no usable sm_100 device
    with pytest.raises(api.RaftError):
        api.break_long_reads(os.path.join(ROOT, "tests", "golden", "edge", "in.fa"), os.path.join(ROOT, "tests", "golden", "edge", "in.asym"),
                             api.AlgoParams(est_cov=1, outputfilename="/tmp/raft_b200_nofallback"))
