"""Host file helpers of the file-level drivers (raft_b200/csrc/file_io.h: sliced pread / pwrite, the mapped writer of an
output slice) compiled into a CPU harness: ragged slices written concurrently by three 'ranks' through a shared mapping
and through pwrite, on tmpfs and on a disk directory; short reads at the end of a file."""
import os
import shutil
import subprocess
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness():
    d = tempfile.mkdtemp(prefix="raft_b200_hio_")
    exe = os.path.join(d, "file_io_harness")
    cxx = "g++" if shutil.which("g++") else "c++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-pthread", "-Wall", os.path.join(HERE, "host", "file_io_harness.cpp"), "-o", exe], check=True)
    yield exe
    shutil.rmtree(d, ignore_errors=True)


def _scratch_dirs():
    dirs = [tempfile.gettempdir()]
    if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > (1 << 30):
        dirs.append("/dev/shm")
    return dirs


@pytest.mark.parametrize("where", _scratch_dirs())
def test_file_io_harness(harness, where):
    d = tempfile.mkdtemp(prefix="raft_b200_hio_", dir=where)
    try:
        r = subprocess.run([harness, d], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        lines = [ln.split() for ln in r.stdout.splitlines() if ln.strip()]
        assert r.returncode == 0, r.stdout
        names = {ln[0] for ln in lines}
        assert {"slices_mapped", "slices_pwrite", "grow_only", "dev_null", "empty_slice", "pread_short_at_eof", "pread_at_eof",
                "map_threads_env"} <= names
        assert all(ln[1] == "ok" for ln in lines), r.stdout
    finally:
        shutil.rmtree(d, ignore_errors=True)
