"""Output-file writer of the raft CLI, variant against variant on one set of sample files (no reference run):
    python tools/io_ab.py [--config C2] [--div 16] [--devices 0,1]
Each variant is a set of environment variables (RAFT_B200_NO_MMAP, RAFT_B200_IO_THREADS); outputs are compared with the first
variant's byte by byte.  Prints one JSON line per variant."""
import argparse
import filecmp
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

VARIANTS = [("pwrite", {"RAFT_B200_NO_MMAP": "1"}), ("mmap", {}), ("mmap4", {"RAFT_B200_IO_THREADS": "4"}),
            ("mmap12", {"RAFT_B200_IO_THREADS": "12"}), ("mmap", {})]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--div", type=float, default=16)
    ap.add_argument("--asymmetric", action="store_true")
    ap.add_argument("--devices", default=None)
    a = ap.parse_args()
    d, fs, fa, pf, (args, n_s, bases, n_reads, paf_bytes) = B.write_sample_files_gpu(a, 0, 1.0 / a.div)
    exe = os.path.join(ROOT, "raft_b200", "raft")
    names = ("coverage.txt", "long_repeats.txt", "long_repeats.bed", "reads.fasta")
    try:
        for k, (label, extra) in enumerate(VARIANTS):
            env = dict(os.environ, RAFT_B200_DEVICE="0", RAFT_B200_TIMING="1", **extra)
            if a.devices:
                env["RAFT_B200_DEVICES"] = a.devices
            own, wall, n, log = B.run_raft_binary(exe, fa, pf, args, os.path.join(d, f"v{k}"), env)
            same = all(filecmp.cmp(os.path.join(d, "v0." + s), os.path.join(d, f"v{k}." + s), shallow=False) for s in names) if k else None
            print(json.dumps({"variant": label, "fs": fs, "t_file_s": own, "wall_s": wall, "n": n, "same_as_first": same}), flush=True)
            if k:
                for s in names:
                    os.remove(os.path.join(d, f"v{k}." + s))
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
