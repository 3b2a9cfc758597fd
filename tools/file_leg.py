"""File -> file leg of bench.py on its own: raft_b200/raft (1 or several GPUs) and oracle/_ref/raft on the same sample files.
    python tools/file_leg.py [--config C2] [--div 16] [--devices 0,1] [--timing]
Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--div", type=float, default=None, help="sample = config / div (default: bench.py's)")
    ap.add_argument("--asymmetric", action="store_true")
    ap.add_argument("--devices", default=None, help="RAFT_B200_DEVICES for our CLI (e.g. 0,1,2,3)")
    ap.add_argument("--timing", action="store_true", help="RAFT_B200_TIMING=1: the CLI prints its phase times on stderr")
    a = ap.parse_args()
    a.scale = 1.0
    if a.div:
        B.SAMPLE_DIV[a.config] = a.div
    if a.devices:
        os.environ["RAFT_B200_DEVICES"] = a.devices
    if a.timing:
        os.environ["RAFT_B200_TIMING"] = "1"
    print(json.dumps(B.file_to_file_leg(a, 0, B.log)))


if __name__ == "__main__":
    main()
