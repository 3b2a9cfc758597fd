"""Host<->device copy rates of this box with pinned memory (context for bench.py's e2e line).

Prints one JSON line: H2D alone, D2H alone, and both directions at once (two streams), GB/s.
"""
import json
import sys

import torch


def rate(fn, nbytes, reps=3):
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def main():
    if "--device" in sys.argv:
        torch.cuda.set_device(int(sys.argv[sys.argv.index("--device") + 1]))
    n = 4 << 30
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    out["h2d_gbs"] = rate(lambda: d_in.copy_(h_in, non_blocking=True), n)
    out["d2h_gbs"] = rate(lambda: h_out.copy_(d_out, non_blocking=True), n)

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)

    out["bidir_each_gbs"] = rate(both, n)
    out["device"] = torch.cuda.current_device()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
