"""Per-source-line summary of an ncu report's source page (needs -lineinfo at compile time and --import-source on):
    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]
Prints the lines holding the most executed warp instructions, with their share of the stall samples."""
import csv
import io
import os
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    cur, hdr, lines = None, None, {}
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1]); continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None or not r[0].isdigit():
            continue
        d = dict(zip(hdr[4:], r[4:]))
        try:
            inst, samp = int(d["Instructions Executed"]), int(d["# Samples"])
            thr = int(d["Thread Instructions Executed"])
        except (KeyError, ValueError):
            continue
        key = (cur, int(r[0]))
        e = lines.setdefault(key, [0, 0, 0, r[1].strip()])
        e[0] += inst; e[1] += samp; e[2] += thr
    tot_i = sum(e[0] for e in lines.values()) or 1
    tot_s = sum(e[1] for e in lines.values()) or 1
    print(f"total warp instructions {tot_i}, samples {tot_s}, avg active threads {sum(e[2] for e in lines.values()) / tot_i:.1f}")
    print("inst%  samp%  thr/inst  file:line  source")
    for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * e[0] / tot_i:5.1f}  {100 * e[1] / tot_s:5.1f}  {e[2] / max(e[0], 1):5.1f}  {f}:{ln}  {e[3][:110]}")


if __name__ == "__main__":
    main()
