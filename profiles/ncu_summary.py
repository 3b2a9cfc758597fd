#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump and an ncu launch list into the tables kept under profiles/."""
import csv
import sys
from collections import OrderedDict, defaultdict

WANT = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "GB"), ("dram__bytes_write.sum", "GB"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "%"), ("launch__registers_per_thread", ""),
        ("launch__occupancy_limit_registers", ""), ("launch__occupancy_limit_shared_mem", ""), ("launch__grid_size", "")]


def raw_table(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    out = []
    for r in data:
        d = OrderedDict(kernel=r[ik].split("(")[0].replace("void ", ""))
        for name, _ in WANT:
            if name in hdr:
                i = hdr.index(name)
                v = float(r[i].replace(",", ""))
                u = units[i]
                if name.startswith("dram__bytes"):
                    v *= {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1.0)
                if name.startswith("gpu__time"):
                    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(u, 1.0)
                d[name] = v
        out.append(d)
    return out


def launch_table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[iu], 1e-6)
        k = r[ik].split("(")[0].replace("void ", "")
        tot[k] += v
        cnt[k] += 1
    return tot, cnt


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "raw":
        t = raw_table(path)
        print("| kernel | ms | dram rd GB | dram wr GB | achieved GB/s | sm thr % | issue % | warps act % | regs | occ(reg/smem) blocks | grid |")
        print("|---|---|---|---|---|---|---|---|---|---|---|")
        for d in t:
            ms = d["gpu__time_duration.sum"]
            rd, wr = d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
            print(f"| {d['kernel']} | {ms:.3f} | {rd:.3f} | {wr:.3f} | {(rd + wr) / ms * 1e3:.0f} | "
                  f"{d.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):.0f} | {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.0f} | "
                  f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.0f} | {d.get('launch__registers_per_thread', 0):.0f} | "
                  f"{d.get('launch__occupancy_limit_registers', 0):.0f}/{d.get('launch__occupancy_limit_shared_mem', 0):.0f} | {d.get('launch__grid_size', 0):.0f} |")
    else:
        tot, cnt = launch_table(path)
        s = sum(tot.values())
        print("| kernel | launches | total ms | share |")
        print("|---|---|---|---|")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            print(f"| {k} | {cnt[k]} | {v:.3f} | {100 * v / s:.1f}% |")
        print(f"| **all** | {sum(cnt.values())} | {s:.3f} | 100% |")
