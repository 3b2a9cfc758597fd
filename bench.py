#!/usr/bin/env python
"""bench.py — PAF overlaps/sec of the fragmentation hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S] [--config C2]

A step = one pass of the whole path over one synthetic input set:
    reads resident -> name table + layout -> PAF tokenise -> coverage scatter + scan -> repeats + cut
    points -> coverage.txt / long_repeats.txt / reads.fasta bytes materialised (window by window).
`value`  : inputs already in HBM, outputs materialised into HBM windows (device-resident).
`e2e`    : same call sequence through the C ABI with pinned HOST inputs and HOST outputs (H2D + D2H inside).
`--impl reference`: the unmodified reference binary (oracle/_ref/raft) on a bounded sample of the same
workload, on the host cores (the program is single-threaded).
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PAF overlaps/sec end-to-end fragmentation"
UNIT = "overlaps/s"
WINDOW = 1 << 30  # device output window (bytes)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
                for k, nm in enumerate(names):
                    if f[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm / cpu baseline
def write_sample_files(d, ds_host_reads, paf_bytes):
    import numpy as np
    r = ds_host_reads
    fa = os.path.join(d, "reads.fa")
    with open(fa, "wb") as fh:
        for i in range(r.n):
            fh.write(b">" + bytes(r.names[r.name_off[i]:r.name_off[i + 1]]) + b"\n")
            fh.write(r.seq[r.seq_off[i]:r.seq_off[i + 1]].tobytes() + b"\n")
    pf = os.path.join(d, "ovl.paf")
    with open(pf, "wb") as fh:
        fh.write(paf_bytes if isinstance(paf_bytes, (bytes, bytearray)) else np.asarray(paf_bytes).tobytes())
    return fa, pf


def make_sample(config, target_overlaps, full_scale_overlaps_per_unit):
    """Bounded sample of the workload for the CPU legs: same generator, smaller genome."""
    import torch
    scale = target_overlaps / full_scale_overlaps_per_unit
    if torch.cuda.is_available():
        from raft_b200 import synth_gpu
        ds = synth_gpu.make_dataset_gpu(config, scale)
        reads = ds.to_host_reads()
        paf = ds.paf.cpu().numpy().tobytes()
        args, n_ovl = ds.args, ds.n_overlaps
        del ds
        torch.cuda.empty_cache()
    else:
        from raft_b200 import synth
        ds = synth.make_dataset(config, scale)
        reads, paf, args, n_ovl = ds.reads, ds.paf, ds.args, ds.n_overlaps
    return reads, paf, args, n_ovl, scale


def run_reference_once(ref_bin, fa, pf, args, d):
    t0 = time.perf_counter()
    p = subprocess.run([ref_bin] + list(args) + ["-o", os.path.join(d, "ref"), fa, pf], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("reference failed: " + p.stdout[-2000:])
    own, n = wall, None
    for line in p.stdout.splitlines():
        if "program completed after" in line:
            own = float(line.split("after")[1].split()[0])
        if "length of alignments" in line:
            n = int(line.split("alignments")[1].strip().rstrip("()"))
    return own, wall, n


OVL_PER_UNIT = {"C2": 2.3e8, "C1": 4.5e5, "C4": 3e7, "C5": 5e7}  # overlaps at scale 1.0, for sample sizing


def reference_arm(a):
    from oracle import oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not O.have_ref():
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle")])
    kind = "reference" if O.have_ref() else "port"
    # ~1.2 M overlaps per step: about 5 s of the single-threaded reference
    reads, paf, args, n_ovl, scale = make_sample(a.config, 1.2e6, OVL_PER_UNIT[a.config])
    d = tempfile.mkdtemp(prefix="raft_ref_")
    try:
        fa, pf = write_sample_files(d, reads, paf)
        times = []
        for it in range(a.warmup + a.steps):
            if kind == "reference":
                own, wall, n = run_reference_once(O.REF_BIN, fa, pf, args, d)
            else:
                t0 = time.perf_counter()
                res = O.run(reads, paf, O.make_params(**_kw(args)))
                own = wall = time.perf_counter() - t0
                n = res.n_rec
            if it >= a.warmup:
                times.append(wall)
        sec = sum(times) / len(times)
        v = n_ovl / sec
        out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32/u8",
               "data": "synthetic",
               "config": {"workload": f"{a.config} synthetic human 32x ONT-Duplex-shaped reads + symmetric all-vs-all PAF, genome scale {a.scale:g} "
                                      f"(raft {' '.join(args)}, defaults -r 50 -l 20000): the workload of the default arm, timed on a bounded sample",
                          "sample": f"the same generator at genome scale {scale:.4g}, file to file",
                          "n_overlaps": n_ovl, "n_reads": reads.n, "bases": int(reads.seq_off[-1]), "flags": " ".join(args)},
               "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "host_cores": os.cpu_count(),
                                "sample": f"{n_ovl} overlaps / {int(reads.seq_off[-1])} bases, file to file, wall clock around the process"},
               "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(out), flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _kw(args):
    m = {"-e": "est_cov", "-r": "reso", "-p": "repeat_length", "-f": "flanking_length", "-v": "overlap_length", "-l": "read_length"}
    return {m[args[k]]: int(args[k + 1]) for k in range(0, len(args), 2)}


def host_can_hold(nbytes, reserve=40 << 30):
    try:
        import psutil
        return psutil.virtual_memory().available > nbytes + reserve
    except Exception:
        return nbytes < (64 << 30)


# ----------------------------------------------------------------------------------------------- our arm
def emit_all_device(ctx, api, win, win2):
    """materialise every output stream into device window buffers, window by window.  The text streams go to `win2`
    and the sequence stream to `win` through the asynchronous API: they run on two CUDA streams of the library, so the
    issue-bound text formatter overlaps the bandwidth-bound gather."""
    total = 0
    sizes = {w: ctx.output_size(w) for w in (api.OUT_COVERAGE, api.OUT_LONG_REPEATS, api.OUT_READS_FASTA)}
    for which in (api.OUT_COVERAGE, api.OUT_LONG_REPEATS):
        for off in range(0, sizes[which], WINDOW):
            ctx.fetch_async(which, off, win2, min(WINDOW, sizes[which] - off))
    n = sizes[api.OUT_READS_FASTA]
    for off in range(0, n, WINDOW):
        ctx.fetch_async(api.OUT_READS_FASTA, off, win, min(WINDOW, n - off))
    ctx.sync()
    return sum(sizes.values())


def ours(a):
    import numpy as np
    import torch
    from raft_b200 import api, synth_gpu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fragmentation path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout; stdout carries the JSON line only
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from raft_b200 import sharded
        return sharded.bench(a, rank, world, local, log)

    t_gen = time.perf_counter()
    ds = synth_gpu.make_dataset_gpu(a.config, a.scale, device=f"cuda:{local}")
    log(f"[bench] generated {ds.meta} in {time.perf_counter() - t_gen:.1f}s")
    p = api.AlgoParams.from_args(ds.args)
    ctx = api.Context(p, local)
    win = torch.empty(WINDOW + 64, dtype=torch.uint8, device=dev)
    win2 = torch.empty(WINDOW + 64, dtype=torch.uint8, device=dev)

    def step_device():
        ctx.set_reads(ds.seq_off, ds.seq, ds.name_off, ds.names)
        ctx.ingest_paf(ds.paf, ds.paf.numel(), last=True)
        st = ctx.run()
        emit_all_device(ctx, api, win, win2)
        return st

    for _ in range(a.warmup):
        st = step_device()
    clocks = ClockSampler(local)
    clocks.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {k: 0.0 for k in ("set_reads", "tokenize", "scatter", "scan", "repeat_cut", "layout", "emit_cov", "emit_rep", "emit_fasta")}
    fasta_ms, fasta_launches, launches = 0.0, 0, 0
    e0.record()
    for _ in range(a.steps):
        st = step_device()
        s2 = ctx.stats()
        launches += s2.kernel_launches
        stage["set_reads"] += s2.ms_set_reads; stage["tokenize"] += s2.ms_tokenize; stage["scatter"] += s2.ms_scatter
        stage["scan"] += s2.ms_scan; stage["repeat_cut"] += s2.ms_repeat_cut; stage["layout"] += s2.ms_layout
        stage["emit_cov"] += s2.ms_emit[0]; stage["emit_rep"] += s2.ms_emit[1]; stage["emit_fasta"] += s2.ms_emit[3]
        fasta_ms += s2.ms_emit[3]; fasta_launches += s2.emit_launches[3]
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / a.steps
    clk = clocks.stop()
    stage = {k: v / a.steps for k, v in stage.items()}
    out_bytes = [int(x) for x in st.out_bytes]
    n_ovl = int(st.n_records)

    # dominant kernel: k_fasta_emit.  Algorithmic bytes = sequence bytes gathered + FASTA bytes written.
    frag = ctx.table(api.TAB_FRAG).reshape(-1, 3)
    gathered = int((frag[:, 2].astype(np.int64) - frag[:, 1]).sum())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_fasta = gathered + out_bytes[3]
    launches_per_step = fasta_launches / a.steps
    ach = alg_fasta / (fasta_ms / a.steps / 1e3) / 1e9 if fasta_ms else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_fasta_emit_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": "k_fasta_emit", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s",
                "alg_bytes_per_launch": alg_fasta / max(launches_per_step, 1), "launches_per_step": launches_per_step,
                "kernel_ms_per_step": fasta_ms / a.steps}
    paf_bytes = int(ds.paf.numel())
    bytes_alg = paf_bytes + gathered + int(ds.names.numel()) + sum(out_bytes)
    path_roof = {"bytes_alg": bytes_alg, "achieved_gbs": bytes_alg / (ms_step / 1e3) / 1e9, "frac": bytes_alg / (ms_step / 1e3) / 1e9 / peak}

    # ---- row f2: device FASTA ingest of the same reads (unwrapped FASTA text generated chunk by chunk in HBM; only the
    # raftgpu_ingest_fasta calls are timed, CUDA events around each)
    fasta_ingest = None
    if not a.no_fasta:
        try:
            ctx.close()
            ctxf = api.Context(p, local)
            chunk_reads = max(1, int(ds.n * (2 << 30) / max(ds.bases, 1)))  # ~2 GiB of text per chunk
            total_text = ds.bases + 39 * ds.n
            tbuf, ms_f = None, 0.0
            # the arena is sized by the hint; drop the resident copy of the bases first when both cannot fit
            free_b, _ = torch.cuda.mem_get_info()
            seq_keep = ds.seq
            if free_b < total_text + (8 << 30):
                ds.seq = None
                seq_keep = None
                torch.cuda.empty_cache()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for timed in (False, True):  # first pass untimed: it pays the one-time cudaMalloc of the ~100 GB arena
                if timed:
                    ctxf.reset()
                ms_f = 0.0
                for r0 in range(0, ds.n, chunk_reads):
                    r1 = min(ds.n, r0 + chunk_reads)
                    tbuf, nb = synth_gpu.gen_fasta_text(ds, r0, r1, tbuf)
                    torch.cuda.synchronize()
                    f0.record()
                    ctxf.ingest_fasta(tbuf, nb, last=(r1 == ds.n), total_hint=total_text)
                    f1.record()
                    torch.cuda.synchronize()
                    ms_f += f0.elapsed_time(f1)
            ok = ctxf.stats().n_reads == 0  # stats are filled by run(); check the read count through a table instead
            nb_off = ctxf.table(api.TAB_BIN_OFF)
            fasta_ingest = {"text_bytes": total_text, "ms": ms_f, "gbs_text": total_text / (ms_f / 1e3) / 1e9,
                            "alg_gbs": (total_text + ds.bases) / (ms_f / 1e3) / 1e9, "reads": int(len(nb_off) - 1),
                            "note": "raftgpu_ingest_fasta on device-resident text, ~2 GiB chunks, second pass over a reset context (arena already allocated); includes layout scans + name table of the last call"}
            ctxf.close()
            del tbuf
            if seq_keep is None:
                ds.seq = synth_gpu.gen_seq(ds, 0, ds.n)
            torch.cuda.empty_cache()
            ctx = api.Context(p, local)
        except Exception as e:  # never lose the headline because of the side measurement
            log(f"[bench] fasta ingest measurement failed: {e}")
            ctx = api.Context(p, local)

    # ---- e2e: pinned host inputs -> C ABI -> host outputs
    e2e = None
    if not a.no_e2e and not host_can_hold(sum(int(t.numel() * t.element_size()) for t in (ds.seq_off, ds.name_off, ds.seq, ds.names, ds.paf))):
        log("[bench] host memory too small to pin the inputs of this workload: e2e skipped")
        a.no_e2e = True
    if not a.no_e2e:
        def pinned(t):  # straight into pinned memory (no pageable intermediate copy)
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h
        hseq_off, hname_off = pinned(ds.seq_off), pinned(ds.name_off)
        hseq, hnames, hpaf = pinned(ds.seq), pinned(ds.names), pinned(ds.paf)
        hout = torch.empty(WINDOW, dtype=torch.uint8, pin_memory=True)
        h2d = sum(int(t.numel() * t.element_size()) for t in (hseq_off, hname_off, hseq, hnames, hpaf))
        # the library now needs its own copy of the inputs in HBM: drop the device-resident set and its context
        ctx.close()
        ds.seq = ds.paf = ds.names = ds.seq_off = ds.name_off = None
        del win, win2
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        log(f"[bench] e2e: torch holds {torch.cuda.memory_reserved() / 1e9:.1f} GB of HBM after releasing the resident inputs")
        ctx = api.Context(p, local)

        # the read arena is uploaded in chunks behind the PAF, overlapping the kernels and the D2H of the outputs
        ctx.set_option(api.OPT_DEFER_SEQ_UPLOAD, 1)

        def step_e2e():
            ctx.set_reads(hseq_off.numpy(), hseq.numpy(), hname_off.numpy(), hnames.numpy())
            ctx.ingest_paf(hpaf.numpy(), hpaf.numel(), last=True)
            ctx.run()
            d2h = 0
            for which in (api.OUT_COVERAGE, api.OUT_LONG_REPEATS, api.OUT_READS_FASTA):
                n = ctx.output_size(which)
                for off in range(0, n, WINDOW):
                    ctx.fetch_into(which, off, hout.data_ptr(), min(WINDOW, n - off))
                d2h += n
            return d2h

        step_e2e()
        torch.cuda.synchronize()
        k = max(1, min(a.steps, 3))
        e0.record()
        for _ in range(k):
            d2h = step_e2e()
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1) / k
        e2e = {"value": n_ovl / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
               "steps": k}
        del hseq, hpaf, hout

    # ---- cpu baseline (rank 0, bounded sample) + parity of the same sample through the CUDA path
    cpu = None
    if not a.no_cpu:
        from oracle import oracle as O
        ctx.close()  # release HBM before the sample is generated and checked
        torch.cuda.empty_cache()
        reads, paf, args, n_s, scale = make_sample(a.config, 4.0e6, OVL_PER_UNIT[a.config])
        d = tempfile.mkdtemp(prefix="raft_cpu_")
        try:
            fa, pf = write_sample_files(d, reads, paf)
            if O.have_ref():
                own, wall, n = run_reference_once(O.REF_BIN, fa, pf, args, d)
                kind = "reference"
                outs = {s: open(os.path.join(d, "ref." + s), "rb").read() for s in ("coverage.txt", "long_repeats.txt", "reads.fasta")}
            else:
                t0 = time.perf_counter()
                res = O.run(reads, paf, O.make_params(**_kw(args)))
                own = wall = time.perf_counter() - t0
                kind = "port"
                outs = {"coverage.txt": res.cov_txt, "long_repeats.txt": res.rep_txt, "reads.fasta": res.fasta}
            c2 = api.Context(api.AlgoParams.from_args(args), local)
            c2.set_reads(reads.seq_off, reads.seq, reads.name_off, reads.names)
            c2.ingest_paf(np.frombuffer(paf, np.uint8), len(paf), last=True)
            c2.run()
            same = all(c2.digest(w) == O.digest(outs[s]) and c2.output_size(w) == len(outs[s]) for w, s in
                       ((api.OUT_COVERAGE, "coverage.txt"), (api.OUT_LONG_REPEATS, "long_repeats.txt"), (api.OUT_READS_FASTA, "reads.fasta")))
            c2.close()
            cpu = {"value": n_s / wall, "unit": UNIT, "cores": 1, "kind": kind, "host_cores": os.cpu_count(), "seconds": wall,
                   "reference_own_timer_s": own, "gbp_per_s": int(reads.seq_off[-1]) / wall / 1e9,
                   "sample": f"{a.config}-shaped at genome scale {scale:.4g}: {n_s} overlaps, {int(reads.seq_off[-1])} bases, file to file",
                   "gpu_outputs_identical_on_sample": bool(same)}
        finally:
            shutil.rmtree(d, ignore_errors=True)

    out = {"metric": METRIC, "value": n_ovl / (ms_step / 1e3), "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32/u8", "data": "synthetic",
           "config": {"workload": f"{a.config} synthetic human 32x ONT-Duplex-shaped reads + symmetric all-vs-all PAF, genome scale {a.scale:g} "
                                  f"({ds.meta['genome'] / 1e9:.2f} Gbp genome; raft {' '.join(ds.args)}, defaults -r 50 -l 20000), 1 B200", "n_overlaps": n_ovl, "n_reads": ds.n, "bases": ds.bases,
                      "paf_bytes": paf_bytes, "out_bytes": {"coverage.txt": out_bytes[0], "long_repeats.txt": out_bytes[1],
                                                                      "reads.fasta": out_bytes[3]},
                      "l2": "inputs and outputs are GBs (>> 126 MB L2); no explicit flush", "sharding": "none (1 GPU)"},
           "gbp_per_s": ds.bases / (ms_step / 1e3) / 1e9, "stage_ms": stage, "roofline": roofline, "path_roofline": path_roof,
           "cpu_baseline": cpu, "e2e": e2e, "fasta_ingest": fasta_ingest, "gpu_launches": launches, "clocks": clk}
    print(json.dumps(out), flush=True)
    ctx.close()


def main():
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL's version banner, child processes) are
    # sent to stderr, and sys.stdout is re-pointed at the saved descriptor for our own print of the result
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale of the config (1.0 = the full human-scale config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fasta", action="store_true", help="skip the device FASTA ingest side measurement")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)


if __name__ == "__main__":
    main()
