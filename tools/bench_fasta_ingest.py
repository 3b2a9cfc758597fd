import sys, time; sys.path.insert(0, '.')
import torch, numpy as np
from raft_b200 import api, synth_gpu
ds = synth_gpu.make_dataset_gpu("C2", 0.05, with_seq=False)
p = api.AlgoParams.from_args(ds.args)
txt, nb = synth_gpu.gen_fasta_text(ds, 0, ds.n)
torch.cuda.synchronize()
print("text bytes", nb, "reads", ds.n)
ctx = api.Context(p)
for rep in range(3):
    t = time.perf_counter(); ctx.ingest_fasta(txt, nb, last=True, total_hint=nb); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"single call: {dt*1e3:.1f} ms  {nb/dt/1e9:.1f} GB/s  set_reads_ms={ctx.stats().ms_set_reads:.1f}")
# chunked
chunk = ds.n // 4
for rep in range(2):
    t = time.perf_counter()
    for r0 in range(0, ds.n, chunk):
        r1 = min(ds.n, r0 + chunk)
        a = int(ds.seq_off[r0]) + 39 * r0; b = int(ds.seq_off[r1]) + 39 * r1
        ctx.ingest_fasta(txt[a:b], b - a, last=(r1 == ds.n), total_hint=nb)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"chunked x4: {dt*1e3:.1f} ms")
