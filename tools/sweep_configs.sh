#!/bin/bash
# Every BASELINE.json config through bench.py at its full size.  N = number of GPUs (default 1).
#   bash tools/sweep_configs.sh [N] [tag]      -> gpurun_out/<tag>_<config>_n<N>.json
N=${1:-1}; TAG=${2:-cfg}
mkdir -p gpurun_out
run() { # name, bench flags...
  name=$1; shift
  if [ "$N" = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N"; fi
  $cmd --steps 5 --warmup 3 "$@" > gpurun_out/${TAG}_${name}_n${N}.json 2> gpurun_out/${TAG}_${name}_n${N}.err
  echo "$name rc=$? $(tail -c 300 gpurun_out/${TAG}_${name}_n${N}.err | tr '\n' ' ' | cut -c1-200)"
}
run C1 --config C1
run C4 --config C4
run C5 --config C5
run C2asym --config C2 --asymmetric
[ "$N" = 1 ] && run C5asym --config C5 --asymmetric
run C2 --config C2
python - "$TAG" "$N" <<'PY'
import json, sys, glob
tag, n = sys.argv[1], sys.argv[2]
for f in sorted(glob.glob(f"gpurun_out/{tag}_*_n{n}.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "no result", e); continue
    st = d.get("stage_ms") or d.get("stage_ms_rank0")
    print(f.split("/")[-1], "ms", round(d["ms_per_step"], 3), "ovl/s %.3g" % d["value"], "path_frac", round(d["path_roofline"]["frac"], 3),
          "parity", d["parity"]["identical"], [c["vs"] for c in d["parity"]["checks"]], "e2e_ms", d["e2e"] and round(d["e2e"]["ms_per_step"], 1),
          "f2f", d.get("file_to_file") and (round(d["file_to_file"]["ours"]["t_file_s"], 2), round(d["file_to_file"]["reference"]["t_file_s"], 2)),
          {k: round(v, 2) for k, v in st.items()})
PY
