#!/bin/bash
# one call on an 8-GPU box: CLI on 3 GPUs (test), concurrent PCIe probe, every config at N = 8, multi-GPU file leg
N=${1:-8}
python -m pytest tests/test_gpu_sharded.py -x -q -k "cli_on_several" 2>&1 | tail -2
nvidia-smi topo -m 2>/dev/null | head -12; free -g | head -2; nproc
for d in $(seq 0 $((N-1))); do python tools/pcie_probe.py --device $d > gpurun_out/n8_pcie_$d.json 2>/dev/null & done; wait
cat gpurun_out/n8_pcie_*.json
bash tools/sweep_configs.sh $N r2n8 2>&1 | tail -9
python tools/file_leg.py --timing --devices $(seq -s, 0 $((N-1))) > gpurun_out/n8_fileleg.json 2> gpurun_out/n8_fileleg.err; grep timing gpurun_out/n8_fileleg.err | tail -6; python -c "
import json; d=json.loads(open('gpurun_out/n8_fileleg.json').read().strip().splitlines()[-1]); print(d['identical'], d['ours'], d['reference']['t_file_s'])"
