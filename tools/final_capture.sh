#!/bin/bash
# Round-end evidence on one GPU: the whole -m gpu suite, sanitizer passes over smoke(), the default bench line, the reference
# arm, the ncu launch list of the default workload and the --set full pages of the three main kernels.
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/fin_tests.txt; cat gpurun_out/fin_tests.txt
for tool in memcheck racecheck; do compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_$tool.txt 2>&1; echo "$tool rc=$?" >> gpurun_out/fin_$tool.txt; tail -3 gpurun_out/fin_$tool.txt; done
python bench.py --steps 20 --warmup 5 > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 --scale 0.1 > gpurun_out/fin_ref.json 2> gpurun_out/fin_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/fin_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1200 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-fasta --no-oracle > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_paf_tokenize|k_scan_cov|k_repeat_cut|k_cov_text|k_fasta_emit" -s 15 -c 7 -o gpurun_out/fin_prof python bench.py --steps 1 --warmup 3 --scale 0.05 --no-e2e --no-cpu --no-oracle --no-fasta > gpurun_out/fin_prof.log 2>&1
ncu -i gpurun_out/fin_prof.ncu-rep --page raw --csv > gpurun_out/fin_prof_raw.csv 2>/dev/null
ls -la gpurun_out/fin_*
