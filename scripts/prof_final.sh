#!/bin/bash
# Final profiles of a round (GPU box): launch list of the bench command, full-set metrics of the big kernels, sanitizer logs.
# usage: bash scripts/prof_final.sh <tag>     -> gpurun_out/<tag>_*
tag=${1:-r02z}
cd /root/repo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --parity-chunks 0 --run-split 1 > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ingest_kernel|place_kernel|lay_apply|lay_reduce|lay_scan|sort_scatter|sort_histogram|validate_text|parse_" -c 120 -o gpurun_out/${tag}_prof python bench.py --total-pairs 2000000 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --parity-chunks 0 --run-split 1 > gpurun_out/${tag}_prof.log 2>&1
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
for tool in memcheck racecheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitizer_case.py > gpurun_out/${tag}_sanitizer_$tool.log 2>&1
  tail -2 gpurun_out/${tag}_sanitizer_$tool.log
done
ls -la gpurun_out/${tag}_*
