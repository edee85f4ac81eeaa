#!/bin/bash
# Final profiles of a round (GPU box): launch list of the bench command, full-set metrics of the kernels, sanitizer logs.
# usage: bash scripts/prof_final.sh <tag> [nosan]     -> gpurun_out/<tag>_*
tag=${1:-r02z}
cd /root/repo
# 1) launch list of the default bench command (one stream so that the launches are attributable)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --parity-chunks 0 --run-split 1 > gpurun_out/${tag}_launches.log 2>&1
# 2) the two big kernels with source counters, on the full configs[1] batch
ncu --set full --clock-control none --import-source on -k regex:"ingest_kernel|place_kernel" -s 6 -c 2 -o gpurun_out/${tag}_k1k4 python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --parity-chunks 0 --run-split 1 > gpurun_out/${tag}_k1k4.log 2>&1
ncu -i gpurun_out/${tag}_k1k4.ncu-rep --page raw --csv > gpurun_out/${tag}_k1k4_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_k1k4.ncu-rep --page source --csv --kernel-name regex:ingest_kernel > gpurun_out/${tag}_k1_source.csv 2>/dev/null
ncu -i gpurun_out/${tag}_k1k4.ncu-rep --page source --csv --kernel-name regex:place_kernel > gpurun_out/${tag}_k4_source.csv 2>/dev/null
# 3) the small kernels (sort, layout, checks, device-side parse), metrics only
ncu --set full --clock-control none -k regex:"lay_|sort_|validate_text|stage_stats|parse_|chunk_summary|scan" -s 30 -c 60 -o gpurun_out/${tag}_small python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --parity-chunks 0 --run-split 1 > gpurun_out/${tag}_small.log 2>&1
ncu -i gpurun_out/${tag}_small.ncu-rep --page raw --csv > gpurun_out/${tag}_small_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_small.ncu-rep
if [ "$2" != "nosan" ]; then
  for tool in memcheck racecheck initcheck synccheck; do
    timeout 600 compute-sanitizer --tool $tool python scripts/sanitizer_case.py > gpurun_out/${tag}_sanitizer_$tool.log 2>&1
    tail -2 gpurun_out/${tag}_sanitizer_$tool.log
  done
fi
ls -la gpurun_out/${tag}_*
