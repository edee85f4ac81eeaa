#!/bin/bash
# Final profiles of a round (GPU box): launch list of the bench command, full-set metrics of the kernels, sanitizer logs.
# Everything is sized to finish in a few minutes and to leave less than 64 MiB behind.
# usage: bash scripts/prof_final.sh <tag> [nosan]     -> gpurun_out/<tag>_*
tag=${1:-r02z}
cd /root/repo
# 1) launch list of the default bench command (one stream so that the launches are attributable)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --parity-chunks 0 --run-split 1 > gpurun_out/${tag}_launches.log 2>&1
# 2) the two big kernels with source counters (2 M pairs = 3 chunks)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ingest_kernel|place_kernel" -s 4 -c 2 -o gpurun_out/${tag}_k1k4 python bench.py --total-pairs 2000000 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --parity-chunks 0 --run-split 1 > gpurun_out/${tag}_k1k4.log 2>&1
ncu -i gpurun_out/${tag}_k1k4.ncu-rep --page raw --csv > gpurun_out/${tag}_k1k4_raw.csv 2>/dev/null
# 3) the small kernels (sort, layout, checks, device-side parse), metrics only
timeout 400 ncu --set full --clock-control none -k regex:"lay_|sort_|validate_text|stage_stats|parse_|chunk_summary" -s 20 -c 16 -o gpurun_out/${tag}_small python bench.py --total-pairs 2000000 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --parity-chunks 0 --run-split 1 > gpurun_out/${tag}_small.log 2>&1
ncu -i gpurun_out/${tag}_small.ncu-rep --page raw --csv > gpurun_out/${tag}_small_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:"parse_" -c 4 -o gpurun_out/${tag}_parse python bench.py --total-pairs 2000000 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --parity-chunks 0 --run-split 1 > gpurun_out/${tag}_parse.log 2>&1
ncu -i gpurun_out/${tag}_parse.ncu-rep --page raw --csv > gpurun_out/${tag}_parse_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_small.ncu-rep gpurun_out/${tag}_parse.ncu-rep
if [ "$2" != "nosan" ]; then
  for tool in memcheck racecheck; do
    FSB_SAN_SMALL=1 timeout 180 compute-sanitizer --tool $tool python scripts/sanitizer_case.py > gpurun_out/${tag}_sanitizer_$tool.log 2>&1
    echo "rc=$?" >> gpurun_out/${tag}_sanitizer_$tool.log
    tail -3 gpurun_out/${tag}_sanitizer_$tool.log
  done
fi
ls -la gpurun_out/${tag}_*
