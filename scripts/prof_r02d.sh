set -x
cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:"ingest_kernel|place_kernel|lay_apply|lay_reduce|sort_scatter|sort_histogram" -s 12 -c 9 -o gpurun_out/r02d_prof python bench.py --total-pairs 2000000 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --parity-chunks 0 > gpurun_out/r02d_prof.log 2>&1
ncu -i gpurun_out/r02d_prof.ncu-rep --page raw --csv > gpurun_out/r02d_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --parity-chunks 0 > gpurun_out/r02d_launch.log 2>&1
tail -3 gpurun_out/r02d_prof.log
ls -la gpurun_out/r02d*
