for g in default 32 64 128; do
  if [ $g = default ]; then unset FSB_L2_FETCH; else export FSB_L2_FETCH=$g; fi
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pack_|signature" -c 4 --csv --log-file gpurun_out/l2exp_$g.csv python bench.py --steps 1 --warmup 0 --pairs 1000000 --no-cpu --e2e-steps 1 > /dev/null 2>&1
done
