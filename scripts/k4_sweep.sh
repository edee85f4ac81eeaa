#!/bin/bash
cd /root/repo
for t in 64 32 16; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DFSB_K4_TILE=$t -shared -o fastore_b200/libfastore_b200.so fastore_b200/csrc/fastore_b200.cu -lcudart 2>&1 | grep -i error
  python bench.py --pairs 5000000 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('tile $t', d['ms_per_step'], d['roofline']['stage_ms'])"
done
