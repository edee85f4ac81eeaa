#!/bin/bash
# builds libfastore_b200.so with several K1 launch configurations on the GPU box and times each
cd /root/repo
for cfg in "8 3" "4 6" "4 5" "4 4" "8 2"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DFSB_K1_WARPS=$1 -DFSB_K1_MINBLOCKS=$2 -Xptxas -v -shared -o fastore_b200/libfastore_b200.so fastore_b200/csrc/fastore_b200.cu -lcudart 2>&1 | grep -A2 "ingest_kernelILi5ELi6" | grep -E "registers|spill" | tr '\n' ' '
  python bench.py --pairs 5000000 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg $1 $2', d['ms_per_step'], d['roofline']['stage_ms'])"
done
