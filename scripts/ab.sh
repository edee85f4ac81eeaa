#!/bin/bash
# A/B of compile-time variants on the GPU box: each argument is a set of -D flags (use "-DNONE" for the default build)
cd /root/repo
for flags in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags -shared -o fastore_b200/libfastore_b200.so fastore_b200/csrc/fastore_b200.cu -lcudart 2>&1 | grep -i " error"
  for rep in 1 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$flags', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
  done
done
