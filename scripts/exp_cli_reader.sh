#!/bin/bash
# A/B of the chunk reader (sliced positioned reads against one fread per request), same box, runs interleaved.
cd /root/repo
python - <<'PY'
import sys, subprocess, time, os, tempfile, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/scripts')
import bench
from fastore_b200 import synth
import binfile_helpers as BF
tmp = tempfile.mkdtemp(dir='/dev/shm')
w = bench.WORKLOADS['c2']
n = 10_000_000
cfg = synth.synth_config(n, w['L'], paired=True, seed=102, **w['synth'])
f1, f2 = os.path.join(tmp, 'a_1.fastq'), os.path.join(tmp, 'a_2.fastq')
t1, t2, _, _ = synth.generate(cfg, threads=16)
t1.tofile(f1); t2.tofile(f2); del t1, t2
out = {}
for rep in range(3):
    for mode in ('sliced', 'fread'):
        env = dict(os.environ); env['FSH_READ_SLICES'] = '4' if mode == 'sliced' else '1'
        t0 = time.time()
        r = subprocess.run([str(BF.CLI), 'e', f'-i{f1} {f2}', f'-o{tmp}/o_{mode}', '-z', '-H', '-q0', '-p8', '-s0', '-b256', '-v', '-P8'], env=env, capture_output=True, text=True)
        dt = time.time() - t0
        line = [l for l in r.stderr.replace('\r', '\n').splitlines() if 'reader finished' in l or 'records in' in l]
        out.setdefault(mode, []).append({'wall_s': round(dt, 3), 'trace': line, 'rc': r.returncode})
print(json.dumps(out, indent=1))
PY
