"""Phase view of one kernel: walks the SASS in address order in windows of N instructions and prints, per
window, the dominant source lines, warp-instructions executed, stall samples and the top stall reasons.

    python scripts/phase_profile.py <lib.so> <ncu_sass.csv> <mangled-name-substring> [divisor] [window]
"""
import csv, re, subprocess, sys, tempfile, os, collections, glob

so, sass_csv, func = sys.argv[1:4]
div = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
win = int(sys.argv[5]) if len(sys.argv) > 5 else 120
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
dis = subprocess.run(['nvdisasm', '-g', '-c'] + glob.glob(tmp + '/*.cubin'), capture_output=True, text=True).stdout.splitlines()
insts = []; infunc = False; cur = ('?', 0)
for ln in dis:
    if ln.startswith('\t.section\t.text.'):
        infunc = func in ln; continue
    if not infunc: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: insts.append((cur, m.group(2)))
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]; ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[hi + 1:] if len(r) > ie and r[ie].isdigit()][:len(insts)]
tot_i = sum(int(r[ie]) for r in data); tot_s = sum(int(r[isamp] or 0) for r in data)
print(f'total warp-instr/unit {tot_i / div:.1f}   samples {tot_s}')
for a in range(0, len(data), win):
    chunk = data[a:a + win]
    ni = sum(int(r[ie]) for r in chunk); ns = sum(int(r[isamp] or 0) for r in chunk)
    st = collections.Counter()
    for r in chunk:
        for i, h in stall_cols:
            if r[i].isdigit(): st[h[6:]] += int(r[i])
    lines = collections.Counter()
    for k in range(a, min(a + win, len(insts))): lines[insts[k][0]] += int(data[k][ie])
    top = ' '.join(f'{f.split(".")[0]}:{l}' for (f, l), _ in lines.most_common(3))
    sts = ' '.join(f'{k}={100 * v / max(ns, 1):.0f}%' for k, v in st.most_common(4))
    print(f'{a:5d} instr {ni / div:7.1f} ({100 * ni / tot_i:4.1f}%)  samples {100 * ns / tot_s:5.1f}%  cyc/instr-ratio {(ns / tot_s) / max(ni / tot_i, 1e-9):4.2f}  [{sts}]  {top}')
