"""Instruction mix of one kernel from `ncu --page source --csv` output (per-warp counts by opcode)."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
warps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tot = 0; byop = collections.Counter(); samp = collections.Counter(); n = 0
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit(): continue
    src = r[ia].strip(); e = int(r[ie]); s = int(r[isamp] or 0)
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else src[:10]
    byop[op] += e; samp[op] += s; tot += e; n += 1
print('static instr', n, 'total warp-instr', tot, 'per warp', tot / warps)
for op, c in byop.most_common(30): print(f'{op:12s} {c / warps:8.1f} per warp   samples {samp[op]}')
