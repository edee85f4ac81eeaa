"""Hot source lines of one kernel from `ncu --page source --print-source cuda --csv` (instructions executed per line)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]
isrc = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
ifile = hdr.index('File') if 'File' in hdr else None
iline = hdr.index('Line') if 'Line' in hdr else (hdr.index('#') if '#' in hdr else None)
tot = 0; items = []
cur = ''
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[ie].isdigit(): 
        if len(r) == 2: cur = r[1]
        continue
    e = int(r[ie]); s = int(r[isamp] or 0); tot += e
    items.append((e, s, (r[iline] if iline is not None else ''), cur[-40:], r[isrc].strip()[:110]))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print('total', tot / div)
for e, s, ln, f, src in sorted(items, reverse=True)[:topn]: print(f'{e / div:9.1f} {s:6d} {f}:{ln} {src}')
