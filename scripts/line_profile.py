"""Per-source-line instruction counts for one kernel.

Joins `nvdisasm -g -c` of the cubin (SASS with //## File/line annotations) with the per-instruction
"Instructions Executed" column of `ncu --page source --csv` (SASS view), by instruction order.

    python scripts/line_profile.py <lib.so> <ncu_sass.csv> <mangled-name-substring> [divisor] [top]
"""
import csv, re, subprocess, sys, tempfile, os, collections, glob

so, sass_csv, func = sys.argv[1:4]
div = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
dis = subprocess.run(['nvdisasm', '-g', '-c'] + glob.glob(tmp + '/*.cubin'), capture_output=True, text=True).stdout.splitlines()
# instructions of the function, each with the (file, line) in force -- innermost inlined location and outermost kernel line
insts = []
infunc = False; cur = ('?', 0); stack = []
for ln in dis:
    if ln.startswith('\t.section\t.text.'):
        infunc = func in ln
        continue
    if not infunc: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: insts.append((cur, m.group(2)))
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]; isrc = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
counts = [(int(r[ie]), int(r[isamp] or 0), r[isrc]) for r in rows[hi + 1:] if len(r) > ie and r[ie].isdigit()]
if len(counts) % len(insts) == 0 and len(counts) != len(insts): counts = counts[:len(insts)]
print(f'# {len(insts)} SASS instructions in cubin, {len(counts)} in profile', file=sys.stderr)
n = min(len(insts), len(counts))
byline = collections.Counter(); bysamp = collections.Counter(); tot = 0
for k in range(n):
    byline[insts[k][0]] += counts[k][0]; bysamp[insts[k][0]] += counts[k][1]; tot += counts[k][0]
print(f'total {tot / div:.1f}')
src_cache = {}
def src(f, l):
    if f not in src_cache:
        p = glob.glob(f'/root/repo/fastore_b200/csrc/**/{f}', recursive=True)
        src_cache[f] = open(p[0]).read().splitlines() if p else []
    s = src_cache[f]
    return s[l - 1].strip()[:100] if 0 < l <= len(s) else ''
for (f, l), c in byline.most_common(top):
    print(f'{c / div:9.1f} {bysamp[(f, l)]:6d}  {f}:{l:<4d} {src(f, l)}')
