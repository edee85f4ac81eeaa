"""Static SASS instruction count per source file/function region of one kernel (nvdisasm -g)."""
import re, subprocess, sys, tempfile, os, glob, collections
so, func = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
dis = subprocess.run(['nvdisasm', '-g', '-c'] + glob.glob(tmp + '/*.cubin'), capture_output=True, text=True).stdout.splitlines()
infunc = False; cur = ('?', 0); cnt = collections.Counter(); tot = 0
for ln in dis:
    if ln.startswith('\t.section\t.text.'):
        infunc = func in ln; continue
    if not infunc: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2)) // 10 * 10); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+', ln): cnt[cur] += 1; tot += 1
print('total', tot)
for k, v in sorted(cnt.items()): 
    if v >= 20: print(f'{k[0]}:{k[1]:<5d} {v}')
