"""Aggregate an ncu source-page CSV (SASS view) by CUDA source line.
  ncu -i rep --page source --csv --kernel-name regex:K > sass.csv
  nvdisasm -g -c file.cubin > dis.txt
  python tools/ncu_by_line.py sass.csv dis.txt <mangled-kernel-substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis_txt, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# --- disassembly: offset -> (file, line) for the requested function
off2line = {}
cur = None
infn = False
for ln in open(dis_txt, errors='replace'):
    if ln.startswith('//---') and '.text.' in ln:
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m and cur is not None:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, ie, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
body = []
for r in rows[2:]:
    if not r or r[0] == 'Kernel Name':
        break
    body.append(r)
base = int(body[0][ia], 16)
agg = defaultdict(lambda: [0, 0, 0])
tot_e = tot_s = 0
for r in body:
    off = int(r[ia], 16) - base
    key = off2line.get(off, ('?', 0))
    e, s = int(r[ie] or 0), int(r[isamp] or 0)
    a = agg[key]
    a[0] += e; a[1] += s; a[2] += 1
    tot_e += e; tot_s += s
print(f'total warp instructions {tot_e}, samples {tot_s}')
srcs = {}
for (f, l), (e, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open('superpoint_transformer_b200/csrc/' + f).read().split('\n')
        except OSError:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ''
    print(f'{100 * e / tot_e:5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% stall  n={n:3d}  {f}:{l}  {text}')
