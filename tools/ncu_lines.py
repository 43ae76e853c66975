"""Attribute the executed instructions / stall samples of an ncu source-page CSV to CUDA source lines.

  ncu -i rep.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all build/rr_integrate.cu.o && nvdisasm -g -c *.cubin > dis.txt
  python tools/ncu_lines.py src.csv dis.txt <mangled-kernel-substring> [top]
"""
import collections
import csv
import os
import re
import sys

src_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l][0]
addr2line = {}
cur = None
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//--------------------- .text'):
        if len(addr2line) > 10:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*);', l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
for i, r in enumerate(data):            # several captured launches: keep the first one's table
    if r and r[0] == "Kernel Name":
        data = data[:i]
        break
ia, ins, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(data[0][ia], 16)
by, bys = collections.Counter(), collections.Counter()
for r in data:
    k = addr2line.get(int(r[ia], 16) - base)
    by[k] += int(r[iex] or 0)
    bys[k] += int(r[ins] or 0)
tot, tots = sum(by.values()), sum(bys.values())
print("instructions", tot, "samples", tots)
cache = {}
for key, c in sorted(by.items(), key=lambda kv: -kv[1])[:top]:
    text = ''
    if key and os.path.exists(key[0]):
        if key[0] not in cache:
            cache[key[0]] = open(key[0]).read().split('\n')
        text = cache[key[0]][key[1] - 1].strip()[:100]
    name = f"{os.path.basename(key[0])}:{key[1]}" if key else "?"
    print(f"{name:28s} exec {100 * c / tot:5.1f}% samp {100 * bys[key] / max(1, tots):5.1f}%  {text}")
