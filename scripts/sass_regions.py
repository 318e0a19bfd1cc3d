#!/usr/bin/env python
"""Where does a warp-specialised kernel spill?  Splits the SASS of one kernel at the role markers
(USETMAXREG.DEALLOC = MMA warps, USETMAXREG.TRY_ALLOC = producers, first LDTM = epilogue) and counts
local-memory traffic and a few instruction classes per region.
   python scripts/sass_regions.py coordx_b200/_build/trunk_fwd_tc.o <mangled-name-substring>"""
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
names = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
funcs = re.findall(r'Function : (\S+)', names)
fn = [f for f in funcs if pat in f][0]
out = subprocess.run(['cuobjdump', '-sass', '-fun', fn, obj], capture_output=True, text=True).stdout
ins = []
for line in out.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
marks = []
seen = set()
for a, t in ins:
    for key, name in (('USETMAXREG.DEALLOC', 'mma'), ('USETMAXREG.TRY_ALLOC', 'producer'), ('LDTM', 'epilogue')):
        if key in t and name not in seen:
            seen.add(name)
            marks.append((a, name))
marks = [(0, 'prologue')] + sorted(marks) + [(1 << 30, 'end')]
print(fn)
for (a0, name), (a1, _) in zip(marks, marks[1:]):
    reg = [t for a, t in ins if a0 <= a < a1]
    cnt = lambda k: sum(1 for t in reg if re.search(k, t))
    maxr = max([int(x) for t in reg for x in re.findall(r'\bR(\d+)\b', t)] or [0])
    print(f'{name:9s} {a0:6x}: {len(reg):5d} instr  maxR {maxr:3d}  STL {cnt("STL"):3d} LDL {cnt("LDL"):3d}  MUFU {cnt("MUFU"):3d} '
          f'FFMA {cnt(r"FFMA"):3d} FMUL {cnt("FMUL"):3d} HFMA2/HMUL2/HADD2 {cnt(r"HFMA2|HMUL2|HADD2"):3d} FHFMA {cnt("FHFMA"):3d} '
          f'LOP3/PRMT/SHF {cnt(r"LOP3|PRMT|SHF"):3d} LDS {cnt(r"LDS"):3d} STS {cnt("STS"):3d} LDC {cnt(r"LDC"):3d} MOV {cnt(r"MOV"):3d}')
