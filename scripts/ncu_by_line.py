#!/usr/bin/env python
"""Aggregate an ncu SASS-page CSV by source line (line info from nvdisasm of the same cubin).
usage: ncu_by_line.py sass.csv dis.txt kernel_mangled_name [top]"""
import csv, re, sys
from collections import defaultdict
sass_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
HELPER_END = 200
# address offset -> line
off2line, cur, inside = {}, None, False
for ln in open(dis):
    if ln.startswith(".text." + kname + ":"):
        inside = True; continue
    if inside and ln.startswith("//--------------------- .text."):
        break
    if not inside: continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        # helper code (other files, lines < 131 of the kernel file) is charged to the last call-site line
        if m.group(1).endswith("rp_kernels.cu") and int(m.group(2)) > HELPER_END: cur = int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ino = hdr.index("stall_no_inst")
base = int(rows[2][ia], 16)
inst, samp, noi, cnt = defaultdict(int), defaultdict(int), defaultdict(int), defaultdict(int)
for r in rows[2:]:
    if len(r) <= ino: continue
    line = off2line.get(int(r[ia], 16) - base)
    inst[line] += int(r[ii]); samp[line] += int(r[isamp]); noi[line] += int(r[ino]); cnt[line] += 1
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total inst {ti:.3e} samples {ts} sass-instrs {sum(cnt.values())}")
print("line   sass   inst%  samp%  noinst%ofsamp")
for line, _ in sorted(samp.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{line!s:>5} {cnt[line]:6d} {100*inst[line]/ti:6.2f} {100*samp[line]/ts:6.2f} {100*noi[line]/max(1,samp[line]):6.1f}")
# coarse ranges
if len(sys.argv) > 5:
    for rng in sys.argv[5:]:
        a, b = map(int, rng.split("-"))
        si = sum(v for k, v in inst.items() if k and a <= k <= b); ss = sum(v for k, v in samp.items() if k and a <= k <= b)
        sc = sum(v for k, v in cnt.items() if k and a <= k <= b)
        print(f"lines {a}-{b}: sass {sc} inst {100*si/ti:.1f}% samples {100*ss/ts:.1f}%")
