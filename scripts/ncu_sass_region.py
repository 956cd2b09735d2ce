#!/usr/bin/env python
"""SASS of the instructions whose outermost kernel line falls in [lo, hi], in address order, with ncu samples.
usage: ncu_sass_region.py sass.csv disi.txt kernel_mangled_name lo hi"""
import csv, re, sys
sass_csv, dis, kname, lo, hi = sys.argv[1:6]
lo, hi = int(lo), int(hi)
info, inside, pend, cur = {}, False, [], None
for ln in open(dis):
    if ln.startswith(".text." + kname + ":"):
        inside = True; continue
    if inside and ln.startswith("//--------------------- .text."):
        break
    if not inside: continue
    m = re.search(r'//## File "(.*?)", line (\d+)( inlined at)?', ln)
    if m:
        pend.append((m.group(1).split("/")[-1], int(m.group(2)), bool(m.group(3)))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m:
        if pend:
            outer = [p for p in pend if not p[2]]
            cur = (outer[-1] if outer else pend[-1], pend[0]); pend = []
        info[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(sass_csv))); hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [(h[6:], hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
for r in rows[2:]:
    if len(r) <= stalls[-1][1]: continue
    off = int(r[ia], 16) - base
    if off not in info: continue
    ((of, ol, _), (inf, inl, _)), txt = info[off]
    if of != "rp_kernels.cu" or not (lo <= ol <= hi): continue
    st = sorted(((int(r[k] or 0), n) for n, k in stalls), reverse=True)[:2]
    print(f"{off:6x} {ol:5d} {inf[:7]:7s}:{inl:<4d} x{int(r[ii])/1e6:8.2f}M s{int(r[isamp]):6d} {st[0][1][:9]:9s} {st[1][1][:9]:9s} {txt[:70]}")
