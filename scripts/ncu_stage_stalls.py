#!/usr/bin/env python
"""Per source-range stall breakdown from an ncu source-page CSV.
Inputs: `ncu -i X.ncu-rep --page source --csv > sass.csv`; `nvdisasm -gi kernels.cubin > disi.txt` (same build).
An instruction is charged to the OUTERMOST kernel line of its inline chain (so helper code lands in the stage
that calls it); __noinline__ helpers (sweep_inverse, lines < 243) are their own stage.
usage: ncu_stage_stalls.py sass.csv disi.txt kernel_mangled_name [--lines lo-hi] lo-hi:name [lo-hi:name ...]"""
import csv, re, sys
from collections import defaultdict
args = sys.argv[1:]
sass_csv, dis, kname = args[:3]
detail = None
rest = args[3:]
if rest and rest[0] == "--lines":
    lo, hi = rest[1].split("-"); detail = (int(lo), int(hi)); rest = rest[2:]
ranges = []
for s in rest:
    r, n = s.split(":"); lo, hi = r.split("-"); ranges.append((int(lo), int(hi), n))
off2line, inside = {}, False
pend = []  # annotations seen since the last instruction
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
            cur = (outer[-1] if outer else pend[-1], pend[0])
            pend = []
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
agg = defaultdict(lambda: defaultdict(int))
dl = defaultdict(lambda: defaultdict(int))
tot_s = tot_i = 0
for r in rows[2:]:
    if len(r) <= stalls[-1][1]: continue
    (of, ol, _), (inf, inl, _) = off2line.get(int(r[ia], 16) - base, (("?", 0, 0), ("?", 0, 0)))
    name = "other"
    for lo, hi, n in ranges:
        if of == "rp_kernels.cu" and lo <= ol <= hi: name = n; break
    a = agg[name]
    a["inst"] += int(r[ii]); a["samp"] += int(r[isamp]); a["sass"] += 1
    tot_s += int(r[isamp]); tot_i += int(r[ii])
    for h, k in stalls: a[h] += int(r[k] or 0)
    if detail and of == "rp_kernels.cu" and detail[0] <= ol <= detail[1]:
        d = dl[(ol, f"{inf[:8]}:{inl}")]
        d["inst"] += int(r[ii]); d["samp"] += int(r[isamp]); d["sass"] += 1
        for h, k in stalls: d[h] += int(r[k] or 0)
print(f"total warp-instr {tot_i:.3e}  samples {tot_s}")
print(f"{'stage':28s} {'sass':>6s} {'inst%':>6s} {'samp%':>6s} {'rel':>5s}  top stalls (% of the stage's samples)")
for name in [n for _, _, n in ranges] + ["other"]:
    a = agg[name]
    if not a["samp"]: continue
    top = sorted(((a[h], h[6:]) for h, _ in stalls), reverse=True)[:5]
    rel = (a["samp"] / tot_s) / max(a["inst"] / tot_i, 1e-12)
    print(f"{name:28s} {a['sass']:6d} {100*a['inst']/tot_i:6.2f} {100*a['samp']/tot_s:6.2f} {rel:5.2f}  " + ", ".join(f"{n} {100*v/a['samp']:.0f}" for v, n in top))
if detail:
    print(f"--- detail, outer lines {detail[0]}-{detail[1]} (outer line, innermost file:line)")
    for key, d in sorted(dl.items(), key=lambda kv: -kv[1]["samp"])[:40]:
        top = sorted(((d[h], h[6:]) for h, _ in stalls), reverse=True)[:3]
        print(f"  {key[0]:5d} {key[1]:16s} sass {d['sass']:4d} inst {100*d['inst']/tot_i:5.2f}% samp {100*d['samp']/tot_s:5.2f}%  " + ", ".join(f"{n} {100*v/max(1,d['samp']):.0f}" for v, n in top))
