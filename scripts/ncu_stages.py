#!/usr/bin/env python
"""Per-stage instruction / stall breakdown of one captured launch of the step kernel.
Inputs: `ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv`; `nvdisasm --print-line-info kernels.cubin >
k.sass` (same build).  An instruction is charged to the outermost rp_kernels.cu line of its inline chain; the two
__noinline__ functions (sweep_inverse, pgs_solve) are split by their own source lines.
usage: ncu_stages.py src.csv k.sass kernel_mangled_prefix rp_kernels.cu"""
import csv, re, sys
from collections import defaultdict

src_csv, dis, kname, cu_src = sys.argv[1:5]
# stage boundaries from marker text in the kernel source of the captured build (first line containing the marker)
src_lines = open(cu_src).read().split("\n")


def line_of(marker, after=0):
    for i, l in enumerate(src_lines):
        if i + 1 > after and marker in l:
            return i + 1
    raise SystemExit(f"marker not found: {marker}")


def ranges(marks, end):
    starts = []
    for m, name in marks:
        starts.append((line_of(m, starts[-1][0] if starts else 0), name))
    return [(lo, (starts[i + 1][0] - 1) if i + 1 < len(starts) else end, name) for i, (lo, name) in enumerate(starts)]


k0 = line_of("rp_step_kernel(RpKernelArgs a) {")
KERNEL = ranges([("rp_step_kernel(RpKernelArgs a) {", "prologue"), ("hand kinematics (by level)", "kinematics"),
                 ("publish S, velocities", "RNE / CRB / scatter M"), ("actuation (lane = actuator)", "actuation, smooth, key rows"),
                 ("hand friction + limit rows", "hand friction / limit row setup"), ("collision (lane = geom of this warp's hand)", "collision"),
                 ("publish what the contact rows read from the other warp", "contact rows + lists"), ("== warm start", "warm start, U, cost"),
                 ("== PGS", "solver call"), ("dual -> primal, Euler", "dual->primal, Euler"), ("    // integrate", "integrate, midi"),
                 ("--- store state", "store state"), ("task epilogue: after_step", "task epilogue")], line_of("Launch order for the next step") - 1)
p0 = line_of("__device__ __noinline__ int pgs_solve(")
PGS = ranges([("__device__ __noinline__ int pgs_solve(", "pgs: prologue"), ("---- friction-loss rows", "pgs: friction rows"),
              ("---- limit rows of keys that carry contacts", "pgs: coupled key rows"), ("---- hand limit rows", "pgs: hand limit rows"),
              ("---- contacts of this hand", "pgs: contacts"), ("if (lane == 0) sm.red[iter & 1][h][1] = improvement;", "pgs: per-sweep barrier + test")],
             line_of("io->fr_f = fr_f; io->lm_f = lm_f; io->va = va;", p0) + 2)
lines = open(dis).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text." + kname)][0]
end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("//--------------------- "))
func, off2, pend, cur = "kernel", {}, [], None
for l in lines[start:end]:
    if l.startswith("$") and l.endswith(":"):
        func = ("pgs_solve" if ("pgs_solve" in l and "pgs_solve_hc" not in l) else "sweep_inverse" if "sweep_inverse" in l
                else "phase barrier (phase_sync)" if "phase_sync" in l else "library routine")
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)( inlined at)?', l)
    if m:
        pend.append((m.group(1).split("/")[-1], int(m.group(2)), bool(m.group(3)))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if m:
        if pend:
            outer = [p for p in pend if p[0] == "rp_kernels.cu"]
            cur = outer[-1] if outer else pend[-1]; pend = []
        off2[int(m.group(1), 16)] = (func, cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [(h[6:], hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
agg = defaultdict(lambda: defaultdict(int))
last_pgs = "pgs: prologue"
tot_s = tot_i = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    f, cl = off2.get(int(r[ia], 16) - base, ("?", ("?", 0, 0)))
    name = f
    if f == "kernel":
        name = "other"
        if cl and cl[0] == "rp_kernels.cu":
            for lo, hi, n in KERNEL:
                if lo <= cl[1] <= hi: name = n; break
    elif f == "pgs_solve":
        if cl and cl[0] == "rp_kernels.cu":
            for lo, hi, n in PGS:
                if lo <= cl[1] <= hi: last_pgs = n; break
        name = last_pgs
    a = agg[name]
    a["inst"] += int(r[iex]); a["samp"] += int(r[isamp]); a["sass"] += 1
    tot_s += int(r[isamp]); tot_i += int(r[iex])
    for h, k in stalls: a[h] += int(r[k] or 0)
print(f"total warp-instr {tot_i:.3e}  samples {tot_s}")
print(f"{'stage':34s} {'sass':>6s} {'inst%':>6s} {'samp%':>6s} {'rel':>5s}  top stalls (% of the stage's samples)")
order = ["phase barrier (phase_sync)"] + [n for _, _, n in KERNEL[:8]] + ["sweep_inverse"] + [n for _, _, n in PGS] + [n for _, _, n in KERNEL[8:]] + ["library routine", "other"]
for name in order:
    a = agg.get(name)
    if not a or not a["samp"]: continue
    top = sorted(((a[h], h) for h, _ in stalls), reverse=True)[:5]
    print(f"{name:34s} {a['sass']:6d} {100*a['inst']/tot_i:6.2f} {100*a['samp']/tot_s:6.2f} {(a['samp']/tot_s)/(a['inst']/tot_i + 1e-12):5.2f}  "
          + ", ".join(f"{h} {100*v/a['samp']:.0f}" for v, h in top if v))
