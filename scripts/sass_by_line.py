#!/usr/bin/env python
"""Static SASS size per source-line range of a kernel: sass_by_line.py lib.so kernel_mangled [ranges a-b ...]"""
import re, subprocess, sys, tempfile, os, collections
so, kname = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if "rp_kernels" in f and "rp_host" not in f][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
inside, cur, cnt, other = False, None, collections.Counter(), collections.Counter()
for ln in dis.split("\n"):
    if ln.startswith(".text." + kname + ":"): inside = True; continue
    if inside and ln.startswith("//--------------------- "): break
    if not inside: continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): cnt[cur] += 1
tot = sum(cnt.values())
print("total", tot)
byfile = collections.Counter()
for (f, l), n in cnt.items(): byfile[f] += n
print(dict(byfile))
for rng in sys.argv[3:]:
    a, b = map(int, rng.split("-"))
    print(f"rp_kernels.cu {a}-{b}: {sum(n for (f, l), n in cnt.items() if f == 'rp_kernels.cu' and a <= l <= b)}")
top = sorted(cnt.items(), key=lambda kv: -kv[1])[:25]
for (f, l), n in top: print(f, l, n)
