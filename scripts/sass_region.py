#!/usr/bin/env python
"""Static SASS of a source-line region (call sites + inlined helpers): sass_region.py lib.so kernel a b"""
import re,subprocess,os,tempfile,collections,sys
so,kname,a,b=sys.argv[1],sys.argv[2],int(sys.argv[3]),int(sys.argv[4])
tmp=tempfile.mkdtemp()
subprocess.check_call(["cuobjdump","-xelf","all",os.path.abspath(so)],cwd=tmp,stdout=subprocess.DEVNULL)
cub=[f for f in os.listdir(tmp) if "rp_kernels" in f and "rp_host" not in f][0]
dis=subprocess.run(["nvdisasm","--print-line-info",os.path.join(tmp,cub)],capture_output=True,text=True).stdout
inside=False;cur=None;curf=None;rows=[]
for ln in dis.split("\n"):
    if ln.startswith(".text."+kname+":"): inside=True;continue
    if inside and ln.startswith("//--------------------- "): break
    if not inside: continue
    m=re.search(r'//## File "(.*?)", line (\d+)',ln)
    if m:
        curf=(os.path.basename(m.group(1)),int(m.group(2)))
        if m.group(1).endswith('rp_kernels.cu') and int(m.group(2))>200: cur=int(m.group(2))
        continue
    m=re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)",ln)
    if m: rows.append((int(m.group(1),16),cur,curf,m.group(2)))
sel=[r for r in rows if r[1] and a<=r[1]<=b]
print("static instrs", len(sel))
agg=collections.Counter()
for r in sel: agg[(r[1], r[2][0] if r[2][0]!='rp_kernels.cu' else '', r[2][1] if r[2][0]!='rp_kernels.cu' else 0)]+=1
for k,v in sorted(agg.items()):
    if v>=12: print(k,v)
ops=collections.Counter(re.sub(r'@!?U?P\d\s+','',r[3]).split()[0].split('.')[0] for r in sel)
print(ops.most_common(30))
