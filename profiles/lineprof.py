"""profiles/lineprof.py CUBIN KERNEL NCU_REP SOURCE [N] -- per-source-line instruction / sample shares of one kernel from an
ncu --set full --import-source on capture (nvdisasm -g maps SASS offsets to lines; needs the cubin of the captured build)."""
import re,csv,collections,sys,subprocess
cubin, kern, rep, srcfile = sys.argv[1:5]
txt=subprocess.run(["nvdisasm","-g","-c",cubin],capture_output=True,text=True).stdout.split('\n')
infn=False; cur=None; off2line={}
for L in txt:
    if L.startswith('.text.'):
        infn = kern in L; continue
    m=re.match(r'\s*\.section\s+\.text\.(\S+)',L)
    if m: infn = kern in m.group(1); continue
    if not infn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',L)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*)',L)
    if m: off2line[int(m.group(1),16)]=cur
rows=list(csv.reader(subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout.splitlines()))
h=rows[1]; ix={n:i for i,n in enumerate(h)}; body=rows[2:]
a0=int(body[0][ix['Address']],16)
agg=collections.Counter(); aggt=collections.Counter(); smp=collections.Counter(); tot=0
for r in body:
    off=int(r[ix['Address']],16)-a0
    key=off2line.get(off)
    v=float(r[ix['Instructions Executed']]); t=float(r[ix['Thread Instructions Executed']]); s=float(r[ix['# Samples']])
    agg[key]+=v; aggt[key]+=t; smp[key]+=s; tot+=v
import os
srcdir=os.path.dirname(srcfile); srcs={}
def line_of(k):
    if not k: return str(k)
    f=os.path.join(srcdir,k[0])
    if f not in srcs:
        try: srcs[f]=open(f).read().split('\n')
        except OSError: srcs[f]=None
    if srcs[f] is None or k[1]-1>=len(srcs[f]): return str(k)
    return (k[0][:12]+': '+srcs[f][k[1]-1].strip())[:110]
ts=sum(smp.values())
print("total warp inst %.3g, mapped %d offsets"%(tot,len(off2line)))
for k,v in agg.most_common(int(sys.argv[5]) if len(sys.argv)>5 else 45):
    txt_ = line_of(k)
    print(f"{100*v/tot:5.1f}% inst {100*smp[k]/ts:5.1f}% smp thr/inst {aggt[k]/max(v,1):5.1f} {k[1] if k else '?':>4} {txt_}")
