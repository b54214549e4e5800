import re,csv,collections,sys
src,kernel=sys.argv[1],sys.argv[2]
dis='/tmp/all_dis.txt'
cur=None; inside=False; lines={}
for ln in open(dis):
    if ln.startswith(".text."): inside = kernel in ln; continue
    if not inside: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m: lines[int(m.group(1),16)]=(cur,m.group(2).strip())
# function ranges by scanning source files
import os
def ranges(path):
    out=[]; name=None; start=None
    for i,l in enumerate(open(path),1):
        m=re.match(r'^(?:template.*\n)?(?:ARB_\w+|static|inline)\s+[\w:<>\*& ]+?\b(\w+)\s*\(', l)
        if m and not l.startswith(' '):
            if name: out.append((start,i-1,name))
            name,start=m.group(1),i
    if name: out.append((start,10**9,name))
    return out
R={f:ranges('/root/repo/arboris-python_b200/csrc/'+f) for f in os.listdir('/root/repo/arboris-python_b200/csrc') if f.endswith(('.cuh','.cu','.h'))}
rows=list(csv.reader(open(src))); hdr=rows[1]
ia,ie,it,isamp=hdr.index("Address"),hdr.index("Instructions Executed"),hdr.index("Thread Instructions Executed"),hdr.index("# Samples")
base=int(rows[2][ia],16)
agg=collections.defaultdict(lambda:[0,0,0]); tot=0; tots=0
for r in rows[2:]:
    off=int(r[ia],16)-base
    k,ins=lines.get(off,(None,""))
    if k is None: name='?'
    else:
        name=k[0]
        for lo,hi,nm in R.get(k[0],[]):
            if lo<=k[1]<=hi: name=nm
    agg[name][0]+=int(r[ie]); agg[name][1]+=int(r[it]); agg[name][2]+=int(r[isamp]); tot+=int(r[ie]); tots+=int(r[isamp])
nw=float(sys.argv[3]) if len(sys.argv)>3 else 2048
print("total warp-instr %d  (%.0f per warp)"%(tot,tot/nw))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:22]:
    print("%-28s %5.1f%% instr  %7.0f instr/warp  %4.1f thr/instr  %5.1f%% samples"%(k,100*v[0]/tot, v[0]/nw, v[1]/max(v[0],1), 100*v[2]/tots))
