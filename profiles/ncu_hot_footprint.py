import csv,re,collections,sys
sass,kern=sys.argv[1],sys.argv[2]
rows=list(csv.reader(open(sass)))
for i,r in enumerate(rows):
    if 'Source' in r and 'Instructions Executed' in r: hdr=r; start=i+1; break
ei=hdr.index('Instructions Executed')
cur=None; inside=False; lines={}
for ln in open('/tmp/dis/arb_fused.sm_100a.txt'):
    if ln.startswith(".text."): inside = kern in ln and 'coop' not in ln; continue
    if not inside: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m: lines[int(m.group(1),16)]=cur
def ranges(path):
    out=[]
    for i,l in enumerate(open(path).read().split('\n'),1):
        m=re.match(r'^(?:ARB_\w+|static|inline|template.*)\s+.*?\b(\w+)\s*\([^;]*$', l)
        if m and not l.startswith(' ') and not l.startswith('#'): out.append((i,m.group(1)))
    return out
R={f:ranges('/root/repo/arboris-python_b200/csrc/'+f) for f in ('arb_smallmat.cuh','arb_constraints.cuh','arb_fused.cuh','arb_artic.cuh','arb_math.cuh','arb_joints.cuh')}
def fn(file,line):
    if file not in R: return file
    name='?'
    for l,n in R[file]:
        if l<=line: name=n
        else: break
    return file.replace('arb_','').replace('.cuh','')+':'+name
hot=collections.Counter(); dyn=collections.Counter(); base=None
thr=int(sys.argv[3]) if len(sys.argv)>3 else 10000
for r in rows[start:]:
    try: a=int(r[0],16) if r[0].startswith('0x') else int(r[0])
    except: continue
    if base is None: base=a
    ex=int(r[ei] or 0); ln=lines.get(a-base); key=fn(*ln) if ln else '?'
    dyn[key]+=ex
    if ex>=thr: hot[key]+=1
tot=sum(dyn.values())
for k,v in hot.most_common(22): print('%-40s %6d instr %6.1f KB  %5.1f%% dyn'%(k,v,v*16/1024,100*dyn[k]/tot))
print('total hot %.1f KB'%(sum(hot.values())*16/1024))
