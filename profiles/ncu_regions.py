import csv,re,collections,sys
sass,dis,kern=sys.argv[1:4]
regions=eval(sys.argv[4])  # list of (lo,hi,name) for arb_fused.cuh
rows=list(csv.reader(open(sass)))
for i,r in enumerate(rows):
    if 'Source' in r and 'Instructions Executed' in r: hdr=r; start=i+1; break
cur=None; inside=False; lines={}
for ln in open(dis):
    if ln.startswith(".text."): inside = kern in ln and 'coop' not in ln; continue
    if not inside: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m: lines[int(m.group(1),16)]=cur
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg=collections.defaultdict(collections.Counter)
base=None; tin=0
for r in rows[start:]:
    try: a=int(r[0],16) if r[0].startswith('0x') else int(r[0])
    except: continue
    if base is None: base=a
    ln=lines.get(a-base); f=ln[0] if ln else '?'
    if f=='arb_fused.cuh':
        f='fused:other'
        for lo,hi,name in regions:
            if lo<=ln[1]<=hi: f='fused:'+name
    for h in stalls:
        v=r[hdr.index(h)]
        if v: agg[f][h]+=int(v)
    n=int(r[hdr.index('Instructions Executed')] or 0); agg[f]['instr']+=n; tin+=n
tot=sum(sum(v for k,v in c.items() if k!='instr') for c in agg.values())
for f,c in sorted(agg.items(), key=lambda kv:-sum(v for k,v in kv[1].items() if k!='instr')):
    s=sum(v for k,v in c.items() if k!='instr')
    top=sorted(((v,k) for k,v in c.items() if k!='instr'),reverse=True)[:4]
    print('%-22s %5.1f%% samples  %5.1f%% instr  '%(f,100*s/tot,100*c['instr']/tin), ' '.join('%s=%.0f%%'%(k[6:],100*v/s) for v,k in top))
