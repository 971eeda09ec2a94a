"""Top SASS lines by stall samples per barrier-delimited region.  python tools/ncu_hot.py src.csv [topN]"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 25
hdr=rows[1]; data=rows[2:]
ia=hdr.index("Source"); ie=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples")
stall_cols=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
reg=0; out=[]
for i,r in enumerate(data):
    op=[t for t in r[ia].split() if not t.startswith('@')]
    o=op[0] if op else ''
    out.append((reg,i,int(r[ie]),int(r[isamp]),r[ia].strip(), {hdr[c]:int(r[c]) for c in stall_cols if r[c] not in ('0','')}))
    if o.startswith('BAR'): reg+=1
for rg in sorted(set(x[0] for x in out)):
    sel=[x for x in out if x[0]==rg]
    tot=sum(x[3] for x in sel)
    if tot < 200: continue
    agg=collections.Counter()
    for x in sel:
        for k,v in x[5].items(): agg[k]+=v
    print(f"--- region {rg}: {len(sel)} SASS, {tot} samples; stalls: {agg.most_common(6)}")
    for x in sorted(sel,key=lambda x:-x[3])[:top]:
        print(f"{x[1]:5d} n={x[2]:8d} s={x[3]:5d} {x[4][:64]:64s} {dict(sorted(x[5].items(), key=lambda kv:-kv[1])[:2])}")
