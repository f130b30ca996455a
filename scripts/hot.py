import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
W=4096.0; T=1800.0
out=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    ex=int(r[ix["Instructions Executed"]] or 0)
    out.append((int(r[ix["Address"]],16), r[ix["Source"]].strip(), ex/W/T, int(r[ix["# Samples"]] or 0),
                {k:int(r[ix[k]] or 0) for k in ["stall_wait","stall_short_sb","stall_no_inst","stall_branch_resolving","stall_math","stall_selected","stall_not_selected","stall_long_sb","stall_dispatch","stall_lg","stall_mio"]}))
base=out[0][0]
tot=sum(o[3] for o in out)
mode=sys.argv[2] if len(sys.argv)>2 else "list"
if mode=="list":
    lo=float(sys.argv[3]) if len(sys.argv)>3 else 0.5
    for a,s,e,n,st in out:
        if e>=lo:
            top=sorted(st.items(),key=lambda kv:-kv[1])[:2]
            print(f"{a-base:#07x} {e:6.2f} {100*n/tot:5.2f}% {s[:70]:70s} {' '.join(f'{k[6:]}={v}' for k,v in top if v)}")
else:
    # bucket by exec count class
    b=collections.defaultdict(lambda: collections.Counter())
    for a,s,e,n,st in out:
        cls = "stage(>=3.5)" if e>=3.5 else ("tick(0.9..3.5)" if e>=0.9 else ("fsw(0.05..0.9)" if e>=0.05 else "rare"))
        op=[o for o in s.split() if not o.startswith("@")][0].split(".")[0]
        b[cls][op]+=e
    for cls,c in b.items():
        t=sum(c.values())
        print(cls, f"total {t:.0f}/tick:", ", ".join(f"{k} {v:.0f}" for k,v in c.most_common(18)))
