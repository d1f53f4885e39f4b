"""Top stalled SASS instructions of every kernel in an `ncu --page source --csv --print-source sass` export: python tools/ncu_source_top.py file.csv [N]"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hidx=[i for i,r in enumerate(rows) if r and r[0]=="Address"]
for k,h in enumerate(hidx):
    hdr=rows[h]; end=hidx[k+1]-1 if k+1<len(hidx) else len(rows)
    data=[r for r in rows[h+1:end] if len(r)==len(hdr)]
    isrc=hdr.index("Source"); isamp=hdr.index("# Samples"); iex=hdr.index("Instructions Executed")
    stalls=[c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot=sum(int(r[isamp]) for r in data)
    print("=== section",k,"total samples",tot,"instr",len(data))
    top=sorted(enumerate(data), key=lambda ir: -int(ir[1][isamp]))[:int(sys.argv[2]) if len(sys.argv)>2 else 30]
    for i,r in sorted(top):
        st=sorted(((int(r[hdr.index(c)]),c[6:]) for c in stalls),reverse=True)[:2]
        print("%5d %6s %9s  %-70s %s"%(i, r[isamp], r[iex], r[isrc].strip()[:70], st))
