import csv, subprocess, sys, io
from collections import Counter
rep=sys.argv[1]; seg=int(sys.argv[2]) if len(sys.argv)>2 else 150
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hi=next(i for i,r in enumerate(rows) if "Source" in r and "Address" in r)
hdr=rows[hi]; data=rows[hi+1:]
ia=hdr.index("Source"); isamp=hdr.index("Warp Stall Sampling (All Samples)"); iex=hdr.index("Instructions Executed")
iw=hdr.index("L1 Wavefronts Shared"); iwi=hdr.index("L1 Wavefronts Shared Ideal")
tot=sum(int(r[isamp]) for r in data); totw=max(1,sum(int(r[iw] or 0) for r in data)); tote=sum(int(r[iex]) for r in data)
print("samples",tot,"wavefronts",totw,"instr",tote,"lines",len(data))
for k in range(0,len(data),seg):
    d=data[k:k+seg]
    s=sum(int(r[isamp]) for r in d); w=sum(int(r[iw] or 0) for r in d); wi=sum(int(r[iwi] or 0) for r in d); e=sum(int(r[iex]) for r in d)
    if s*200<tot and e*200<tote: continue
    ops=Counter()
    for r in d:
        t=r[ia].split(); ops[(t[1] if t[0].startswith('@') else t[0]).split('.')[0]]+=int(r[iex])
    top=', '.join(f"{o}:{100*n/max(e,1):.0f}%" for o,n in ops.most_common(5))
    print(f"{k:5d} samples {100*s/tot:5.1f}%  instr {100*e/tote:5.1f}% wavefronts {100*w/totw:5.1f}% (ideal/act {wi/max(w,1):.2f})  {top}")
