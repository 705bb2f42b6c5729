import csv, sys
rows=list(csv.DictReader(open(sys.argv[1])))
tot=sum(float(r['ms']) for r in rows)
print("total ms",round(tot,1), "classes",len(rows))
acc=0
nf=lambda l:(l+1)*(l+2)//2
nat=lambda a,b,c,d: nf(a)*nf(b)+nf(c)*nf(d)+nf(a)*nf(c)+nf(a)*nf(d)+nf(b)*nf(c)+nf(b)*nf(d)
for r in rows[:int(sys.argv[2]) if len(sys.argv)>2 else 40]:
    acc+=float(r['ms'])
    c=r['class']; li,lj,lk,ll=int(c[1]),int(c[2]),int(c[4]),int(c[5])
    N=nf(li)*nf(lj)*nf(lk)*nf(ll); q=float(r['quartets']); ms=float(r['ms'])
    print(c,"N=%4d"%N,"ms %8.1f"%ms,"q %.2e"%q,"TF %6.3f"%float(r['tflops']),"ns/q %7.2f"%(ms*1e6/q),"atom/s %.2e"%(q*nat(li,lj,lk,ll)/ms*1e3),"cum %.1f%%"%(100*acc/tot))
