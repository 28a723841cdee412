import csv,sys,subprocess
rep=sys.argv[1]; n=int(sys.argv[2]) if len(sys.argv)>2 else 40; scale_envturns=float(sys.argv[3]) if len(sys.argv)>3 else 262144*25
out=subprocess.run(['ncu','-i',rep,'--page','source','--print-source','sass,cuda','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Line No"][0]
hdr=rows[hi]; iex=hdr.index("Instructions Executed"); ism=hdr.index("# Samples")
agg={}; smp={}
for r in rows[hi+1:]:
    if len(r)<iex+1: continue
    if r[0]!="" and r[2]=="-":
        try:
            k=(int(r[0]),r[1][:88]); agg[k]=agg.get(k,0)+int(r[iex]); smp[k]=smp.get(k,0)+int(r[ism])
        except: pass
tot=sum(agg.values()); ts=sum(smp.values())
print("total warp-inst",tot,"per env-turn %.1f"%(tot/scale_envturns),"per warp-turn(8 envs) %.0f"%(tot/scale_envturns*8))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:n]:
    print(f"{v/tot*100:5.1f}% smp {smp[k]/ts*100:4.1f}% {v/scale_envturns*8:7.1f}/wt L{k[0]:>4d} {k[1]}")
