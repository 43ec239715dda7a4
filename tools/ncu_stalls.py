import csv, sys, collections, subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=rows[1]; si=h.index("# Samples"); src=h.index("Source")
stall_cols=[i for i,c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
agg=collections.Counter(); data=[]
for idx,r in enumerate(rows[2:]):
    try: n=int(r[si])
    except: continue
    data.append((n,idx,r))
    for i in stall_cols:
        try: agg[h[i]]+=int(r[i] or 0)
        except: pass
tot=sum(agg.values()); print("samples",tot,"lines",len(data))
print(", ".join(f"{k[6:]}={100*v/tot:.1f}%" for k,v in agg.most_common(8)))
for n,idx,r in sorted(data,key=lambda x:-x[0])[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    st=sorted(((int(r[i] or 0),h[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{idx:5d} {n:6d} {100*n/tot:5.1f}% {r[src][:64]:64s} {st}")
