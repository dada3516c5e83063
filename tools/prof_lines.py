import csv,sys,subprocess,collections,io
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Line No"][0]
hdr=rows[hi]
iW=hdr.index("L1 Wavefronts Shared"); iX=hdr.index("L1 Wavefronts Shared Excessive"); iS=hdr.index("# Samples"); iE=hdr.index("Instructions Executed")
B=4096
agg={}; ops=collections.Counter(); cur=None
for r in rows[hi+1:]:
    if len(r)<len(hdr): continue
    if r[2]=='-':
        try: cur=int(r[0])
        except: cur=None; continue
        agg[cur]=(float(r[iW] or 0),float(r[iX] or 0),float(r[iS] or 0),float(r[iE] or 0),r[1]); continue
    t=r[3].strip().split()
    if not t: continue
    op=t[1] if t[0].startswith('@') else t[0]
    try: ops[op.split('.')[0]]+=float(r[iE] or 0)/B
    except ValueError: pass
print("ops/scenario:", ', '.join(f"{k}:{v:.0f}" for k,v in ops.most_common(22)), " total", sum(ops.values()))
tot=sum(v[0] for v in agg.values()); print("wavefronts/scen", tot/B, "excess", sum(v[1] for v in agg.values())/B, "samples", sum(v[2] for v in agg.values()))
n=int(sys.argv[2]) if len(sys.argv)>2 else 25
print("-- top by shared wavefronts")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:n]:
    print(k, f"wave {v[0]/B:6.0f} exc {v[1]/B:5.0f} samp {v[2]:5.0f} inst {v[3]/B:6.0f} |", v[4][:95])
print("-- top by samples")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][2])[:n]:
    print(k, f"wave {v[0]/B:6.0f} exc {v[1]/B:5.0f} samp {v[2]:5.0f} inst {v[3]/B:6.0f} |", v[4][:95])
