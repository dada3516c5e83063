import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); H=rows[0]
want=['gpu__time_duration.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__shared_mem_per_block_dynamic','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum']
for w in want:
    for i,h in enumerate(H):
        if h==w: print(w, [r[i] for r in rows[2:]])
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
o=[]; cur=None; H=None
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>5 and r[0]=='Line No': H=r; continue
    if H and len(r)==len(H) and r[0]!='':
        try: ln=int(r[0])
        except: continue
        d=dict(zip(H,r))
        o.append((cur,ln,int(d['# Samples'] or 0),int(d['Instructions Executed'] or 0),r[1][:100],int(d.get('stall_barrier',0) or 0)))
tot_s=sum(x[2] for x in o); tot_i=sum(x[3] for x in o)
src=open('/root/repo/compressor-mpc_b200/csrc/step_kernel.cuh').read().splitlines()
marks=[(i+1,l.strip()) for i,l in enumerate(src) if '// ---- phase' in l]
def phase(ln):
    name='helpers(frag/tile)'
    for m,l in marks:
        if ln>=m: name=l[8:40]
    return name
agg={}
for x in o:
    key = phase(x[1]) if x[0]=='step_kernel.cuh' else x[0]
    a=agg.setdefault(key,[0,0,0]); a[0]+=x[2]; a[1]+=x[3]; a[2]+=x[5]
print("total inst", tot_i, "per scenario", tot_i/4096)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print(f"{k:40s} inst={v[1]/tot_i:6.1%} samples={v[0]/tot_s:6.1%} barrier={v[2]/tot_s:6.1%}")
n=int(sys.argv[2]) if len(sys.argv)>2 else 16
for x in sorted(o,key=lambda x:-x[2])[:n]:
    print(f"{x[0]:16s}:{x[1]:4d} inst={x[3]/tot_i:6.1%} samp={x[2]/tot_s:6.1%} bar={x[5]/tot_s:6.1%}| {x[4]}")
