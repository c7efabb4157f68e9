import csv, subprocess, sys, io
from collections import Counter
rep=sys.argv[1]; kname=sys.argv[2]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw))); hdr=rows[0]
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","smsp__inst_executed.sum","smsp__cycles_active.avg","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","lts__t_sector_hit_rate.pct","l1tex__t_sector_hit_rate.pct","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum","l1tex__t_requests_pipe_lsu_mem_global_op_st.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","lts__t_bytes.sum","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers","l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts.sum","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    if len(r)<len(hdr) or kname not in r[hdr.index("Kernel Name")]: continue
    print("==", r[hdr.index("Kernel Name")][:70])
    for w in want:
        if w in hdr: print(f"   {w} = {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
    st=[(float(r[i]),h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio","")) for i,h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and r[i]]
    print("   stalls:", ", ".join(f"{n}={v:.2f}" for v,n in sorted(st,reverse=True)[:7]))
    break
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kname],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hdr=rows[1]; si=hdr.index("Source"); ei=hdr.index("Instructions Executed"); wi=hdr.index("Warp Stall Sampling (All Samples)")
data=[]; ops=Counter()
for k,r in enumerate(rows[2:]):
    if len(r)<len(hdr):
        if r and r[0]=="Kernel Name": break
        continue
    try: data.append((int(r[wi]),k,r[si][:90],int(r[ei])))
    except: pass
tot=sum(d[0] for d in data); ti=sum(d[3] for d in data)
print("   SASS lines",len(data),"warp instr",ti,"samples",tot)
for w,k,s,e in sorted(data,reverse=True)[:int(sys.argv[3]) if len(sys.argv)>3 else 14]: print(f"   {100*w/tot:5.1f}%  line {k:5d} exec {e:8d}  {s}")
for d in data:
    op=d[2].split()[0] if not d[2].startswith('@') else d[2].split()[1]
    ops[op.split('.')[0]]+=d[3]
print("   ops:", ", ".join(f"{o}={100*c/ti:.1f}%" for o,c in ops.most_common(12)))
