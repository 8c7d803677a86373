import csv, sys, subprocess
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","smsp__cycles_active.avg","sm__cycles_elapsed.max","smsp__inst_executed.sum","launch__grid_size","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","lts__t_sectors_op_read.sum","lts__t_sectors_op_write.sum","lts__t_sectors_srcunit_tex_op_read.sum","lts__t_sectors_srcunit_tex_op_write.sum"]
seen=set()
for r in rows[2:]:
    rec=dict(zip(hdr,r))
    k=rec["Kernel Name"][:40]
    if k in seen: continue
    seen.add(k)
    print("==",k)
    for w in want:
        if w in rec: print("  ",w, rec[w], units[hdr.index(w)])
    for i,h in enumerate(hdr):
        if 'issue_stalled' in h and 'ratio' in h and float(r[i] or 0)>0.3:
            print("     stall", h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""), r[i])
for kern in sys.argv[2:]:
    src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kern],capture_output=True,text=True).stdout
    rows=list(csv.reader(src.splitlines()))
    hdr=rows[1]
    iS=hdr.index("Source"); iSamp=hdr.index("# Samples"); iInst=hdr.index("Instructions Executed")
    data=[]
    for r in rows[2:]:
        if len(r)<len(hdr) or r[0] in ("Kernel Name","Address"): break
        data.append((r[iS].strip(), int(r[iSamp] or 0), int(r[iInst] or 0)))
    print("##",kern,len(data),"SASS; executed",sum(d[2] for d in data),"samples",sum(d[1] for d in data))
    cum_i=cum_s=0; last=(0,0,0)
    for idx,(s,sa,i) in enumerate(data):
        cum_i+=i; cum_s+=sa
        if any(k in s for k in ("BAR.SYNC","RED.","ATOMG","EXIT","RET")) or idx%250==0 or idx==len(data)-1:
            print(f"  {idx:5d} inst+{cum_i-last[1]:8d} samp+{cum_s-last[2]:4d} | {s[:70]}")
            last=(idx,cum_i,cum_s)
    top=sorted(data,key=lambda d:-d[1])[:12]
    print("  top sampled:", [(t[0][:40],t[1]) for t in top])
