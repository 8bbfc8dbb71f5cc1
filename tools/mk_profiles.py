import csv, io, json, subprocess, hashlib, sys
def raw(rep):
    out = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
    r=list(csv.reader(io.StringIO(out))); h=r[0]; u=r[1]; v=r[-1]
    return {k:(x,un) for k,x,un in zip(h,v,u)}
KEYS=["Kernel Name","gpu__time_duration.sum","launch__grid_size","launch__block_size","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","sm__warps_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","dram__bytes_read.sum.per_second","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","lts__t_sector_hit_rate.pct","l1tex__t_sector_hit_rate.pct","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio","smsp__inst_executed_op_shared_ld.sum" ]
def table(rep, title):
    m=raw(rep)
    lines=["### %s" % title, "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m: lines.append("| %s | %s | %s |" % (k, m[k][0][:110], m[k][1]))
    return "\n".join(lines), m
def sass_mix(rep):
    out = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(io.StringIO(out)))
    h=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
    hdr=rows[h]; ii=hdr.index("Instructions Executed"); ai=hdr.index("Source")
    mix={}
    tot=0
    for r in rows[h+1:]:
        if len(r)!=len(hdr): continue
        c=float(r[ii] or 0); tot+=c
        op=r[ai].strip().split()[0] if r[ai].strip() else "?"
        if op.startswith("@"): op=r[ai].strip().split()[1]
        op=op.split(".")[0]
        mix[op]=mix.get(op,0)+c
    return tot, mix
out=[]
out.append("# Round 2: ncu --set full captures (B200, --clock-control none), summarised by tools/mk_profiles.py from gpurun_out/*.ncu-rep\n")
t,m=table("gpurun_out/r2_star_2p28.ncu-rep","star_kernel<kGlobal, bulk staging, funnel> on the bench graph, n = 2^28, degree 28, random f (tools/star_variants.py --one 28)")
out.append(t)
rd=float(m["dram__bytes_read.sum"][0]); wr=float(m["dram__bytes_write.sum"][0])
ru=m["dram__bytes_read.sum"][1]; wu=m["dram__bytes_write.sum"][1]
scale={"Gbyte":1e9,"Mbyte":1e6,"Kbyte":1e3,"byte":1,"Tbyte":1e12}
rdb=rd*scale[ru]; wrb=wr*scale[wu]
alg=4*28*2**28+17*2**28
out.append("\nDRAM traffic %.2f GB read + %.2f GB written = %.3f x the algorithmic %.2f GB (round 1: 1.38 x)\n" % (rdb/1e9, wrb/1e9, (rdb+wrb)/alg, alg/1e9))
sha=hashlib.sha256(open("shgo_b200/csrc/star.cu","rb").read()).hexdigest()[:16]
json.dump({"log2n":28,"n_gpus":1,"dram_bytes_read":rdb,"dram_bytes_write":wrb,"kernel":m["Kernel Name"][0],"star_cu_sha16":sha,
           "duration_ms_under_ncu":float(m["gpu__time_duration.sum"][0]),
           "source":"profiles/r2_kernels.md (ncu --set full of tools/star_variants.py --one 28, round 2)"}, open("profiles/r2_star_traffic.json","w"), indent=1)
for fn,title,npts,d in (("rastrigin","Rastrigin 6-D, Sobol n = 2^26 (config 2's objective)",2**26,6),("ackley","Ackley 10-D, Sobol n = 2^26 (config 3's objective)",2**26,10),("cattle","cattle-feed 4-D with g1, g2, Sobol n = 2^26 (config 4)",2**26,4),("st","Styblinski-Tang 8-D, Sobol n = 2^26 (config 5's objective)",2**26,8)):
    rep="gpurun_out/r2_eval_%s.ncu-rep" % fn
    t,m=table(rep,"eval_kernel: "+title); out.append("\n"+t)
    tot,mix=sass_mix(rep)
    fp=["DFMA","DADD","DMUL","DSETP","MUFU"]
    per=lambda op: mix.get(op,0)*32/npts
    out.append("\nSASS per point (warp instructions x 32 / points): total %.1f; " % (tot*32/npts) + ", ".join("%s %.1f" % (op, per(op)) for op in fp) + "; per point-coordinate: FP64 (DFMA+DADD+DMUL) %.1f" % ((per("DFMA")+per("DADD")+per("DMUL"))/d))
open("profiles/r2_kernels.md","w").write("\n".join(out)+"\n")
print("\n".join(out)[:6000])
