import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
keys=['Kernel Name','gpu__time_duration.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed_pipe_fp64.sum','sm__cycles_elapsed.max','smsp__average_warp_latency_issue_stalled','smsp__average_warps_issue_stalled','local','lmem','smsp__pcsamp_warps_issue_stalled','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_shared','sm__inst_executed_pipe_lsu']
for r in rows[2:]:
    print('----')
    for i,h in enumerate(hdr):
        if any(h==k or (k in h and ('stalled' in k or k in ('local','lmem','smsp__inst_executed_op_shared','sm__inst_executed_pipe_lsu'))) for k in keys):
            if r[i] not in ('0','','n/a'): print(h, units[i], r[i])
