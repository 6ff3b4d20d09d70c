import csv,sys,subprocess
out=subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; vals=rows[2]
d=dict(zip(hdr,vals))
want=['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.per_cycle_active','smsp__issue_active.avg.per_cycle_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum']
for k in want:
    print(k, d.get(k))
for k,v in d.items():
    if 'issue_stalled' in k and 'per_issue_active' in k and float(v or 0)>0.05: print(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), v)
