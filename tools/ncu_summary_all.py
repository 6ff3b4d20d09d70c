import csv,sys,subprocess
out=subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.per_cycle_active','smsp__issue_active.avg.per_cycle_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__warps_eligible.avg.per_cycle_active']
for vals in rows[2:]:
    d=dict(zip(hdr,vals))
    print('==',d.get('Kernel Name'), d.get('launch__grid_size'))
    for k in want: print('  ',k,d.get(k))
    st=[(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),float(v)) for k,v in d.items() if 'issue_stalled' in k and 'per_issue_active' in k and v and float(v)>0.05]
    print('  stalls',sorted(st,key=lambda kv:-kv[1]))
