import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(out.splitlines()))
h=r[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sectors.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for row in r[2:]:
    print(row[h.index('Kernel Name')][:70])
    for k in want:
        if k in h: print('  ',k,row[h.index(k)],r[1][h.index(k)])
    for i,k in enumerate(h):
        if 'issue_stalled' in k and 'per_issue_active' in k and 'not_issued' not in k:
            try:
                v=float(row[i])
                if v>0.3: print('   stall',k.split('issue_stalled_')[1].split('_per_issue')[0],v)
            except: pass
