#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers we track per kernel variant."""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'launch__registers_per_thread','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'sm__cycles_elapsed.avg','sm__cycles_active.avg','smsp__warps_eligible.avg.per_cycle_active','smsp__average_warp_latency_per_inst_issued.ratio',
 'lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__cycles_active.avg.pct_of_peak_sustained_elapsed']
STALL = 'smsp__average_warps_issue_stalled_'
def main(path):
    raw = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')][:50]
        print('==', path, name)
        for k in KEYS:
            if k in hdr: print(f'  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
        st = [(float(r[i]),h[len(STALL):-len("_per_issue_active.ratio")]) for i,h in enumerate(hdr) if h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and r[i]]
        print('  stalls (warps per issue):', ', '.join(f'{n}={v:.2f}' for v,n in sorted(st,reverse=True)[:8]))
        ni = [(float(r[i]),h) for i,h in enumerate(hdr) if h.startswith('smsp__pcsamp_warps_issue_stalled_') and h.endswith('_not_issued') and r[i]]
        tot = sum(v for v,_ in ni) or 1
        print('  not-issued samples:', ', '.join(f'{h[len("smsp__pcsamp_warps_issue_stalled_"):-len("_not_issued")]}={100*v/tot:.0f}%' for v,h in sorted(ni,reverse=True)[:6]))
if __name__ == '__main__':
    for p in sys.argv[1:]: main(p)
