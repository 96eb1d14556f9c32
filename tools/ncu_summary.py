"""Summarise an .ncu-rep (raw page) into JSON + print the key roofline counters.  Runs where there is no GPU."""
import csv, io, json, subprocess, sys
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg.per_second', 'smsp__warps_eligible.avg.per_cycle_active', 'lts__t_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg']
def main(path, out=None):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {'kernel': vals[hdr.index('Kernel Name')]}
        for i, h in enumerate(hdr):
            if h in KEEP or ('issue_stalled' in h and h.endswith('.ratio')):
                d[h] = {'value': vals[i], 'unit': units[i]}
        res.append(d)
    if out:
        json.dump(res, open(out, 'w'), indent=1, sort_keys=True)
    for d in res:
        print(d['kernel'][:110])
        for k in KEEP:
            if k in d: print(f"   {k:78s} {d[k]['value']:>16s} {d[k]['unit']}")
        st = sorted(((float(v['value'].replace(',', '')), k) for k, v in d.items() if 'issue_stalled' in k), reverse=True)[:6]
        for v, k in st: print(f"   stall {k.split('issue_stalled_')[1].split('_per')[0]:28s} {v:8.3f}")
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
