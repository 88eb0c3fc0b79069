"""Print the headline metrics of an .ncu-rep (first profiled launch)."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'launch__registers_per_thread ',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum ', 'sm__inst_executed_pipe_fp64.sum ',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max ', 'smsp__average_warp_latency_issue_stalled', 'smsp__average_warps_issue_stalled',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ',
        'sass__inst_executed_local', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__pcsamp_warps_issue_stalled', 'launch__shared_mem_per_block', 'sm__inst_executed_pipe_lsu', 'gpu__dram_throughput.avg.pct']
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for h, u, v in zip(hdr, units, vals):
        if any(k.strip() in h for k in KEYS) and 'pcsamp' not in h:
            print(f'{h:88s} {v:>18s} {u}')
    # stall reasons
    print('--- warp stall (pc sampling) ---')
    st = [(h, float(v.replace(',', ''))) for h, v in zip(hdr, vals) if 'smsp__pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and v.replace(',', '').replace('.', '').isdigit()]
    tot = sum(v for _, v in st) or 1
    for h, v in sorted(st, key=lambda t: -t[1])[:10]:
        print(f'{h.replace("smsp__pcsamp_warps_issue_stalled_", ""):40s} {100*v/tot:6.1f}%')
if __name__ == '__main__':
    main(sys.argv[1])
