"""Per-kernel counters from an `ncu -i X.ncu-rep --page raw --csv` export, in the layout of profiles/r*_ncu_k2_k3_k4.txt:
    python tools/ncu_summary.py gpurun_out/k234_raw.csv > profiles/rNN_ncu_k2_k3_k4.txt"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_config_size', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum',
        'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio'] + \
       ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % r for r in
        ('barrier', 'long_scoreboard', 'math_pipe_throttle', 'short_scoreboard', 'sleeping', 'wait', 'lg_throttle',
         'mio_throttle', 'no_instruction')]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {}
for i, h in enumerate(hdr):
    for w in WANT:
        if h == w or h.endswith('.' + w):
            col.setdefault(w, []).append(i)
kn = hdr.index('Kernel Name')
for r in rows[2:]:
    print('=====')
    print('Kernel Name |  | ' + r[kn])
    for w in WANT:
        for i in col.get(w, []):
            if r[i] != '':
                print('{} | {} | {}'.format(w, units[i], r[i]))
                break
