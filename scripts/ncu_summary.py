"""Condenses an `ncu --set full` report into the handful of numbers the roofline argument uses.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [title] > profiles/rNN_ncu_x.txt"""
import csv, subprocess, sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = [
    ('gpu__time_duration.sum', 'kernel time'),
    ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active (% of elapsed)'),
    ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'shared-memory reads by the tensor core (% of peak)'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'shared-memory LSU wavefronts (% of peak)'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput (% of peak)'),
    ('dram__bytes_read.sum', 'DRAM bytes read'), ('dram__bytes_write.sum', 'DRAM bytes written'),
    ('lts__t_bytes.sum', 'L2 bytes'), ('lts__t_sector_hit_rate.pct', 'L2 hit rate'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput (% of peak)'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy (%)'),
    ('launch__registers_per_thread', 'registers / thread'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'),
    ('smsp__inst_executed.sum', 'warp instructions'),
]
print('# %s' % title)
print('# source: ncu --set full --clock-control none (%s); per-launch values, cold caches' % rep.split('/')[-1])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print('\nkernel: %s' % d.get('Kernel Name', '?')[:150])
    for key, label in WANT:
        cands = [h for h in hdr if h.endswith(key)]
        if cands:
            print('  %-55s %s %s' % (label, d[cands[0]], u[cands[0]]))
