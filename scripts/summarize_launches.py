"""Summarise an ncu launch list (csv, one row per launch x metric) of one profiled step.

  ncu --nvtx --nvtx-include "profiled/" --clock-control none --csv \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --log-file gpurun_out/launches.csv python scripts/profile_step.py --batch 512
  python scripts/summarize_launches.py gpurun_out/launches.csv profiles/r01_launch_shares.txt profiles/traffic.json

Writes the per-kernel time shares (text) and the DRAM traffic per launch of the tcgen05 conv
kernels (json, read by bench.py for `roofline.traffic`)."""
import collections, csv, json, re, sys

src, out_txt, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src, newline='') as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
iname, imet, ival, iunit, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('ID')
SCALE = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6, 'second': 1e3,
         'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
per = collections.defaultdict(lambda: collections.defaultdict(float))
launches = collections.defaultdict(set)
for r in rd:
    if len(r) <= ival:
        continue
    name = re.sub(r'^void\s+', '', r[iname])
    name = re.sub(r'\(.*$', '', name).replace('simclr::', '').replace('<unnamed>::', '').replace('tc::', '')
    short = re.sub(r'<.*$', '', name)
    v = float(r[ival].replace(',', '')) * SCALE.get(r[iunit], 1.0)
    per[short][r[imet]] += v
    launches[short].add(r[iid])
tot = sum(d['gpu__time_duration.sum'] for d in per.values())
n = sum(len(s) for s in launches.values())
with open(out_txt, 'w') as fh:
    fh.write('ncu launch list of ONE eager pretrain step (ResNet-50 1x, 512 samples = 1024 views, 224x224, bf16 tcgen05 path)\n')
    fh.write('per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n')
    fh.write('total kernel time %.2f ms over %d launches\n\n' % (tot, n))
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
        t = d['gpu__time_duration.sum']
        gb = (d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)) / 1e9
        fh.write('%-44s launches=%4d %10.3f ms %6.1f%%   dram %8.2f GB\n' % (k[:44], len(launches[k]), t, 100 * t / tot, gb))
conv = [k for k in per if k.startswith(('igemm_kernel', 'wgrad_kernel', 'halo3x3_kernel', 'halo3x3_wgrad_kernel', 'stem7x7_kernel', 'stem7x7_wgrad_kernel'))]
cl = sum(len(launches[k]) for k in conv)
cb = sum(per[k].get('dram__bytes_read.sum', 0.0) + per[k].get('dram__bytes_write.sum', 0.0) for k in conv)
ct = sum(per[k]['gpu__time_duration.sum'] for k in conv)
json.dump({'kernel': 'igemm_kernel/wgrad_kernel/halo3x3_kernel/halo3x3_wgrad_kernel/stem7x7_kernel/stem7x7_wgrad_kernel', 'launches': cl, 'dram_bytes_per_launch': cb / max(cl, 1),
           'dram_bytes_per_step': cb, 'share_of_step_kernel_time': ct / tot,
           'workload': 'ResNet-50 1x, batch 512, 224x224, bf16', 'source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum'},
          open(out_json, 'w'), indent=1)
print(open(out_txt).read())
