#!/bin/bash
# Round evidence on ONE GPU (run under gpurun): ncu launch list of a step, ncu --set full of the dominant kernels,
# then the wider configs of BASELINE.json (cfg4: ResNet-50 2x, cfg5: ResNet-152 2x SK) at the largest per-GPU batch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${NCU_TIMEOUT:-240}
timeout $T ncu --nvtx --nvtx-include "profiled/" --clock-control none --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --log-file gpurun_out/r2_launches.csv python scripts/profile_step.py --batch 512 > gpurun_out/r2_profile_step.log 2>&1
echo "launch list exit $?"
cap() {  # name, kernel regex, layer index, kind
  timeout 150 ncu --set full --clock-control none --import-source on -k "regex:$2" -c 1 -o gpurun_out/$1 \
    python scripts/layer_bench.py --views 256 --layers $3 --only $4 --fused_stats --reps 1 > gpurun_out/$1.log 2>&1
  echo "ncu $1 exit $?"
}
cap r2_full_halo_fprop halo3x3_kernel 3 fprop
cap r2_full_halo_wgrad halo3x3_wgrad 3 wgrad
cap r2_full_igemm_3x3 igemm_kernel 16 fprop
cap r2_full_igemm_1x1 igemm_kernel 1 fprop
cap r2_full_wgrad wgrad_kernel 16 wgrad
cap r2_full_stem_fprop stem7x7_kernel 0 fprop
cap r2_full_stem_wgrad stem7x7_wgrad 0 wgrad
timeout 200 ncu --set full --clock-control none -k "regex:bn_apply_kernel|bn_reduce_kernel|bn_bwd_apply_kernel" -c 3 -o gpurun_out/r2_full_bn \
  python scripts/bn_bench.py > gpurun_out/r2_full_bn.log 2>&1
echo "ncu bn exit $?"
for B in 512 256; do
  timeout 300 python bench.py --width_multiplier 2 --batch $B --learning_rate 0.1 --learning_rate_scaling sqrt --steps 5 --warmup 3 \
    --no_cpu_baseline --no_secondary > gpurun_out/r2_bench_cfg4_b$B.json 2> gpurun_out/r2_bench_cfg4_b$B.err
  if [ -s gpurun_out/r2_bench_cfg4_b$B.json ]; then echo "cfg4 batch $B ok"; break; fi
  tail -n 2 gpurun_out/r2_bench_cfg4_b$B.err | cut -c1-200
done
for B in 256 128 64; do
  timeout 400 python bench.py --resnet_depth 152 --width_multiplier 2 --sk_ratio 0.0625 --num_proj_layers 3 --batch $B --steps 3 --warmup 3 \
    --no_cpu_baseline --no_secondary > gpurun_out/r2_bench_cfg5_b$B.json 2> gpurun_out/r2_bench_cfg5_b$B.err
  if [ -s gpurun_out/r2_bench_cfg5_b$B.json ]; then echo "cfg5 batch $B ok"; break; fi
  tail -n 2 gpurun_out/r2_bench_cfg5_b$B.err | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
