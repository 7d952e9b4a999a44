#!/bin/bash
# bring-up of the stem wgrad kernel: which descriptor / TMEM-lane variant reproduces the reference?
cd "$(dirname "$0")/.."
best=""
for v in 0 1 2 3; do
  SIMCLR_STEM_WG_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_tc.py -q -k "wgrad_bf16" 2>&1 | tail -1 > gpurun_out/r2_stem_wg_v$v.txt
  echo "variant $v: $(cat gpurun_out/r2_stem_wg_v$v.txt)"
  if grep -q "passed" gpurun_out/r2_stem_wg_v$v.txt && ! grep -q "failed" gpurun_out/r2_stem_wg_v$v.txt && [ -z "$best" ]; then best=$v; fi
done
echo "best variant: ${best:-none}"
if [ -z "$best" ]; then export SIMCLR_TC_STEM_WGRAD=0; else export SIMCLR_STEM_WG_VARIANT=$best; fi
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -4
timeout 200 python scripts/kernel_times.py > gpurun_out/r2_kt1f.log 2>&1
head -22 gpurun_out/kernel_times.txt | cut -c1-50,105-150
timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_secondary > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err
grep "timed region" gpurun_out/r2_bench_j.err
