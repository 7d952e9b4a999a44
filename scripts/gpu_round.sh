#!/bin/bash
# One GPU-box visit: BN/step parity for the current build, CUPTI kernel split, ncu of chosen layers.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "=== bn + step tests ==="
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -q -m gpu -x -k "batchnorm or step or sk or se_layer" 2>&1 | tail -5
echo "=== kernel times ==="
timeout 300 python scripts/kernel_times.py > gpurun_out/kt.log 2>&1; echo "exit $?"
echo "=== ncu layers ${NCU_LAYERS:-3} ==="
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'igemm_kernel|wgrad_kernel' -c ${NCU_COUNT:-6} \
  -f -o gpurun_out/layers python scripts/layer_bench.py --views 256 --layers ${NCU_LAYERS:-3} --reps 1 --fused_stats > gpurun_out/ncu_layers.log 2>&1
echo "exit $?"; ls -la gpurun_out/*.ncu-rep
} > gpurun_out/round.txt 2>&1
tail -30 gpurun_out/round.txt
