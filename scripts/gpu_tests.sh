#!/bin/bash
# Runs the GPU parity suites in separate processes (a trapped tcgen05 kernel
# kills its CUDA context; isolation keeps the other suites' results).
# Usage (on the GPU box, from the repo root): bash scripts/gpu_tests.sh [suite ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
SUITES="${@:-kernels tc_fprop tc_dgrad tc_wgrad tc_tf32 tc_large step}"
for s in $SUITES; do
  case $s in
    kernels)  cmd="tests/test_gpu_kernels.py" ;;
    tc_fprop) cmd="tests/test_gpu_tc.py -k fprop" ;;
    tc_dgrad) cmd="tests/test_gpu_tc.py -k dgrad" ;;
    tc_wgrad) cmd="tests/test_gpu_tc.py -k wgrad" ;;
    tc_tf32)  cmd="tests/test_gpu_tc.py -k tf32" ;;
    tc_large) cmd="tests/test_gpu_tc.py -k large" ;;
    step)     cmd="tests/test_gpu_step.py -s" ;;
    *) echo "unknown suite $s"; continue ;;
  esac
  echo "=== $s ===" | tee -a gpurun_out/summary.txt
  timeout 900 python -m pytest $cmd -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/test_$s.log 2>&1
  echo "exit $?" >> gpurun_out/test_$s.log
  tail -n 4 gpurun_out/test_$s.log | tee -a gpurun_out/summary.txt
done
