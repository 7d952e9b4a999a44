#!/bin/bash
# 2-GPU validation: multi-rank parity tests, same-box N=1 / N=2 bench lines with the per-rank peer-wait counters
cd "$(dirname "$0")/.."
timeout 120 python -m pytest tests/test_gpu_tc.py -q -k "wgrad_bf16" 2>&1 | tail -1
timeout 120 python -m pytest tests/test_gpu_kernels.py -q -k "blur or pool" 2>&1 | tail -1
timeout 500 python -m pytest tests/test_gpu_multi.py -x -q -k "bf16-peer-graph or fp32-syncbn-peer" 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no_cpu_baseline --no_secondary > gpurun_out/r2_two_n2.json 2> gpurun_out/r2_two_n2.err
grep "timed region" gpurun_out/r2_two_n2.err | head -2
timeout 200 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_secondary > gpurun_out/r2_two_n1.json 2> gpurun_out/r2_two_n1.err
grep "timed region" gpurun_out/r2_two_n1.err
timeout 150 python scripts/kernel_times.py > gpurun_out/r2_kt1g.log 2>&1
grep -E "total|stem7x7|blur|wgrad_kernel  " gpurun_out/kernel_times.txt | cut -c1-50,105-150 | head -8
