#!/bin/bash
# same-box N=1 / N=4 bench lines (per-rank peer-wait counters in the N=4 line)
cd "$(dirname "$0")/.."
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 10 --warmup 3 --no_cpu_baseline --no_secondary > gpurun_out/r2_four_n4.json 2> gpurun_out/r2_four_n4.err
grep "timed region" gpurun_out/r2_four_n4.err | head -1
timeout 150 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_secondary > gpurun_out/r2_four_n1.json 2> gpurun_out/r2_four_n1.err
grep "timed region" gpurun_out/r2_four_n1.err
