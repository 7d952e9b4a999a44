#!/bin/bash
# First-pass benchmark on the GPU box: bench line + (optional) ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -n 15 gpurun_out/bench.err
