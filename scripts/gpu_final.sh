#!/bin/bash
# Round-end validation on ONE GPU: smoke, the whole GPU test suite, the default bench line, then the evidence captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 2>&1 | tail -15 | tee gpurun_out/r2_final_pytest.txt
timeout 400 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
grep -E "timed region|e2e done" gpurun_out/r2_bench_final.err | tail -2
NCU_TIMEOUT=200 bash scripts/gpu_evidence.sh 2>&1 | tail -20
