#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of one profiled eager step, summarised into profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --nvtx --nvtx-include "profiled/" --clock-control none --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --log-file gpurun_out/launches.csv python scripts/profile_step.py --batch 512 > gpurun_out/profile_step.log 2>&1
echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv gpurun_out/launch_shares.txt gpurun_out/traffic.json | head -20
