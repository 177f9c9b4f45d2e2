#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md for the bench workload (1 GPU).
# Usage (under gpurun):  bash profiles/run_ncu.sh <tag>
# Writes gpurun_out/<tag>_launches.csv (every launch of one whole device-resident step with its
# device time) and gpurun_out/<tag>_spmv.ncu-rep (--set full capture of the dominant kernel).
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
# a step launches ~4.8k kernels; the warm-up step comes first, then the timed one
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4780 -c 4820 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pcg_spmv -s 40 -c 3 \
    -f -o gpurun_out/${TAG}_spmv $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_pcg_update -s 40 -c 2 \
    -f -o gpurun_out/${TAG}_update $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "update capture rc=$?"
tail -2 gpurun_out/${TAG}_plain.log | cut -c1-600
