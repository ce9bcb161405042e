#!/bin/bash
# Re-tune fp16 accumulate on the new kernel, then the full harness-metric sweep (fp32 accumulate) on one GPU.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1g.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
timeout 1200 $DC grid 16 0 1 2.0 > gpurun_out/grid_fp16.csv 2>> $LOG
echo "grid16 rc=$?" >> $LOG
for s in "4096 4096 4096" "8192 8192 8192" "2048 11008 4096" "64 4096 64"; do
  timeout 300 $DC check 32 -1 $s >> $LOG 2>&1
done
timeout 2400 $DC wallgrid 32 0 1 0.3 5 15 > gpurun_out/wallgrid_fp32.txt 2>> $LOG
echo "wallgrid rc=$?" >> $LOG
echo DONE >> $LOG
tail -3 $LOG; tail -2 gpurun_out/wallgrid_fp32.txt; du -sh gpurun_out
