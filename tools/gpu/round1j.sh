#!/bin/bash
# Re-tune both accumulators with interleaved candidate timing (same clock / L2 state for every candidate).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1j.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
timeout 1500 $DC grid 32 0 1 2.0 > gpurun_out/grid_fp32.csv 2>> $LOG
echo "grid32 rc=$?" >> $LOG
timeout 1500 $DC grid 16 0 1 2.0 > gpurun_out/grid_fp16.csv 2>> $LOG
echo "grid16 rc=$?" >> $LOG
echo DONE >> $LOG
tail -3 $LOG; wc -l gpurun_out/grid_fp*.csv; du -sh gpurun_out
