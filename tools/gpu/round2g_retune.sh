#!/bin/bash
# Second tuning pass (one B200, another box than the first pass): fp32 accumulation for every shape >= 1 GFLOP with the
# portable alternatives always in the shortlist and stream-K wherever the tile count is not a wave multiple; then the
# first wall-metric pass for fp16 accumulation (its table was still the round-1 event-time one).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round2g.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
# the largest problems (>= 200 GFLOP): every pair configuration incl. the 512x256 tiles, all rasterisation widths, no shortlist
B200_TUNE_CFGS=3,6,4,26,27,28,0 B200_TUNE_KEEP_ALL=1 timeout 420 $DC grid 32 0 1 3.0 200 1e30 wall > gpurun_out/grid_fp32_wall_r2_huge.csv 2>> $LOG; echo "fp32 huge rc=$?" >> $LOG
timeout 600 $DC grid 32 0 1 3.0 1 1e30 wall > gpurun_out/grid_fp32_wall_r2_pass2.csv 2>> $LOG; echo "fp32 pass2 rc=$?" >> $LOG
B200_TUNE_CFGS=3,6,4,26,27,28,0 B200_TUNE_KEEP_ALL=1 timeout 420 $DC grid 16 0 1 3.0 200 1e30 wall > gpurun_out/grid_fp16_wall_r2_huge.csv 2>> $LOG; echo "fp16 huge rc=$?" >> $LOG
timeout 600 $DC grid 16 0 1 3.0 20 1e30 wall > gpurun_out/grid_fp16_wall_r2_big.csv 2>> $LOG; echo "fp16 big rc=$?" >> $LOG
timeout 420 $DC grid 16 0 1 3.0 0 20 wall > gpurun_out/grid_fp16_wall_r2_small.csv 2>> $LOG; echo "fp16 small rc=$?" >> $LOG
wc -l gpurun_out/grid_fp*_wall_r2_*.csv >> $LOG
tail -8 $LOG
