#!/bin/bash
# Full harness-metric sweeps with the interleaved-tuned table: fp32 then fp16 accumulate, adaptive auto-tuning rounds.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1k.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $LOG 2>&1
timeout 2400 $DC wallgrid 32 0 1 0.3 -25 0 > gpurun_out/wallgrid_fp32.txt 2>> $LOG
echo "wallgrid32 rc=$?" >> $LOG
timeout 2400 $DC wallgrid 16 0 1 0.3 -25 0 > gpurun_out/wallgrid_fp16.txt 2>> $LOG
echo "wallgrid16 rc=$?" >> $LOG
echo DONE >> $LOG
tail -4 $LOG; tail -1 gpurun_out/wallgrid_fp32.txt; tail -1 gpurun_out/wallgrid_fp16.txt; du -sh gpurun_out
