#!/bin/bash
# Fourth pass: re-verify after the MMA-loop / producer / split-K-reduction rewrite, then tune: full-grid sweeps
# for both accumulators (isolated launches), plus the harness metric on the BASELINE shapes.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1d.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
run() { timeout 300 $DC "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
for acc in 32 16; do
  for cfg in 0 1 2 3 4 5 6; do
    run check $acc $cfg 256 256 64
    run check $acc $cfg 1024 1536 1024
    run check $acc $cfg 200 328 72
  done
  for cfg in 2 1 0 5; do
    run check $acc $cfg 64 64 16384 0 32
    run check $acc $cfg 256 512 12288 0 8
    run check $acc $cfg 200 328 1096 0 5
  done
  run check $acc -1 4096 4096 4096
  run check $acc -1 8192 8192 8192
done
echo "=== pytest" >> $LOG
timeout 900 python -m pytest tests -m gpu -x -q >> $LOG 2>&1
echo "pytest rc=$?" >> $LOG
echo "=== sweeps" >> $LOG
run sweep 32 4096 4096 4096 20
run sweep 32 1024 1024 1024 50
echo "=== grid" >> $LOG
timeout 900 $DC grid 32 0 1 2.0 > gpurun_out/grid_fp32.csv 2>> $LOG
echo "grid32 rc=$?" >> $LOG
timeout 900 $DC grid 16 0 1 2.0 > gpurun_out/grid_fp16.csv 2>> $LOG
echo "grid16 rc=$?" >> $LOG
echo DONE >> $LOG
tail -5 $LOG
