#!/bin/bash
# Multicast-cluster / BN=32 / cluster split-K bring-up, then re-tune.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1f.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
run() { timeout 300 $DC "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
for acc in 32 16; do
  for cfg in 7 8 9 10 11 12 13 14 15 16 17 18 19; do
    run check $acc $cfg 256 512 128
    run check $acc $cfg 1024 1536 1024
    run check $acc $cfg 200 328 72
    run check $acc $cfg 128 4096 4096
  done
  for cfg in 0 1 2 3 4 5 6; do
    run check $acc $cfg 1024 1536 1024
    run check $acc $cfg 200 328 72
  done
  for cfg in 2 1 0; do
    run check $acc $cfg 256 512 4096 0 -2
    run check $acc $cfg 200 328 1096 0 -4
    run check $acc $cfg 128 128 8192 0 -8
    run check $acc $cfg 64 64 16384 0 32
  done
done
echo "=== sweeps" >> $LOG
for s in "4096 4096 4096 10" "1024 1024 1024 30" "128 4096 4096 20" "512 512 2048 30" "1024 1024 8192 20"; do
  run sweep 32 $s
done
echo "=== grid" >> $LOG
timeout 1200 $DC grid 32 0 1 2.0 > gpurun_out/grid_fp32.csv 2>> $LOG
echo "grid32 rc=$?" >> $LOG
echo DONE >> $LOG
tail -3 $LOG
du -sh gpurun_out
