#!/bin/bash
# First-light script for the B200 box: exactness of every configuration on a ladder of shapes, then
# event timings. One process per case so a trap in one configuration cannot poison the rest.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run() { timeout 120 $DC "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
# ladder: single tile/single k-block -> multi k-block -> multi tile -> edges -> big
for acc in 32 16; do
  for cfg in 2 1 0 5 4 3 6; do
    run check $acc $cfg 256 256 64
    run check $acc $cfg 256 256 512
    run check $acc $cfg 1024 1536 1024
    run check $acc $cfg 200 328 72
    run check $acc $cfg 1000 1000 1000
  done
done
for acc in 32 16; do
  run check $acc -1 64 4096 64
  run check $acc -1 4096 4096 4096
  run check $acc -1 2048 11008 4096
  run check $acc -1 8192 8192 8192
  run check $acc -1 64 64 16384
done
for cfg in 0 1 3 4 6; do
  run time 32 $cfg 4096 4096 4096 20
done
run time 32 3 8192 8192 8192 10
run time 16 3 8192 8192 8192 10
run time 32 4 8192 8192 8192 10
run time 32 0 8192 8192 8192 10
run time 32 -1 2048 11008 4096 20
run time 32 -1 64 4096 64 50
echo DONE >> $LOG
tail -5 $LOG
