#!/bin/bash
# CTA-pair multicast clusters: address probe, exactness, then the harness metric with forced configurations on
# the compute-bound BASELINE shapes.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1h.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
run() { timeout 300 $DC "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
run probe
for acc in 32 16; do
  for cfg in 20 21 22 23 24 25; do
    run check $acc $cfg 512 512 128
    run check $acc $cfg 1024 1536 1024
    run check $acc $cfg 200 328 72
    run check $acc $cfg 4096 4096 1024
  done
  for cfg in 3 4 9 11 13; do run check $acc $cfg 1024 1536 1024; done
done
echo "=== forced wall" >> $LOG
for s in "4096 4096 4096" "8192 8192 8192" "2048 11008 4096" "8192 4096 2048"; do
  for f in "3,8,1" "3,4,1" "4,1,1" "6,4,1" "20,8,1" "20,4,1" "21,8,1" "21,4,1" "22,4,1" "23,4,1" "24,4,1" "25,4,1" "19,8,1" "18,8,1"; do
    echo "force $f" >> $LOG
    B200_HGEMM_FORCE=$f timeout 300 $DC wall 32 $s 0.4 3 8 >> $LOG 2>&1
  done
done
echo DONE >> $LOG
tail -3 $LOG; du -sh gpurun_out
