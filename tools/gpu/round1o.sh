#!/bin/bash
# 8-warp epilogue: exactness ladder + harness-metric comparison on the shapes measured in round1n.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1o.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
run() { timeout 300 $DC "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
for acc in 32 16; do
  for cfg in 0 1 2 3 4 5 6 7 9 11 12 13 18 19 20; do
    run check $acc $cfg 1024 1536 512
    run check $acc $cfg 200 328 72
  done
  for cfg in 2 1 0 5; do
    run check $acc $cfg 256 512 4096 0 -4
    run check $acc $cfg 200 328 1096 0 5
  done
  run check $acc -1 4096 4096 4096
  run check $acc -1 8192 8192 512
  run check $acc -1 64 4096 64
done
echo "=== wall" >> $LOG
for s in "64 64 64" "512 512 512" "512 512 2048" "1024 1024 1024" "4096 4096 1024" "4096 4096 4096" "8192 8192 512" "2048 11008 4096" "8192 8192 8192" "16384 4096 256" "8192 1024 64" "4096 4096 256"; do
  timeout 300 $DC wall 32 $s 0.3 3 8 >> $LOG 2>&1
done
timeout 300 $DC wall 16 8192 8192 512 0.3 3 8 >> $LOG 2>&1
timeout 300 $DC wall 16 4096 4096 4096 0.3 3 8 >> $LOG 2>&1
echo DONE >> $LOG
grep -c PASS $LOG; grep -E "FAIL|exit|watchdog" $LOG | head -5
grep -E "^WALL" $LOG | sed 's/samples=.*speedup/ speedup/' | cut -d, -f1-9
