#!/bin/bash
# L2 cache hints + serpentine rasterisation: exactness spot checks and harness-metric A/B.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1i.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
run() { timeout 300 $DC "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
for acc in 32 16; do
  for cfg in 0 2 3 4 9 20; do run check $acc $cfg 2048 3072 512; run check $acc $cfg 200 328 72; done
  run check $acc -1 8192 128 16384
  run check $acc -1 128 8192 16384
  run check $acc -1 4096 4096 4096
done
echo "=== hints A/B" >> $LOG
for s in "8192 128 16384" "128 8192 16384" "12288 256 16384" "64 8192 12288" "256 4096 8192" "16384 64 4096" "128 16384 8192"; do
  echo "hints on" >> $LOG;  timeout 300 $DC wall 32 $s 0.4 3 8 >> $LOG 2>&1
  echo "hints off" >> $LOG; B200_HGEMM_NO_CACHE_HINTS=1 timeout 300 $DC wall 32 $s 0.4 3 8 >> $LOG 2>&1
done
echo "=== forced skinny" >> $LOG
for s in "8192 128 16384" "128 8192 16384"; do
  for f in "2,0,1" "1,0,1" "1,0,-2" "12,0,1" "4,0,1" "0,0,1" "15,0,1" "7,0,1" "16,0,1" "8,0,1"; do
    echo "force $f" >> $LOG
    B200_HGEMM_FORCE=$f timeout 300 $DC wall 32 $s 0.3 3 8 >> $LOG 2>&1
  done
done
echo "=== big" >> $LOG
for s in "8192 8192 8192:3,8,1" "8192 8192 8192:3,4,1" "4096 4096 4096:4,1,1" "4096 4096 4096:6,4,1" "4096 4096 4096:3,8,1" "16384 16384 4096:3,8,1" "2048 11008 4096:3,8,1"; do
  echo "force ${s#*:}" >> $LOG
  B200_HGEMM_FORCE=${s#*:} timeout 300 $DC wall 32 ${s%%:*} 0.4 3 8 >> $LOG 2>&1
done
echo DONE >> $LOG
tail -3 $LOG; du -sh gpurun_out
