#!/bin/bash
# Second bring-up pass: full exactness ladder, configuration sweeps on the headline shapes, and
# one ncu --set full capture of our kernel next to cuBLAS on 8192^3.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1b.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
run() { timeout 180 $DC "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
for acc in 32 16; do
  for cfg in 0 1 2 3 4 5 6; do
    run check $acc $cfg 256 256 64
    run check $acc $cfg 256 256 512
    run check $acc $cfg 1024 1536 1024
    run check $acc $cfg 200 328 72
    run check $acc $cfg 1000 1000 1000
  done
  run check $acc -1 64 4096 64
  run check $acc -1 4096 4096 4096
  run check $acc -1 2048 11008 4096
  run check $acc -1 8192 8192 8192
  run check $acc -1 64 64 16384
  run check $acc -1 16384 64 64
done
echo "=== sweeps" >> $LOG
run sweep 32 4096 4096 4096 20
run sweep 32 8192 8192 8192 10
run sweep 16 8192 8192 8192 10
run sweep 32 2048 11008 4096 20
run sweep 32 2048 2048 2048 30
run sweep 32 1024 1024 1024 50
run sweep 32 16384 16384 16384 3
run sweep 32 512 512 512 50
run sweep 32 128 16384 1024 50
run sweep 32 16384 4096 256 30
run sweep 32 256 512 12288 50
run sweep 32 64 64 64 100
run sweep 32 64 64 16384 50
run sweep 32 8192 1024 64 50
echo "=== ncu" >> $LOG
timeout 300 ncu --set full --clock-control none --import-source on -s 2 -c 12 -f -o gpurun_out/prof_8192_cfg3 \
   $DC time 32 3 8192 8192 8192 1 >> $LOG 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_4096.csv \
   $DC time 32 -1 4096 4096 4096 5 >> $LOG 2>&1
echo DONE >> $LOG
tail -3 $LOG
