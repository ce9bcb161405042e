#!/bin/bash
# Third pass: split-K exactness, the full-grid configuration sweep (isolated launches), the harness metric on
# representative shapes, then the GPU test-suite and a first bench line.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1c.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
run() { timeout 300 $DC "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
for acc in 32 16; do
  for cfg in 2 1 0 5; do
    run check $acc $cfg 64 64 16384 0 32
    run check $acc $cfg 256 512 12288 0 8
    run check $acc $cfg 200 328 1096 0 5
    run check $acc $cfg 128 128 2048 0 64
  done
  run check $acc -1 64 64 16384
  run check $acc -1 256 512 12288
done
echo "=== grid" >> $LOG
timeout 900 $DC grid 32 0 1 2.0 > gpurun_out/grid_fp32.csv 2>> $LOG
echo "grid rc=$?" >> $LOG
echo "=== wall" >> $LOG
for s in "64 4096 64" "64 64 64" "512 512 512" "1024 1024 1024" "2048 2048 2048" "4096 4096 4096" "8192 8192 8192" \
         "2048 11008 4096" "64 64 16384" "256 512 12288" "128 16384 1024" "16384 4096 256" "8192 1024 64" "16384 16384 16384"; do
  run wall 32 $s 1.0 10 30
done
run wall 16 8192 8192 8192 1.0 10 30
echo "=== pytest" >> $LOG
timeout 900 python -m pytest tests -m gpu -x -q >> $LOG 2>&1
echo "pytest rc=$?" >> $LOG
echo "=== bench" >> $LOG
timeout 300 python bench.py --steps 500 --warmup 10 >> $LOG 2>&1
echo "bench rc=$?" >> $LOG
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 >> $LOG 2>&1
echo DONE >> $LOG
tail -5 $LOG
