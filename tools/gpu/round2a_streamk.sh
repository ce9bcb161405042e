#!/bin/bash
# Round 2, first GPU call: validate the round-2 candidate branch (stream-K: exact in its first light, owner path
# reworked since; config 26 = 512x256 pair tile: never run on a GPU before this script).
#   1. regression ladder of the plain schedule (the loops were rewritten around WorkIter);
#   2. stream-K exactness: every plain config x {tail, tail+wave} x shapes with 1..13 contributors per tile;
#   3. A/B timings on the wave-quantised shapes that motivated it (tile count = 0.865 of a wave multiple);
#   4. the GPU test-suite.
# One process per case so that a trap in one configuration cannot poison the rest.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round2a.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run() { timeout 120 $DC "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
echo "== 1. plain regression" >> $LOG
for acc in 32 16; do
  for cfg in 0 1 2 3 4 5 6 12 10 18 19 20 26 27 28 29 30; do
    run check $acc $cfg 1024 1536 1024
    run check $acc $cfg 1000 1000 1000
  done
  run check $acc -1 4096 4096 4096
  run check $acc -1 64 64 16384
  run check $acc 1 256 512 2048 0 8
  run check $acc 1 256 512 2048 0 -4
done
echo "== 2. stream-K exactness" >> $LOG
for acc in 32 16; do
  for cfg in 0 1 2 3 4 5 6; do
    for sk in 100 101; do
      run check $acc $cfg 512 768 4096 0 $sk
      run check $acc $cfg 1000 1224 2048 0 $sk
      run check $acc $cfg 512 8192 8192 0 $sk
      run check $acc $cfg 4096 4096 4096 8 $sk
    done
  done
done
grep -c PASS $LOG >> $LOG; grep -c "FAIL\|exit" $LOG >> $LOG
echo "== 3. A/B timings (same process order: plain, tail, tail+wave)" >> $LOG
ab() { for sk in 1 100 101; do run time "$1" "$2" "$3" "$4" "$5" 30 "$6" $sk; done; }
ab 32 3 512 8192 8192 0
ab 32 3 1024 4096 8192 0
ab 32 4 512 4096 8192 0
ab 32 3 1024 8192 8192 0
ab 32 6 4096 4096 4096 8
ab 32 3 4096 4096 4096 8
ab 32 3 2048 11008 4096 8
ab 32 4 12288 2048 4096 8
ab 32 0 1024 1024 4096 0
ab 32 4 1024 1024 4096 0
ab 16 3 512 8192 8192 0
ab 16 3 4096 4096 4096 8
echo "== 3b. config 26 (512x256 pair tile, one accumulator stage) against config 3 on large shapes" >> $LOG
for acc in 32 16; do run check $acc 26 4096 4096 4096 8; run check $acc 26 1000 1224 2048; run check $acc 26 8192 8192 8192 8; done
for shape in "8192 8192 8192" "16384 16384 16384" "16384 16384 4096" "4096 12288 16384" "4096 4096 4096" "8192 8192 2048"; do
  for cfg in 3 26 27 28; do run time 32 $cfg $shape 10 8 1; done
done
echo "== 3c. sustain (2 s): burst vs power-capped, headline shapes, candidate configs" >> $LOG
for spec in "6 4096 4096 4096 8 1" "6 4096 4096 4096 8 100" "3 4096 4096 4096 8 1" "3 4096 4096 4096 8 100" "20 4096 4096 4096 8 1" "21 4096 4096 4096 8 1" "29 4096 4096 4096 8 1" "26 4096 4096 4096 8 1" \
            "3 8192 8192 8192 8 1" "26 8192 8192 8192 8 1" "27 8192 8192 8192 8 1" "29 8192 8192 8192 8 1" "3 2048 11008 4096 8 1" "6 2048 11008 4096 8 1"; do
  set -- $spec
  run sustain 32 $1 $2 $3 $4 2.0 $5 $6
done
echo "== 3d. bench.py (pipelined e2e)" >> $LOG
timeout 600 python bench.py --steps 200 --warmup 10 --cpu_seconds 2 >> $LOG 2>&1
echo "== 4. pytest" >> $LOG
timeout 1200 python -m pytest tests -m gpu -x -q >> $LOG 2>&1; echo "pytest rc=$?" >> $LOG
grep -E "FAIL|exit|watchdog|TIME|SUSTAIN|pytest rc|passed|failed|\"metric\"" $LOG | cut -c1-400 | tail -120
