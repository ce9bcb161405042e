#!/bin/bash
# Round 2, second GPU call (one B200): (1) the full GPU test-suite with the round-2 additions (bf16, operator, top-of-grid
# parity, cooperative split-K, programmatic dependent launch), (2) A/B of programmatic dependent launch and of the
# cooperative launch attribute, (3) bench.py as the driver runs it (with the sweep leg), (4) the wall-metric tuner over
# the whole grid for fp32 accumulation, large shapes first.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round2d.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run() { echo "## $*" >> $LOG; timeout 180 "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
echo "== 0. sanity (new library: PDL, cooperative split-K, bf16 build)" >> $LOG
run $DC check 32 -1 4096 4096 4096
run $DC check 32 -1 64 64 16384
run $DC check 16 -1 1000 1000 1000
run $DC check 32 1 256 512 2048 0 8
echo "== 1. pytest -m gpu" >> $LOG
timeout 1500 python -m pytest tests -m gpu -x -q >> $LOG 2>&1; echo "pytest rc=$?" >> $LOG
echo "== 2. programmatic dependent launch A/B (back-to-back CUDA-event timing; NO_PDL=1 is the old launch)" >> $LOG
for spec in "6 4096 4096 4096 8" "21 4096 4096 4096 8" "3 4096 4096 4096 8" "26 8192 8192 8192 8" "3 8192 8192 8192 8" \
            "3 2048 11008 4096 8" "21 2048 11008 4096 8" "-1 1024 1024 1024 0" "-1 256 2048 2048 0" "-1 512 512 512 0" "-1 4096 2048 1024 0"; do
  set -- $spec
  run $DC time 32 $1 $2 $3 $4 50 $5 1
  echo "## NO_PDL" >> $LOG; B200_HGEMM_NO_PDL=1 timeout 180 $DC time 32 $1 $2 $3 $4 50 $5 1 >> $LOG 2>&1
done
run $DC time 16 26 8192 8192 8192 20 8 1
run $DC time 16 3 8192 8192 8192 20 8 1
echo "== 2b. cooperative launch attribute: cost on the workspace split-K shapes" >> $LOG
for spec in "64 64 16384" "128 64 16384" "64 128 8192"; do
  set -- $spec
  run $DC time 32 -1 $1 $2 $3 200
  echo "## NO_COOPERATIVE" >> $LOG; B200_HGEMM_NO_COOPERATIVE=1 timeout 180 $DC time 32 -1 $1 $2 $3 200 >> $LOG 2>&1
  run $DC time 32 2 $1 $2 $3 200 0 16
  echo "## NO_COOPERATIVE" >> $LOG; B200_HGEMM_NO_COOPERATIVE=1 timeout 180 $DC time 32 2 $1 $2 $3 200 0 16 >> $LOG 2>&1
done
echo "== 2c. the harness protocol on the BASELINE shapes (pairs, 0.3 s per auto-tuning pair, reference tuning rounds)" >> $LOG
for spec in "4096 4096 4096" "2048 11008 4096" "64 4096 64" "1024 1024 2048"; do set -- $spec; run $DC wall 32 $1 $2 $3 0.3 50 100; done
echo "== 3. bench.py (driver-style 20 steps, then 200)" >> $LOG
timeout 600 python bench.py --steps 20 --warmup 3 --cpu_seconds 2 > gpurun_out/bench_r2d_20.json 2>> $LOG; tail -c 3000 gpurun_out/bench_r2d_20.json >> $LOG
timeout 600 python bench.py --steps 200 --warmup 10 --cpu_seconds 2 --sweep none > gpurun_out/bench_r2d_200.json 2>> $LOG
echo "== 4. wall-metric tuner, fp32 accumulation: >= 20 GFLOP first, then the rest" >> $LOG
timeout 700 $DC grid 32 0 1 3.0 20 1e30 wall > gpurun_out/grid_fp32_wall_r2_big.csv 2>> $LOG; echo "tuner(big) rc=$?" >> $LOG
timeout 420 $DC grid 32 0 1 3.0 0 20 wall > gpurun_out/grid_fp32_wall_r2_small.csv 2>> $LOG; echo "tuner(small) rc=$?" >> $LOG
wc -l gpurun_out/grid_fp32_wall_r2_*.csv >> $LOG
grep -E "FAIL|exit|watchdog|TIME |WALL,|pytest rc|passed|failed|tuner|NO_" $LOG | cut -c1-420 | tail -150
