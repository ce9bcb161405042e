#!/bin/bash
# Round 2, one B200: (1) the two test files that changed, (2) where the time of the mid-K kernels goes (per-CTA phase
# timestamps, isolated-launch A/B of the set-up experiments), (3) bench.py with the NUMA-local host buffers, (4) the
# torch.matmul column of the whole grid through the harness's Python loop (pyharness engine).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round2f.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check; DT=cuda_l2_b200/lib/dev_check_trace; DE=cuda_l2_b200/lib/dev_check_early; DS=cuda_l2_b200/lib/dev_check_split
run() { echo "## $*" >> $LOG; timeout 180 "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
echo "== 1. pytest (changed files)" >> $LOG
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_harness.py -x -q >> $LOG 2>&1; echo "pytest rc=$?" >> $LOG
echo "== 2. traces: dispatcher's choice, warm and cold" >> $LOG
for shape in "1024 1024 2048" "256 2048 2048" "512 2048 1024" "4096 2048 1024" "1024 1024 1024" "2048 2048 2048" "4096 4096 4096"; do
  run $DT trace 32 -1 $shape
  run $DT trace 32 -1 $shape 0 1 1
done
echo "== 2b. set-up experiments, isolated launches (TIME-ISOLATED is what a synchronising caller sees)" >> $LOG
for spec in "2 1024 1024 2048 0" "12 256 2048 2048 0" "1 512 2048 1024 0" "3 4096 2048 1024 8" "4 2048 2048 2048 8"; do
  set -- $spec
  run $DC time 32 $1 $2 $3 $4 200 $5; run $DE time 32 $1 $2 $3 $4 200 $5; run $DS time 32 $1 $2 $3 $4 200 $5
done
echo "== 2c. try_wait suspend hint (20 us): sustained throughput, ours normal / hint build" >> $LOG
DH=cuda_l2_b200/lib/dev_check_hint
for spec in "26 8192 8192 8192 8" "3 4096 4096 4096 8" "3 2048 11008 4096 8"; do
  set -- $spec
  run $DC sustain 32 $1 $2 $3 $4 1.5 $5 1; run $DH sustain 32 $1 $2 $3 $4 1.5 $5 1
done
echo "== 3. bench.py" >> $LOG
timeout 600 python bench.py --steps 20 --warmup 3 --cpu_seconds 2 > gpurun_out/bench_r2f_20.json 2>> $LOG; tail -c 1500 gpurun_out/bench_r2f_20.json >> $LOG
echo "== 4. torch.matmul column (pyharness, whole grid, 0.1 s per shape)" >> $LOG
rm -rf gpurun_out/farm_matmul_fp32
timeout 900 python farm_sweep.py --gpus 1 --acc_precise fp32 --engine pyharness --perf_funcs matmul --seconds 0.1 \
    --base_dir gpurun_out/farm_matmul_fp32 --out_dir gpurun_out/eval_matmul >> $LOG 2>&1; echo "matmul rc=$?" >> $LOG
wc -l gpurun_out/farm_matmul_fp32/*.jsonl >> $LOG
grep -E "FAIL|exit|watchdog|TIME|TRACE|SUSTAIN|pytest rc|passed|failed|matmul rc|median" $LOG | cut -c1-260 | tail -150
