#!/bin/bash
# The round's evidence (one B200, ~10 minutes): bench.py as the driver runs it, the ncu launch list of that command,
# `ncu --set full` captures (reports brought back for the CSV exports under profiles/), a server-mode sample, the GPU
# test-suite and smoke().
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round2z.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
echo "== 1. bench.py (driver-style)" >> $LOG
timeout 240 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_20.json 2>> $LOG; echo "bench20 rc=$?" >> $LOG
echo "== 2. ncu launch list of the bench command (serialised, cold-cache: shares only)" >> $LOG
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --e2e_steps 2 --cpu_seconds 0.5 --sweep none --sustained_seconds 0 > gpurun_out/r2_bench_under_ncu.json 2>> $LOG
echo "== 3. ncu --set full" >> $LOG
timeout 200 ncu --set full --clock-control none --import-source on -k regex:hgemm_tn -s 5 -c 1 -f -o gpurun_out/r2_prof_bench_4096 \
    python bench.py --steps 20 --warmup 3 --e2e_steps 1 --cpu_seconds 0.2 --sweep none --sustained_seconds 0 >> $LOG 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:hgemm_tn -s 2 -c 1 -f -o gpurun_out/r2_prof_8192_fp16 $DC time 16 -1 8192 8192 8192 2 >> $LOG 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
    --clock-control none -s 4 -c 8 --csv --log-file gpurun_out/r2_ncu_2048_11008_4096_vs_cublas.csv $DC time 32 -1 2048 11008 4096 2 >> $LOG 2>&1
echo "== 4. more bench lines" >> $LOG
timeout 200 python bench.py --mnk 8192_8192_8192 --acc fp16 --steps 300 --sweep none --cpu_seconds 3 > gpurun_out/r2_bench_8192_fp16.json 2>> $LOG
timeout 240 python bench.py > gpurun_out/r2_bench_n1.json 2>> $LOG; echo "bench rc=$?" >> $LOG
echo "== 5. server mode (qps 100) on the stratified sample: the harness's Python loop on the C-ABI libraries" >> $LOG
rm -rf gpurun_out/farm_server_fp32
timeout 200 python farm_sweep.py --gpus 1 --acc_precise fp32 --engine pyharness --perf_funcs auto --mode server --target_qps 100 --seconds 0.4 \
    --shapes "$(cat profiles/r2_harness_sample_shapes.txt)" --base_dir gpurun_out/farm_server_fp32 --out_dir gpurun_out/eval_server --tag _sample >> $LOG 2>&1
echo "server rc=$?" >> $LOG
echo "== 6. pytest -m gpu + smoke" >> $LOG
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1; echo "smoke rc=$?" >> $LOG
timeout 330 python -m pytest tests -m gpu -x -q >> $LOG 2>&1; echo "pytest rc=$?" >> $LOG
grep -E "pytest rc|passed|failed|smoke|bench.* rc|server rc" $LOG; tail -c 400 gpurun_out/r2_bench_n1_20.json
