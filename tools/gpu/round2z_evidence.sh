#!/bin/bash
# The round's evidence (one B200): full GPU test-suite, smoke(), bench.py as the driver runs it, the ncu launch list of
# that command, `ncu --set full` captures of the kernels behind BASELINE configs 2-4 (reports brought back for the CSV
# exports under profiles/).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round2z.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
echo "== 1. pytest -m gpu + smoke" >> $LOG
timeout 1800 python -m pytest tests -m gpu -x -q >> $LOG 2>&1; echo "pytest rc=$?" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1; echo "smoke rc=$?" >> $LOG
echo "== 2. bench.py" >> $LOG
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_20.json 2>> $LOG; echo "bench20 rc=$?" >> $LOG
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2>> $LOG; echo "bench rc=$?" >> $LOG
timeout 900 python bench.py --mnk 8192_8192_8192 --acc fp16 --steps 300 --sweep none > gpurun_out/r2_bench_8192_fp16.json 2>> $LOG
timeout 900 python bench.py --mnk 2048_11008_4096 --steps 1000 --sweep none > gpurun_out/r2_bench_2048_11008_4096.json 2>> $LOG
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference.json 2>> $LOG
echo "== 2b. server mode (qps 100) on the stratified sample, both accumulators: the harness's Python loop on the C-ABI libraries" >> $LOG
for acc in fp32 fp16; do
  rm -rf gpurun_out/farm_server_$acc
  timeout 600 python farm_sweep.py --gpus 1 --acc_precise $acc --engine pyharness --perf_funcs auto --mode server --target_qps 100 --seconds 0.4 \
      --shapes "$(cat profiles/r2_harness_sample_shapes.txt)" --base_dir gpurun_out/farm_server_$acc --out_dir gpurun_out/eval_server --tag _sample >> $LOG 2>&1
  echo "server $acc rc=$?" >> $LOG
done
echo "== 3. ncu launch list of the bench command (serialised, cold-cache: shares only)" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --e2e_steps 2 --cpu_seconds 0.5 --sweep none --sustained_seconds 0 > gpurun_out/r2_bench_under_ncu.json 2>> $LOG
echo "== 4. ncu --set full" >> $LOG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hgemm_tn -s 5 -c 1 -f -o gpurun_out/r2_prof_bench_4096 \
    python bench.py --steps 20 --warmup 3 --e2e_steps 1 --cpu_seconds 0.2 --sweep none --sustained_seconds 0 >> $LOG 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hgemm_tn -s 2 -c 1 -f -o gpurun_out/r2_prof_8192_fp16 $DC time 16 -1 8192 8192 8192 2 >> $LOG 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hgemm_tn -s 2 -c 1 -f -o gpurun_out/r2_prof_2048_11008_4096 $DC time 32 -1 2048 11008 4096 2 >> $LOG 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hgemm_tn -s 2 -c 1 -f -o gpurun_out/r2_prof_16384 $DC time 32 -1 16384 16384 16384 2 >> $LOG 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
    --clock-control none -k regex:nvjet -s 1 -c 2 --csv --log-file gpurun_out/r2_ncu_cublas_ref.csv $DC time 32 -1 4096 4096 4096 2 >> $LOG 2>&1
ls -la gpurun_out/*.ncu-rep >> $LOG 2>&1
grep -E "pytest rc|passed|failed|smoke|bench.* rc" $LOG; tail -c 600 gpurun_out/r2_bench_n1_20.json
