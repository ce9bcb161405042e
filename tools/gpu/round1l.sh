#!/bin/bash
# Final single-GPU validation of the round: smoke, the GPU test-suite (C ABI + JIT harness), the bench lines,
# ncu evidence for bench.py, and one run of the reference-style per-shape driver in server mode.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1l.log
: > $LOG
echo "=== smoke" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1; echo "smoke rc=$?" >> $LOG
echo "=== pytest" >> $LOG
timeout 1500 python -m pytest tests -m gpu -x -q >> $LOG 2>&1; echo "pytest rc=$?" >> $LOG
echo "=== bench" >> $LOG
timeout 600 python bench.py > gpurun_out/bench_n1.json 2>> $LOG; echo "bench rc=$?" >> $LOG
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> $LOG
timeout 600 python bench.py --mnk 8192_8192_8192 --acc fp16 --steps 300 > gpurun_out/bench_8192_fp16.json 2>> $LOG
echo "=== ncu" >> $LOG
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --e2e_steps 2 --cpu_seconds 0.5 >> $LOG 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hgemm_tn -s 5 -c 1 -f -o /tmp/prof_bench \
    python bench.py --steps 20 --warmup 3 --e2e_steps 2 --cpu_seconds 0.5 >> $LOG 2>&1
ncu -i /tmp/prof_bench.ncu-rep --page raw --csv > gpurun_out/ncu_bench_full_raw.csv 2>> $LOG
ncu -i /tmp/prof_bench.ncu-rep --page details --csv > gpurun_out/ncu_bench_full_details.csv 2>> $LOG
echo "=== eval_one_file (server mode)" >> $LOG
timeout 1200 ./eval_one_file.sh --mnk 2048_11008_4096 --acc_precise fp32 --device_type b200 --warmup_seconds 1 \
    --benchmark_seconds 3 --base_dir gpurun_out/eval_2048_11008_4096_server --gpu_device_id 0 --mode server --target_qps 100 \
    > gpurun_out/eval_one_file_server.log 2>&1; echo "eval rc=$?" >> $LOG
rm -rf gpurun_out/eval_2048_11008_4096_server/*.o gpurun_out/eval_2048_11008_4096_server/*.so gpurun_out/eval_2048_11008_4096_server/.ninja* gpurun_out/eval_2048_11008_4096_server/build.ninja
echo DONE >> $LOG
tail -5 $LOG; du -sh gpurun_out
