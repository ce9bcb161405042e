#!/bin/bash
# The REAL reference-style harness (eval_one_file.sh: JIT-built torch extension, 0/1 check, one fresh process per baseline,
# summarize_result.py) on a stratified sample of the grid, cuBLASLt-auto-tuning pair only, farmed over 2 GPUs; then the
# fp16-accumulate sweep in the C++ restatement of the protocol.
#   tools/gpu/round2k_harness_sample.sh <benchmark seconds per pair>
cd "$(dirname "$0")/../.." || exit 1
SEC=${1:-1.5}
mkdir -p gpurun_out
LOG=gpurun_out/round2k.log
: > $LOG
echo "== 8192^3 fp16 accumulate (BASELINE config 3) in the harness protocol with the 512x256 tile" >> $LOG
CUDA_VISIBLE_DEVICES=0 timeout 300 cuda_l2_b200/lib/dev_check wall 16 8192 8192 8192 0.3 50 100 >> $LOG 2>&1
CUDA_VISIBLE_DEVICES=0 B200_HGEMM_FORCE=3,8,1 timeout 300 cuda_l2_b200/lib/dev_check wall 16 8192 8192 8192 0.3 50 100 >> $LOG 2>&1
rm -rf gpurun_out/farm_harness_fp32
timeout 2400 python farm_sweep.py --gpus 2 --acc_precise fp32 --engine harness --perf_funcs auto --seconds $SEC \
    --shapes "$(cat profiles/r2_harness_sample_shapes.txt)" --base_dir gpurun_out/farm_harness_fp32 --out_dir gpurun_out/eval_harness --tag _harness_sample >> $LOG 2>&1
echo "harness farm rc=$?" >> $LOG
grep "^WALL" $LOG | sed "s/.*cfg=/cfg=/" | cut -c1-60; grep -o "speedup_vs_lt_auto_max=[0-9.]*" $LOG | head -4; tail -c 1500 $LOG
