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
rm -rf gpurun_out/farm_harness_fp32
timeout 2400 python farm_sweep.py --gpus 2 --acc_precise fp32 --engine harness --perf_funcs auto --seconds $SEC \
    --shapes "$(cat profiles/r2_harness_sample_shapes.txt)" --base_dir gpurun_out/farm_harness_fp32 --out_dir gpurun_out/eval_harness --tag _harness_sample >> $LOG 2>&1
echo "harness farm rc=$?" >> $LOG
tail -c 1500 $LOG
