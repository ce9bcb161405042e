#!/bin/bash
# The 1001-shape sweep in the harness's protocol (dev_check wallgrid: one (baseline, ours) pair at a time, fresh operands
# per iteration, zero-filled output, reference 50+100 auto-tuning rounds), farmed over the GPUs of the box.
#   tools/gpu/round2e_sweep.sh <fp32|fp16> <gpus> <seconds per auto-tuning pair> [tag] [auto-tuning warm,timed rounds]
cd "$(dirname "$0")/../.." || exit 1
ACC=${1:-fp32}; GPUS=${2:-1}; SEC=${3:-0.12}; TAG=${4:-r2}; ROUNDS=${5:-50,100}
mkdir -p gpurun_out
LOG=gpurun_out/sweep_${ACC}_${TAG}.log
: > $LOG
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
BITS=32; [ "$ACC" = fp16 ] && BITS=16
for shape in "4096 4096 4096" "2048 11008 4096" "8192 8192 8192" "1024 1024 2048" "64 64 16384" "1000 1224 2048"; do
  CUDA_VISIBLE_DEVICES=0 timeout 120 cuda_l2_b200/lib/dev_check check $BITS -1 $shape >> $LOG 2>&1 || echo "  -> check failed: $shape" >> $LOG
done
rm -rf gpurun_out/farm_${ACC}_${TAG}
timeout 3000 python farm_sweep.py --gpus $GPUS --acc_precise $ACC --seconds $SEC --tune_rounds $ROUNDS --engine wallgrid \
    --base_dir gpurun_out/farm_${ACC}_${TAG} --out_dir gpurun_out/eval_${TAG} >> $LOG 2>&1
echo "farm rc=$?" >> $LOG
tail -c 2500 $LOG
