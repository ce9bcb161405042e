#!/bin/bash
# Harness-protocol trial (2 GPUs) of the 256x256 CTA-pair tile with the default rasterisation (config 3, group_m 8, no
# K-decomposition) on the tensor-bound shapes that sit below 0.985 of cuBLASLt-auto-tuning-max with another choice: the
# partial re-sweep showed 18/19 -> 3 gaining 3-6 % and 3 -> 6 losing 3 %, which the rotation-based tuner cannot resolve.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round2i.log
: > $LOG
rm -rf gpurun_out/farm_fp32_trial3
B200_HGEMM_FORCE=3,8,1 timeout 2400 python farm_sweep.py --gpus 2 --acc_precise fp32 --seconds 0.12 --tune_rounds 50,100 --engine wallgrid \
    --shapes_file profiles/r2_trial_cfg3_shapes.txt --base_dir gpurun_out/farm_fp32_trial3 --out_dir gpurun_out/eval_trial3 >> $LOG 2>&1
echo "farm rc=$?" >> $LOG
tail -c 1200 $LOG
