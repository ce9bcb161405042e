#!/bin/bash
# Partial re-sweep (harness protocol, 2 GPUs) of the fp32 shapes whose tuned entry changed after the second tuning pass,
# merged with the full sweep of the first table; before it, on GPU 0, the cooperative + programmatic-launch attribute pair.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round2h.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
echo "== stream-K launches: cooperative only vs cooperative + programmatic stream serialisation" >> $LOG
for spec in "3 4096 4096 4096 4 101" "3 2048 11008 4096 4 100"; do
  set -- $spec
  CUDA_VISIBLE_DEVICES=0 timeout 120 $DC time 32 $1 $2 $3 $4 50 $5 $6 >> $LOG 2>&1
  echo "## COOP_PDL=1" >> $LOG; CUDA_VISIBLE_DEVICES=0 B200_HGEMM_COOP_PDL=1 timeout 120 $DC time 32 $1 $2 $3 $4 50 $5 $6 >> $LOG 2>&1
  CUDA_VISIBLE_DEVICES=0 B200_HGEMM_COOP_PDL=1 timeout 120 $DC check 32 $1 $2 $3 $4 $5 $6 >> $LOG 2>&1
done
for shape in "4096 4096 4096" "16384 16384 16384" "12288 12288 12288" "1024 1024 4096"; do
  CUDA_VISIBLE_DEVICES=0 timeout 200 $DC check 32 -1 $shape >> $LOG 2>&1 || echo "  -> check failed: $shape" >> $LOG
done
rm -rf gpurun_out/farm_fp32_r2b
timeout 2400 python farm_sweep.py --gpus 2 --acc_precise fp32 --seconds 0.12 --tune_rounds 50,100 --engine wallgrid \
    --shapes_file profiles/r2_resweep_fp32_shapes.txt --base_dir gpurun_out/farm_fp32_r2b --out_dir gpurun_out/eval_r2b >> $LOG 2>&1
echo "farm rc=$?" >> $LOG
grep -E "TIME|CHECK|COOP|check failed|farm rc|sweep wall" $LOG | cut -c1-300; tail -c 1500 $LOG
