#!/bin/bash
# Two-GPU check of the multi-GPU paths: bench.py under torchrun and the farmed sweep on a handful of shapes.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1m.log
: > $LOG
nvidia-smi -L >> $LOG 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 500 --warmup 5 > gpurun_out/bench_n2.json 2>> $LOG; echo "bench2 rc=$?" >> $LOG
timeout 150 python farm_sweep.py --gpus 2 --acc_precise fp32 --seconds 0.2 --tune_rounds 3,6 --limit 3 \
    --base_dir gpurun_out/farm2 --out_dir gpurun_out/farm2_results >> $LOG 2>&1; echo "farm2 rc=$?" >> $LOG
echo DONE >> $LOG
tail -6 $LOG; cut -c1-600 gpurun_out/bench_n2.json
