#!/bin/bash
# Confirmation (2 GPUs): bench.py under torchrun at N = 2 (the sweep leg sharded over two ranks), then the full fp32 sweep
# in the harness protocol with the FINAL table — the number to quote.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round2j.log
: > $LOG
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
echo "== bench.py --gpus 2 under torchrun" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2_20.json 2>> $LOG; echo "bench n2 rc=$?" >> $LOG
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_n2.json 2>> $LOG; echo "reference n2 rc=$?" >> $LOG
tail -c 2500 gpurun_out/r2_bench_n2_20.json >> $LOG
echo "== full fp32 sweep, final table" >> $LOG
bash tools/gpu/round2e_sweep.sh fp32 2 0.12 final >> $LOG 2>&1
tail -c 2500 $LOG
