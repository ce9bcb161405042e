#!/bin/bash
# A/B of the MMA-issue-loop and producer-prefetch variants + light ncu metric lists of mid-size shapes (ours vs
# cuBLAS). Keep gpurun_out small: only CSV/text goes there.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1e.log
: > $LOG
for v in A B C D; do
  DC=cuda_l2_b200/lib/variants/$v/dev_check
  echo "=== variant $v" >> $LOG
  for s in "4096 4096 4096 10" "1024 1024 1024 30" "128 4096 4096 20" "512 512 2048 30"; do
    timeout 300 $DC sweep 32 $s >> $LOG 2>&1
  done
done
echo "=== ncu" >> $LOG
DC=cuda_l2_b200/lib/dev_check
M=gpu__time_duration.sum,sm__cycles_elapsed.max,launch__grid_size,launch__block_size,launch__cluster_size,launch__cluster_dim_x,launch__cluster_dim_y,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
for s in "128 4096 4096" "1024 1024 8192" "512 512 2048" "1024 1024 1024" "64 64 16384" "4096 4096 4096" "16384 16384 16384"; do
  n=$(echo $s | tr ' ' 'x')
  timeout 300 ncu --metrics $M --clock-control none -s 2 -c 16 --csv --log-file gpurun_out/ncu_mid_$n.csv $DC time 32 -1 $s 1 >> $LOG 2>&1
done
echo DONE >> $LOG
tail -3 $LOG
du -sh gpurun_out
