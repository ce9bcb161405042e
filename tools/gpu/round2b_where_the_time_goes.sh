#!/bin/bash
# Round 2 diagnosis run (about 6 minutes of box time): why the tensor-bound class sits at 0.97x of auto-tuned cuBLASLt
# in the harness while it is at parity in isolated timing.
#   1. sustain: burst vs power-capped throughput of our kernel and cuBLAS on the same shape, with clocks / watts sampled;
#      configs that move fewer bytes per FLOP (CTA pairs, multicast) against those the round-1 tuner picked;
#   2. trace: per-CTA phase timestamps (ramp, k-block rate, tail) of the headline shapes;
#   3. ncu: SM clock, tensor-pipe activity, L2 / DRAM traffic of ours vs the cuBLAS kernel.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round2b.log
: > $LOG
python -c "from cuda_l2_b200 import build; build.build_trace(); build.build_wait_hint(2000); build.build_early_tma(); build.build_split_setup()" >> $LOG 2>&1
DC=cuda_l2_b200/lib/dev_check
DT=cuda_l2_b200/lib/dev_check_trace
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown \
    --format=csv -lms 100 > gpurun_out/round2b_clocks.csv &
SMI=$!
run() { echo "## $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc" >> $LOG; }
echo "== 1. sustain (3 s each; timestamps let the clocks log be lined up)" >> $LOG
for spec in "3 8192 8192 8192 8" "0 8192 8192 8192 16" "20 8192 8192 8192 8" "21 8192 8192 8192 8" \
            "6 4096 4096 4096 8" "3 4096 4096 4096 8" "24 4096 4096 4096 8" \
            "0 4096 12288 16384 16" "3 4096 12288 16384 8" \
            "2 128 8192 16384 0" "7 128 8192 16384 0" "8 128 8192 16384 0" "1 128 8192 16384 0"; do
  set -- $spec
  date +%T.%N >> $LOG
  run $DC sustain 32 $1 $2 $3 $4 3.0 $5 1
done
run $DC sustain 32 1 128 8192 16384 3.0 0 -2
echo "== 1b. experiments: polling hint build, L2 promotion of the operand maps, config 26, stream-K on skinny shapes" >> $LOG
DH=cuda_l2_b200/lib/dev_check_hint
for spec in "3 8192 8192 8192 8" "6 4096 4096 4096 8" "2 128 8192 16384 0"; do set -- $spec; run $DH sustain 32 $1 $2 $3 $4 3.0 $5 1; done
for promo in 0 2; do
  for spec in "3 8192 8192 8192 8" "2 128 8192 16384 0"; do set -- $spec; echo "## L2 promotion $promo" >> $LOG; B200_HGEMM_L2_PROMOTION=$promo timeout 120 $DC sustain 32 $1 $2 $3 $4 3.0 $5 1 >> $LOG 2>&1; done
done
for spec in "26 8192 8192 8192 8" "26 16384 16384 16384 8" "3 16384 16384 16384 8" "26 4096 12288 16384 8"; do set -- $spec; run $DC sustain 32 $1 $2 $3 $4 3.0 $5 1; done
for spec in "1 128 8192 16384" "0 128 8192 16384" "2 128 8192 16384" "3 12288 256 12288" "4 12288 256 12288"; do set -- $spec; run $DC sustain 32 $1 $2 $3 $4 3.0 0 100; done
echo "== 2. trace" >> $LOG
run $DT trace 32 6 4096 4096 4096 8 1
run $DT trace 32 3 4096 4096 4096 8 1
run $DT trace 32 3 8192 8192 8192 8 1
run $DT trace 32 3 512 8192 8192 0 1
run $DT trace 32 2 128 8192 16384 0 1
run $DT trace 32 2 1024 1024 1024 0 1
# the K = 1024..2048 family (7-9 us kernels, won on only 24-33 % of the shapes, by 0.5-2 us): where do the fixed costs sit?
run $DT trace 32 -1 256 2048 2048
run $DT trace 32 -1 1024 1024 2048
run $DT trace 32 -1 512 2048 1024
run $DT trace 32 -1 4096 2048 1024
run $DT trace 32 -1 256 2048 2048 0 1 1
# the same launches from the state the harness's rotation leaves behind (caches and TLBs hold other kernels' data):
# the skinny HBM-bound shapes lose 15-20 % between back-to-back and isolated timing, cuBLAS only 5 %
run $DT trace 32 2 128 8192 16384 0 1 1
run $DT trace 32 2 1024 1024 1024 0 1 1
run $DT trace 32 6 4096 4096 4096 8 1 1
run $DT trace 32 3 512 8192 8192 0 1 1
echo "== 2b. first loads before the set-up barrier (dev_check_early) on the mid-K family, single-CTA configs" >> $LOG
DE=cuda_l2_b200/lib/dev_check_early
for spec in "2 1024 1024 2048" "2 256 2048 2048" "1 512 2048 1024" "2 1024 1024 1024" "12 256 256 1024"; do
  set -- $spec
  run $DC check 32 $1 $2 $3 $4; run $DE check 32 $1 $2 $3 $4
  run $DC time 32 $1 $2 $3 $4 200; run $DE time 32 $1 $2 $3 $4 200
done
echo "== 2c. TMEM allocation behind the first set-up barrier (dev_check_split), single CTAs and pairs" >> $LOG
DS=cuda_l2_b200/lib/dev_check_split
for spec in "2 1024 1024 2048 0" "4 256 2048 2048 0" "3 4096 2048 1024 8" "6 4096 4096 4096 8"; do
  set -- $spec
  run $DS check 32 $1 $2 $3 $4 $5
  run $DC time 32 $1 $2 $3 $4 200 $5; run $DS time 32 $1 $2 $3 $4 200 $5
done
kill $SMI
echo "== 3. ncu (serialised, cold: compare shapes of the numbers)" >> $LOG
M="gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu.sum,smsp__inst_executed.sum,launch__grid_size,launch__cluster_size"
for spec in "-1 4096 4096 4096" "-1 512 8192 8192" "-1 128 8192 16384" "-1 8192 8192 8192"; do
  set -- $spec
  timeout 300 ncu --metrics $M --clock-control none -s 4 -c 16 --csv --log-file gpurun_out/round2b_ncu_$2x$3x$4.csv \
      $DC time 32 $1 $2 $3 $4 2 >> $LOG 2>&1
done
grep -E "SUSTAIN|TRACE|median|exit" $LOG | cut -c1-220 | tail -120
