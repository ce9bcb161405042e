#!/bin/bash
# Final fp32 harness-metric sweep of the round (table tuned on the 8-warp-epilogue kernel).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
DC=cuda_l2_b200/lib/dev_check
timeout 2400 $DC wallgrid 32 0 1 0.3 -25 0 > gpurun_out/wallgrid_fp32.txt 2> gpurun_out/round1p.log
echo "wallgrid32 rc=$?" >> gpurun_out/round1p.log
tail -1 gpurun_out/wallgrid_fp32.txt; du -sh gpurun_out
