#!/bin/bash
# Round 2: re-tune both accumulators in the harness's power state (`dev_check grid ... wall`: event-time shortlist, then a
# harness-like rotation ranked by mean TFLOP/s), one share of the shape grid per visible GPU.
#   bash tools/gpu/round2c_retune.sh [budget_ms=3.0] [accs="32 16"]
# Output: gpurun_out/grid_wall_fp{32,16}.csv  ->  python tools/tune_b200.py <both files>; python __graft_entry__.py
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
BUDGET=${1:-3.0}
ACCS=${2:-"32 16"}
NG=$(nvidia-smi -L | wc -l)
DC=cuda_l2_b200/lib/dev_check
for acc in $ACCS; do
  pids=""
  for g in $(seq 0 $((NG - 1))); do
    CUDA_VISIBLE_DEVICES=$g $DC grid $acc $g $NG $BUDGET 0 1e30 wall > gpurun_out/grid_wall_fp${acc}_part$g.csv 2> gpurun_out/grid_wall_fp${acc}_part$g.err &
    pids="$pids $!"
  done
  rc=0
  for p in $pids; do wait $p || rc=1; done
  cat gpurun_out/grid_wall_fp${acc}_part*.csv | grep "^GRID," > gpurun_out/grid_wall_fp${acc}.csv
  echo "acc $acc: $(wc -l < gpurun_out/grid_wall_fp${acc}.csv) shapes tuned, rc=$rc, failures: $(grep -c GRIDFAIL gpurun_out/grid_wall_fp${acc}_part*.csv | paste -sd' ')"
  rm -f gpurun_out/grid_wall_fp${acc}_part*.csv
done
