#!/bin/bash
# First light of the stream-K branch inside the last ~2 GPU-minutes of round 1: a handful of exactness checks
# (stream-K on single CTAs and pairs, one and many contributors, tail + data-parallel), the plain / split-K
# regression, then A/B timings on wave-quantised shapes. One short process per case.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1r.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
run() { timeout 15 $DC "$@" >> $LOG 2>&1; rc=$?; [ $rc -ne 0 ] && echo "  -> exit $rc : $*" >> $LOG; }
run check 32 3 512 8192 8192 0 100
run check 32 0 512 768 4096 0 100
run check 16 3 512 8192 8192 0 100
run check 32 6 4096 4096 4096 8 100
run check 32 6 4096 4096 4096 8 101
run check 32 2 1000 1224 2048 0 100
run check 16 1 1000 1224 2048 0 101
run check 32 -1 4096 4096 4096
run check 32 1 256 512 2048 0 8
run check 32 1 256 512 2048 0 -4
run check 32 10 1024 1536 1024
run check 16 4 1000 1000 1000
for sk in 1 100; do run time 32 3 512 8192 8192 30 0 $sk; done
for sk in 1 100 101; do run time 32 6 4096 4096 4096 30 8 $sk; done
for sk in 1 100; do run time 32 2 128 8192 16384 30 0 $sk; done
for sk in 1 100; do run time 32 3 12288 256 12288 30 0 $sk; done
for sk in 1 100; do run time 32 0 128 8192 16384 30 0 $sk; done
for sk in 1 100; do run time 32 3 1024 4096 8192 30 0 $sk; done
cat $LOG
