#!/bin/bash
# Re-run of the GPU test-suite (harness JIT tests first, then the C-ABI parity tests).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1q.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -x -q >> $LOG 2>&1; echo "pytest rc=$?" >> $LOG
tail -6 $LOG
