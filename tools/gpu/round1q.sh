#!/bin/bash
# Re-run of the GPU test-suite after the shared-static fix (harness JIT tests first, then the C-ABI parity tests).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1q.log
: > $LOG
timeout 1500 python -m pytest tests -m gpu -q >> $LOG 2>&1; echo "pytest rc=$?" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1; echo "smoke rc=$?" >> $LOG
tail -15 $LOG
