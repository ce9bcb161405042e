#!/bin/bash
# Where does the wall-minus-kernel time go? Host-call duration of our launch vs the library calls.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/round1n.log
: > $LOG
DC=cuda_l2_b200/lib/dev_check
for s in "64 64 64" "512 512 512" "512 512 2048" "1024 1024 1024" "4096 4096 1024" "4096 4096 4096" "8192 8192 512" "2048 11008 4096"; do
  timeout 300 $DC wall 32 $s 0.3 3 8 >> $LOG 2>&1
done
for f in "2,0,1" "1,0,-4" "1,0,4" "7,0,1" "4,0,1"; do
  echo "force $f" >> $LOG
  B200_HGEMM_FORCE=$f timeout 300 $DC wall 32 512 512 2048 0.3 3 8 >> $LOG 2>&1
done
echo DONE >> $LOG
grep -E "^WALL|force" $LOG | sed 's/samples=.*speedup/ speedup/' 
