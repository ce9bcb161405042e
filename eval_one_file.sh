#!/bin/bash
# Evaluate ONE (M,N,K): exactness check, then the b200 kernel against each of the seven baselines
# (one fresh process per baseline, random order), then the speed-up summary.
# Flags are the reference's (eval_one_file.sh:14-59):
#   --mnk M_N_K --acc_precise fp32|fp16 --device_type b200 --warmup_seconds S --benchmark_seconds S
#   --base_dir DIR --gpu_device_id I --mode offline|server [--target_qps Q]
#   (extras: --seed N; --perf_funcs a,b,... restricts the baselines timed — default all seven — so that a sweep that
#    only needs the cuBLASLt-auto-tuning pair does not pay for seven process start-ups per shape)
# Unlike the reference, a FAILED correctness check (not only an exception) stops the run.
set -u
MODE=offline; SEED=0; TARGET_QPS=""
PERF_FUNCS="hgemm_cublas_tn,hgemm_cublas_nn,hgemm_cublaslt_heuristic_tn,hgemm_cublaslt_heuristic_nn,hgemm_cublaslt_auto_tuning_tn,hgemm_cublaslt_auto_tuning_nn,matmul"
while [ $# -gt 0 ]; do
  case "$1" in
    --mnk) MNK=$2;; --acc_precise) ACC_PRECISE=$2;; --device_type) DEVICE_TYPE=$2;;
    --warmup_seconds) WARMUP_SECONDS=$2;; --benchmark_seconds) BENCHMARK_SECONDS=$2;;
    --base_dir) BASE_DIR=$2;; --gpu_device_id) GPU_DEVICE_ID=$2;; --mode) MODE=$2;;
    --target_qps) TARGET_QPS=$2;; --seed) SEED=$2;; --perf_funcs) PERF_FUNCS=$2;;
    *) echo "Unknown option: $1"; exit 1;;
  esac
  shift 2
done
for v in MNK ACC_PRECISE DEVICE_TYPE WARMUP_SECONDS BENCHMARK_SECONDS BASE_DIR GPU_DEVICE_ID; do
  if [ -z "${!v:-}" ]; then echo "missing --$(echo $v | tr 'A-Z' 'a-z')"; exit 1; fi
  echo "$v: ${!v}"
done
if [ "$MODE" = "server" ] && [ -z "$TARGET_QPS" ]; then echo "--mode server needs --target_qps"; exit 1; fi
cd "$(dirname "$0")" || exit 1
mkdir -p "$BASE_DIR"
rm -f "$BASE_DIR"/benchmark_result_*.json
COMMON=(--mnk "$MNK" --acc_precise "$ACC_PRECISE" --device_type "$DEVICE_TYPE" --base_dir "$BASE_DIR"
        --gpu_device_id "$GPU_DEVICE_ID" --seed "$SEED")

python zero_one_correctness_check.py "${COMMON[@]}" || { echo "Error: correctness check did not pass. Exiting..."; exit 1; }

echo "Executing hgemm benchmark with shuffled perf_funcs..."
for func in $(shuf -e ${PERF_FUNCS//,/ }); do
  echo "---------------------------------------------------------"
  echo ">>> Running benchmark for: $func"
  if [ "$MODE" = "server" ]; then
    python benchmarking_server.py "${COMMON[@]}" --warmup_seconds "$WARMUP_SECONDS" \
      --benchmark_seconds "$BENCHMARK_SECONDS" --perf_func "$func" --target_qps "$TARGET_QPS"
  else
    python benchmarking_offline.py "${COMMON[@]}" --warmup_seconds "$WARMUP_SECONDS" \
      --benchmark_seconds "$BENCHMARK_SECONDS" --perf_func "$func"
  fi || { echo "Error: Benchmark failed at perf_func: $func. Exiting..."; exit 1; }
done
python summarize_result.py --base_dir "$BASE_DIR" --acc_precise "$ACC_PRECISE" --device_type "$DEVICE_TYPE"
echo "All benchmarks completed successfully!"
