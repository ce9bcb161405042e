// cuBLASLt "AutoTuning" comparator for --acc_precise fp16 — THE bar the b200 kernels are measured against:
// up to 100 heuristic candidates timed over 50 warm-up + 100 measured rounds, best median kept.
// Stands in for cublas/fp16/hgemm_cublaslt_auto_tuning.cu:108-546.
#include "cuda_l2_b200/csrc/baselines.cuh"
#include "b200_raw_api.h"

static b200bl::LtAutoTune<16> g_tune;

int b200raw_lt_autotune_init() { return g_tune.init(); }
void b200raw_lt_autotune_destroy() { g_tune.destroy(); }
int b200raw_lt_autotune_find(int layout, int M, int N, int K) {
  return g_tune.find(layout ? b200bl::kTN : b200bl::kNN, M, N, K);
}
int b200raw_lt_autotune_gemm(int layout, const void* A, const void* B, void* C, int M, int N, int K) {
  return g_tune.gemm(layout ? b200bl::kTN : b200bl::kNN, static_cast<const __half*>(A), static_cast<const __half*>(B),
                     static_cast<__half*>(C), M, N, K);
}
