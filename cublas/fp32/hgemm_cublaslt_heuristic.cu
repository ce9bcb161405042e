// cuBLASLt "heuristic" comparator for --acc_precise fp32: first algorithm cublasLtMatmulAlgoGetHeuristic
// returns, descriptors cached per shape. Stands in for cublas/fp32/hgemm_cublaslt_heuristic.cu:65-217.
#include "cuda_l2_b200/csrc/baselines.cuh"
#include "b200_raw_api.h"

static b200bl::LtHeuristic<32> g_heur;

int b200raw_lt_heuristic_init() { return g_heur.init(); }
void b200raw_lt_heuristic_destroy() { g_heur.destroy(); }
int b200raw_lt_heuristic_gemm(int layout, const void* A, const void* B, void* C, int M, int N, int K) {
  return g_heur.gemm(layout ? b200bl::kTN : b200bl::kNN, static_cast<const __half*>(A), static_cast<const __half*>(B),
                     static_cast<__half*>(C), M, N, K);
}
