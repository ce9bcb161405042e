// cuBLAS comparator for --acc_precise fp32 (cublasGemmEx, NN and TN).
// Stands in for the reference's cublas/fp32/hgemm_cublas.cu:41-68 (same library call and compute type);
// implementation shared in cuda_l2_b200/csrc/baselines.cuh.
#include "cuda_l2_b200/csrc/baselines.cuh"
#include "b200_raw_api.h"

static b200bl::Cublas<32> g_cublas;

int b200raw_cublas_init() { return g_cublas.init(); }
void b200raw_cublas_destroy() { g_cublas.destroy(); }
int b200raw_cublas_gemm(int layout, const void* A, const void* B, void* C, int M, int N, int K) {
  return g_cublas.gemm(layout ? b200bl::kTN : b200bl::kNN, static_cast<const __half*>(A), static_cast<const __half*>(B),
                       static_cast<__half*>(C), M, N, K);
}
