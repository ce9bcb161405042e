/* b200_baselines.h — C ABI of the COMPARATORS (libb200_baselines.so): the same GEMM through cuBLAS and
 * cuBLASLt, so that bench.py, the tuner and the tests can time the library against our kernel without
 * a torch JIT build. These are not part of the accelerated path.
 *
 * They mirror the reference's baseline entry points (pybind/hgemm_a100_fp32.cc:9-25):
 *   init_cublas_handle / hgemm_cublas_{nn,tn}                       cublas/<acc>/hgemm_cublas.cu:41-68
 *   init_cublaslt_handle_v1 / hgemm_cublaslt_heuristic_{nn,tn}      cublas/<acc>/hgemm_cublaslt_heuristic.cu:65-217
 *   init_cublaslt_handle_v2 / find_best_algo_{nn,tn}_v2_torch /
 *   hgemm_cublaslt_auto_tuning_{nn,tn}                              cublas/<acc>/hgemm_cublaslt_auto_tuning.cu:108-546
 *
 * acc_bits: 32 (CUBLAS_COMPUTE_32F) or 16 (CUBLAS_COMPUTE_16F).  layout: 0 = NN (B is row-major [K,N]),
 * 1 = TN (B is K-major [N,K]).  All pointers are device pointers to fp16. Work is issued on the
 * legacy default stream, as in the reference. Returns 0 on success.
 */
#ifndef B200_BASELINES_H_
#define B200_BASELINES_H_
#ifdef __cplusplus
extern "C" {
#endif

int b200_bl_init(int acc_bits);
void b200_bl_destroy(int acc_bits);
int b200_bl_cublas(int acc_bits, int layout, const void* A, const void* B, void* C, int M, int N, int K);
int b200_bl_lt_heuristic(int acc_bits, int layout, const void* A, const void* B, void* C, int M, int N, int K);
/* warm_rounds / bench_rounds <= 0 select the reference's 50 / 100. */
int b200_bl_lt_autotune_find(int acc_bits, int layout, int M, int N, int K, int warm_rounds, int bench_rounds);
int b200_bl_lt_autotune(int acc_bits, int layout, const void* A, const void* B, void* C, int M, int N, int K);
/* number of candidates the last find() examined and the winner's median time in ms */
int b200_bl_lt_autotune_info(int acc_bits, int layout, int* candidates, float* best_ms);

#ifdef __cplusplus
}
#endif
#endif /* B200_BASELINES_H_ */
