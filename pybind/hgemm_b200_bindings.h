// torch-extension surface of one `hgemm_lib` build for --device_type b200.
//
// Exports exactly the 15 functions the reference's pybind/hgemm_<dev>_<acc>.cc exports
// (pybind/hgemm_a100_fp32.cc:9-27 declarations, :29-52 registration), with the same names, argument
// meaning and error behaviour (a C++ exception -> Python RuntimeError):
//   init_cublas_handle, destroy_cublas_handle, hgemm_cublas_nn, hgemm_cublas_tn,
//   init_cublaslt_handle_v1, destroy_cublaslt_handle_v1, hgemm_cublaslt_heuristic_nn/_tn,
//   init_cublaslt_handle_v2, destroy_cublaslt_handle_v2, find_best_algo_nn_v2_torch, find_best_algo_tn_v2_torch,
//   hgemm_cublaslt_auto_tuning_nn/_tn, and cuda_l2_b200_<acc>(a, b, b_col_major, c).
// The including .cc defines B200_CUDA_L2_NAME (the exported kernel name) before including this file.
#pragma once
#include <torch/extension.h>

#include <stdexcept>
#include <string>

#include "b200_raw_api.h"

namespace b200_bind {

inline void require(bool ok, const char* msg) {
  if (!ok) throw std::runtime_error(msg);
}
inline void check_half_cuda(const torch::Tensor& t, const char* name) {
  if (t.scalar_type() != torch::kHalf) throw std::runtime_error(std::string(name) + " must be a torch.half tensor");
  if (!t.is_cuda()) throw std::runtime_error(std::string(name) + " must live on the GPU");
  if (!t.is_contiguous()) throw std::runtime_error(std::string(name) + " must be contiguous");
  if (t.dim() != 2) throw std::runtime_error(std::string(name) + " must be 2-D");
}
struct Dims { int M, N, K; };
// a [M,K]; bmat labelled [K,N] (row-major b, or b_col_major whose storage is [N,K]); c [M,N]
inline Dims dims_of(const torch::Tensor& a, const torch::Tensor& bmat, const torch::Tensor& c) {
  check_half_cuda(a, "a");
  check_half_cuda(bmat, "b");
  check_half_cuda(c, "c");
  Dims d{int(a.size(0)), int(bmat.size(1)), int(a.size(1))};
  require(bmat.size(0) == d.K, "Tensor size mismatch!");
  require(c.size(0) == d.M && c.size(1) == d.N, "Tensor size mismatch!");
  return d;
}
inline void lib_ok(int status, const char* what) {
  if (status != 0) throw std::runtime_error(std::string(what) + " failed with status " + std::to_string(status));
}

}  // namespace b200_bind

void init_cublas_handle() { b200_bind::lib_ok(b200raw_cublas_init(), "cublasCreate"); }
void destroy_cublas_handle() { b200raw_cublas_destroy(); }
void hgemm_cublas_nn(torch::Tensor a, torch::Tensor b, torch::Tensor b_col_major, torch::Tensor c) {
  auto d = b200_bind::dims_of(a, b, c);
  b200_bind::lib_ok(b200raw_cublas_gemm(0, a.data_ptr(), b.data_ptr(), c.data_ptr(), d.M, d.N, d.K), "cublasGemmEx");
}
void hgemm_cublas_tn(torch::Tensor a, torch::Tensor b, torch::Tensor b_col_major, torch::Tensor c) {
  auto d = b200_bind::dims_of(a, b_col_major, c);
  b200_bind::lib_ok(b200raw_cublas_gemm(1, a.data_ptr(), b_col_major.data_ptr(), c.data_ptr(), d.M, d.N, d.K), "cublasGemmEx");
}

void init_cublaslt_handle_v1() { b200_bind::lib_ok(b200raw_lt_heuristic_init(), "cublasLtCreate"); }
void destroy_cublaslt_handle_v1() { b200raw_lt_heuristic_destroy(); }
void hgemm_cublaslt_heuristic_nn(torch::Tensor a, torch::Tensor b, torch::Tensor b_col_major, torch::Tensor c) {
  auto d = b200_bind::dims_of(a, b, c);
  b200_bind::lib_ok(b200raw_lt_heuristic_gemm(0, a.data_ptr(), b.data_ptr(), c.data_ptr(), d.M, d.N, d.K), "cublasLtMatmul");
}
void hgemm_cublaslt_heuristic_tn(torch::Tensor a, torch::Tensor b, torch::Tensor b_col_major, torch::Tensor c) {
  auto d = b200_bind::dims_of(a, b_col_major, c);
  b200_bind::lib_ok(b200raw_lt_heuristic_gemm(1, a.data_ptr(), b_col_major.data_ptr(), c.data_ptr(), d.M, d.N, d.K), "cublasLtMatmul");
}

void init_cublaslt_handle_v2() { b200_bind::lib_ok(b200raw_lt_autotune_init(), "cublasLtCreate"); }
void destroy_cublaslt_handle_v2() { b200raw_lt_autotune_destroy(); }
void find_best_algo_nn_v2_torch(int M, int N, int K) { b200_bind::lib_ok(b200raw_lt_autotune_find(0, M, N, K), "cuBLASLt auto-tuning (NN)"); }
void find_best_algo_tn_v2_torch(int M, int N, int K) { b200_bind::lib_ok(b200raw_lt_autotune_find(1, M, N, K), "cuBLASLt auto-tuning (TN)"); }
void hgemm_cublaslt_auto_tuning_nn(torch::Tensor a, torch::Tensor b, torch::Tensor b_col_major, torch::Tensor c) {
  auto d = b200_bind::dims_of(a, b, c);
  b200_bind::lib_ok(b200raw_lt_autotune_gemm(0, a.data_ptr(), b.data_ptr(), c.data_ptr(), d.M, d.N, d.K), "cublasLtMatmul");
}
void hgemm_cublaslt_auto_tuning_tn(torch::Tensor a, torch::Tensor b, torch::Tensor b_col_major, torch::Tensor c) {
  auto d = b200_bind::dims_of(a, b_col_major, c);
  b200_bind::lib_ok(b200raw_lt_autotune_gemm(1, a.data_ptr(), b_col_major.data_ptr(), c.data_ptr(), d.M, d.N, d.K), "cublasLtMatmul");
}

// The kernel under test. Reads `a` and `b_col_major`, writes only `c`; launched on the legacy default
// stream like every reference kernel (kernels/a100_F32F16F16F32/4096_4096_4096.cu:275-278), asynchronously.
void B200_CUDA_L2_NAME(torch::Tensor a, torch::Tensor b, torch::Tensor b_col_major, torch::Tensor c) {
  auto d = b200_bind::dims_of(a, b_col_major, c);
  b200_bind::check_half_cuda(b, "b");
  const int st = b200_hgemm_shape_entry(a.data_ptr(), b_col_major.data_ptr(), c.data_ptr(), d.M, d.N, d.K, nullptr);
  if (st != 0) throw std::runtime_error(std::string("b200 hgemm: ") + b200_hgemm_shape_strerror(st));
}

#define B200_STR2(x) #x
#define B200_STR(x) B200_STR2(x)
#define B200_DEF(m, f) m.def(#f, &f, #f)

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  B200_DEF(m, init_cublas_handle);
  B200_DEF(m, destroy_cublas_handle);
  B200_DEF(m, hgemm_cublas_nn);
  B200_DEF(m, hgemm_cublas_tn);
  B200_DEF(m, init_cublaslt_handle_v1);
  B200_DEF(m, destroy_cublaslt_handle_v1);
  B200_DEF(m, hgemm_cublaslt_heuristic_nn);
  B200_DEF(m, hgemm_cublaslt_heuristic_tn);
  B200_DEF(m, init_cublaslt_handle_v2);
  B200_DEF(m, destroy_cublaslt_handle_v2);
  B200_DEF(m, find_best_algo_nn_v2_torch);
  B200_DEF(m, find_best_algo_tn_v2_torch);
  B200_DEF(m, hgemm_cublaslt_auto_tuning_nn);
  B200_DEF(m, hgemm_cublaslt_auto_tuning_tn);
  m.def(B200_STR(B200_CUDA_L2_NAME), &B200_CUDA_L2_NAME, B200_STR(B200_CUDA_L2_NAME));
}
