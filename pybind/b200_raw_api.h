// Raw (torch-free) entry points that the .cu translation units of one `hgemm_lib` build define and
// that pybind/hgemm_b200_fp{32,16}.cc wraps with the reference's torch::Tensor signatures.
// Keeping torch out of the .cu files is what makes a per-shape rebuild take seconds.
#pragma once

// layout: 0 = NN (B row-major [K,N]), 1 = TN (B K-major [N,K]); all pointers are device fp16
int b200raw_cublas_init();
void b200raw_cublas_destroy();
int b200raw_cublas_gemm(int layout, const void* A, const void* B, void* C, int M, int N, int K);

int b200raw_lt_heuristic_init();
void b200raw_lt_heuristic_destroy();
int b200raw_lt_heuristic_gemm(int layout, const void* A, const void* B, void* C, int M, int N, int K);

int b200raw_lt_autotune_init();
void b200raw_lt_autotune_destroy();
int b200raw_lt_autotune_find(int layout, int M, int N, int K);
int b200raw_lt_autotune_gemm(int layout, const void* A, const void* B, void* C, int M, int N, int K);

// the shape-specialised kernel of this build (kernels/b200_<acc dir>/<M>_<N>_<K>.cu)
extern "C" int b200_hgemm_shape_entry(const void* A, const void* B_kmajor, void* C, int M, int N, int K,
                                      void* stream);
extern "C" const char* b200_hgemm_shape_strerror(int status);
