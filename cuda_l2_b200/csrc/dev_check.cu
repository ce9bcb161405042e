// dev_check — standalone bring-up / tuning tool for libb200_hgemm.so (developer tool, not product).
//
//   dev_check check <acc_bits> <cfg|-1> <M> <N> <K>          exactness vs an independent GPU checker
//   dev_check time  <acc_bits> <cfg|-1> <M> <N> <K> [iters]  CUDA-event timing (+ cuBLAS for scale)
//   dev_check sweep <acc_bits> <M> <N> <K> [iters]           time every config and group_m variant
//
// Inputs are small integers, so every product and partial sum is exact in fp16 and fp32: any
// mismatch is a kernel bug, never rounding. C is surrounded by guard bands to catch stray writes.
#include <cublas_v2.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200_hgemm.h"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

__global__ void fill_ternary(__half* p, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t x = uint32_t(i) * 2654435761u ^ seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  p[i] = __float2half(float(int(x % 3u) - 1));
}
__global__ void fill_normalish(__half* p, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t x = uint32_t(i) * 2654435761u ^ seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  p[i] = __float2half((float(x & 0xffff) / 65536.f - 0.5f) * 2.f);
}
__global__ void fill_u16(uint16_t* p, size_t n, uint16_t v) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}
// independent checker: one thread per C element, fp32 accumulation, RN to fp16
__global__ void naive_tn(const __half* A, const __half* Bt, __half* C, int M, int N, int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= M || n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __half2float(A[size_t(m) * K + k]) * __half2float(Bt[size_t(n) * K + k]);
  C[size_t(m) * N + n] = __float2half_rn(acc);
}
__global__ void compare(const uint16_t* a, const uint16_t* b, size_t n, unsigned long long* nbad, unsigned long long* first) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint16_t x = a[i], y = b[i];
  if (x == y) return;
  if ((x & 0x7fff) == 0 && (y & 0x7fff) == 0) return;   // +0 vs -0
  atomicAdd(nbad, 1ull);
  atomicMin(first, (unsigned long long)i);
}
__global__ void check_guard(const uint16_t* p, size_t n, uint16_t v, unsigned long long* nbad) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n && p[i] != v) atomicAdd(nbad, 1ull);
}

static inline dim3 g1(size_t n) { return dim3(unsigned((n + 255) / 256)); }

struct Problem {
  int M, N, K;
  __half *A, *Bt, *Cbuf, *C, *Cref;
  static constexpr size_t kGuard = 16384;
  void alloc(int m, int n, int k) {
    M = m; N = n; K = k;
    CK(cudaMalloc(&A, size_t(M) * K * 2));
    CK(cudaMalloc(&Bt, size_t(N) * K * 2));
    CK(cudaMalloc(&Cbuf, (size_t(M) * N + 2 * kGuard) * 2));
    CK(cudaMalloc(&Cref, size_t(M) * N * 2));
    C = Cbuf + kGuard;
    fill_ternary<<<g1(size_t(M) * K), 256>>>(A, size_t(M) * K, 0x1234567u);
    fill_ternary<<<g1(size_t(N) * K), 256>>>(Bt, size_t(N) * K, 0x89abcdeu);
    CK(cudaDeviceSynchronize());
  }
  void reset_c() {
    fill_u16<<<g1(size_t(M) * N + 2 * kGuard), 256>>>((uint16_t*)Cbuf, size_t(M) * N + 2 * kGuard, 0x7bffu);
  }
  void release() { cudaFree(A); cudaFree(Bt); cudaFree(Cbuf); cudaFree(Cref); }
};

static cublasHandle_t g_blas;
static void cublas_tn(const Problem& p, __half* out) {
  // row-major C = A * Bt^T  <=>  column-major C^T[N,M] = Bt(op T)[N,K] * A^T[K,M]
  const float alpha = 1.f, beta = 0.f;
  cublasStatus_t s = cublasGemmEx(g_blas, CUBLAS_OP_T, CUBLAS_OP_N, p.N, p.M, p.K, &alpha, p.Bt, CUDA_R_16F, p.K,
                                  p.A, CUDA_R_16F, p.K, &beta, out, CUDA_R_16F, p.N, CUBLAS_COMPUTE_32F,
                                  CUBLAS_GEMM_DEFAULT_TENSOR_OP);
  if (s != CUBLAS_STATUS_SUCCESS) { printf("cublas error %d\n", int(s)); exit(3); }
}

static int run_ours(int acc, int cfg, const Problem& p, int group_m = 0) {
  if (cfg < 0)
    return acc == 32 ? b200_hgemm_f32acc(p.A, nullptr, p.Bt, p.C, p.M, p.N, p.K, nullptr)
                     : b200_hgemm_f16acc(p.A, nullptr, p.Bt, p.C, p.M, p.N, p.K, nullptr);
  return b200_hgemm_run_config(acc, cfg, p.A, p.Bt, p.C, p.M, p.N, p.K, group_m, 0, nullptr);
}

static int do_check(int acc, int cfg, int M, int N, int K) {
  Problem p; p.alloc(M, N, K);
  const bool use_naive = double(M) * N * K <= 2.2e10;
  if (use_naive) {
    dim3 b(32, 8), g((N + 31) / 32, (M + 7) / 8);
    naive_tn<<<g, b>>>(p.A, p.Bt, p.Cref, M, N, K);
  } else {
    cublas_tn(p, p.Cref);
  }
  CK(cudaDeviceSynchronize());
  p.reset_c();
  CK(cudaDeviceSynchronize());
  int st = run_ours(acc, cfg, p);
  cudaError_t e = cudaDeviceSynchronize();
  int sel = cfg < 0 ? b200_hgemm_select_config(acc, M, N, K) : cfg;
  if (st != 0 || e != cudaSuccess) {
    printf("CHECK acc=%d cfg=%d(%d) %dx%dx%d  LAUNCH-FAIL status=%d (%s) sync=%s\n", acc, cfg, sel, M, N, K, st,
           b200_hgemm_strerror(st), cudaGetErrorString(e));
    return 1;
  }
  unsigned long long *d, h[3] = {0, ~0ull, 0};
  CK(cudaMalloc(&d, 24));
  CK(cudaMemcpy(d, h, 24, cudaMemcpyHostToDevice));
  compare<<<g1(size_t(M) * N), 256>>>((uint16_t*)p.C, (uint16_t*)p.Cref, size_t(M) * N, d, d + 1);
  check_guard<<<g1(Problem::kGuard), 256>>>((uint16_t*)p.Cbuf, Problem::kGuard, 0x7bffu, d + 2);
  check_guard<<<g1(Problem::kGuard), 256>>>((uint16_t*)(p.C + size_t(M) * N), Problem::kGuard, 0x7bffu, d + 2);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost));
  const bool ok = h[0] == 0 && h[2] == 0;
  printf("CHECK acc=%d cfg=%d(%d) %dx%dx%d  %s mismatches=%llu first=(%lld,%lld) guard_bad=%llu checker=%s\n", acc, cfg,
         sel, M, N, K, ok ? "PASS" : "FAIL", h[0], h[0] ? (long long)(h[1] / N) : -1LL,
         h[0] ? (long long)(h[1] % N) : -1LL, h[2], use_naive ? "naive" : "cublas");
  if (!ok && h[0]) {
    // dump a small corner of both matrices around the first mismatch to make layout bugs readable
    size_t r0 = h[1] / N, c0 = (h[1] % N) & ~size_t(7);
    std::vector<__half> a(8), b(8);
    for (size_t r = r0; r < r0 + 4 && r < size_t(M); ++r) {
      CK(cudaMemcpy(a.data(), p.C + r * N + c0, 16, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(b.data(), p.Cref + r * N + c0, 16, cudaMemcpyDeviceToHost));
      printf("   row %zu col %zu..: got", r, c0);
      for (int i = 0; i < 8; ++i) printf(" %g", __half2float(a[i]));
      printf(" | want");
      for (int i = 0; i < 8; ++i) printf(" %g", __half2float(b[i]));
      printf("\n");
    }
  }
  cudaFree(d);
  p.release();
  fflush(stdout);
  return ok ? 0 : 1;
}

template <class F>
static float time_ms(F&& f, int iters, int warm = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < warm; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < iters; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a); cudaEventDestroy(b);
  return ms / iters;
}

static void alloc_random(Problem& p, int M, int N, int K) {
  p.alloc(M, N, K);
  fill_normalish<<<g1(size_t(M) * K), 256>>>(p.A, size_t(M) * K, 0x1234567u);
  fill_normalish<<<g1(size_t(N) * K), 256>>>(p.Bt, size_t(N) * K, 0x89abcdeu);
  CK(cudaDeviceSynchronize());
}

static int do_time(int acc, int cfg, int M, int N, int K, int iters) {
  Problem p; alloc_random(p, M, N, K);
  const double flops = 2.0 * M * N * K;
  int st = run_ours(acc, cfg, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (st != 0 || e != cudaSuccess) { printf("TIME launch fail %d %s\n", st, cudaGetErrorString(e)); return 1; }
  float ours = time_ms([&] { run_ours(acc, cfg, p); }, iters);
  float blas = time_ms([&] { cublas_tn(p, p.Cref); }, iters);
  int sel = cfg < 0 ? b200_hgemm_select_config(acc, M, N, K) : cfg;
  printf("TIME acc=%d cfg=%d(%d) %dx%dx%d  ours %.2f us %.1f TFLOP/s | cublas(fp32acc) %.2f us %.1f TFLOP/s | ratio %.3f\n",
         acc, cfg, sel, M, N, K, ours * 1e3, flops / ours * 1e-9, blas * 1e3, flops / blas * 1e-9, blas / ours);
  p.release();
  fflush(stdout);
  return 0;
}

static int do_sweep(int acc, int M, int N, int K, int iters) {
  Problem p; alloc_random(p, M, N, K);
  const double flops = 2.0 * M * N * K;
  float blas = time_ms([&] { cublas_tn(p, p.Cref); }, iters);
  printf("SWEEP acc=%d %dx%dx%d cublas %.2f us %.1f TFLOP/s\n", acc, M, N, K, blas * 1e3, flops / blas * 1e-9);
  const int ncfg = b200_hgemm_num_configs();
  const int gms[] = {1, 2, 4, 8, 16, 32};
  for (int c = 0; c < ncfg; ++c) {
    int bn, st_, cg; b200_hgemm_config_info(c, &bn, &st_, &cg);
    if (cg == 2 && M <= 128) continue;
    for (int gm : gms) {
      const int nm = (M + 128 * cg - 1) / (128 * cg);
      if (gm > 1 && gm / 2 >= nm) continue;   // wider than the problem: same schedule as the previous one
      int st = run_ours(acc, c, p, gm);
      cudaError_t e = cudaDeviceSynchronize();
      if (st != 0 || e != cudaSuccess) { printf("  cfg %d gm %d FAIL %d %s\n", c, gm, st, cudaGetErrorString(e)); return 1; }
      float t = time_ms([&] { run_ours(acc, c, p, gm); }, iters, 3);
      printf("  cfg=%d (BN=%d st=%d cg=%d) gm=%-2d  %.2f us  %.1f TFLOP/s  vs cublas %.3f\n", c, bn, st_, cg, gm,
             t * 1e3, flops / t * 1e-9, blas / t);
    }
  }
  p.release();
  fflush(stdout);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: see source header\n"); return 64; }
  CK(cudaSetDevice(0));
  cublasCreate(&g_blas);
  std::string mode = argv[1];
  if (mode == "check" && argc >= 7)
    return do_check(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]));
  if (mode == "time" && argc >= 7)
    return do_time(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), argc > 7 ? atoi(argv[7]) : 20);
  if (mode == "sweep" && argc >= 6)
    return do_sweep(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), argc > 6 ? atoi(argv[6]) : 20);
  printf("bad arguments\n");
  return 64;
}
