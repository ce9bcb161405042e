// Comparators for the benchmark harness: the same GEMM through cuBLAS and cuBLASLt.
// These are library calls, not part of the accelerated path. Semantics follow the reference's baselines
//   cublas/{fp32,fp16}/hgemm_cublas.cu:41-68                 cublasGemmEx, NN and TN, COMPUTE_32F / COMPUTE_16F
//   cublas/{fp32,fp16}/hgemm_cublaslt_heuristic.cu:65-217     first of 4 heuristic algorithms, cached descriptors
//   cublas/{fp32,fp16}/hgemm_cublaslt_auto_tuning.cu:108-306  up to 100 heuristic candidates, 50 warm-up +
//        100 timed rounds in shuffled order on fresh random data, best median wins
// but the code is ours: one templated implementation, raw device pointers (no torch in this file).
//
// Row-major C[M,N] = A[M,K] * B[K,N] is computed as column-major C^T = B^T * A^T:
//   NN: B row-major [K,N]  == column-major [N,K] ld N, no transpose
//   TN: B K-major   [N,K]  == column-major [K,N] ld K, transposed
//
// kAccBits = 32: fp32 compute/scale type; 16: fp16 compute and fp16 alpha/beta.
#pragma once
#include <cublasLt.h>
#include <cublas_v2.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

namespace b200bl {

enum Layout { kNN = 0, kTN = 1 };

inline size_t env_mb(const char* name, size_t dflt_bytes) {
  const char* v = std::getenv(name);
  return v ? size_t(std::strtoull(v, nullptr, 10)) << 20 : dflt_bytes;
}

template <int kAccBits>
struct Scalars {
  float one_f = 1.f, zero_f = 0.f;
  __half one_h = __float2half(1.f), zero_h = __float2half(0.f);
  const void* alpha() const { return kAccBits == 32 ? (const void*)&one_f : (const void*)&one_h; }
  const void* beta() const { return kAccBits == 32 ? (const void*)&zero_f : (const void*)&zero_h; }
  static constexpr cublasComputeType_t compute() { return kAccBits == 32 ? CUBLAS_COMPUTE_32F : CUBLAS_COMPUTE_16F; }
  static constexpr cudaDataType_t scale() { return kAccBits == 32 ? CUDA_R_32F : CUDA_R_16F; }
};

// ---------------------------------------------------------------------------------------- cuBLAS
template <int kAccBits>
struct Cublas {
  cublasHandle_t h = nullptr;
  int init() {
    if (h) return 0;
    if (cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) { h = nullptr; return 1; }
    cublasSetMathMode(h, CUBLAS_TENSOR_OP_MATH);
    return 0;
  }
  void destroy() { if (h) { cublasDestroy(h); h = nullptr; } }
  int gemm(Layout lay, const __half* A, const __half* B, __half* C, int M, int N, int K) {
    if (init()) return 1;
    Scalars<kAccBits> s;
    cublasStatus_t st = cublasGemmEx(h, lay == kTN ? CUBLAS_OP_T : CUBLAS_OP_N, CUBLAS_OP_N, N, M, K, s.alpha(), B,
                                     CUDA_R_16F, lay == kTN ? K : N, A, CUDA_R_16F, K, s.beta(), C, CUDA_R_16F, N,
                                     Scalars<kAccBits>::compute(), CUBLAS_GEMM_DEFAULT_TENSOR_OP);
    return st == CUBLAS_STATUS_SUCCESS ? 0 : int(st);
  }
};

// ---------------------------------------------------------------------------------------- cuBLASLt plans
template <int kAccBits>
struct LtPlan {
  cublasLtMatmulDesc_t op = nullptr;
  cublasLtMatrixLayout_t lb = nullptr, la = nullptr, lc = nullptr;
  cublasLtMatmulAlgo_t algo{};
  bool has_algo = false;
  int M = 0, N = 0, K = 0;

  void release() {
    if (op) cublasLtMatmulDescDestroy(op);
    if (lb) cublasLtMatrixLayoutDestroy(lb);
    if (la) cublasLtMatrixLayoutDestroy(la);
    if (lc) cublasLtMatrixLayoutDestroy(lc);
    op = nullptr; lb = la = lc = nullptr; has_algo = false; M = N = K = 0;
  }
  int describe(Layout lay, int m, int n, int k) {
    release();
    if (cublasLtMatmulDescCreate(&op, Scalars<kAccBits>::compute(), Scalars<kAccBits>::scale()) != CUBLAS_STATUS_SUCCESS)
      return 1;
    cublasOperation_t ta = lay == kTN ? CUBLAS_OP_T : CUBLAS_OP_N, tb = CUBLAS_OP_N;
    cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta));
    cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb));
    if (lay == kTN) cublasLtMatrixLayoutCreate(&lb, CUDA_R_16F, k, n, k);
    else cublasLtMatrixLayoutCreate(&lb, CUDA_R_16F, n, k, n);
    cublasLtMatrixLayoutCreate(&la, CUDA_R_16F, k, m, k);
    cublasLtMatrixLayoutCreate(&lc, CUDA_R_16F, n, m, n);
    M = m; N = n; K = k;
    return 0;
  }
  int candidates(cublasLtHandle_t h, size_t ws_bytes, int want, std::vector<cublasLtMatmulHeuristicResult_t>& out) {
    cublasLtMatmulPreference_t pref = nullptr;
    cublasLtMatmulPreferenceCreate(&pref);
    cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes));
    out.resize(want);
    int got = 0;
    cublasStatus_t st = cublasLtMatmulAlgoGetHeuristic(h, op, lb, la, lc, lc, pref, want, out.data(), &got);
    cublasLtMatmulPreferenceDestroy(pref);
    out.resize(st == CUBLAS_STATUS_SUCCESS ? got : 0);
    return int(out.size());
  }
  int run(cublasLtHandle_t h, const cublasLtMatmulAlgo_t* a, const __half* A, const __half* B, __half* C, void* ws,
          size_t ws_bytes, cudaStream_t stream) {
    Scalars<kAccBits> s;
    cublasStatus_t st = cublasLtMatmul(h, op, s.alpha(), B, lb, A, la, s.beta(), C, lc, C, lc, a, ws, ws_bytes, stream);
    return st == CUBLAS_STATUS_SUCCESS ? 0 : int(st);
  }
};

// ---------------------------------------------------------------------------------------- heuristic (V1)
template <int kAccBits>
struct LtHeuristic {
  cublasLtHandle_t h = nullptr;
  void* ws = nullptr;
  // The reference asks for "20 GB" with int arithmetic that wraps to 0 (hgemm_cublaslt_heuristic.cu:18),
  // so its heuristic baseline effectively runs without workspace. Same default here; override with
  // B200_BL_HEUR_WORKSPACE_MB to give the heuristic path a real workspace.
  size_t ws_bytes = 0;
  LtPlan<kAccBits> plan[2];
  int init() {
    if (h) return 0;
    if (cublasLtCreate(&h) != CUBLAS_STATUS_SUCCESS) { h = nullptr; return 1; }
    ws_bytes = env_mb("B200_BL_HEUR_WORKSPACE_MB", 0);
    if (ws_bytes && cudaMalloc(&ws, ws_bytes) != cudaSuccess) { ws = nullptr; ws_bytes = 0; }
    return 0;
  }
  void destroy() {
    plan[0].release(); plan[1].release();
    if (h) { cublasLtDestroy(h); h = nullptr; }
    if (ws) { cudaFree(ws); ws = nullptr; }
  }
  int gemm(Layout lay, const __half* A, const __half* B, __half* C, int M, int N, int K) {
    if (init()) return 1;
    LtPlan<kAccBits>& p = plan[lay];
    if (!(p.has_algo && p.M == M && p.N == N && p.K == K)) {
      if (p.describe(lay, M, N, K)) return 2;
      std::vector<cublasLtMatmulHeuristicResult_t> c;
      if (p.candidates(h, ws_bytes, 4, c) == 0) return 3;
      p.algo = c[0].algo;
      p.has_algo = true;
    }
    return p.run(h, &p.algo, A, B, C, ws, ws_bytes, 0);
  }
};

// ---------------------------------------------------------------------------------------- auto-tuning (V2)
static __global__ void bl_fill_uniform(__half* p, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t x = uint32_t(i) * 2654435761u + seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  p[i] = __float2half(float(x >> 8) * (2.0f / 16777216.0f) - 1.0f);   // uniform in [-1, 1)
}

template <int kAccBits>
struct LtAutoTune {
  cublasLtHandle_t h = nullptr;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  LtPlan<kAccBits> plan[2];
  int last_candidates[2] = {0, 0};
  float last_best_ms[2] = {0.f, 0.f};
  int init() {
    if (h) return 0;
    if (cublasLtCreate(&h) != CUBLAS_STATUS_SUCCESS) { h = nullptr; return 1; }
    ws_bytes = env_mb("B200_BL_TUNE_WORKSPACE_MB", size_t(20) << 30);   // reference: 20 GB (auto_tuning.cu:23)
    if (cudaMalloc(&ws, ws_bytes) != cudaSuccess) {
      cudaGetLastError();
      ws_bytes = size_t(1) << 30;
      if (cudaMalloc(&ws, ws_bytes) != cudaSuccess) { cudaGetLastError(); ws = nullptr; ws_bytes = 0; }
    }
    return 0;
  }
  void destroy() {
    plan[0].release(); plan[1].release();
    if (h) { cublasLtDestroy(h); h = nullptr; }
    if (ws) { cudaFree(ws); ws = nullptr; }
  }
  // Time every candidate `bench_rounds` times (after `warm_rounds`), each round on new random operands and in
  // a new random order, preceded by one untimed call; keep the candidate with the smallest median.
  int find(Layout lay, int M, int N, int K, int warm_rounds = 50, int bench_rounds = 100, int max_algos = 100) {
    if (init()) return 1;
    LtPlan<kAccBits>& p = plan[lay];
    if (p.describe(lay, M, N, K)) return 2;
    std::vector<cublasLtMatmulHeuristicResult_t> cand;
    const int n = p.candidates(h, ws_bytes, max_algos, cand);
    last_candidates[lay] = n;
    if (n == 0) return 3;
    __half *a = nullptr, *b = nullptr, *c = nullptr;
    if (cudaMalloc(&a, size_t(M) * K * 2) != cudaSuccess || cudaMalloc(&b, size_t(K) * N * 2) != cudaSuccess ||
        cudaMalloc(&c, size_t(M) * N * 2) != cudaSuccess) {
      cudaFree(a); cudaFree(b); cudaFree(c);
      return 4;
    }
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<std::vector<float>> t(n);
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::mt19937 rng(std::random_device{}());
    std::vector<char> usable(n, 1);
    for (int r = 0; r < warm_rounds + bench_rounds; ++r) {
      const size_t na = size_t(M) * K, nb = size_t(K) * N;
      bl_fill_uniform<<<unsigned((na + 255) / 256), 256, 0, s>>>(a, na, 0x9e3779b9u * uint32_t(2 * r + 1));
      bl_fill_uniform<<<unsigned((nb + 255) / 256), 256, 0, s>>>(b, nb, 0x85ebca6bu * uint32_t(2 * r + 2));
      std::shuffle(order.begin(), order.end(), rng);
      p.run(h, &cand[order[n - 1]].algo, a, b, c, ws, ws_bytes, s);
      cudaStreamSynchronize(s);
      for (int i = 0; i < n; ++i) {
        const int id = order[i];
        if (!usable[id]) continue;
        cudaEventRecord(e0, s);
        const int st = p.run(h, &cand[id].algo, a, b, c, ws, ws_bytes, s);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        if (st) { usable[id] = 0; continue; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= warm_rounds) t[id].push_back(ms);
      }
    }
    int best = -1;
    float best_ms = 0.f;
    for (int id = 0; id < n; ++id) {
      if (!usable[id] || t[id].empty()) continue;
      std::vector<float>& v = t[id];
      std::sort(v.begin(), v.end());
      const size_t h2 = v.size() / 2;
      const float med = (v.size() % 2) ? v[h2] : 0.5f * (v[h2 - 1] + v[h2]);
      if (best < 0 || med < best_ms) { best = id; best_ms = med; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamDestroy(s);
    cudaFree(a); cudaFree(b); cudaFree(c);
    if (best < 0) return 5;
    p.algo = cand[best].algo;
    p.has_algo = true;
    last_best_ms[lay] = best_ms;
    return 0;
  }
  int gemm(Layout lay, const __half* A, const __half* B, __half* C, int M, int N, int K) {
    LtPlan<kAccBits>& p = plan[lay];
    if (!h || !(p.has_algo && p.M == M && p.N == N && p.K == K)) return 6;   // find() must come first
    return p.run(h, &p.algo, A, B, C, ws, ws_bytes, 0);
  }
};

}  // namespace b200bl
