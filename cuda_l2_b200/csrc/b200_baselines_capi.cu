// libb200_baselines.so — C-ABI wrappers over baselines.cuh (declared in include/b200_baselines.h).
#include "../../include/b200_baselines.h"

#include "baselines.cuh"

namespace {
template <int kAcc>
struct Set {
  b200bl::Cublas<kAcc> blas;
  b200bl::LtHeuristic<kAcc> heur;
  b200bl::LtAutoTune<kAcc> tune;
};
Set<32> g32;
Set<16> g16;
inline const __half* H(const void* p) { return static_cast<const __half*>(p); }
inline __half* H(void* p) { return static_cast<__half*>(p); }
inline b200bl::Layout L(int l) { return l ? b200bl::kTN : b200bl::kNN; }
}  // namespace

#define DISPATCH(expr32, expr16) (acc_bits == 32 ? (expr32) : acc_bits == 16 ? (expr16) : -1)

extern "C" {

int b200_bl_init(int acc_bits) {
  return DISPATCH(g32.blas.init() | g32.heur.init() | g32.tune.init(), g16.blas.init() | g16.heur.init() | g16.tune.init());
}
void b200_bl_destroy(int acc_bits) {
  if (acc_bits == 32) { g32.blas.destroy(); g32.heur.destroy(); g32.tune.destroy(); }
  if (acc_bits == 16) { g16.blas.destroy(); g16.heur.destroy(); g16.tune.destroy(); }
}
int b200_bl_cublas(int acc_bits, int layout, const void* A, const void* B, void* C, int M, int N, int K) {
  return DISPATCH(g32.blas.gemm(L(layout), H(A), H(B), H(C), M, N, K), g16.blas.gemm(L(layout), H(A), H(B), H(C), M, N, K));
}
int b200_bl_lt_heuristic(int acc_bits, int layout, const void* A, const void* B, void* C, int M, int N, int K) {
  return DISPATCH(g32.heur.gemm(L(layout), H(A), H(B), H(C), M, N, K), g16.heur.gemm(L(layout), H(A), H(B), H(C), M, N, K));
}
int b200_bl_lt_autotune_find(int acc_bits, int layout, int M, int N, int K, int warm_rounds, int bench_rounds) {
  if (warm_rounds <= 0) warm_rounds = 50;
  if (bench_rounds <= 0) bench_rounds = 100;
  return DISPATCH(g32.tune.find(L(layout), M, N, K, warm_rounds, bench_rounds),
                  g16.tune.find(L(layout), M, N, K, warm_rounds, bench_rounds));
}
int b200_bl_lt_autotune(int acc_bits, int layout, const void* A, const void* B, void* C, int M, int N, int K) {
  return DISPATCH(g32.tune.gemm(L(layout), H(A), H(B), H(C), M, N, K), g16.tune.gemm(L(layout), H(A), H(B), H(C), M, N, K));
}
int b200_bl_lt_autotune_info(int acc_bits, int layout, int* candidates, float* best_ms) {
  const int l = layout ? 1 : 0;
  if (acc_bits == 32) { if (candidates) *candidates = g32.tune.last_candidates[l]; if (best_ms) *best_ms = g32.tune.last_best_ms[l]; return 0; }
  if (acc_bits == 16) { if (candidates) *candidates = g16.tune.last_candidates[l]; if (best_ms) *best_ms = g16.tune.last_best_ms[l]; return 0; }
  return -1;
}

}  // extern "C"
