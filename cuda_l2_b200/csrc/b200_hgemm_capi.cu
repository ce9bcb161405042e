// libb200_hgemm.so — C-ABI entry points declared in include/b200_hgemm.h.
// Holds every kernel configuration and the per-shape dispatcher. No torch, no CUTLASS, no cuBLAS.
#include "../../include/b200_hgemm.h"

#include <atomic>
#include <cstdlib>

#include "hgemm_configs.cuh"
#include "hgemm_dispatch.cuh"

namespace {

std::atomic<unsigned long long> g_launches{0};

template <bool kAccF32>
int run_config(int id, const void* A, const void* Bt, void* C, int M, int N, int K, int group_m, int max_ctas,
               int splits, cudaStream_t s) {
  using namespace b200;
  int st;
  switch (id) {
#define B200_CASE(ID, BN, STAGES, CG, CM, CN, MR)                                                      \
  case ID:                                                                                         \
    st = host::launch<Config<BN, STAGES, CG, kAccF32, CM, CN, MR>>(A, Bt, C, M, N, K, s, group_m, max_ctas, splits); \
    break;
    B200_HGEMM_CONFIGS(B200_CASE)
#undef B200_CASE
    default:
      return host::kBadConfig;
  }
  if (st == host::kOk) g_launches.fetch_add(1, std::memory_order_relaxed);
  return st;
}

template <class Cfg>
int schedule_units(int M, int N, int K, int splits, int num_sms, int worker, int* units, int max_units,
                   int* num_workers, int* sk_tiles, int* contributors) {
  using namespace b200;
  const host::Plan p = host::make_plan<Cfg>(M, N, K, num_sms / Cfg::CLUSTER_CTAS, splits);
  if (num_workers) *num_workers = p.workers;
  if (sk_tiles) *sk_tiles = p.sk_tiles;
  if (worker < 0 || worker >= p.workers) return host::kBadShape;
  WorkIter it(worker, p.workers, p.num_tiles, p.nkb, p.splits, p.sk_tiles);
  WorkUnit u;
  int n = 0;
  while (it.next(u)) {
    if (n < max_units && units) {
      units[3 * n] = u.tile; units[3 * n + 1] = u.kb0; units[3 * n + 2] = u.kb1;
      if (contributors)
        contributors[n] = (p.sk_tiles && u.kb0 == 0 && u.kb1 < p.nkb)
                              ? streamk_contributors(worker, p.workers, p.sk_tiles * p.nkb, u.tile, p.nkb) : 0;
    }
    ++n;
  }
  return n;
}

int run(int acc_bits, int id, const void* A, const void* Bt, void* C, int M, int N, int K, int group_m,
        int max_ctas, int splits, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (acc_bits == 32) return run_config<true>(id, A, Bt, C, M, N, K, group_m, max_ctas, splits, s);
  if (acc_bits == 16) return run_config<false>(id, A, Bt, C, M, N, K, group_m, max_ctas, splits, s);
  return b200::host::kBadConfig;
}

}  // namespace

extern "C" {

int b200_hgemm_num_configs(void) { return b200::kNumConfigs; }

int b200_hgemm_config_info(int config_id, int* bn, int* stages, int* cta_group) {
  switch (config_id) {
#define B200_CASE(ID, BN, STAGES, CG, CM, CN, MR)  \
  case ID:                             \
    if (bn) *bn = BN;                  \
    if (stages) *stages = STAGES;      \
    if (cta_group) *cta_group = CG;    \
    return 0;
    B200_HGEMM_CONFIGS(B200_CASE)
#undef B200_CASE
    default:
      return b200::host::kBadConfig;
  }
}

int b200_hgemm_config_cluster(int config_id, int* cluster_m, int* cluster_n) {
  switch (config_id) {
#define B200_CASE(ID, BN, STAGES, CG, CM, CN, MR) \
  case ID:                                    \
    if (cluster_m) *cluster_m = CM;           \
    if (cluster_n) *cluster_n = CN;           \
    return 0;
    B200_HGEMM_CONFIGS(B200_CASE)
#undef B200_CASE
    default:
      return b200::host::kBadConfig;
  }
}

int b200_hgemm_schedule_units(int config_id, int M, int N, int K, int splits, int num_sms, int worker, int* units,
                              int max_units, int* num_workers, int* sk_tiles, int* contributors) {
  if (M <= 0 || N <= 0 || K <= 0 || num_sms <= 0) return b200::host::kBadShape;
  switch (config_id) {
#define B200_CASE(ID, BN, STAGES, CG, CM, CN, MR)                                                                   \
  case ID:                                                                                                          \
    return schedule_units<b200::Config<BN, STAGES, CG, true, CM, CN, MR>>(M, N, K, splits, num_sms, worker, units, \
                                                                          max_units, num_workers, sk_tiles, contributors);
    B200_HGEMM_CONFIGS(B200_CASE)
#undef B200_CASE
    default:
      return b200::host::kBadConfig;
  }
}

int b200_hgemm_config_m_rep(int config_id) {
  switch (config_id) {
#define B200_CASE(ID, BN, STAGES, CG, CM, CN, MR) \
  case ID:                                        \
    return MR;
    B200_HGEMM_CONFIGS(B200_CASE)
#undef B200_CASE
    default:
      return b200::host::kBadConfig;
  }
}

int b200_hgemm_select_config(int acc_bits, int M, int N, int K) {
  if (acc_bits != 32 && acc_bits != 16) return b200::host::kBadConfig;
  if (M <= 0 || N <= 0 || K <= 0) return b200::host::kBadShape;
  return b200::dispatch::select(acc_bits, M, N, K).config_id;
}

int b200_hgemm_select(int acc_bits, int M, int N, int K, int* config_id, int* group_m, int* splits) {
  if (acc_bits != 32 && acc_bits != 16) return b200::host::kBadConfig;
  if (M <= 0 || N <= 0 || K <= 0) return b200::host::kBadShape;
  const b200::dispatch::Choice ch = b200::dispatch::select(acc_bits, M, N, K);
  if (config_id) *config_id = ch.config_id;
  if (group_m) *group_m = ch.group_m;
  if (splits) *splits = ch.splits;
  return 0;
}

int b200_hgemm_run_config(int acc_bits, int config_id, const void* A, const void* B_kmajor, void* C, int M,
                          int N, int K, int group_m, int max_ctas, int splits, void* stream) {
  return run(acc_bits, config_id, A, B_kmajor, C, M, N, K, group_m, max_ctas, splits, stream);
}

int b200_hgemm_f32acc(const void* A, const void* /*B_rowmajor*/, const void* B_kmajor, void* C, int M, int N,
                      int K, void* stream) {
  int st = b200::host::validate(A, B_kmajor, C, M, N, K);
  if (st) return st;
  const b200::dispatch::Choice ch = b200::dispatch::select(32, M, N, K);
  return run(32, ch.config_id, A, B_kmajor, C, M, N, K, ch.group_m, 0, ch.splits, stream);
}

int b200_hgemm_f16acc(const void* A, const void* /*B_rowmajor*/, const void* B_kmajor, void* C, int M, int N,
                      int K, void* stream) {
  int st = b200::host::validate(A, B_kmajor, C, M, N, K);
  if (st) return st;
  const b200::dispatch::Choice ch = b200::dispatch::select(16, M, N, K);
  return run(16, ch.config_id, A, B_kmajor, C, M, N, K, ch.group_m, 0, ch.splits, stream);
}

int b200_hgemm_host(int acc_bits, const void* hA, const void* hB_kmajor, void* hC, int M, int N, int K) {
  if (!hA || !hB_kmajor || !hC) return b200::host::kNullPointer;
  if (M <= 0 || N <= 0 || K <= 0) return b200::host::kBadShape;
  // device scratch grows monotonically and is reused across calls
  static thread_local void* dbuf = nullptr;
  static thread_local size_t dcap = 0;
  const size_t a_bytes = size_t(M) * K * 2, b_bytes = size_t(N) * K * 2, c_bytes = size_t(M) * N * 2;
  auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t need = up(a_bytes) + up(b_bytes) + up(c_bytes);
  if (need > dcap) {
    if (dbuf) cudaFree(dbuf);
    dbuf = nullptr; dcap = 0;
    cudaError_t e = cudaMalloc(&dbuf, need);
    if (e != cudaSuccess) return int(e);
    dcap = need;
  }
  char* dA = static_cast<char*>(dbuf);
  char* dB = dA + up(a_bytes);
  char* dC = dB + up(b_bytes);
  if (acc_bits != 32 && acc_bits != 16) return b200::host::kBadConfig;
  auto gemm = [&](const void* a, void* c, int m, cudaStream_t s) {
    return acc_bits == 32 ? b200_hgemm_f32acc(a, nullptr, dB, c, m, N, K, s) : b200_hgemm_f16acc(a, nullptr, dB, c, m, N, K, s);
  };
  cudaError_t e;

  // Large problems are PCIe time: B goes first, then A in row blocks; the GEMM of block i runs while block i+1 is on
  // its way in and block i-1 on its way out (PCIe is full duplex), on three private streams joined before returning.
  // Row blocks are independent GEMMs (C_i = A_i * B), so the result does not depend on the blocking.
  constexpr int kBlocks = 4;
  const bool pipelined = M >= kBlocks * 256 && (a_bytes + c_bytes) >= (size_t(8) << 20) &&
                         !(std::getenv("B200_HGEMM_HOST_UNPIPELINED"));
  if (!pipelined) {
    if ((e = cudaMemcpyAsync(dA, hA, a_bytes, cudaMemcpyHostToDevice, 0)) != cudaSuccess) return int(e);
    if ((e = cudaMemcpyAsync(dB, hB_kmajor, b_bytes, cudaMemcpyHostToDevice, 0)) != cudaSuccess) return int(e);
    int st = gemm(dA, dC, M, nullptr);
    if (st) return st;
    if ((e = cudaMemcpyAsync(hC, dC, c_bytes, cudaMemcpyDeviceToHost, 0)) != cudaSuccess) return int(e);
    e = cudaStreamSynchronize(0);
    return e == cudaSuccess ? 0 : int(e);
  }

  struct Lanes { cudaStream_t in = nullptr, run = nullptr, out = nullptr; cudaEvent_t b_in = nullptr, a_in[kBlocks] = {}, done[kBlocks] = {}; };
  static thread_local Lanes L;
  if (!L.in) {
    if ((e = cudaStreamCreateWithFlags(&L.in, cudaStreamNonBlocking)) != cudaSuccess) return int(e);
    if ((e = cudaStreamCreateWithFlags(&L.run, cudaStreamNonBlocking)) != cudaSuccess) return int(e);
    if ((e = cudaStreamCreateWithFlags(&L.out, cudaStreamNonBlocking)) != cudaSuccess) return int(e);
    if ((e = cudaEventCreateWithFlags(&L.b_in, cudaEventDisableTiming)) != cudaSuccess) return int(e);
    for (int i = 0; i < kBlocks; ++i) {
      if ((e = cudaEventCreateWithFlags(&L.a_in[i], cudaEventDisableTiming)) != cudaSuccess) return int(e);
      if ((e = cudaEventCreateWithFlags(&L.done[i], cudaEventDisableTiming)) != cudaSuccess) return int(e);
    }
  }
  // work queued by the caller on the legacy default stream (non-blocking streams do not wait for it on their own)
  if ((e = cudaStreamSynchronize(0)) != cudaSuccess) return int(e);
  if ((e = cudaMemcpyAsync(dB, hB_kmajor, b_bytes, cudaMemcpyHostToDevice, L.in)) != cudaSuccess) return int(e);
  if ((e = cudaEventRecord(L.b_in, L.in)) != cudaSuccess) return int(e);
  if ((e = cudaStreamWaitEvent(L.run, L.b_in, 0)) != cudaSuccess) return int(e);
  const int rows_per = ((M + kBlocks - 1) / kBlocks + 127) / 128 * 128;   // whole 128-row tiles per block
  for (int i = 0; i < kBlocks; ++i) {
    const int r0 = i * rows_per, rows = std::min(rows_per, M - r0);
    if (rows <= 0) break;
    const size_t a_off = size_t(r0) * K * 2, c_off = size_t(r0) * N * 2;
    if ((e = cudaMemcpyAsync(dA + a_off, static_cast<const char*>(hA) + a_off, size_t(rows) * K * 2, cudaMemcpyHostToDevice, L.in)) != cudaSuccess) return int(e);
    if ((e = cudaEventRecord(L.a_in[i], L.in)) != cudaSuccess) return int(e);
    if ((e = cudaStreamWaitEvent(L.run, L.a_in[i], 0)) != cudaSuccess) return int(e);
    int st = gemm(dA + a_off, dC + c_off, rows, L.run);
    if (st) { cudaDeviceSynchronize(); return st; }
    if ((e = cudaEventRecord(L.done[i], L.run)) != cudaSuccess) return int(e);
    if ((e = cudaStreamWaitEvent(L.out, L.done[i], 0)) != cudaSuccess) return int(e);
    if ((e = cudaMemcpyAsync(static_cast<char*>(hC) + c_off, dC + c_off, size_t(rows) * N * 2, cudaMemcpyDeviceToHost, L.out)) != cudaSuccess) return int(e);
  }
  if ((e = cudaStreamSynchronize(L.out)) != cudaSuccess) return int(e);   // the last copy out is behind everything else
  if ((e = cudaStreamSynchronize(L.run)) != cudaSuccess) return int(e);
  e = cudaStreamSynchronize(L.in);
  return e == cudaSuccess ? 0 : int(e);
}

#ifdef B200_HGEMM_TRACE
// Developer-only entry points of libb200_hgemm_trace.so (see kTraceSlots in hgemm_sm100.cuh); not part of the C ABI.
static unsigned long long* g_trace_dev = nullptr;
static int g_trace_ctas = 0;
int b200_hgemm_trace_slots(void) { return b200::kTraceSlots; }
int b200_hgemm_trace_arm(int max_ctas) {   // zero the buffer and point the kernels at it (max_ctas <= 0: disarm)
  unsigned long long* none = nullptr;
  if (max_ctas <= 0) return int(cudaMemcpyToSymbol(b200::g_trace_buf, &none, sizeof(none)));
  const size_t bytes = size_t(max_ctas) * b200::kTraceSlots * 2 * sizeof(unsigned long long);
  if (max_ctas > g_trace_ctas) {
    if (g_trace_dev) cudaFree(g_trace_dev);
    g_trace_dev = nullptr; g_trace_ctas = 0;
    cudaError_t e = cudaMalloc(&g_trace_dev, bytes);
    if (e != cudaSuccess) return int(e);
    g_trace_ctas = max_ctas;
  }
  cudaError_t e = cudaMemset(g_trace_dev, 0, bytes);
  if (e != cudaSuccess) return int(e);
  return int(cudaMemcpyToSymbol(b200::g_trace_buf, &g_trace_dev, sizeof(g_trace_dev)));
}
int b200_hgemm_trace_read(unsigned long long* out, int ctas) {
  if (!g_trace_dev || ctas > g_trace_ctas) return b200::host::kBadShape;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return int(e);
  return int(cudaMemcpy(out, g_trace_dev, size_t(ctas) * b200::kTraceSlots * 2 * sizeof(unsigned long long),
                        cudaMemcpyDeviceToHost));
}
#endif

unsigned long long b200_hgemm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* b200_hgemm_strerror(int status) { return b200::host::status_string(status); }

}  // extern "C"
