// The family of kernel configurations compiled into libb200_hgemm.so and referenced by the
// generated per-shape translation units (kernels/b200_*/<M>_<N>_<K>.cu).
//   X(id, BN, STAGES, CTA_GROUP)
// Stage counts fill the 227 KB of shared memory left after the 32 KB epilogue staging area.
#pragma once
#include "hgemm_host.cuh"

#define B200_HGEMM_CONFIGS(X) \
  X(0, 256, 4, 1)             \
  X(1, 128, 6, 1)             \
  X(2, 64, 8, 1)              \
  X(3, 256, 6, 2)             \
  X(4, 128, 8, 2)             \
  X(5, 192, 4, 1)             \
  X(6, 192, 6, 2)

namespace b200 {
constexpr int kNumConfigs = 7;
}
