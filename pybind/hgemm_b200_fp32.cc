// torch extension entry for --device_type b200 --acc_precise fp32.
// Counterpart of the reference's pybind/hgemm_a100_fp32.cc: same 15 exported functions, the kernel symbol
// is cuda_l2_b200_fp32 (the harness dispatches on that name: benchmarking_utils.py:41).
#define B200_CUDA_L2_NAME cuda_l2_b200_fp32
#include "hgemm_b200_bindings.h"
