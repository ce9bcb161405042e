// torch extension entry for --device_type b200 --acc_precise fp16.
// Counterpart of the reference's pybind/hgemm_a100_fp16.cc: same 15 exported functions, the kernel symbol
// is cuda_l2_b200_fp16 (the harness dispatches on that name: benchmarking_utils.py:41).
#define B200_CUDA_L2_NAME cuda_l2_b200_fp16
#include "hgemm_b200_bindings.h"
