#include "gemm_tc.cuh"

namespace gatx {
int launch_gemm_tc_tn(const float*, int64_t, const float*, int64_t, float*, int64_t, int, int, int, bool,
                      cudaStream_t) {
  return -1;
}
int launch_gemm_tc_atb(const float*, int64_t, const float*, int64_t, float*, int64_t, int, int, int64_t, float*,
                       size_t, cudaStream_t) {
  return -1;
}
}  // namespace gatx
