// temporary stubs (replaced as the real kernels land)
#include "common.cuh"
bool supports_mul_mat_id(const b200_op *) { return false; }
int op_mul_mat_id(b200_ctx *, const b200_op *) { return B200_ERR_UNSUPPORTED; }
int launch_gemm_i8(b200_ctx *, int, const uint8_t *, size_t, int64_t, int64_t, const uint8_t *, int64_t, float *, size_t, bool *handled) { *handled = false; return B200_OK; }
