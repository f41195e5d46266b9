// als_tc.cu -- tensor-core K1 (rank 64 / 128).  Placeholder until the tcgen05 kernel lands: reports
// "unsupported" so FMS_ALS_KERNEL_AUTO keeps using the generic CUDA-core kernel.
#include "common.cuh"

namespace fms {
int als_tc_supported(const fms_ctx *, int) { return 0; }
int als_launch_half_step_tc(fms_ctx *ctx, int, float, int64_t, int64_t) {
    return fail(ctx, FMS_ERR_UNSUPPORTED, "tensor-core ALS kernel not available for rank %d", ctx->als.rank);
}
int als_tc_refresh_planes(fms_ctx *, int) { return FMS_OK; }
}  // namespace fms
