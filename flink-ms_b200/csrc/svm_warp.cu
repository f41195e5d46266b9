// svm_warp.cu -- register-resident warp-per-block SDCA sweep.  Placeholder until the kernel lands.
#include "common.cuh"

namespace fms {
int svm_warp_supported(const fms_ctx *) { return 0; }
int svm_launch_sweep_warp(fms_ctx *ctx, float *, int64_t) {
    return fail(ctx, FMS_ERR_UNSUPPORTED, "warp SVM sweep not available");
}
}  // namespace fms
