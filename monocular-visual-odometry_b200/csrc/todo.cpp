// Entry points declared in include/mvo.h whose kernels are not written yet.
// Each returns MVO_ERR_UNSUPPORTED loudly; entries move out of this file as stages land.
#include "mvo_internal.h"
#define TODO(ctx, name) return mvo_fail(ctx, MVO_ERR_UNSUPPORTED, name ": not implemented yet")
extern "C" {
int mvo_bundle_adjustment(mvo_ctx *ctx, double *, int, float *, int, const int32_t *, const int32_t *, const float *, int, const double *, const double *, int, int, double *) { TODO(ctx, "mvo_bundle_adjustment"); }
int mvo_optimize_single_frame(mvo_ctx *ctx, double *, float *, const float *, int, const double *, int, int) { TODO(ctx, "mvo_optimize_single_frame"); }
}
