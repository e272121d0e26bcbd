// Entry points declared in include/mvo.h whose kernels are not written yet.
// Each returns MVO_ERR_UNSUPPORTED loudly; entries move out of this file as stages land.
#include "mvo_internal.h"
#define TODO(ctx, name) return mvo_fail(ctx, MVO_ERR_UNSUPPORTED, name ": not implemented yet")
extern "C" {
int mvo_solve_pnp_ransac(mvo_ctx *ctx, const float *, const float *, int, const double *, double *, double *, int32_t *, int *) { TODO(ctx, "mvo_solve_pnp_ransac"); }
int mvo_pnp_last_hypotheses(mvo_ctx *ctx, double *, int32_t *, int, int *) { TODO(ctx, "mvo_pnp_last_hypotheses"); }
int mvo_pnp_refine(mvo_ctx *ctx, const float *, const float *, int, const double *, double *, double *) { TODO(ctx, "mvo_pnp_refine"); }
int mvo_bundle_adjustment(mvo_ctx *ctx, double *, int, float *, int, const int32_t *, const int32_t *, const float *, int, const double *, const double *, int, int, double *) { TODO(ctx, "mvo_bundle_adjustment"); }
int mvo_optimize_single_frame(mvo_ctx *ctx, double *, float *, const float *, int, const double *, int, int) { TODO(ctx, "mvo_optimize_single_frame"); }
}
