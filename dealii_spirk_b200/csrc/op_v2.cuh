// Matrix-free cell operator, variant 2 (fused-epilogue fast path).  See DESIGN.md.
#pragma once
#include "op_v1.cuh"

namespace spirk
{
  enum V2Mode
  {
    V2_APPLY    = 0, // dst = A src
    V2_RESIDUAL = 1, // dst = rhs - A src
    V2_CHEB     = 2  // dst = src + f1 (src - x_old) + f2 dinv (rhs - A src)
  };

  // returns SPIRK_ERR_UNSUPPORTED when the level / operator shape is not covered; the caller
  // then uses the general variant-1 kernels.
  inline int v2_apply(spirk_ctx *, const Geo &, const spirk_opdesc *, V2Mode, double *, const double *, const double *,
                      const double *, const double *, long long, const double *, const double *)
  {
    return SPIRK_ERR_UNSUPPORTED;
  }
} // namespace spirk
