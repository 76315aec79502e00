// Plane-streaming cell operator (op_v3.cuh), coupled pairs of blocks with fused epilogues (V2_RESIDUAL, V2_CHEB):
// residual and Chebyshev smoother of the complex level operators (operator.h:616-665).  Explicit instantiations in their
// own translation unit (see v3_mode3.cu for why every mode is compiled on its own).
#define SPIRK_V3_INSTANTIATE
#include "op_v3.cuh"

namespace spirk
{
  template int v3_launch_mode<4, 8, 8, V2_RESIDUAL, 4, 2>(spirk_ctx *, V3Args &);
  template int v3_launch_mode<4, 4, 4, V2_RESIDUAL, 4, 2>(spirk_ctx *, V3Args &);
  template int v3_launch_mode<4, 8, 8, V2_CHEB, 4, 2>(spirk_ctx *, V3Args &);
  template int v3_launch_mode<4, 4, 4, V2_CHEB, 4, 2>(spirk_ctx *, V3Args &);

  // this unit's copy of the 1-D tables
  int v3_upload_constants_mode5(const FeConst *all)
  {
    SPIRK_CUDA(cudaMemcpyToSymbol(c_fe, all, sizeof(FeConst) * (SPIRK_MAX_DEGREE + 1)));
    return SPIRK_OK;
  }
} // namespace spirk
