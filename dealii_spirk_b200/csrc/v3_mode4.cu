// Plane-streaming cell operator (op_v3.cuh), mode V2_CHEB_FIRST: explicit instantiations in their own translation unit.
// The kernels sit on a register cliff; a multi-threaded compile of one big unit (nvcc --split-compile) produced
// different spill placement from build to build (0.24 ms vs 0.48 ms per launch), so every mode is compiled
// single-threaded and the units are compiled side by side (dealii_spirk_b200/build.py).
#define SPIRK_V3_INSTANTIATE
#include "op_v3.cuh"

namespace spirk
{
  template int v3_launch_mode<4, 8, 8, V2_CHEB_FIRST, 4, 1>(spirk_ctx *, V3Args &);
  template int v3_launch_mode<4, 8, 8, V2_CHEB_FIRST, 2, 1>(spirk_ctx *, V3Args &);
  template int v3_launch_mode<4, 4, 4, V2_CHEB_FIRST, 4, 1>(spirk_ctx *, V3Args &);

  // this unit's copy of the 1-D tables
  int v3_upload_constants_mode4(const FeConst *all)
  {
    SPIRK_CUDA(cudaMemcpyToSymbol(c_fe, all, sizeof(FeConst) * (SPIRK_MAX_DEGREE + 1)));
    return SPIRK_OK;
  }
} // namespace spirk
