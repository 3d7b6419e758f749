// mma3d.cuh -- geometry shared by the two tensor-core kernel families for d = 3, m <= 6: mma3d.cu (FP64 DMMA and legacy
// TF32 mma.sync) and tc5.cu (fp32 plans on tcgen05 / TMEM).  Both sweep a TILE of T x T grid cells (T = 17 - (2m+2)), whose
// tap boxes lie in a footprint of 16 x 16 pencils along z, in ascending order of the nodes' lowest tap u2.
#pragma once

#include "common.cuh"

namespace nfftcu {

struct MmaParams {
  int n0, n1, n2;
  int T;            // tile edge: 17 - W
  int NT0, NT1;
  int zseg;         // work units per tile along z
  const double *img;   // window images of mma3d.cu (kImgDoubles per batch) or null
  long long M;
  int m;
  int deg;          // Horner length (fitted polynomial degree)
};

MmaParams mma3d_params(const nfftcu_ctx *c);   // mma3d.cu

// tc5.cu: fp32 plans, tcgen05.mma kind::tf32 with the grid window and the accumulators in tensor memory
bool tc5_selected(const nfftcu_ctx *c);                      // the plan's B / B^T run on the tcgen05 kernels
int tc5_build(nfftcu_ctx *c, const MmaParams &P);            // plan time: batch table, chunks, operand images
int tc5_interp(nfftcu_ctx *c, void *f_dev);
int tc5_spread(nfftcu_ctx *c, const void *f_dev);

}  // namespace nfftcu
