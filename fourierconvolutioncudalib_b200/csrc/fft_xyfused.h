// x+y passes of a z plane fused through distributed shared memory (fft_xyfused.cu).
#pragma once
#include "fft_kernels.h"

namespace fcb200 {

struct XYArgs {
    const float* in_real;   // forward: [nplanes][ny][nx]
    float* out_real;        // inverse
    float2* spec;           // [nplanes][ny][xcp], y-transformed (natural ky) after forward / before inverse
    Geometry g;
    AxisPlanDev Px, Py;
    const float2* twx;
    int nplanes;
};

// false: shape not covered (caller runs the separate x and y passes)
bool launch_xy_fwd_cluster(const XYArgs& a, cudaStream_t st);
bool launch_yx_inv_cluster(const XYArgs& a, cudaStream_t st);

}  // namespace fcb200
