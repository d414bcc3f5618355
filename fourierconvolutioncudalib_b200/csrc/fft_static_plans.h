// Shared by the compile-time specialised translation units (fft_static_col.cu, fft_static_x.cu):
// the list of static plans and small dispatch helpers.
#pragma once
#include <algorithm>
#include <cstdlib>

#include "fft_static.cuh"
#include "fft_kernels.h"

namespace fcb200 {
namespace {

template <class P>
bool plan_matches(const AxisPlanDev& d)
{
    if (d.L != P::L || d.ns != P::ns || d.generic) return false;
    const int r[4] = {P::R0, P::R1, P::R2, P::R3};
    for (int i = 0; i < P::ns; ++i)
        if (d.radix[i] != r[i]) return false;
    return true;
}


bool static_enabled()
{
    static const bool on = [] {
        const char* e = std::getenv("FCB200_STATIC");
        return !(e && std::atoi(e) == 0);
    }();
    return on;
}

// The radix sequences are exactly what the planner (fc_plan.cu: factorize) produces.

int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}


// The radix sequences are exactly what the planner (fc_plan.cu: factorize) produces.
typedef SPlan<32, 8, 4> P32;
typedef SPlan<64, 8, 8> P64;
typedef SPlan<128, 16, 8> P128;
typedef SPlan<192, 8, 8, 3> P192;
typedef SPlan<256, 16, 16> P256;
typedef SPlan<384, 16, 8, 3> P384;
typedef SPlan<512, 8, 8, 8> P512;
typedef SPlan<1024, 16, 16, 4> P1024;
typedef SPlan<2048, 16, 16, 8> P2048;   // 64-byte row segments (4 column pairs): a 128-byte tile would need 256 KB
// planning style 1 (fused z axis): L = 256 as (8,8,4)
typedef SPlan<256, 8, 8, 4> P256b;
// planning style 2 (x axis): half-length 512 of nx = 1024 rows as (16,4,8)
typedef SPlan<512, 16, 4, 8> P512x;
// 7-smooth extents of caller-padded volumes (BASELINE configs 2-4 padded: 270, 300, 420, 448, 560) and
// the half-lengths of their x transforms.  Lengths with four small prime factors run as two stages of fat composite
// radices (fc_plan.cu: factorize has the measurements).
typedef SPlan<560, 16, 5, 7> P560;      // y axis
typedef SPlan<560, 28, 20> P560z;       // fused z axis
typedef SPlan<448, 8, 8, 7> P448;       // fused z axis
typedef SPlan<448, 16, 28> P448y;       // y axis
typedef SPlan<420, 20, 21> P420;
typedef SPlan<300, 20, 15> P300;
typedef SPlan<270, 18, 15> P270;
typedef SPlan<280, 8, 5, 7> P280;
typedef SPlan<224, 8, 4, 7> P224;
typedef SPlan<210, 2, 3, 5, 7> P210;
typedef SPlan<150, 10, 15> P150;
typedef SPlan<135, 9, 15> P135;


}  // namespace
}  // namespace fcb200
