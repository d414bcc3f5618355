// In-library multi-GPU orchestration (single process, peer access): slab-decomposed single volumes and batches of
// independent blocks dealt over several devices.  See fc_multi.cu.
#pragma once
#include <functional>

#include "fc_plan.h"

namespace fcb200 {

// ---- pieces of fc_api.cu that the multi-device paths reuse ------------------------------------------
bool is_device_ptr(const void* p, int dev);
bool prepare_psf(ConvPlan& p, const float* kernel, bool k_dev, const int* pdims, bool save_memory, cudaStream_t st);
struct BatchKinds {
    bool any_pageable = false, any_device = false, any_host = false;
};
BatchKinds classify_batch(float* const* ims, int n, int dev);
// Pipelined batch on ONE device; `next` hands out block indices (-1: none left).
void batch_core(float* const* ims, const std::function<int()>& next, BatchKinds kinds, int nx, int ny, int nz,
                const float* kernel, const int* pdims, int dev, bool save_memory, const PadGeom* pad = nullptr);

// ---- fc_multi.cu ----------------------------------------------------------------------------------------
// One volume [d2][d1][d0] cut in z slabs over the devices `devs` (rank r = devs[r]; a device may appear more than
// once: emulated ranks, used by the single-GPU tests).  Exactly one of `im` / `slabs` is given:
//   im    : the whole volume, host pointer (pinned, registered or pageable)
//   slabs : slabs[r] = rank r's z slab, device pointer on devs[r]
// kernel: host pointer or device pointer (any device with peer access).  Synchronous.
void slab_convolve(float* im, float* const* slabs, const int* imDim, const float* kernel, const int* kernelDim,
                   const int* devs, int ndev);
// Per-rank device times of the most recent slab_convolve on (imDim, devs): ms[4*r + {0,1,2,3}] =
// {x+y forward, fused z incl. the wait for the peers' rows, y+x inverse incl. the wait, whole call}.
int slab_last_timing(const int* imDim, const int* devs, int ndev, float* ms, int cap);
// n host blocks of one shape and one PSF over several devices: one pipelined batch per device (batch_core), all
// fed from one shared counter.  blocks_per_dev (optional, ndev ints) receives how many blocks each device took.
void batch_multi(float* const* ims, int n, const int* imDim, const float* kernel, const int* kernelDim, const int* devs,
                 int ndev, int* blocks_per_dev);
// devices that a volume of imDim should be spread over when it does not fit on devCUDA alone (SaveMemory routing):
// devCUDA first, then every device with mutual peer access; empty when slab mode is off / pointless
std::vector<int> slab_devices_for(const int* imDim, int devCUDA, bool host_pointer);
void release_multi();

}  // namespace fcb200
