// extern "C" boundary: the reference's exported functions (include/convolution3Dfft.h) plus the
// extensions of include/fcb200_ext.h.  Host orchestration only; all arithmetic is in fft_kernels.cu.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "convolution3Dfft.h"
#include "fc_api_util.h"
#include "fc_multi.h"
#include "fc_plan.h"
#include "fcb200_ext.h"

using namespace fcb200;

namespace fcb200 {
thread_local std::string g_last_error;
std::atomic<int> g_error_mode{0};
}  // namespace fcb200

namespace fcb200 {

// Is `p` device memory usable from device `dev`?  (extension: the reference only takes host pointers)
bool is_device_ptr(const void* p, int dev)
{
    cudaPointerAttributes attr{};
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (attr.type == cudaMemoryTypeDevice) {
        if (attr.device != dev) throw std::runtime_error("fcb200: device pointer belongs to another device");
        return true;
    }
    return attr.type == cudaMemoryTypeManaged;
}

// PSF taps (host or device pointer) -> PSF spectrum in plan.d_H (or the SaveMemory window buffers).
// Host-pointer taps are remembered: when the next call brings the same bytes for the same shapes, d_H is
// still their spectrum and the three PSF passes are skipped (FCB200_PSF_CACHE=0 turns
// it off, bench.py does so to time the stateless path).  Returns true when the window (SaveMemory) path is set up.
bool prepare_psf(ConvPlan& p, const float* kernel, bool k_dev, const int* pdims, bool save_memory, cudaStream_t st)
{
    static const bool cache_on = env_flag("FCB200_PSF_CACHE", true);
    // On-the-fly PSF spectrum (windows of 16 / 32 / 64 PSF planes) is the SaveMemory path, and InPlace takes it too where
    // the on-the-fly fused pass ITSELF is as fast as the one that reads a materialised spectrum (FCB200_OTF_INPLACE: unset =
    // this rule, 0 = never, 1 = wherever a window exists; read per call):
    //   nz = 256, 16-plane window (C3): 0.132 against 0.142 ms, and no PSF z pass: C3 0.588 -> 0.517 ms, 256^3 0.182 -> 0.156
    //   nz = 384, window <= 32 planes (a C4 block): 0.147 against 0.144 ms: 384^3 0.574 -> 0.500 ms when the spectrum has to
    //   be built, within 1 % of a cached materialised spectrum when it has not -- and 231 MB less memory
    //   nz <= 128, 16-plane window (launch-bound volumes, C1): the PSF chain on the side stream -- three small launches -- is
    //   longer than the image's x and y passes it runs next to; without the PSF z launch 64^3 45.5 -> 38.7 us, 128^3 49.9 -> 45.5
    // Elsewhere the materialised spectrum stays: it is cached across calls with the same host taps (the deconvolution
    // pattern) and its fused pass is the faster one (512: 0.367 against 0.469 ms).  The rule depends on the shapes only, so
    // host and device pointers give bit-identical results.
    bool otf_inplace = false;
    if (!save_memory) {
        const char* e = std::getenv("FCB200_OTF_INPLACE");
        if (e) otf_inplace = std::atoi(e) != 0;
        else if (p.g.nz <= 128 || p.g.nz == 256 || p.g.nz == 384)
            otf_inplace = psf_window_applies(p, pdims, st) && p.psf_window_planes <= (p.g.nz == 384 ? 32 : 16);
    }
    const size_t ktaps = (size_t)pdims[0] * pdims[1] * pdims[2];
    auto same_taps = [&](bool valid, const int* dims, const std::vector<float>& taps) {
        return !k_dev && cache_on && valid && std::memcmp(dims, pdims, sizeof(int) * 6) == 0 && taps.size() == ktaps &&
               std::memcmp(taps.data(), kernel, ktaps * sizeof(float)) == 0;
    };
    const bool window = (save_memory || otf_inplace) && psf_window_applies(p, pdims, st);
    if (window && same_taps(p.hwin_valid, p.hwin_dims, p.hwin_taps)) return true;
    if (!window && same_taps(p.h_valid, p.h_dims, p.h_taps)) return false;
    const float* d_kernel = kernel;
    if (!k_dev) {
        if (ktaps > p.kernel_cap) {
            cudaFree(p.d_kernel);
            p.d_kernel = nullptr;
            p.kernel_cap = 0;
            FC_CUDA(cudaMalloc(&p.d_kernel, ktaps * sizeof(float)));
            p.kernel_cap = ktaps;
        }
        FC_CUDA(cudaMemcpyAsync(p.d_kernel, kernel, ktaps * sizeof(float), cudaMemcpyHostToDevice, st));
        d_kernel = p.d_kernel;
    }
    if (window && run_psf_window(p, d_kernel, pdims, st)) {
        if (!k_dev && cache_on) {
            p.hwin_taps.assign(kernel, kernel + ktaps);
            std::memcpy(p.hwin_dims, pdims, sizeof(int) * 6);
            p.hwin_valid = true;
        }
        return true;
    }
    run_psf_spectrum(p, d_kernel, pdims, st);
    if (!k_dev && cache_on) {
        p.h_taps.assign(kernel, kernel + ktaps);
        std::memcpy(p.h_dims, pdims, sizeof(int) * 6);
        p.h_valid = true;
    }
    return false;
}

}  // namespace fcb200

namespace {

// The PSF passes (three small, partly launch-bound kernels) run on a side stream next to the image's x/y
// passes and join before the fused z pass.  Off while per-pass profiling is on (events need serial passes).
struct PsfSide {
    ConvPlan& p;
    cudaStream_t st, sp;
    bool overlap;
    PsfSide(ConvPlan& plan, cudaStream_t main) : p(plan), st(main), sp(main)
    {
        static const bool overlap_on = env_flag("FCB200_PSF_OVERLAP", true);
        overlap = overlap_on && !profile_enabled();
        if (!overlap) return;
        if (!p.s_psf) FC_CUDA(cudaStreamCreateWithFlags(&p.s_psf, cudaStreamNonBlocking));
        if (!p.ev_psf_fork) FC_CUDA(cudaEventCreateWithFlags(&p.ev_psf_fork, cudaEventDisableTiming));
        if (!p.ev_psf_done) FC_CUDA(cudaEventCreateWithFlags(&p.ev_psf_done, cudaEventDisableTiming));
        sp = p.s_psf;
        FC_CUDA(cudaEventRecord(p.ev_psf_fork, st));      // after whatever still reads the PSF buffers on st
        FC_CUDA(cudaStreamWaitEvent(sp, p.ev_psf_fork, 0));
    }
    cudaStream_t stream() const { return sp; }
    void done() { if (overlap) FC_CUDA(cudaEventRecord(p.ev_psf_done, sp)); }
    void join() { if (overlap) FC_CUDA(cudaStreamWaitEvent(st, p.ev_psf_done, 0)); }
};

void convolve_core(float* im, int nx, int ny, int nz, const float* kernel, const int* pdims, int dev,
                   bool force_async, cudaStream_t user_stream, bool save_memory = false)
{
    DeviceGuard guard(dev);
    for (int i = 0; i < 3; ++i)
        if (pdims[i] > pdims[i + 3]) throw std::runtime_error("fcb200: kernel larger than image");
    auto plan = get_plan(dev, nx, ny, nz);
    std::lock_guard<std::mutex> lock(plan->mu);
    ConvPlan& p = *plan;

    const HostMem im_kind = force_async ? HostMem::Device : classify_pointer(im, dev);
    const bool im_dev = im_kind == HostMem::Device;
    const bool k_dev = force_async || is_device_ptr(kernel, dev);
    cudaStream_t st = force_async ? user_stream : (im_dev ? (cudaStream_t)0 : p.stream);
    // pageable buffers (what JNA hands over) go through pinned slots filled by the copy threads
    static const bool staging_on = env_flag("FCB200_STAGING", true);
    const bool staged = im_kind == HostMem::Pageable && staging_on;

    workspace_acquire(p, st);
    struct Release {
        ConvPlan& p;
        cudaStream_t st;
        ~Release() { try { workspace_release(p, st); } catch (...) {} }
    } release_on_exit{p, st};
    PsfSide psf(p, st);
    const bool window = prepare_psf(p, kernel, k_dev, pdims, save_memory, psf.stream());
    psf.done();
    auto join_psf = [&] { psf.join(); };

    // Pinned host image: the volume travels in z chunks on its own copy streams; x+y forward of chunk c runs
    // while chunk c+1 is still on the wire (and the PSF passes run under chunk 0), y+x inverse of chunk c+1
    // runs while chunk c is already going home.  Only the fused z pass needs the whole volume.
    static const int chunks_env = [] {
        const char* e = std::getenv("FCB200_E2E_CHUNKS");
        return e ? std::atoi(e) : 8;
    }();
    const int nch = (int)std::max<long long>(
        1, std::min<long long>(std::min(8, chunks_env), std::min<long long>(nz, (long long)(p.real_bytes() >> 25))));
    if (im_kind == HostMem::Pinned && nch > 1) {
        if (!p.d_real) FC_CUDA(device_alloc_retry(&p.d_real, p.real_bytes()));
        if (!p.s_h2d) FC_CUDA(cudaStreamCreateWithFlags(&p.s_h2d, cudaStreamNonBlocking));
        if (!p.s_d2h) FC_CUDA(cudaStreamCreateWithFlags(&p.s_d2h, cudaStreamNonBlocking));
        for (cudaEvent_t& e : p.ev_chunk)
            if (!e) FC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const size_t plane = (size_t)ny * nx;
        const int per = (nz + nch - 1) / nch;
        auto z0_of = [&](int c) { return std::min(nz, c * per); };
        for (int c = 0; c < nch; ++c) {
            const int z0 = z0_of(c), n = z0_of(c + 1) - z0;
            if (n <= 0) continue;
            FC_CUDA(cudaMemcpyAsync(p.d_real + z0 * plane, im + z0 * plane, n * plane * sizeof(float),
                                    cudaMemcpyHostToDevice, p.s_h2d));
            FC_CUDA(cudaEventRecord(p.ev_chunk[c], p.s_h2d));
        }
        for (int c = 0; c < nch; ++c) {
            const int z0 = z0_of(c), n = z0_of(c + 1) - z0;
            if (n <= 0) continue;
            FC_CUDA(cudaStreamWaitEvent(st, p.ev_chunk[c], 0));
            run_xy_forward_planes(p, p.d_real, z0, n, st);
        }
        join_psf();
        run_z_fused(p, window, st);
        for (int c = 0; c < nch; ++c) {
            const int z0 = z0_of(c), n = z0_of(c + 1) - z0;
            if (n <= 0) continue;
            run_yx_inverse_planes(p, p.d_real, z0, n, st);
            FC_CUDA(cudaEventRecord(p.ev_chunk[8 + c], st));
            FC_CUDA(cudaStreamWaitEvent(p.s_d2h, p.ev_chunk[8 + c], 0));
            FC_CUDA(cudaMemcpyAsync(im + z0 * plane, p.d_real + z0 * plane, n * plane * sizeof(float),
                                    cudaMemcpyDeviceToHost, p.s_d2h));
        }
        FC_CUDA(cudaStreamSynchronize(p.s_d2h));
        FC_CUDA(cudaStreamSynchronize(st));
        return;
    }

    float* d_im = im;
    if (!im_dev) {
        if (!p.d_real) FC_CUDA(device_alloc_retry(&p.d_real, p.real_bytes()));
        if (staged) p.stager.upload(p.d_real, im, p.real_bytes(), st);
        else FC_CUDA(cudaMemcpyAsync(p.d_real, im, p.real_bytes(), cudaMemcpyHostToDevice, st));
        d_im = p.d_real;
    } else {
        check_real_alignment(im, nx);
    }
    run_xy_forward_planes(p, d_im, 0, nz, st);
    join_psf();
    run_z_fused(p, window, st);
    run_yx_inverse_planes(p, d_im, 0, nz, st);
    if (!im_dev) {
        if (staged) p.stager.download(im, p.d_real, p.real_bytes(), st);
        else FC_CUDA(cudaMemcpyAsync(im, p.d_real, p.real_bytes(), cudaMemcpyDeviceToHost, st));
    }
    if (!force_async) FC_CUDA(cudaStreamSynchronize(st));
}

// ------------------------------------------------------------------------------------------------
// In-library padding (the step the reference leaves to its callers, src/convolution3Dfft.h:39,
// :54, and that its tests do on the host: tests/padd_utils.h:99-171 + the sub-view read-back of
// tests/test_fixtures.hpp:254-268).  The caller hands over the UNPADDED volume; it is embedded at offsets
// kernelDim/2 in a padded grid (zeros or mirror), convolved there exactly like convolution3DfftCUDAInPlace
// would convolve the caller-padded volume, and the interior is written back.  What the library gains from
// knowing about the padding: only the unpadded bytes cross PCIe, zero z-halo planes are never transformed
// forward (their spectrum planes are cleared instead), and no halo plane is transformed back.
// ------------------------------------------------------------------------------------------------
// Device-resident unpadded volume d_src convolved on the padded grid of plan p, fused form (the x passes pad and
// crop, see fft_xpass.cuh); before_z() runs between the forward passes and the fused z pass (PSF join).
template <typename F>
void padded_convolve_device(ConvPlan& p, float* d_src, const PadGeom& g, cudaStream_t st, F&& before_z, bool window)
{
    const size_t splane = (size_t)p.g.ny * p.g.xcp;
    const int z_lo = g.oz, z_hi = g.oz + g.sz;
    run_xy_forward_planes(p, d_src, z_lo, g.sz, st, &g);
    const int r0[2] = {0, z_hi}, rn[2] = {z_lo, g.pz - z_hi};
    for (int i = 0; i < 2; ++i) {
        if (rn[i] <= 0) continue;
        if (g.mode == 0) FC_CUDA(cudaMemsetAsync(p.d_spec + r0[i] * splane, 0, rn[i] * splane * sizeof(float2), st));
        else run_xy_forward_planes(p, d_src, r0[i], rn[i], st, &g);
    }
    before_z();
    run_z_fused(p, window, st);
    run_yx_inverse_planes(p, d_src, z_lo, g.sz, st, &g);
}

PadGeom make_pad_geom(const int* imDim, const int* kernelDim, int mode, int policy, int* pd)
{
    if (mode != 0 && mode != 1) throw std::runtime_error("fcb200: padding mode must be 0 (zero) or 1 (mirror)");
    if (policy != 0 && policy != 1) throw std::runtime_error("fcb200: padding policy must be 0 (exact) or 1 (7-smooth)");
    padded_extents(imDim, kernelDim, policy, pd);
    return PadGeom{imDim[0], imDim[1], imDim[2], pd[0], pd[1], pd[2],
                   kernelDim[0] / 2, kernelDim[1] / 2, kernelDim[2] / 2, mode};
}

void padded_core(float* im, const int* imDim, const float* kernel, const int* kernelDim, int mode, int policy, int dev,
                 bool force_async, cudaStream_t user_stream)
{
    int pd[3];
    const PadGeom g = make_pad_geom(imDim, kernelDim, mode, policy, pd);
    const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], pd[0], pd[1], pd[2]};
    DeviceGuard guard(dev);
    auto plan = get_plan(dev, pd[0], pd[1], pd[2]);
    std::lock_guard<std::mutex> lock(plan->mu);
    ConvPlan& p = *plan;

    const HostMem im_kind = force_async ? HostMem::Device : classify_pointer(im, dev);
    const bool im_dev = im_kind == HostMem::Device;
    const bool k_dev = force_async || is_device_ptr(kernel, dev);
    cudaStream_t st = force_async ? user_stream : (im_dev ? (cudaStream_t)0 : p.stream);
    static const bool staging_on = env_flag("FCB200_STAGING", true);
    const bool staged = im_kind == HostMem::Pageable && staging_on;

    // Fused (default): the x passes assemble the padded rows while loading the unpadded volume and store only the
    // interior back, so no padded real volume exists.  FCB200_PAD_FUSED=0: separate embed / crop kernels.
    static const bool fused = env_flag("FCB200_PAD_FUSED", true);
    if (!fused && !p.d_real) FC_CUDA(device_alloc_retry(&p.d_real, p.real_bytes()));
    const size_t src_plane = (size_t)g.sy * g.sx, src_elems = src_plane * g.sz, src_bytes = src_elems * sizeof(float);
    float* d_src = im;
    if (!im_dev) {
        if (src_elems > p.unpadded_cap) {
            cudaFree(p.d_unpadded);
            p.d_unpadded = nullptr;
            p.unpadded_cap = 0;
            FC_CUDA(device_alloc_retry(&p.d_unpadded, src_bytes));
            p.unpadded_cap = src_elems;
        }
        d_src = p.d_unpadded;
    }

    workspace_acquire(p, st);
    struct Release {
        ConvPlan& p;
        cudaStream_t st;
        ~Release() { try { workspace_release(p, st); } catch (...) {} }
    } release_on_exit{p, st};
    PsfSide psf(p, st);
    const bool window = prepare_psf(p, kernel, k_dev, pdims, false, psf.stream());
    psf.done();

    const size_t splane = (size_t)p.g.ny * p.g.xcp;
    const int z_lo = g.oz, z_hi = g.oz + g.sz;          // source planes sit in padded planes [z_lo, z_hi)
    auto forward_planes = [&](int pz0, int pn) {        // x+y forward of padded planes [pz0, pz0+pn)
        if (fused) {
            run_xy_forward_planes(p, d_src, pz0, pn, st, &g);
        } else {
            run_pad_embed(d_src, p.d_real, g, pz0, pn, st);
            run_xy_forward_planes(p, p.d_real, pz0, pn, st);
        }
    };
    auto inverse_planes = [&](int z0, int n) {          // y+x inverse of source planes [z0, z0+n), cropped into d_src
        if (fused) {
            run_yx_inverse_planes(p, d_src, z_lo + z0, n, st, &g);
        } else {
            run_yx_inverse_planes(p, p.d_real, z_lo + z0, n, st);
            run_pad_crop(p.d_real, d_src, g, z0, n, st);
        }
    };
    // forward spectrum planes of the z halo: zero volume planes have a zero spectrum; mirrored ones are built
    // from the (complete) source and transformed
    auto halo_forward = [&] {
        const int r0[2] = {0, z_hi}, rn[2] = {z_lo, g.pz - z_hi};
        for (int i = 0; i < 2; ++i) {
            if (rn[i] <= 0) continue;
            if (mode == 0) {
                FC_CUDA(cudaMemsetAsync(p.d_spec + r0[i] * splane, 0, rn[i] * splane * sizeof(float2), st));
            } else {
                forward_planes(r0[i], rn[i]);
            }
        }
    };

    static const int chunks_env = [] {
        const char* e = std::getenv("FCB200_E2E_CHUNKS");
        return e ? std::atoi(e) : 8;
    }();
    const int nch = (int)std::max<long long>(
        1, std::min<long long>(std::min(8, chunks_env), std::min<long long>(g.sz, (long long)(src_bytes >> 25))));
    if (im_kind == HostMem::Pinned && nch > 1) {
        // the unpadded volume travels in z chunks; embedding + x/y forward of chunk c run while chunk c+1 is on
        // the wire, y/x inverse + crop of chunk c+1 run while chunk c goes home
        if (!p.s_h2d) FC_CUDA(cudaStreamCreateWithFlags(&p.s_h2d, cudaStreamNonBlocking));
        if (!p.s_d2h) FC_CUDA(cudaStreamCreateWithFlags(&p.s_d2h, cudaStreamNonBlocking));
        for (cudaEvent_t& e : p.ev_chunk)
            if (!e) FC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const int per = (g.sz + nch - 1) / nch;
        auto z0_of = [&](int c) { return std::min(g.sz, c * per); };
        for (int c = 0; c < nch; ++c) {
            const int z0 = z0_of(c), n = z0_of(c + 1) - z0;
            if (n <= 0) continue;
            FC_CUDA(cudaMemcpyAsync(d_src + z0 * src_plane, im + z0 * src_plane, n * src_plane * sizeof(float),
                                    cudaMemcpyHostToDevice, p.s_h2d));
            FC_CUDA(cudaEventRecord(p.ev_chunk[c], p.s_h2d));
        }
        if (mode == 0) halo_forward();
        for (int c = 0; c < nch; ++c) {
            const int z0 = z0_of(c), n = z0_of(c + 1) - z0;
            if (n <= 0) continue;
            FC_CUDA(cudaStreamWaitEvent(st, p.ev_chunk[c], 0));
            forward_planes(z_lo + z0, n);
        }
        if (mode != 0) halo_forward();
        psf.join();
        run_z_fused(p, window, st);
        for (int c = 0; c < nch; ++c) {
            const int z0 = z0_of(c), n = z0_of(c + 1) - z0;
            if (n <= 0) continue;
            inverse_planes(z0, n);
            FC_CUDA(cudaEventRecord(p.ev_chunk[8 + c], st));
            FC_CUDA(cudaStreamWaitEvent(p.s_d2h, p.ev_chunk[8 + c], 0));
            FC_CUDA(cudaMemcpyAsync(im + z0 * src_plane, d_src + z0 * src_plane, n * src_plane * sizeof(float),
                                    cudaMemcpyDeviceToHost, p.s_d2h));
        }
        FC_CUDA(cudaStreamSynchronize(p.s_d2h));
        FC_CUDA(cudaStreamSynchronize(st));
        return;
    }

    if (!im_dev) {
        if (staged) p.stager.upload(d_src, im, src_bytes, st);
        else FC_CUDA(cudaMemcpyAsync(d_src, im, src_bytes, cudaMemcpyHostToDevice, st));
    }
    forward_planes(z_lo, g.sz);
    halo_forward();
    psf.join();
    run_z_fused(p, window, st);
    inverse_planes(0, g.sz);
    if (!im_dev) {
        if (staged) p.stager.download(im, d_src, src_bytes, st);
        else FC_CUDA(cudaMemcpyAsync(im, d_src, src_bytes, cudaMemcpyDeviceToHost, st));
    }
    if (!force_async) FC_CUDA(cudaStreamSynchronize(st));
}

// ------------------------------------------------------------------------------------------------
// Batch of independent volumes of one shape with one PSF (BASELINE config 4: 64 blocks of 384^3).
// The PSF spectrum is computed once; block b+1 is uploaded and block b-1 downloaded while block b is
// convolved: three device image buffers, one copy stream per direction (PCIe is full duplex), one compute
// stream.  Pinned / registered host buffers are DMA'd directly; pageable ones are staged chunk by chunk
// (upload on a helper thread, download on the calling thread).
// ------------------------------------------------------------------------------------------------
void batch_resources(ConvPlan& p)
{
    if (!p.d_real) FC_CUDA(device_alloc_retry(&p.d_real, p.real_bytes()));
    p.d_ring[0] = p.d_real;
    for (int i = 1; i < 3; ++i)
        if (!p.d_ring[i]) FC_CUDA(device_alloc_retry(&p.d_ring[i], p.real_bytes()));
    if (!p.s_h2d) FC_CUDA(cudaStreamCreateWithFlags(&p.s_h2d, cudaStreamNonBlocking));
    if (!p.s_d2h) FC_CUDA(cudaStreamCreateWithFlags(&p.s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 3; ++i) {
        if (!p.ev_up[i]) FC_CUDA(cudaEventCreateWithFlags(&p.ev_up[i], cudaEventDisableTiming));
        if (!p.ev_comp[i]) FC_CUDA(cudaEventCreateWithFlags(&p.ev_comp[i], cudaEventDisableTiming));
        if (!p.ev_down[i]) FC_CUDA(cudaEventCreateWithFlags(&p.ev_down[i], cudaEventDisableTiming));
    }
}

}  // namespace

namespace fcb200 {

// `next` hands out the index of the next block to convolve (-1: none left).  A single-device batch counts
// 0..n-1; fcb200_convolve_batch_multi gives every device's pipeline the SAME counter, so a device that finishes
// early simply takes more blocks (work stealing at block granularity, at most three blocks in flight per device).
void batch_core(float* const* ims, const std::function<int()>& next, BatchKinds kinds, int nx, int ny, int nz,
                const float* kernel, const int* pdims, int dev, bool save_memory, const PadGeom* pad)
{
    DeviceGuard guard(dev);
    for (int i = 0; i < 3; ++i)
        if (pdims[i] > pdims[i + 3]) throw std::runtime_error("fcb200: kernel larger than image");
    auto plan = get_plan(dev, nx, ny, nz);
    std::lock_guard<std::mutex> lock(plan->mu);
    ConvPlan& p = *plan;
    // padded batches: (nx, ny, nz) is the padded grid of the plan, the blocks and the ring buffers hold UNPADDED volumes
    const size_t bytes = pad ? (size_t)pad->sx * pad->sy * pad->sz * sizeof(float) : p.real_bytes();
    cudaStream_t st = p.stream;
    workspace_acquire(p, st);
    struct Release {
        ConvPlan& p;
        cudaStream_t st;
        ~Release() { try { workspace_release(p, st); } catch (...) {} }
    } release_on_exit{p, st};

    const bool window = prepare_psf(p, kernel, is_device_ptr(kernel, dev), pdims, save_memory, st);
    auto convolve = [&](float* d) {
        if (pad) padded_convolve_device(p, d, *pad, st, [] {}, window);
        else if (window) run_convolve_window(p, d, st);
        else run_convolve(p, d, st);
    };
    if (kinds.any_device) {   // device-resident blocks: nothing to overlap
        if (kinds.any_host) throw std::runtime_error("fcb200: a batch must be all host or all device pointers");
        for (int b = next(); b >= 0; b = next()) {
            if (classify_pointer(ims[b], dev) != HostMem::Device)
                throw std::runtime_error("fcb200: a batch must be all host or all device pointers");
            check_real_alignment(ims[b], nx);
            convolve(ims[b]);
        }
        FC_CUDA(cudaStreamSynchronize(st));
        return;
    }
    batch_resources(p);
    static const bool staging_on = env_flag("FCB200_STAGING", true);

    if (!kinds.any_pageable || !staging_on) {
        // three streams ordered by events; the host only waits for a ring slot to drain before it takes another block
        for (int count = 0;; ++count) {
            const int s = count % 3;
            if (count >= 3) FC_CUDA(cudaEventSynchronize(p.ev_down[s]));
            const int b = next();
            if (b < 0) break;
            FC_CUDA(cudaMemcpyAsync(p.d_ring[s], ims[b], bytes, cudaMemcpyHostToDevice, p.s_h2d));
            FC_CUDA(cudaEventRecord(p.ev_up[s], p.s_h2d));
            FC_CUDA(cudaStreamWaitEvent(st, p.ev_up[s], 0));
            convolve(p.d_ring[s]);
            FC_CUDA(cudaEventRecord(p.ev_comp[s], st));
            FC_CUDA(cudaStreamWaitEvent(p.s_d2h, p.ev_comp[s], 0));
            FC_CUDA(cudaMemcpyAsync(ims[b], p.d_ring[s], bytes, cudaMemcpyDeviceToHost, p.s_d2h));
            FC_CUDA(cudaEventRecord(p.ev_down[s], p.s_d2h));
        }
        FC_CUDA(cudaStreamSynchronize(p.s_d2h));
        FC_CUDA(cudaStreamSynchronize(st));
        return;
    }

    // pageable blocks: an uploader thread takes block after block and stages it; this thread enqueues the
    // convolutions and drains the results.  Host-side counters order the event records against the waits that use them.
    p.stager.prepare();
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> order;   // blocks in the order the uploader took them
    bool upload_done = false;
    int downloaded = 0;
    std::string upload_error;
    std::thread uploader([&] {
        try {
            FC_CUDA(cudaSetDevice(dev));
            for (int i = 0;; ++i) {
                const int s = i % 3;
                {
                    std::unique_lock<std::mutex> l(mu);
                    cv.wait(l, [&] { return downloaded >= i - 2; });   // ring slot drained
                }
                const int b = next();
                if (b < 0) break;
                p.stager.upload(p.d_ring[s], ims[b], bytes, p.s_h2d);
                FC_CUDA(cudaEventRecord(p.ev_up[s], p.s_h2d));
                {
                    std::lock_guard<std::mutex> l(mu);
                    order.push_back(b);
                }
                cv.notify_all();
            }
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> l(mu);
            upload_error = e.what();
        }
        {
            std::lock_guard<std::mutex> l(mu);
            upload_done = true;
        }
        cv.notify_all();
    });
    std::string error;
    try {
        int next_conv = 0;   // next uploaded block whose convolution has not been enqueued yet
        for (int i = 0;; ++i) {
            bool finished = false;
            for (;;) {
                int up;
                {
                    std::unique_lock<std::mutex> l(mu);
                    if (next_conv <= i) cv.wait(l, [&] { return (int)order.size() > next_conv || upload_done; });
                    up = (int)order.size();
                    if (!upload_error.empty()) throw std::runtime_error(upload_error);
                    finished = upload_done && up <= i;
                }
                if (next_conv >= up || next_conv > i + 1) break;
                const int s = next_conv % 3;
                FC_CUDA(cudaStreamWaitEvent(st, p.ev_up[s], 0));
                convolve(p.d_ring[s]);
                FC_CUDA(cudaEventRecord(p.ev_comp[s], st));
                ++next_conv;
            }
            if (finished) break;
            int b;
            {
                std::lock_guard<std::mutex> l(mu);
                b = order[(size_t)i];
            }
            const int s = i % 3;
            FC_CUDA(cudaStreamWaitEvent(p.s_d2h, p.ev_comp[s], 0));
            p.stager.download(ims[b], p.d_ring[s], bytes, p.s_d2h);
            {
                std::lock_guard<std::mutex> l(mu);
                downloaded = i + 1;
            }
            cv.notify_all();
        }
    } catch (const std::exception& e) {
        error = e.what();
        std::lock_guard<std::mutex> l(mu);
        downloaded = 1 << 30;   // release the uploader
        cv.notify_all();
    }
    uploader.join();
    cudaStreamSynchronize(p.s_h2d);
    cudaStreamSynchronize(p.s_d2h);
    cudaStreamSynchronize(st);
    if (!error.empty()) throw std::runtime_error(error);
}

BatchKinds classify_batch(float* const* ims, int n, int dev)
{
    BatchKinds k{};
    for (int b = 0; b < n; ++b) {
        const HostMem m = classify_pointer(ims[b], dev);
        k.any_pageable = k.any_pageable || m == HostMem::Pageable;
        k.any_device = k.any_device || m == HostMem::Device;
        k.any_host = k.any_host || m != HostMem::Device;
    }
    return k;
}

}  // namespace fcb200

namespace {

void batch_single(float* const* ims, int n, int nx, int ny, int nz, const float* kernel, const int* pdims, int dev,
                  bool save_memory, const PadGeom* pad = nullptr)
{
    if (n <= 0) return;
    int counter = 0;
    batch_core(ims, [&] { return counter < n ? counter++ : -1; }, classify_batch(ims, n, dev), nx, ny, nz, kernel, pdims,
               dev, save_memory, pad);
}

// host spectrum [nz][ny][xc] (what numpy.fft.rfftn returns)  <->  device layout [nz][ny][xcp]
void download_spectrum(ConvPlan& p, const float2* d_spec, float* out, cudaStream_t st)
{
    const Geometry& g = p.g;
    std::vector<float2> h((size_t)g.nz * g.ny * g.xcp);
    FC_CUDA(cudaMemcpyAsync(h.data(), d_spec, h.size() * sizeof(float2), cudaMemcpyDeviceToHost, st));
    FC_CUDA(cudaStreamSynchronize(st));
    // device rows are pair-planar: floats (re_2c, re_2c+1, im_2c, im_2c+1) per column pair
    float2* o = reinterpret_cast<float2*>(out);
    const float* hf = reinterpret_cast<const float*>(h.data());
    for (size_t r = 0; r < (size_t)g.nz * g.ny; ++r)
        for (int k = 0; k < g.xc; ++k) {
            const size_t f = r * g.xcp * 2 + (size_t)((k >> 1) << 2) + (k & 1);
            o[r * g.xc + k] = make_float2(hf[f], hf[f + 2]);
        }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// reference ABI
// ------------------------------------------------------------------------------------------------
void convolution3DfftCUDAInPlace(imageType* im, int* imDim, imageType* kernel, int* kernelDim, int devCUDA)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        // reference: stack_shape = reverse(imDim) => nx = imDim[0]; placement gets (kernelDim, imDim)
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
        convolve_core(im, imDim[0], imDim[1], imDim[2], kernel, pdims, devCUDA, false, nullptr);
    });
}

void convolution3DfftCUDAInPlaceSaveMemory(imageType* im, int* imDim, imageType* kernel, int* kernelDim, int devCUDA)
{
    // Same contract as InPlace, same numbers up to fp32 round-off, but the image-sized PSF spectrum is never
    // materialised when the placed PSF spans <= 16 z planes ("SaveMemory").
    // A HOST volume that does not fit on devCUDA (or FCB200_SLAB=1) is spread in z slabs over every device with
    // peer access to devCUDA: the 3D transposes become peer stores over NVLink and each device builds only its own
    // slab of the PSF spectrum (fc_multi.cu) -- BASELINE config 5, which overflows `int` in the reference
    // (src/convolution3Dfft.cu:423-436).
    guarded([&] {
        check_dims(imDim, kernelDim);
        if (classify_pointer(im, devCUDA) != HostMem::Device) {
            const std::vector<int> devs = slab_devices_for(imDim, devCUDA, true);
            if (devs.size() >= 2) {
                slab_convolve(im, nullptr, imDim, kernel, kernelDim, devs.data(), (int)devs.size());
                return;
            }
        }
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
        convolve_core(im, imDim[0], imDim[1], imDim[2], kernel, pdims, devCUDA, false, nullptr, true);
    });
}

void fcb200_convolve_batch(imageType* const* ims, int n, const int* imDim, const imageType* kernel, const int* kernelDim,
                           int devCUDA)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        if (!ims && n > 0) throw std::runtime_error("fcb200: ims is NULL");
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
        batch_single(ims, n, imDim[0], imDim[1], imDim[2], kernel, pdims, devCUDA, false);
    });
}

void fcb200_padded_extents(const int* imDim, const int* kernelDim, int policy, int* padDim)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        if (!padDim) throw std::runtime_error("fcb200: padDim is NULL");
        if (policy != 0 && policy != 1) throw std::runtime_error("fcb200: padding policy must be 0 (exact) or 1 (7-smooth)");
        padded_extents(imDim, kernelDim, policy, padDim);
    });
}

void fcb200_convolve_padded(imageType* im, const int* imDim, const imageType* kernel, const int* kernelDim, int mode,
                            int policy, int devCUDA)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        padded_core(im, imDim, kernel, kernelDim, mode, policy, devCUDA, false, nullptr);
    });
}

void fcb200_convolve_batch_padded(imageType* const* ims, int n, const int* imDim, const imageType* kernel,
                                  const int* kernelDim, int mode, int policy, int devCUDA)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        if (!ims && n > 0) throw std::runtime_error("fcb200: ims is NULL");
        int pd[3];
        const PadGeom g = make_pad_geom(imDim, kernelDim, mode, policy, pd);
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], pd[0], pd[1], pd[2]};
        batch_single(ims, n, pd[0], pd[1], pd[2], kernel, pdims, devCUDA, false, &g);
    });
}

void fcb200_convolve_padded_device_async(imageType* im_dev, const int* imDim, const imageType* kernel_dev,
                                         const int* kernelDim, int mode, int policy, int devCUDA, void* stream)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        padded_core(im_dev, imDim, kernel_dev, kernelDim, mode, policy, devCUDA, true, (cudaStream_t)stream);
    });
}

void fcb200_convolve_device_async_savememory(imageType* im_dev, const int* imDim, const imageType* kernel_dev,
                                             const int* kernelDim, int devCUDA, void* stream)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
        convolve_core(im_dev, imDim[0], imDim[1], imDim[2], kernel_dev, pdims, devCUDA, true, (cudaStream_t)stream, true);
    });
}

void fcb200_convolve_device_async(imageType* im_dev, const int* imDim, const imageType* kernel_dev,
                                  const int* kernelDim, int devCUDA, void* stream)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
        convolve_core(im_dev, imDim[0], imDim[1], imDim[2], kernel_dev, pdims, devCUDA, true, (cudaStream_t)stream);
    });
}

// Legacy entry points: the reference versions were written for cuFFT's removed "native" layout and
// are flagged as not ported (src/convolution3Dfft.cu:215, :298).  Implemented here with their
// intended, self-consistent legacy convention (imDim[2] fastest for image, kernel and placement).
imageType* convolution3DfftCUDA(imageType* im, int* imDim, imageType* kernel, int* kernelDim, int devCUDA)
{
    return guarded([&]() -> imageType* {
        check_dims(imDim, kernelDim);
        const size_t n = (size_t)imDim[0] * imDim[1] * imDim[2];
        std::unique_ptr<imageType[]> out(new imageType[n]);
        std::memcpy(out.get(), im, n * sizeof(imageType));
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
        convolve_core(out.get(), imDim[2], imDim[1], imDim[0], kernel, pdims, devCUDA, false, nullptr);
        return out.release();
    });
}

imageType* convolution3DfftCUDA_test(imageType* im, int* imDim, imageType* kernel, int devCUDA)
{
    return guarded([&]() -> imageType* {
        check_dims(imDim, nullptr);
        const int nx = imDim[2], ny = imDim[1], nz = imDim[0];
        const size_t n = (size_t)nx * ny * nz;
        std::unique_ptr<imageType[]> out(new imageType[n]);
        DeviceGuard guard(devCUDA);
        auto plan = get_plan(devCUDA, nx, ny, nz);
        std::lock_guard<std::mutex> lock(plan->mu);
        ConvPlan& p = *plan;
        cudaStream_t st = p.stream;
        if (!p.d_real) FC_CUDA(device_alloc_retry(&p.d_real, p.real_bytes()));
        // kernel is already image-sized and used as is (no shift), reference :253, :270
        FC_CUDA(cudaMemcpyAsync(p.d_real, kernel, n * sizeof(float), cudaMemcpyHostToDevice, st));
        ensure_full_workspace(p);
        p.h_valid = false;
        run_forward(p, p.d_real, p.d_H, 3, st);
        FC_CUDA(cudaMemcpyAsync(p.d_real, im, n * sizeof(float), cudaMemcpyHostToDevice, st));
        run_convolve(p, p.d_real, st);
        FC_CUDA(cudaMemcpyAsync(out.get(), p.d_real, n * sizeof(float), cudaMemcpyDeviceToHost, st));
        FC_CUDA(cudaStreamSynchronize(st));
        return out.release();
    });
}

void fcb200_free_result(imageType* p) { delete[] p; }

// ---- device queries (reference: src/standardCUDAfunctions.cu:13-71) ---------------------------
int getNumDevicesCUDA(void)
{
    return guarded([&] {
        int count = 0;
        FC_CUDA(cudaGetDeviceCount(&count));
        return count;
    });
}

int getCUDAcomputeCapabilityMajorVersion(int devCUDA)
{
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, devCUDA) != cudaSuccess) {
        cudaGetLastError();
        return 0;  // the reference returns 0 when the query fails (unchecked cuDeviceComputeCapability)
    }
    return v;
}

int getCUDAcomputeCapabilityMinorVersion(int devCUDA)
{
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, devCUDA) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return v;
}

int selectDeviceWithHighestComputeCapability(void)
{
    return guarded([&] {
        int n = 0;
        FC_CUDA(cudaGetDeviceCount(&n));
        int best = 0, value = -1;  // first device wins ties; -1 when there is none (reference :13-33)
        for (int d = 0; d < n; ++d) {
            int meta = 10 * getCUDAcomputeCapabilityMajorVersion(d) + getCUDAcomputeCapabilityMinorVersion(d);
            if (meta > best) {
                best = meta;
                value = d;
            }
        }
        return value;
    });
}

void getNameDeviceCUDA(int devCUDA, char* name)
{
    guarded([&] {
        cudaDeviceProp prop;
        FC_CUDA(cudaGetDeviceProperties(&prop, devCUDA));
        std::memcpy(name, prop.name, sizeof(char) * 256);  // exactly 256 bytes, reference :58-64
    });
}

long long int getMemDeviceCUDA(int devCUDA)
{
    return guarded([&] {
        cudaDeviceProp prop;
        FC_CUDA(cudaGetDeviceProperties(&prop, devCUDA));
        return (long long int)prop.totalGlobalMem;
    });
}

int cuda_version(void) { return CUDART_VERSION; }

long long fcb200_workspace_bytes(const int* imDim)
{
    Geometry g = make_geometry(imDim[0], imDim[1], imDim[2]);
    const long long spec = (long long)g.nz * g.ny * g.xcp * (long long)sizeof(float2);
    const long long tables = 16LL * ((long long)g.M + g.ny + g.nz) + 8LL * (g.M + 1);
    return 2 * spec + tables;
}

int gpu_mem_needed_mb(int* shape, int len)
{
    return guarded([&] {
        if (!shape || len < 1 || len > 3) throw std::runtime_error("fcb200: gpu_mem_needed_mb needs len in {1,2,3}");
        // cufftEstimate{1,2,3}d(shape...) treats the LAST extent as the fastest one (reference :587-603)
        int dims[3] = {1, 1, 1};
        for (int i = 0; i < len; ++i) dims[i] = shape[len - 1 - i];
        check_dims(dims, nullptr);
        return (int)(fcb200_workspace_bytes(dims) / (1 << 20));
    });
}

// ------------------------------------------------------------------------------------------------
// extensions
// ------------------------------------------------------------------------------------------------
const char* fcb200_last_error(void) { return g_last_error.c_str(); }
void fcb200_set_error_mode(int mode) { g_error_mode.store(mode == 1 ? 1 : 0); }

int fcb200_plan_radices(int L, int* radices, int* generic) { return fcb200_plan_radices_style(L, 0, radices, generic); }

int fcb200_plan_radices_style(int L, int style, int* radices, int* generic)
{
    return guarded([&] {
        if (L < 1) throw std::runtime_error("fcb200: L must be >= 1");
        bool gen = false;
        std::vector<int> r = factorize(L, &gen, style);
        for (size_t i = 0; i < r.size(); ++i) radices[i] = r[i];
        if (generic) *generic = gen ? 1 : 0;
        return (int)r.size();
    });
}

void fcb200_plan_tables(int L, int* rev, int* pos, float* tw) { fcb200_plan_tables_style(L, 0, rev, pos, tw); }

void fcb200_plan_tables_style(int L, int style, int* rev, int* pos, float* tw)
{
    guarded([&] {
        bool gen = false;
        std::vector<int> r = factorize(L, &gen, style), hrev, hpos;
        std::vector<float2> htw;
        build_tables(L, r, hrev, hpos, htw);
        if (rev) std::memcpy(rev, hrev.data(), sizeof(int) * L);
        if (pos) std::memcpy(pos, hpos.data(), sizeof(int) * L);
        if (tw) std::memcpy(tw, htw.data(), sizeof(float2) * L);
    });
}

int fcb200_plan_rader(int p, int* radices, int* perm, int* iperm, float* bf, float* bi)
{
    return guarded([&] {
        RaderTables r;
        if (!build_rader(p, r)) return 0;
        if (radices)
            for (size_t i = 0; i < r.radix.size(); ++i) radices[i] = r.radix[i];
        if (perm) std::memcpy(perm, r.perm.data(), sizeof(int) * r.n);
        if (iperm) std::memcpy(iperm, r.iperm.data(), sizeof(int) * r.n);
        if (bf) std::memcpy(bf, r.bf.data(), sizeof(float2) * r.n);
        if (bi) std::memcpy(bi, r.bi.data(), sizeof(float2) * r.n);
        return (int)r.radix.size();
    });
}

int fcb200_spectrum_pitch(int nx) { return make_geometry(nx, 1, 1).xcp; }

long long fcb200_psf_active_rows(const int* imDim, const int* kernelDim, int* rows, long long cap)
{
    return guarded([&] {
        check_dims(imDim, kernelDim);
        std::vector<int> r = psf_active_rows(imDim, kernelDim, imDim[0]);
        if (rows)
            for (long long i = 0; i < (long long)r.size() && i < cap; ++i) rows[i] = r[(size_t)i];
        return (long long)r.size();
    });
}

void fcb200_debug_rfft3(const imageType* im, const int* imDim, float* spec, int passes, int devCUDA)
{
    guarded([&] {
        check_dims(imDim, nullptr);
        DeviceGuard guard(devCUDA);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2]);
        std::lock_guard<std::mutex> lock(plan->mu);
        ConvPlan& p = *plan;
        if (!p.d_real) FC_CUDA(device_alloc_retry(&p.d_real, p.real_bytes()));
        FC_CUDA(cudaMemcpyAsync(p.d_real, im, p.real_bytes(), cudaMemcpyHostToDevice, p.stream));
        run_forward(p, p.d_real, p.d_spec, passes, p.stream);
        download_spectrum(p, p.d_spec, spec, p.stream);
    });
}

void fcb200_debug_irfft3(const float* spec, const int* imDim, imageType* out, int devCUDA)
{
    guarded([&] {
        check_dims(imDim, nullptr);
        DeviceGuard guard(devCUDA);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2]);
        std::lock_guard<std::mutex> lock(plan->mu);
        ConvPlan& p = *plan;
        const Geometry& g = p.g;
        std::vector<float2> h((size_t)g.nz * g.ny * g.xcp, make_float2(0.f, 0.f));
        const float2* s = reinterpret_cast<const float2*>(spec);
        float* hf = reinterpret_cast<float*>(h.data());
        for (size_t r = 0; r < (size_t)g.nz * g.ny; ++r)
            for (int k = 0; k < g.xc; ++k) {
                const size_t f = r * g.xcp * 2 + (size_t)((k >> 1) << 2) + (k & 1);
                hf[f] = s[r * g.xc + k].x;
                hf[f + 2] = s[r * g.xc + k].y;
            }
        if (!p.d_real) FC_CUDA(device_alloc_retry(&p.d_real, p.real_bytes()));
        FC_CUDA(cudaMemcpyAsync(p.d_spec, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice, p.stream));
        run_inverse(p, p.d_spec, p.d_real, p.stream);
        FC_CUDA(cudaMemcpyAsync(out, p.d_real, p.real_bytes(), cudaMemcpyDeviceToHost, p.stream));
        FC_CUDA(cudaStreamSynchronize(p.stream));
    });
}

void fcb200_debug_psf_spectrum(const imageType* kernel, const int* kernelDim, const int* imDim, float* spec, int devCUDA)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        DeviceGuard guard(devCUDA);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2]);
        std::lock_guard<std::mutex> lock(plan->mu);
        ConvPlan& p = *plan;
        const size_t ktaps = (size_t)kernelDim[0] * kernelDim[1] * kernelDim[2];
        if (ktaps > p.kernel_cap) {
            cudaFree(p.d_kernel);
            p.d_kernel = nullptr;
            p.kernel_cap = 0;
            FC_CUDA(cudaMalloc(&p.d_kernel, ktaps * sizeof(float)));
            p.kernel_cap = ktaps;
        }
        FC_CUDA(cudaMemcpyAsync(p.d_kernel, kernel, ktaps * sizeof(float), cudaMemcpyHostToDevice, p.stream));
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
        run_psf_spectrum(p, p.d_kernel, pdims, p.stream);
        download_spectrum(p, p.d_H, spec, p.stream);
    });
}

// ---- slab-decomposed single volume ---------------------------------------------------------------
void fcb200_slab_xy_forward(const imageType* real_slab, float* zslab_spec, float* send, const int* imDim, int nzl, int nzp,
                            int nyl, int devCUDA, void* stream)
{
    guarded([&] {
        check_dims(imDim, nullptr);
        DeviceGuard guard(devCUDA);
        check_real_alignment(real_slab, imDim[0]);
        check_spec_alignment(zslab_spec);
        check_spec_alignment(send);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2], false);
        std::lock_guard<std::mutex> lock(plan->mu);
        run_slab_xy_forward(*plan, real_slab, reinterpret_cast<float2*>(zslab_spec), reinterpret_cast<float2*>(send), nzl,
                            nyl, (cudaStream_t)stream, nullptr, 0, nzp);
    });
}

void fcb200_slab_z_fused(float* yslab_spec, const float* H_yslab, const int* imDim, int nyl, int devCUDA, void* stream)
{
    guarded([&] {
        check_dims(imDim, nullptr);
        DeviceGuard guard(devCUDA);
        check_spec_alignment(yslab_spec);
        check_spec_alignment(H_yslab);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2], false);
        std::lock_guard<std::mutex> lock(plan->mu);
        run_slab_z_fused(*plan, reinterpret_cast<float2*>(yslab_spec), reinterpret_cast<const float2*>(H_yslab), nyl,
                         (cudaStream_t)stream);
    });
}

void fcb200_slab_yx_inverse(const float* recv, float* zslab_spec, imageType* real_slab, const int* imDim, int nzl, int nzp,
                            int nyl, int devCUDA, void* stream)
{
    guarded([&] {
        check_dims(imDim, nullptr);
        DeviceGuard guard(devCUDA);
        check_real_alignment(real_slab, imDim[0]);
        check_spec_alignment(zslab_spec);
        check_spec_alignment(recv);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2], false);
        std::lock_guard<std::mutex> lock(plan->mu);
        run_slab_yx_inverse(*plan, reinterpret_cast<const float2*>(recv), reinterpret_cast<float2*>(zslab_spec), real_slab,
                            nzl, nyl, (cudaStream_t)stream, nzp);
    });
}

void fcb200_slab_xy_forward_peer(const imageType* real_slab, float* zslab_spec, void* const* peer_yslabs, const int* imDim,
                                 int nzl, int nzp, int nyl, int rank, int devCUDA, void* stream)
{
    guarded([&] {
        check_dims(imDim, nullptr);
        DeviceGuard guard(devCUDA);
        check_real_alignment(real_slab, imDim[0]);
        check_spec_alignment(zslab_spec);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2], false);
        std::lock_guard<std::mutex> lock(plan->mu);
        run_slab_xy_forward(*plan, real_slab, reinterpret_cast<float2*>(zslab_spec), nullptr, nzl, nyl, (cudaStream_t)stream,
                            reinterpret_cast<float2* const*>(peer_yslabs), rank, nzp);
    });
}

void fcb200_slab_z_fused_peer(float* yslab_spec, const float* H_yslab, void* const* peer_recv, const int* imDim, int nzp,
                              int nyl, int rank, int devCUDA, void* stream)
{
    guarded([&] {
        check_dims(imDim, nullptr);
        DeviceGuard guard(devCUDA);
        check_spec_alignment(yslab_spec);
        check_spec_alignment(H_yslab);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2], false);
        std::lock_guard<std::mutex> lock(plan->mu);
        run_slab_z_fused(*plan, reinterpret_cast<float2*>(yslab_spec), reinterpret_cast<const float2*>(H_yslab), nyl,
                         (cudaStream_t)stream, reinterpret_cast<float2* const*>(peer_recv), rank, nzp);
    });
}

void* fcb200_device_malloc(long long bytes, int devCUDA)
{
    return guarded([&]() -> void* {
        DeviceGuard guard(devCUDA);
        void* p = nullptr;
        FC_CUDA(cudaMalloc(&p, (size_t)bytes));
        FC_CUDA(cudaMemset(p, 0, (size_t)bytes));   // ragged slabs leave the pad rows / planes of a block unwritten
        return p;
    });
}

void fcb200_device_free(void* p, int devCUDA)
{
    guarded([&] {
        DeviceGuard guard(devCUDA);
        FC_CUDA(cudaFree(p));
    });
}

void fcb200_ipc_get_handle(void* dev_ptr, char* handle64)
{
    guarded([&] {
        cudaIpcMemHandle_t h;
        FC_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        std::memcpy(handle64, &h, 64);
    });
}

void* fcb200_ipc_open_handle(const char* handle64, int devCUDA)
{
    return guarded([&]() -> void* {
        DeviceGuard guard(devCUDA);
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, 64);
        void* p = nullptr;
        FC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        return p;
    });
}

void fcb200_ipc_close_handle(void* mapped_ptr, int devCUDA)
{
    guarded([&] {
        DeviceGuard guard(devCUDA);
        FC_CUDA(cudaIpcCloseMemHandle(mapped_ptr));
    });
}

long long fcb200_slab_psf_scratch_elems(const int* imDim, const int* kernelDim, int devCUDA)
{
    return guarded([&] {
        check_dims(imDim, kernelDim);
        DeviceGuard guard(devCUDA);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2], false);
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
        return (long long)psf_slab_scratch_elems(*plan, pdims);
    });
}

void fcb200_slab_psf(const imageType* kernel_dev, const int* kernelDim, const int* imDim, int y0, int nyl, float* H_yslab,
                     float* scratch, int devCUDA, void* stream)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        for (int i = 0; i < 3; ++i)
            if (kernelDim[i] > imDim[i]) throw std::runtime_error("fcb200: kernel larger than image");
        DeviceGuard guard(devCUDA);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2], false);
        std::lock_guard<std::mutex> lock(plan->mu);
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
        run_slab_psf(*plan, kernel_dev, pdims, y0, nyl, reinterpret_cast<float2*>(H_yslab),
                     reinterpret_cast<float2*>(scratch), (cudaStream_t)stream);
    });
}

// ---- single-process multi-GPU entry points (fc_multi.cu) ----------------------------------------------
void fcb200_convolve_slab(imageType* im, const int* imDim, const imageType* kernel, const int* kernelDim, const int* devs,
                          int ndev)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        if (!im) throw std::runtime_error("fcb200: im is NULL");
        slab_convolve(im, nullptr, imDim, kernel, kernelDim, devs, ndev);
    });
}

void fcb200_convolve_slab_device(imageType* const* slabs, const int* imDim, const imageType* kernel, const int* kernelDim,
                                 const int* devs, int ndev)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        if (!slabs) throw std::runtime_error("fcb200: slabs is NULL");
        slab_convolve(nullptr, slabs, imDim, kernel, kernelDim, devs, ndev);
    });
}

int fcb200_slab_last_timing(const int* imDim, const int* devs, int ndev, float* ms, int cap)
{
    return guarded([&] {
        check_dims(imDim, nullptr);
        return slab_last_timing(imDim, devs, ndev, ms, cap);
    });
}

int fcb200_slab_devices(const int* imDim, int devCUDA, int* devs, int cap)
{
    return guarded([&] {
        check_dims(imDim, nullptr);
        const std::vector<int> d = slab_devices_for(imDim, devCUDA, true);
        for (int i = 0; i < (int)d.size() && i < cap; ++i) devs[i] = d[(size_t)i];
        cudaSetDevice(devCUDA);
        return (int)d.size();
    });
}

void fcb200_convolve_batch_multi(imageType* const* ims, int n, const int* imDim, const imageType* kernel,
                                 const int* kernelDim, const int* devs, int ndev, int* blocks_per_dev)
{
    guarded([&] {
        check_dims(imDim, kernelDim);
        if (!ims && n > 0) throw std::runtime_error("fcb200: ims is NULL");
        batch_multi(ims, n, imDim, kernel, kernelDim, devs, ndev, blocks_per_dev);
    });
}

int fcb200_psf_window_planes(const int* imDim, const int* kernelDim, int devCUDA)
{
    return guarded([&] {
        check_dims(imDim, kernelDim);
        DeviceGuard guard(devCUDA);
        auto plan = get_plan(devCUDA, imDim[0], imDim[1], imDim[2]);
        std::lock_guard<std::mutex> lock(plan->mu);
        const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
        return psf_window_applies(*plan, pdims, plan->stream) ? plan->psf_window_planes : 0;
    });
}

void fcb200_release(void)
{
    release_multi();
    release_all_plans();
}
void fcb200_profile_enable(int on) { profile_enable(on); }
int fcb200_profile_read(float* ms_sum, long long* counts, int n) { return profile_read(ms_sum, counts, n); }
long long fcb200_launch_count(void) { return launch_count(); }
