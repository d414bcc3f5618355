/* Extensions of the B200-native library that have no counterpart in the reference ABI.
 * They exist for (1) device-resident, stream-ordered use (benchmarks, multi-GPU tile scheduling),
 * (2) CPU-side tests of the planner, (3) pass-level debugging / parity tests.
 * Plain C ABI: pointers and sizes only. */
#ifndef FCB200_EXT_H
#define FCB200_EXT_H

#include "convolution3Dfft.h"

/* Same arithmetic as convolution3DfftCUDAInPlace (reference src/convolution3Dfft.cu:399-578) on
 * DEVICE pointers, enqueued on `stream` (a cudaStream_t; NULL = legacy default stream) without any
 * host synchronisation.  The PSF spectrum is recomputed on every call, like the reference does. */
FUNCTION_PREFIX void fcb200_convolve_device_async(imageType* im_dev, const int* imDim, const imageType* kernel_dev,
                                                 const int* kernelDim, int devCUDA, void* stream);

/* Same for convolution3DfftCUDAInPlaceSaveMemory (PSF spectrum derived on the fly). */
FUNCTION_PREFIX void fcb200_convolve_device_async_savememory(imageType* im_dev, const int* imDim,
                                                            const imageType* kernel_dev, const int* kernelDim,
                                                            int devCUDA, void* stream);

/* In-library padding -- the step the reference leaves to its callers (TODO at src/convolution3Dfft.h:39, :54;
 * done on the host by its tests: zero_padd::insert_at_offsets, tests/padd_utils.h:99-171, result read back through
 * the same sub-view, tests/test_fixtures.hpp:254-268).
 * `im` is the UNPADDED volume (imDim[0] fastest; host or device pointer), overwritten with the cropped result.
 * It is embedded at offsets kernelDim[i]/2 in a grid of fcb200_padded_extents() voxels and convolved there exactly
 * as convolution3DfftCUDAInPlace would convolve the caller-padded volume (same PSF placement on the padded grid).
 *   mode   0: zeros outside (the reference's zero_padd)      1: mirror, numpy.pad(mode="reflect")
 *   policy 0: padDim[i] = imDim[i] + 2*(kernelDim[i]/2) (tests/padd_utils.h:12-24)
 *          1: that, rounded up to the next 7-smooth size (fastest extent even): every FFT stage is radix 2..16/3/5/7
 * Only the unpadded bytes cross PCIe; zero halo planes are not transformed forward and no halo plane is
 * transformed back. */
FUNCTION_PREFIX void fcb200_padded_extents(const int* imDim, const int* kernelDim, int policy, int* padDim);
FUNCTION_PREFIX void fcb200_convolve_padded(imageType* im, const int* imDim, const imageType* kernel, const int* kernelDim,
                                           int mode, int policy, int devCUDA);
/* n unpadded blocks of one shape, one PSF: fcb200_convolve_batch (below) with the padding done in the library --
 * the multi-view deconvolution pattern (blocks + halo) in one call; transfers of neighbouring blocks overlap. */
FUNCTION_PREFIX void fcb200_convolve_batch_padded(imageType* const* ims, int n, const int* imDim, const imageType* kernel,
                                                 const int* kernelDim, int mode, int policy, int devCUDA);
/* device pointers, stream-ordered, no host synchronisation */
FUNCTION_PREFIX void fcb200_convolve_padded_device_async(imageType* im_dev, const int* imDim, const imageType* kernel_dev,
                                                        const int* kernelDim, int mode, int policy, int devCUDA,
                                                        void* stream);

/* Planner introspection (pure host code; callable without a GPU).
 * fcb200_plan_radices: writes the stage radices of a length-L transform (at most 16), returns the
 * number of stages; *generic is 1 when a prime factor above 23 is present (those stages run as direct sums between
 * two tile buffers; 2..16, the primes 11, 13, 17, 19, 23 and the composite radices 14, 18, 20, 21, 25, 28 of the
 * two-stage plans run as register butterflies). */
FUNCTION_PREFIX int fcb200_plan_radices(int L, int* radices, int* generic);
/* rev[p]: frequency at position p after the forward transform; pos = inverse permutation;
 * tw: L interleaved (re,im) roots exp(-2*pi*i*t/L).  Any pointer may be NULL. */
FUNCTION_PREFIX void fcb200_plan_tables(int L, int* rev, int* pos, float* tw);
/* Same with an explicit planning style: 0 = fewest stages, radix 16 allowed (x and y axes);
 * 1 = z axis, whose fused forward-multiply-inverse kernel is register-bound: L = 256 is planned as
 * (8,8,4) instead of (16,16);  2 = x axis: L = 1024 is planned as (16,8,8) for the row-wise kernel;
 * 3 = like 0 without the measured two-stage plans (300 = 20 * 15, 420 = 20 * 21, ...): sub-transforms of the Rader stage.
 * Styles 0, 1 and 2 each carry their own measured exceptions (csrc/fc_plan.cu: factorize). */
FUNCTION_PREFIX int fcb200_plan_radices_style(int L, int style, int* radices, int* generic);
FUNCTION_PREFIX void fcb200_plan_tables_style(int L, int style, int* rev, int* pos, float* tw);
/* Rader tables of a prime p (pure host code): the length-p DFT as a cyclic convolution of length n = p - 1 done with an
 * n-point transform of register radices.  Returns the number of stages of that transform (0: p - 1 is not smooth enough,
 * the direct sum is used) and writes, when the pointers are not NULL: its radices (at most 8), perm[m] = g^m mod p,
 * iperm[q] = g^(-q) mod p, and the spectra (n interleaved complex values each, divided by n, in the position order of the
 * n-point decimation-in-frequency transform, see fcb200_plan_tables(n)) of b[t] = exp(-+2*pi*i*g^(-t)/p). */
FUNCTION_PREFIX int fcb200_plan_rader(int p, int* radices, int* perm, int* iperm, float* bf, float* bi);
/* Spectrum row pitch (complex elements) used for a volume whose fastest extent is nx. */
FUNCTION_PREFIX int fcb200_spectrum_pitch(int nx);
/* Bytes of device workspace a cached plan for imDim holds (spectrum + PSF spectrum + tables). */
FUNCTION_PREFIX long long fcb200_workspace_bytes(const int* imDim);
/* On-the-fly PSF spectrum: when every z plane of the placed PSF (reference placement, src/convolution3Dfft.cu:128-166)
 * that holds a tap lies inside a window of 16, 32 or 64 consecutive planes (mod imDim[2]) and the z length has a
 * kernel for it, the fused z pass derives the PSF spectrum tile by tile from those planes and no image-sized PSF
 * spectrum is allocated, written or read.  Returns the window size the library uses for (imDim, kernelDim) on
 * devCUDA, 0 when it materialises the spectrum instead. */
FUNCTION_PREFIX int fcb200_psf_window_planes(const int* imDim, const int* kernelDim, int devCUDA);
/* Rows (z*d1+y) of the padded PSF volume that hold at least one tap (PSF pruning); returns count.
 * rows may be NULL. */
FUNCTION_PREFIX long long fcb200_psf_active_rows(const int* imDim, const int* kernelDim, int* rows, long long cap);

/* Pass-level debug entry points (host pointers).  Spectra are [d2][d1][d0/2+1] interleaved complex
 * in NATURAL frequency order, i.e. numpy.fft.rfftn of the [d2][d1][d0] volume.
 * passes: 1 = x only, 2 = x and y, 3 = full 3D. */
FUNCTION_PREFIX void fcb200_debug_rfft3(const imageType* im, const int* imDim, float* spec, int passes, int devCUDA);
/* Unnormalised inverse of the full 3D spectrum (what cufftExecC2R returns: N * irfftn). */
FUNCTION_PREFIX void fcb200_debug_irfft3(const float* spec, const int* imDim, imageType* out, int devCUDA);
/* Full 3D spectrum of the placed PSF (zero-pad + circular shift fused into the x pass). */
FUNCTION_PREFIX void fcb200_debug_psf_spectrum(const imageType* kernel, const int* kernelDim, const int* imDim,
                                              float* spec, int devCUDA);

/* Per-pass device timing with CUDA events recorded on the launching stream around every pass.
 * Pass ids: 0 psf_clear, 1 psf_x, 2 psf_y, 3 psf_z, 4 x_fwd, 5 y_fwd, 6 z_fused, 7 y_inv, 8 x_inv.
 * fcb200_profile_read waits for the recorded events, writes the summed milliseconds and launch counts
 * per pass id (n entries), resets the record and returns the number of pass ids (9). */
FUNCTION_PREFIX void fcb200_profile_enable(int on);
FUNCTION_PREFIX int fcb200_profile_read(float* ms_sum, long long* counts, int n);

/* ---- slab-decomposed single volume (one process per GPU; the transposes between the two
 * decompositions are done by the caller, e.g. NCCL all-to-all) -----------------------------------
 * Rank g of P owns the z planes [g*nzp, g*nzp + nzl) of the real volume [d2][d1][d0] and, after the exchange,
 * the ky rows [g*nyl, g*nyl + rows) of every plane, with the block pitches nzp = ceil(d2/P) and nyl = ceil(d1/P)
 * equal on all ranks.  Extents need not be divisible by P (ragged slabs): the last rank then owns fewer planes
 * (nzl < nzp) and fewer rows; the pad rows / planes of its blocks are carried along and never read back.
 * All pointers are DEVICE pointers; everything is enqueued on `stream` without host synchronisation.
 *   real slab           [nzl][d1][d0]   float
 *   z-slab spectrum     [nzl][d1][xcp]  complex (xcp = fcb200_spectrum_pitch(d0)), scratch for the passes
 *   exchange buffer     [P][nzp][nyl][xcp] complex: block p goes to / comes from rank p
 *   y-slab spectrum     [P*nzp][nyl][xcp] complex, planes 0..d2-1 valid
 * fcb200_slab_xy_forward : x + y forward passes; the y pass writes the exchange (send) buffer directly.
 * fcb200_slab_z_fused    : forward z, multiply by the PSF-spectrum slab and 1/N, inverse z, in place.
 * fcb200_slab_yx_inverse : y + x inverse passes; the y pass reads the exchange (receive) buffer directly.
 * fcb200_slab_psf        : the rank's y-slab [y0, min(y0+nyl, d1)) of the PSF spectrum, row pitch nyl (placement fused, pruned);
 *                          scratch must hold fcb200_slab_psf_scratch_elems() complex values. */
FUNCTION_PREFIX void fcb200_slab_xy_forward(const imageType* real_slab, float* zslab_spec, float* send, const int* imDim,
                                           int nzl, int nzp, int nyl, int devCUDA, void* stream);
FUNCTION_PREFIX void fcb200_slab_z_fused(float* yslab_spec, const float* H_yslab, const int* imDim, int nyl, int devCUDA,
                                        void* stream);
FUNCTION_PREFIX void fcb200_slab_yx_inverse(const float* recv, float* zslab_spec, imageType* real_slab, const int* imDim,
                                           int nzl, int nzp, int nyl, int devCUDA, void* stream);
/* Batch of n independent volumes of one shape convolved in place with ONE PSF (the multi-view deconvolution /
 * BASELINE config 4 pattern: reference callers loop over convolution3DfftCUDAInPlace, src/convolution3Dfft.h:56).
 * ims[b]: host pointers (pinned, registered or pageable) or device pointers on devCUDA, all of one kind.
 * The PSF spectrum is computed once; upload of block b+1, convolution of block b and download of block b-1
 * overlap (three device buffers, one copy stream per direction).  Results are identical to n separate calls. */
FUNCTION_PREFIX void fcb200_convolve_batch(imageType* const* ims, int n, const int* imDim, const imageType* kernel,
                                          const int* kernelDim, int devCUDA);

/* Fused compute + exchange over NVLink / NVSwitch peer memory (no NCCL on the data path):
 * fcb200_slab_xy_forward_peer : like fcb200_slab_xy_forward, but the y pass stores every output row straight
 *     into the y-slab buffer [P*nzp][nyl][xcp] of the rank that owns it; peer_yslabs is a DEVICE array of P
 *     pointers (entry p = rank p's y-slab buffer as mapped in this process, own entry included).
 * fcb200_slab_z_fused_peer    : like fcb200_slab_z_fused, but the last inverse stage stores every output plane
 *     straight into its owner's receive buffer [P][nzp][nyl][xcp] (block `rank`); peer_recv as above.
 * The caller separates the phases with a stream-ordered barrier across ranks.
 * fcb200_device_malloc / fcb200_ipc_*: plain cudaMalloc'ed buffers and CUDA IPC handles (64 bytes) so that
 * one-process-per-GPU callers can map each other's exchange buffers. */
FUNCTION_PREFIX void fcb200_slab_xy_forward_peer(const imageType* real_slab, float* zslab_spec, void* const* peer_yslabs,
                                                const int* imDim, int nzl, int nzp, int nyl, int rank, int devCUDA,
                                                void* stream);
FUNCTION_PREFIX void fcb200_slab_z_fused_peer(float* yslab_spec, const float* H_yslab, void* const* peer_recv,
                                             const int* imDim, int nzp, int nyl, int rank, int devCUDA, void* stream);
FUNCTION_PREFIX void* fcb200_device_malloc(long long bytes, int devCUDA);
FUNCTION_PREFIX void fcb200_device_free(void* p, int devCUDA);
FUNCTION_PREFIX void fcb200_ipc_get_handle(void* dev_ptr, char* handle64);
FUNCTION_PREFIX void* fcb200_ipc_open_handle(const char* handle64, int devCUDA);
FUNCTION_PREFIX void fcb200_ipc_close_handle(void* mapped_ptr, int devCUDA);
FUNCTION_PREFIX long long fcb200_slab_psf_scratch_elems(const int* imDim, const int* kernelDim, int devCUDA);
FUNCTION_PREFIX void fcb200_slab_psf(const imageType* kernel_dev, const int* kernelDim, const int* imDim, int y0, int nyl,
                                    float* H_yslab, float* scratch, int devCUDA, void* stream);

/* ---- single-process multi-GPU (one host process, several devices with peer access; no NCCL, no IPC) -------------
 * The reference's only multi-GPU hook is cudaSetDevice(devCUDA) (src/convolution3Dfft.cu:409): its callers run one
 * host thread per GPU.  These entry points do the spreading inside the library.
 *
 * fcb200_convolve_slab: ONE volume (host pointer, imDim[0] fastest) convolved in place like convolution3DfftCUDAInPlace,
 *   cut in z slabs over devs[0..ndev) (rank r = devs[r]; extents need not be divisible by ndev).  One worker thread
 *   per device; x+y forward -> the y pass stores every ky row straight into its owner's buffer over NVLink -> fused
 *   z pass with the rank's own slab of the PSF spectrum -> planes stored straight back -> y+x inverse.  The phases are
 *   ordered by CUDA events across devices.  Volumes of 2^31 voxels and more are fine (64-bit sizes throughout).
 *   convolution3DfftCUDAInPlaceSaveMemory routes here by itself when a host volume does not fit on devCUDA
 *   (FCB200_SLAB=1 forces it, =0 forbids it; FCB200_SLAB_DEVICES="0,1,.." picks the devices).
 * fcb200_convolve_slab_device: the same with the volume already resident: slabs[r] = device pointer on devs[r] of
 *   rank r's planes [r*nzp, min(d2, (r+1)*nzp)), nzp = ceil(d2/ndev).
 * fcb200_slab_last_timing: device times of the most recent call for (imDim, devs): ms[4*r + 0..3] = x+y forward,
 *   fused z (incl. waiting for the peers' rows), y+x inverse (incl. waiting), whole call; returns ranks written.
 * fcb200_slab_devices: the device list SaveMemory would use for a host volume of imDim on devCUDA (0 = single device).
 * fcb200_convolve_batch_multi: fcb200_convolve_batch over several devices -- one pipelined batch per device, all
 *   taking blocks from one shared counter (a device that finishes early takes more); blocks_per_dev (optional,
 *   ndev ints) receives how many blocks each device convolved.  Host blocks only. */
FUNCTION_PREFIX void fcb200_convolve_slab(imageType* im, const int* imDim, const imageType* kernel, const int* kernelDim,
                                         const int* devs, int ndev);
FUNCTION_PREFIX void fcb200_convolve_slab_device(imageType* const* slabs, const int* imDim, const imageType* kernel,
                                                const int* kernelDim, const int* devs, int ndev);
FUNCTION_PREFIX int fcb200_slab_last_timing(const int* imDim, const int* devs, int ndev, float* ms, int cap);
FUNCTION_PREFIX int fcb200_slab_devices(const int* imDim, int devCUDA, int* devs, int cap);
FUNCTION_PREFIX void fcb200_convolve_batch_multi(imageType* const* ims, int n, const int* imDim, const imageType* kernel,
                                                const int* kernelDim, const int* devs, int ndev, int* blocks_per_dev);

/* Frees every cached plan and its device workspace on all devices. */
FUNCTION_PREFIX void fcb200_release(void);
/* Number of kernels this library has launched since load (for benchmark accounting). */
FUNCTION_PREFIX long long fcb200_launch_count(void);

#endif /* FCB200_EXT_H */
