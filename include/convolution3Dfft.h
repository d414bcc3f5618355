/* C ABI of the B200-native 3D FFT convolution library -- drop-in for
 * StephanPreibisch/FourierConvolutionCUDALib (reference header: src/convolution3Dfft.h).
 *
 * Every entry point below keeps the reference's name, argument order, argument meaning and return
 * type, so a JNA/ctypes/cgo binding written against the reference binds unchanged.  The citation
 * on each declaration is the reference interface it replaces.
 *
 * Conventions (identical to the reference, src/convolution3Dfft.cu:394-396,417-421):
 *   - im / kernel are dense fp32 arrays; imDim = {d0,d1,d2} with d0 the FASTEST running index of
 *     im (im[x + d0*(y + d1*z)]); kernelDim = {k0,k1,k2} with k2 the fastest index of kernel.
 *   - im is overwritten with the circular convolution on the grid the caller passed (callers pad).
 *   - pointers may be host pointers (reference behaviour) or -- extension, detected with
 *     cudaPointerGetAttributes -- device pointers on devCUDA, which skips the PCIe copies.
 * Errors: recoverable failures (out of memory, bad shape, CUDA errors) throw std::runtime_error
 * across the C boundary exactly where the reference throws for cuFFT failures
 * (src/book.h:112-123, used at src/convolution3Dfft.cu:519-547); the library never calls exit().
 * fcb200_last_error() additionally returns the message for callers that cannot catch C++.
 */
#ifndef __CONVOLUTION_3D_FFT_H__
#define __CONVOLUTION_3D_FFT_H__

#include "FourierConvolutionCUDALib_Export.h"

#ifdef __cplusplus
#  define FUNCTION_PREFIX extern "C" FourierConvolutionCUDALib_EXPORT
#else
#  define FUNCTION_PREFIX FourierConvolutionCUDALib_EXPORT
#endif

/* public constants of the installed reference header (src/convolution3Dfft.h:17-21) */
typedef float imageType;
static const int MAX_THREADS_CUDA = 1024;
static const int MAX_BLOCKS_CUDA = 65535;
static const int dimsImage = 3;

/* ---- the hot path ------------------------------------------------------------------------- */

/* replaces src/convolution3Dfft.h:56 (impl. src/convolution3Dfft.cu:399-578) */
FUNCTION_PREFIX void convolution3DfftCUDAInPlace(imageType* im, int* imDim, imageType* kernel, int* kernelDim, int devCUDA);

/* replaces src/convolution3Dfft.h:58-64 (declaration commented out in the reference snapshot; the
 * surviving device code is src/convolution3Dfft.cu:65-125).  Same contract as InPlace ("like
 * convolution3DfftCUDAInPlace but ... less memory"): here it is numerically identical to InPlace and
 * never materialises the image-sized PSF spectrum (see DESIGN.md "SaveMemory"). */
FUNCTION_PREFIX void convolution3DfftCUDAInPlaceSaveMemory(imageType* im, int* imDim, imageType* kernel, int* kernelDim, int devCUDA);

/* ---- legacy out-of-place entry points (returned buffer is new[]-allocated; free with
 *      fcb200_free_result or delete[]) ------------------------------------------------------- */

/* replaces src/convolution3Dfft.h:29-32 (impl. :216-292); kernel already image-sized, imDim[2] fastest */
FUNCTION_PREFIX imageType* convolution3DfftCUDA_test(imageType* im, int* imDim, imageType* kernel, int devCUDA);
/* replaces src/convolution3Dfft.h:41-45 (impl. :299-391); imDim[2] fastest (legacy convention) */
FUNCTION_PREFIX imageType* convolution3DfftCUDA(imageType* im, int* imDim, imageType* kernel, int* kernelDim, int devCUDA);

/* ---- device queries (src/convolution3Dfft.h:66-72, impl. src/standardCUDAfunctions.cu:13-71) */
FUNCTION_PREFIX int selectDeviceWithHighestComputeCapability(void);
FUNCTION_PREFIX int getCUDAcomputeCapabilityMinorVersion(int devCUDA);
FUNCTION_PREFIX int getCUDAcomputeCapabilityMajorVersion(int devCUDA);
FUNCTION_PREFIX int getNumDevicesCUDA(void);
FUNCTION_PREFIX void getNameDeviceCUDA(int devCUDA, char* name);   /* writes exactly 256 bytes */
FUNCTION_PREFIX long long int getMemDeviceCUDA(int devCUDA);
FUNCTION_PREFIX int cuda_version(void);                            /* src/convolution3Dfft.cu:581-584 */

/* replaces src/convolution3Dfft.h:84 (impl. :587-603, cufftEstimate{1,2,3}d): MiB of device workspace
 * this library's own plan needs for a real-to-complex transform of `shape` (len in {1,2,3}). */
FUNCTION_PREFIX int gpu_mem_needed_mb(int* shape, int len);

/* ---- extensions (not in the reference; see include/fcb200_ext.h for the rest) ------------- */
/* message of the last failure on the calling thread ("" if the last call succeeded) */
FUNCTION_PREFIX const char* fcb200_last_error(void);
/* 0 (default): failures throw std::runtime_error like the reference; 1: failures only record the
 * message (for callers such as ctypes/JNA that cannot catch a C++ exception) */
FUNCTION_PREFIX void fcb200_set_error_mode(int mode);
FUNCTION_PREFIX void fcb200_free_result(imageType* p);

#endif /* __CONVOLUTION_3D_FFT_H__ */
