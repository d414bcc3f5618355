/* Export macro for the drop-in FourierConvolutionCUDALib replacement.
 *
 * Replaces: /root/reference/src/FourierConvolutionCUDALib_Export.h:5-22 (CMake GenerateExportHeader
 * output used by src/convolution3Dfft.h:4-13).  Same macro names so code that includes the reference
 * header keeps compiling; on ELF platforms the macro marks default visibility (the library is built
 * with -fvisibility=hidden so that only the C ABI is exported).
 */
#ifndef FourierConvolutionCUDALib_EXPORT_H
#define FourierConvolutionCUDALib_EXPORT_H

#if defined(FourierConvolutionCUDALib_BUILT_AS_STATIC)
#  define FourierConvolutionCUDALib_EXPORT
#  define FOURIERCONVOLUTIONCUDALIB_NO_EXPORT
#elif defined(_WIN32)
#  if defined(FourierConvolutionCUDALib_EXPORTS)
#    define FourierConvolutionCUDALib_EXPORT __declspec(dllexport)
#  else
#    define FourierConvolutionCUDALib_EXPORT __declspec(dllimport)
#  endif
#  define FOURIERCONVOLUTIONCUDALIB_NO_EXPORT
#else
#  define FourierConvolutionCUDALib_EXPORT __attribute__((visibility("default")))
#  define FOURIERCONVOLUTIONCUDALIB_NO_EXPORT __attribute__((visibility("hidden")))
#endif

#if defined(_WIN32)
#  define FOURIERCONVOLUTIONCUDALIB_DEPRECATED __declspec(deprecated)
#else
#  define FOURIERCONVOLUTIONCUDALIB_DEPRECATED __attribute__((deprecated))
#endif
#define FOURIERCONVOLUTIONCUDALIB_DEPRECATED_EXPORT FourierConvolutionCUDALib_EXPORT FOURIERCONVOLUTIONCUDALIB_DEPRECATED
#define FOURIERCONVOLUTIONCUDALIB_DEPRECATED_NO_EXPORT FOURIERCONVOLUTIONCUDALIB_NO_EXPORT FOURIERCONVOLUTIONCUDALIB_DEPRECATED

#endif /* FourierConvolutionCUDALib_EXPORT_H */
