// TEST / BENCHMARK INFRASTRUCTURE (oracle side, never on the product path).
//
// Times the DEVICE work of the reference's convolution3DfftCUDAInPlace with device-resident inputs, so that the
// reference's cuFFT build can be compared like for like with the product's device-resident number.  The reference's
// own translation unit is compiled in unmodified from where it lies (#include of /root/reference/src/convolution3Dfft.cu
// via -I, nothing is copied): the kernels launched below ARE the reference's fftShiftKernel and
// modulateAndNormalize_kernel, and the FFTs are the cuFFT plans it creates (cufftPlan3d(d2, d1, d0), :519, :544).
//
// One step = the device operations of /root/reference/src/convolution3Dfft.cu:442-547 in the reference's order:
//   cudaMemset shifted_kernel (:443) . fftShiftKernel (:454) . cudaMemset kernelPaddedCUDA (:469)
//   . dense rows -> FFTW-padded rows (:474-486) . cufftExecR2C image (:523) . cufftExecR2C PSF (:525)
//   . modulateAndNormalize_kernel (:530) . cufftExecC2R (:547)
// Left out, all in the reference's favour: cudaMalloc/cudaFree, the two cufftPlan3d per call (plans are created once
// here), the host<->device copies and host row loops, and the d2*d1 blocking D2D row copies of :474-486, which are
// issued as ONE cudaMemcpy2DAsync here.
#include "convolution3Dfft.cu"

extern "C" int ref_device_time(const int* imDim, const int* kernelDim, int steps, int warmup, float* ms_out /*[4]*/)
{
    const int d0 = imDim[0], d1 = imDim[1], d2 = imDim[2];
    const size_t N = (size_t)d0 * d1 * d2, Nc = (size_t)d2 * d1 * (d0 / 2 + 1);
    const size_t K = (size_t)kernelDim[0] * kernelDim[1] * kernelDim[2];
    cufftComplex *imCUDA = NULL, *kernelPaddedCUDA = NULL;
    imageType *shifted = NULL, *kernelCUDA = NULL;
    if (cudaMalloc(&imCUDA, Nc * sizeof(cufftComplex)) != cudaSuccess) return 1;
    if (cudaMalloc(&kernelPaddedCUDA, Nc * sizeof(cufftComplex)) != cudaSuccess) return 1;
    if (cudaMalloc(&shifted, N * sizeof(imageType)) != cudaSuccess) return 1;
    if (cudaMalloc(&kernelCUDA, K * sizeof(imageType)) != cudaSuccess) return 1;
    cudaMemset(imCUDA, 0x3c, Nc * sizeof(cufftComplex));     // finite floats; FFT timing does not depend on the data
    cudaMemset(kernelCUDA, 0x3c, K * sizeof(imageType));
    cufftHandle fwd, inv;
    if (cufftPlan3d(&fwd, d2, d1, d0, CUFFT_R2C) != CUFFT_SUCCESS) return 2;
    if (cufftPlan3d(&inv, d2, d1, d0, CUFFT_C2R) != CUFFT_SUCCESS) return 2;
    cudaEvent_t e[5];
    for (cudaEvent_t& x : e) cudaEventCreate(&x);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int it = 0; it < warmup + steps; ++it) {
        cudaEventRecord(e[0]);
        // PSF preparation (:442-491)
        cudaMemset(shifted, 0, N * sizeof(imageType));
        int numThreads = closest_multiplier((int)std::min((size_t)MAX_THREADS_CUDA, K));
        int numBlocks = (int)std::min((long long int)MAX_BLOCKS_CUDA, (long long int)(K + numThreads - 1) / numThreads);
        fftShiftKernel<<<numBlocks, numThreads>>>(kernelCUDA, shifted, kernelDim[0], kernelDim[1], kernelDim[2], d0, d1, d2);
        cudaMemset(kernelPaddedCUDA, 0, Nc * sizeof(cufftComplex));
        cudaMemcpy2DAsync(kernelPaddedCUDA, (size_t)(d0 / 2 + 1) * sizeof(cufftComplex), shifted, (size_t)d0 * sizeof(imageType),
                          (size_t)d0 * sizeof(imageType), (size_t)d1 * d2, cudaMemcpyDeviceToDevice, 0);
        cudaEventRecord(e[1]);
        // forward transforms (:523, :525)
        cufftExecR2C(fwd, (cufftReal*)imCUDA, imCUDA);
        cufftExecR2C(fwd, (cufftReal*)kernelPaddedCUDA, kernelPaddedCUDA);
        cudaEventRecord(e[2]);
        // multiply (:527-536)
        numThreads = (int)std::min((size_t)MAX_THREADS_CUDA, Nc);
        numBlocks = (int)std::min((long long int)MAX_BLOCKS_CUDA, (long long int)((Nc - 1 + numThreads) / numThreads));
        modulateAndNormalize_kernel<<<numBlocks, numThreads>>>(imCUDA, kernelPaddedCUDA, (long long int)Nc, 1.0f / (float)N);
        cudaEventRecord(e[3]);
        // inverse (:547)
        cufftExecC2R(inv, imCUDA, (cufftReal*)imCUDA);
        cudaEventRecord(e[4]);
        if (cudaEventSynchronize(e[4]) != cudaSuccess) return 3;
        if (it >= warmup)
            for (int i = 0; i < 4; ++i) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e[i], e[i + 1]);
                acc[i] += ms;
            }
    }
    for (int i = 0; i < 4; ++i) ms_out[i] = acc[i] / (float)std::max(1, steps);
    for (cudaEvent_t& x : e) cudaEventDestroy(x);
    cufftDestroy(fwd);
    cufftDestroy(inv);
    cudaFree(imCUDA);
    cudaFree(kernelPaddedCUDA);
    cudaFree(shifted);
    cudaFree(kernelCUDA);
    return cudaGetLastError() == cudaSuccess ? 0 : 4;
}
