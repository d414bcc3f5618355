/* CPU oracle, plain C.  TEST INFRASTRUCTURE ONLY -- only tests/, smoke() and bench.py's
 * cpu_baseline leg may link or call this; the product never does.
 *
 * Restates the reference test-suite's host-side direct 3D convolution:
 *   /root/reference/tests/test_algorithms.hpp:10-58  (template convolve)
 * Same loop nest (z,y,x / kz,ky,kx), same flipped-kernel indexing, same float accumulator.
 * Arrays are [z][y][x] row-major (x fastest), like boost::multi_array's default storage order
 * used by the reference fixtures (tests/image_stack_utils.h).
 *
 * fc_oracle_convolve      : single thread, as written in the reference.
 * fc_oracle_convolve_omp  : the same arithmetic with the outer z loop shared between threads
 *                           (the reference has no threading on this loop; this is the
 *                           "all host cores" flavour BASELINE.md 2.2 asks for).
 */
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void conv_plane(const float* image, const int* ishape, const float* kernel, const int* kshape,
                       float* result, const int* offset, int image_z)
{
    const int hk0 = kshape[0] / 2, hk1 = kshape[1] / 2, hk2 = kshape[2] / 2;
    const size_t iy = (size_t)ishape[2], iz = (size_t)ishape[1] * ishape[2];
    const size_t ky = (size_t)kshape[2], kz = (size_t)kshape[1] * kshape[2];
    for (int image_y = offset[1]; image_y < ishape[1] - offset[1]; ++image_y) {
        for (int image_x = offset[2]; image_x < ishape[2] - offset[2]; ++image_x) {
            float value = 0.f;
            for (int kernel_z = 0; kernel_z < kshape[0]; ++kernel_z)
                for (int kernel_y = 0; kernel_y < kshape[1]; ++kernel_y)
                    for (int kernel_x = 0; kernel_x < kshape[2]; ++kernel_x) {
                        float kernel_value = kernel[(size_t)(kshape[0] - 1 - kernel_z) * kz +
                                                    (size_t)(kshape[1] - 1 - kernel_y) * ky +
                                                    (size_t)(kshape[2] - 1 - kernel_x)];
                        float image_value = image[(size_t)(image_z - hk0 + kernel_z) * iz +
                                                  (size_t)(image_y - hk1 + kernel_y) * iy +
                                                  (size_t)(image_x - hk2 + kernel_x)];
                        value += kernel_value * image_value;
                    }
            result[(size_t)image_z * iz + (size_t)image_y * iy + (size_t)image_x] = value;
        }
    }
}

/* ishape/kshape/offset are {z,y,x}; result must be pre-filled by the caller (the reference copies
 * the padded image into it first, tests/test_fixtures.hpp:212-215). */
void fc_oracle_convolve(const float* image, const int* ishape, const float* kernel, const int* kshape,
                        float* result, const int* offset)
{
    if ((size_t)ishape[0] * ishape[1] * ishape[2] == 0) return;
    for (int image_z = offset[0]; image_z < ishape[0] - offset[0]; ++image_z)
        conv_plane(image, ishape, kernel, kshape, result, offset, image_z);
}

/* returns the number of threads used; z_begin/z_end bound the outer loop so a caller can time a
 * bounded sample of a large volume. */
int fc_oracle_convolve_omp(const float* image, const int* ishape, const float* kernel, const int* kshape,
                           float* result, const int* offset, int z_begin, int z_end)
{
    int nthreads = 1;
    if ((size_t)ishape[0] * ishape[1] * ishape[2] == 0) return nthreads;
    if (z_begin < offset[0]) z_begin = offset[0];
    if (z_end > ishape[0] - offset[0]) z_end = ishape[0] - offset[0];
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int image_z = z_begin; image_z < z_end; ++image_z)
        conv_plane(image, ishape, kernel, kshape, result, offset, image_z);
    return nthreads;
}

int fc_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
