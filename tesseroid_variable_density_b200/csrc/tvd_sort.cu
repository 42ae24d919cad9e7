// tvd_sort.cu -- ordering of the near-field pair keys ((point << 32) | piece).
//
// Bookkeeping, not arithmetic: the far-field sweep appends pairs with atomics, so their order
// is not reproducible; sorting them makes Phase C's per-point summation order (piece order) a
// function of the model only.  Uses the CUB radix sort shipped with the CUDA toolkit, limited
// to the key bits that are actually populated.  Kept in its own translation unit because CUB
// dominates compile time.
#include <cub/device/device_radix_sort.cuh>
#include "tvd_internal.cuh"

cudaError_t tvd_sort_keys(void *d_temp, size_t &temp_bytes, const unsigned long long *in,
                          unsigned long long *out, int64_t n, int end_bit, cudaStream_t stream)
{
    if (n > 0x7fffffffLL)
        return cudaErrorInvalidValue;
    cudaError_t err = cub::DeviceRadixSort::SortKeys(d_temp, temp_bytes, in, out, (int)n, 0,
                                                     end_bit, stream);
    if (d_temp != nullptr)
        tvd_count_launch(4);
    return err;
}

cudaError_t tvd_sort_pairs(void *d_temp, size_t &temp_bytes, const unsigned long long *keys_in,
                           unsigned long long *keys_out, const int32_t *vals_in,
                           int32_t *vals_out, int64_t n, cudaStream_t stream)
{
    if (n > 0x7fffffffLL)
        return cudaErrorInvalidValue;
    cudaError_t err = cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, keys_in, keys_out,
                                                      vals_in, vals_out, (int)n, 0, 64, stream);
    if (d_temp != nullptr)
        tvd_count_launch(8);
    return err;
}
