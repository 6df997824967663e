// fsg_sort.cu — stable LSD radix sort of (bin id, slot) pairs: the key half of the reference's
// thrust::sort_by_key (solver.cu:181, solver-unidyn.cu:331).  The 64-byte records are NOT carried
// through the radix passes (the reference drags 340 B through every pass); k_reorder gathers them
// once.  Only the low `bits` bits of the keys are sorted (bin ids are < numcells + 1).
// This library radix sort runs where NOTHING is known about the order: the first step after an upload, and the stage API
// (the caller's key array).  Every later step sorts an almost sorted array with the hand-written kernels of fsg_nsort.cu.
#include "fsg_internal.cuh"

#include <cub/device/device_radix_sort.cuh>

size_t fsg_sort_temp_bytes(int64_t n, int bits)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned *)nullptr, (unsigned *)nullptr, (const int *)nullptr,
                                    (int *)nullptr, n, 0, bits);
    return bytes;
}

cudaError_t fsg_sort_pairs(void *tmp, size_t tmp_bytes, const int *keys_in, int *keys_out, const int *vals_in,
                           int *vals_out, int64_t n, int bits, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const unsigned *)keys_in, (unsigned *)keys_out, vals_in,
                                           vals_out, n, 0, bits, s);
}

// full-range signed keys (stage API: the caller's key array is sorted exactly as thrust::sort_by_key<int> would)
size_t fsg_sort_int_temp_bytes(int64_t n)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int *)nullptr, (int *)nullptr, (const int *)nullptr, (int *)nullptr, n);
    return bytes;
}
cudaError_t fsg_sort_pairs_int(void *tmp, size_t tmp_bytes, const int *keys_in, int *keys_out, const int *vals_in,
                               int *vals_out, int64_t n, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, 32, s);
}

