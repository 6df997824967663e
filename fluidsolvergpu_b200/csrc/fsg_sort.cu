// fsg_sort.cu — stable LSD radix sort of (bin id, slot) pairs: the key half of the reference's
// thrust::sort_by_key (solver.cu:181, solver-unidyn.cu:331).  The 64-byte records are NOT carried
// through the radix passes (the reference drags 340 B through every pass); k_reorder gathers them
// once.  Only the low `bits` bits of the keys are sorted (bin ids are < numcells + 1).
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

// ------------------------------------------------------------------------------------------------
// The same stable sort for a key array that is ALMOST sorted already.  After a step the slots are in the order of the
// previous step's sorted keys, and only the particles that changed bin (a fraction of a percent per step at the
// reference's time step) are out of place.  With composite keys (bin id << 32 | slot) — whose order IS the stable
// order by bin id — the sort becomes:
//   partition  slots whose key did not change ("stayers": a sorted subsequence, kept in order) | the others ("movers")
//   radix sort of the movers only
//   merge      of the two sorted sequences, then split into the key and slot arrays k_reorder reads
// ≈ 48 B of traffic per particle instead of the 4 radix passes' ≈ 68 B at a third of the kernel time.  The result is
// VERIFIED on the device (strictly increasing composites); the caller falls back to the radix sort when the check
// fails (stale previous keys) or when too many particles moved.  Two small device-to-host reads per call.
// ------------------------------------------------------------------------------------------------
#include <cub/device/device_merge.cuh>
#include <cub/device/device_partition.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

typedef unsigned long long u64;
struct NsComposite {
    const int *keys;
    __host__ __device__ u64 operator()(int k) const { return ((u64)(unsigned)keys[k] << 32) | (u64)(unsigned)k; }
};
struct NsStayer {
    const int *prev;
    __host__ __device__ bool operator()(const u64 &c) const { return prev[(int)(unsigned)(c & 0xffffffffull)] == (int)(c >> 32); }
};
struct NsLess {
    __host__ __device__ bool operator()(const u64 &a, const u64 &b) const { return a < b; }
};
typedef thrust::transform_iterator<NsComposite, thrust::counting_iterator<int>, u64> NsIn;

size_t fsg_nsort_temp_bytes(int64_t n, int64_t movers_cap, int bits)
{
    size_t b1 = 0, b2 = 0, b3 = 0;
    NsIn in(thrust::counting_iterator<int>(0), NsComposite{nullptr});
    cub::DevicePartition::If(nullptr, b1, in, (u64 *)nullptr, (int *)nullptr, (int)n, NsStayer{nullptr});
    cub::DeviceRadixSort::SortKeys(nullptr, b2, (const u64 *)nullptr, (u64 *)nullptr, movers_cap, 0, 32 + bits);
    cub::DeviceMerge::MergeKeys(nullptr, b3, (const u64 *)nullptr, (int)n, (const u64 *)nullptr, (int)movers_cap, (u64 *)nullptr, NsLess{});
    size_t m = b1 > b2 ? b1 : b2;
    return m > b3 ? m : b3;
}

__global__ void __launch_bounds__(256)
k_nsort_split(const u64 *__restrict__ merged, int *__restrict__ keys, int *__restrict__ vals, int64_t n, int *bad)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const u64 v = merged[k];
    if (k > 0 && merged[k - 1] >= v) atomicOr(bad, 1);        // not the sorted order: the caller redoes it with the radix sort
    keys[k] = (int)(v >> 32);
    vals[k] = (int)(unsigned)(v & 0xffffffffull);
}

cudaError_t fsg_sort_nearly_sorted(void *tmp, size_t tmp_bytes, const int *keys_new, const int *keys_prev, int *keys_out, int *vals_out,
                                   unsigned long long *buf_a, unsigned long long *buf_b, unsigned long long *buf_c, int64_t movers_cap,
                                   int *dflags, int64_t n, int bits, cudaStream_t s, bool *done)
{
    *done = false;
    if (n <= 0 || n > 0x7fffffffll) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(dflags, 0, 2 * sizeof(int), s);
    if (e != cudaSuccess) return e;
    NsIn in(thrust::counting_iterator<int>(0), NsComposite{keys_new});
    size_t tb = tmp_bytes;
    e = cub::DevicePartition::If(tmp, tb, in, buf_a, dflags, (int)n, NsStayer{keys_prev}, s);
    if (e != cudaSuccess) return e;
    int nsel = 0;
    e = cudaMemcpyAsync(&nsel, dflags, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    const int64_t movers = n - nsel;
    if (movers > movers_cap) return cudaSuccess;               // too many particles changed bin: not "nearly sorted"
    const u64 *sorted_movers = buf_a + nsel;
    if (movers > 0) {
        tb = tmp_bytes;
        e = cub::DeviceRadixSort::SortKeys(tmp, tb, (const u64 *)(buf_a + nsel), buf_c, movers, 0, 32 + bits, s);
        if (e != cudaSuccess) return e;
        sorted_movers = buf_c;
    }
    tb = tmp_bytes;
    e = cub::DeviceMerge::MergeKeys(tmp, tb, (const u64 *)buf_a, nsel, sorted_movers, (int)movers, buf_b, NsLess{}, s);
    if (e != cudaSuccess) return e;
    k_nsort_split<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(buf_b, keys_out, vals_out, n, dflags + 1);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int bad = 0;
    e = cudaMemcpyAsync(&bad, dflags + 1, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    *done = bad == 0;
    return cudaSuccess;
}
