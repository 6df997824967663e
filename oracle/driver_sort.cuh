/*
 * driver_sort.cuh — the ONE substitution the reference drivers need to compile with the CUDA 12.9 toolkit:
 * thrust::sort_by_key<int, Particle> (solver.cu:181, solver-unidyn.cu:331) instantiates radix-sort kernels whose static shared
 * memory exceeds 48 KB with 340-byte values (CUB 2.8), so ptxas rejects it on every architecture.  oracle/Makefile replaces that
 * one call, in a temporary copy of the driver, by fsg_driver_sort: the same stable sort by key on (key, slot) pairs followed by a
 * gather of the records — the same permutation.  TEST INFRASTRUCTURE ONLY; force-included (-include) into the patched driver.
 */
#pragma once
#include <thrust/copy.h>
#include <thrust/device_ptr.h>
#include <thrust/device_vector.h>
#include <thrust/gather.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

template <typename Record>
static void fsg_driver_sort(thrust::device_ptr<int> keys, long n, thrust::device_ptr<Record> records)
{
    if (n <= 0) return;
    thrust::device_vector<int> perm(n);
    thrust::sequence(perm.begin(), perm.end());
    thrust::stable_sort_by_key(keys, keys + n, perm.begin());
    thrust::device_vector<Record> tmp(n);
    thrust::gather(perm.begin(), perm.end(), records, tmp.begin());
    thrust::copy(tmp.begin(), tmp.end(), records);
}
