// One CUDA CTA as the "Backend" of the host/device templates (octree_core.h, bow_core.h): thread index,
// barrier, shared-memory atomics and a block-wide exclusive scan.  The host twin is HostBackend
// (octree_core.h), which runs the same template single-threaded for the CPU tests.
#pragma once
#include <cuda_runtime.h>

namespace orbb200 {

struct CtaBackend {
  int* smem_ints;  // [0..3] user slots, [4..4+32] scan partials, [40] total
  __device__ int tid() const { return threadIdx.x; }
  __device__ int nthreads() const { return blockDim.x; }
  __device__ void sync() { __syncthreads(); }
  __device__ int atomic_add(int* p, int v) { return atomicAdd(p, v); }
  __device__ void atomic_max64(unsigned long long* p, unsigned long long v) { atomicMax(p, v); }
  __device__ int* shared_int(int i) { return smem_ints + i; }
  // In-place exclusive scan of d[0..n) by the whole CTA; every thread gets the total.
  __device__ int exclusive_scan(int* d, int n) {
    __syncthreads();
    const int nt = blockDim.x, t = threadIdx.x;
    const int chunk = (n + nt - 1) / nt;
    const int lo = min(t * chunk, n), hi = min(lo + chunk, n);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += d[i];
    // block exclusive scan of `sum`
    const int lane = t & 31, warp = t >> 5;
    int incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    int* wsum = smem_ints + 4;
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int nw = (nt + 31) >> 5;
      int v = lane < nw ? wsum[lane] : 0;
      int inc2 = v;
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc2, o);
        if (lane >= o) inc2 += u;
      }
      if (lane < nw) wsum[lane] = inc2 - v;
      if (lane == 31) smem_ints[40] = inc2;
    }
    __syncthreads();
    int acc = wsum[warp] + incl - sum;
    for (int i = lo; i < hi; i++) {
      int v = d[i];
      d[i] = acc;
      acc += v;
    }
    const int total = smem_ints[40];
    __syncthreads();
    return total;
  }
};

}  // namespace orbb200
