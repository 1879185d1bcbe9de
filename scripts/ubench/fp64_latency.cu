// Dependent-issue latencies on one warp of one SM (cycles per operation in a dependent chain), for modelling the
// single-CTA LDL^T kernel: DFMA, DMUL, fp64 reciprocal by float seed + 2 Newton steps, LDS (double), STS->LDS,
// DMMA m8n8k4 (accumulator chain), __syncthreads at 512 threads, bar.sync of 128 threads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/fp64_latency scripts/ubench/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }

__global__ void lat_kernel(double* out, long long* cyc, double seed, int iters) {
  __shared__ double sm[1024];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 1024; i += blockDim.x) sm[i] = seed + i * 1e-9;
  __syncthreads();
  double a = seed + lane * 1e-12, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // DFMA chain
  if (tid < 32) {
    t0 = clk();
    for (int i = 0; i < iters; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    t1 = clk(); if (tid == 0) cyc[0] = (t1 - t0) / (4 * iters);
    // DMUL chain
    t0 = clk();
    for (int i = 0; i < iters; i++) { a = a * b; a = a * b; a = a * b; a = a * b; }
    t1 = clk(); if (tid == 0) cyc[1] = (t1 - t0) / (4 * iters);
    // reciprocal chain: float seed + 2 Newton
    t0 = clk();
    for (int i = 0; i < iters; i++) {
      double d = a; double inv = (double)__frcp_rn((float)d); inv = inv * (2.0 - d * inv); inv = inv * (2.0 - d * inv); a = inv + 1.5;
    }
    t1 = clk(); if (tid == 0) cyc[2] = (t1 - t0) / iters;
    // LDS chain (pointer chasing through doubles holding indices)
    for (int i = lane; i < 1024; i += 32) sm[i] = (double)((i * 37 + 11) & 1023);
    __syncwarp();
    int idx = lane;
    t0 = clk();
    for (int i = 0; i < iters; i++) { idx = (int)sm[idx]; idx = (int)sm[idx]; idx = (int)sm[idx]; idx = (int)sm[idx]; }
    t1 = clk(); if (tid == 0) cyc[3] = (t1 - t0) / (4 * iters);
    a += idx;
    // plain LDS latency without the F2I: use int view
    int* smi = reinterpret_cast<int*>(sm);
    for (int i = lane; i < 2048; i += 32) smi[i] = (i * 37 + 11) & 2047;
    __syncwarp();
    idx = lane;
    t0 = clk();
    for (int i = 0; i < iters; i++) { idx = smi[idx]; idx = smi[idx]; idx = smi[idx]; idx = smi[idx]; }
    t1 = clk(); if (tid == 0) cyc[4] = (t1 - t0) / (4 * iters);
    a += idx;
    // DMMA accumulator chain
    double c0 = a, c1 = a * 0.5;
    t0 = clk();
    for (int i = 0; i < iters; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(b), "d"(c));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(b), "d"(c));
    }
    t1 = clk(); if (tid == 0) cyc[5] = (t1 - t0) / (2 * iters);
    a += c0 + c1;
    // STS -> LDS round trip (same thread)
    t0 = clk();
    for (int i = 0; i < iters; i++) { sm[lane] = a; a = sm[lane] + 1.0; sm[lane] = a; a = sm[lane] + 1.0; }
    t1 = clk(); if (tid == 0) cyc[6] = (t1 - t0) / (2 * iters);
    // 8 independent DFMA chains (throughput of one warp)
    double v[8];
    for (int k = 0; k < 8; k++) v[k] = a + k;
    t0 = clk();
    for (int i = 0; i < iters; i++)
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = fma(v[k], b, c);
    t1 = clk(); if (tid == 0) cyc[7] = (t1 - t0) * 100 / (8 * iters);   // x100
    for (int k = 0; k < 8; k++) a += v[k];
  }
  __syncthreads();
  // __syncthreads cost with the whole CTA arriving together
  t0 = clk();
  for (int i = 0; i < iters; i++) { __syncthreads(); __syncthreads(); }
  t1 = clk(); if (tid == 0) cyc[8] = (t1 - t0) / (2 * iters);
  if (tid < 128) {
    t0 = clk();
    for (int i = 0; i < iters; i++) { asm volatile("bar.sync 1, 128;" ::: "memory"); asm volatile("bar.sync 1, 128;" ::: "memory"); }
    t1 = clk(); if (tid == 0) cyc[9] = (t1 - t0) / (2 * iters);
  }
  // global load latency (dependent chain through L2)
  if (tid == 0) out[0] = a;
}

__global__ void gl_kernel(const int* __restrict__ p, int* out, long long* cyc, int iters) {
  int idx = threadIdx.x;
  long long t0 = clk();
  for (int i = 0; i < iters; i++) { idx = p[idx]; idx = p[idx]; }
  long long t1 = clk();
  if (threadIdx.x == 0) { cyc[10] = (t1 - t0) / (2 * iters); out[0] = idx; }
}

int main() {
  double* out; long long* cyc; int* p; int* io;
  cudaMalloc(&out, 64); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&io, 64);
  const int N = 1 << 20;
  cudaMalloc(&p, N * 4);
  int* h = new int[N];
  for (int i = 0; i < N; i++) h[i] = (int)(((long long)i * 7919 + 104729) % N);
  cudaMemcpy(p, h, N * 4, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; rep++) {
    lat_kernel<<<1, 512>>>(out, cyc, 1.2345, 2000);
    gl_kernel<<<1, 32>>>(p, io, cyc, 2000);
    cudaDeviceSynchronize();
  }
  long long hc[16];
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  printf("cycles per dependent op: DFMA %lld  DMUL %lld  rcp(seed+2 Newton)+add %lld  LDS.f64+F2I %lld  LDS.32 %lld  DMMA m8n8k4 %lld  "
         "STS->LDS %lld  DFMA issue (8 chains) %.2f  __syncthreads(512) %lld  bar.sync(128) %lld  LDG (L2 hit, dependent) %lld\n",
         hc[0], hc[1], hc[2], hc[3], hc[4], hc[5], hc[6], hc[7] / 100.0, hc[8], hc[9], hc[10]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
