// Micro-benchmarks that guide the Riccati kernel design (not product code): dependent-issue latencies on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/lat tools/ubench/lat.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double *out, long long *cyc, int mode, int iters) {
  extern __shared__ double sm[];
  double a = 1.0 + threadIdx.x * 1e-9, b = 0.5, c0 = 0, c1 = 0, d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) for (int i = 0; i < iters; i++) dmma(c0, c1, a, b);                                    // dependent DMMA chain
  if (mode == 1) for (int i = 0; i < iters; i++) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); }             // 2 chains
  if (mode == 2) for (int i = 0; i < iters; i++) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); dmma(e0, e1, a, b); dmma(f0, f1, a, b); } // 4 chains
  if (mode == 3) for (int i = 0; i < iters; i++) c0 = fma(c0, a, b);                                      // DFMA chain
  if (mode == 4) for (int i = 0; i < iters; i++) c0 = rsqrt(c0 + 2.0);                                    // library rsqrt chain
  if (mode == 5) for (int i = 0; i < iters; i++) { double y = (double)rsqrtf((float)(c0 + 2.0)); double t = (c0 + 2.0) * y, e = fma(-t, y, 1.0); y = fma(0.5 * y, e, y); t = (c0 + 2.0) * y; e = fma(-t, y, 1.0); c0 = fma(0.5 * y, e, y); }
  if (mode == 6) { int idx = threadIdx.x & 31; for (int i = 0; i < iters; i++) { idx = (int)sm[idx] + (threadIdx.x & 31) - 1; } c0 = idx; } // dependent shared load
  if (mode == 7) for (int i = 0; i < iters; i++) __syncthreads();                                         // barrier
  if (mode == 8) for (int i = 0; i < iters; i++) { sm[threadIdx.x] = c0; __syncwarp(); c0 = sm[threadIdx.x ^ 1] + 1.0; __syncwarp(); } // st -> syncwarp -> ld round trip
  if (mode == 9) for (int i = 0; i < iters; i++) c0 = 1.0 / (c0 + 2.0);                                   // division chain
  if (mode == 10) for (int i = 0; i < iters; i++) { double x = c0 + 2.0; double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); double e = fma(-x, y, 1.0); y = fma(y, e, y); e = fma(-x, y, 1.0); c0 = fma(y, e, y); } // fast reciprocal
  if (mode == 11) for (int i = 0; i < iters; i++) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
  if (mode == 12) for (int i = 0; i < iters; i++) c0 = __shfl_xor_sync(0xffffffffu, c0, 1) + 1.0;        // double shuffle chain
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const char *names[] = {"dmma chain", "dmma 2 chains (per pair)", "dmma 4 chains (per quad)", "dfma chain", "rsqrt lib", "rsqrtf+2 newton", "dependent LDS", "__syncthreads 256thr",
                         "st/syncwarp/ld/syncwarp", "fp64 divide", "rcp.approx+2 newton", "bar.sync named 256", "shfl double"};
  for (int thr : {32, 256}) for (int mode = 0; mode < 13; mode++) {
    int iters = 2000;
    k<<<1, thr, 16384>>>(out, cyc, mode, iters); cudaDeviceSynchronize();
    k<<<1, thr, 16384>>>(out, cyc, mode, iters); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  %-28s %8.1f cycles/iter\n", thr, names[mode], (double)h / iters);
  }
  // throughput: all SMs, 8 warps x 4 chains
  return 0;
}
