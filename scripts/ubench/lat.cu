// Dependent-chain latencies on B200 (cycles per op, one warp): DFMA, DMUL, DMMA m8n8k4, MUFU.RSQ64H, SHFL.IDX (64-bit), LDS.64
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = a;
  __syncthreads();
  double x = a + threadIdx.x * 1e-9, y = b;
  long long t0, t1;
  // DFMA
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DMUL
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = x * y; x = x * y; x = x * y; x = x * y; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // DMMA dependent through the accumulator
  double c0 = x, c1 = y;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // DMMA dependent through the A operand (accumulator of one feeds A of the next)
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(c1), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(c1), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(c1), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(c1), "d"(b));
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  x += c0 + c1;
  // rsqrt seed
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x)); asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x));
    asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x)); asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x));
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // library rsqrt
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = rsqrt(x); x = rsqrt(x); x = rsqrt(x); x = rsqrt(x); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // 64-bit shuffle
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 5) & 31); x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 7) & 31);
    x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 9) & 31); x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 11) & 31);
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // LDS.64 pointer chase
  int idx = threadIdx.x & 63;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    double v = sm[idx]; idx = (idx + (int)v) & 63; v = sm[idx]; idx = (idx + (int)v) & 63;
    v = sm[idx]; idx = (idx + (int)v) & 63; v = sm[idx]; idx = (idx + (int)v) & 63;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  // library exp and log chains
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = exp(-fabs(x)); x = exp(-fabs(x)); x = exp(-fabs(x)); x = exp(-fabs(x)); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = log(x + 2.0); x = log(x + 2.0); x = log(x + 2.0); x = log(x + 2.0); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
  // __syncthreads round trip with 8 warps
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[10] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + idx;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
  const char* names[] = {"DFMA", "DMUL", "DMMA(acc chain)", "DMMA(A chain)", "MUFU.RSQ64H seed", "rsqrt() lib", "SHFL 64-bit", "LDS.64+I2F chase", "exp() lib", "log() lib", "__syncthreads"};
  for (int threads : {32, 256}) {
    const int iters = 2000;
    k<<<1, threads>>>(out, cyc, iters, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads=%d:", threads);
    for (int i = 0; i < 11; ++i) printf("  %s %.1f", names[i], (double)h[i] / (4.0 * iters));
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
