// Cycles per iteration of the tile-product inner loop (4 LDS.64 + 2 DMMA + pointer bumps) for 1..8 warps of one CTA,
// and of a variant that issues the next iteration's loads before the current DMMAs.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int swz(int c) { return ((c & 2) ? 6 : 0) ^ (c & 4); }
__device__ __forceinline__ int el(int r, int c) { return c * 8 + (r ^ swz(c)); }
__global__ void k(double* out, long long* cyc, int iters, int mode) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 104 * 64; i += blockDim.x) sm[i] = 1e-3 * (i % 17);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), g = lane >> 2, t = lane & 3;
  const int pg = (g & 1) ? 4 + (g >> 1) : (g >> 1);
  const int a0 = el(g, t), a1 = el(g, 4 + t), p0i = el(t, pg), p1i = el(4 + t, pg);
  double p0 = 0, p1 = 0, q0 = 0, q1 = 0;
  const double* Xr = sm + 64 * warp;
  const double* Wk = sm + 64 * (8 + warp);
  long long t0 = clock64();
  if (mode == 0) {
    for (int rep = 0; rep < iters; ++rep) {
      const double* xr = Xr; const double* wk = Wk;
      for (int K = 1; K < 13; ++K) {
        mma(p0, p1, xr[a0], wk[p0i]);
        mma(q0, q1, xr[a1], wk[p1i]);
        xr += 64; wk += 64 * (K + 1) % (64 * 40);
      }
    }
  } else {
    for (int rep = 0; rep < iters; ++rep) {
      const double* xr = Xr; const double* wk = Wk;
      double xa = xr[a0], xb = xr[a1], wa = wk[p0i], wb = wk[p1i];
      for (int K = 1; K < 13; ++K) {
        xr += 64; wk += 64 * (K + 1) % (64 * 40);
        const double nxa = xr[a0], nxb = xr[a1], nwa = wk[p0i], nwb = wk[p1i];
        mma(p0, p1, xa, wa);
        mma(q0, q1, xb, wb);
        xa = nxa; xb = nxb; wa = nwa; wb = nwb;
      }
    }
  }
  long long t1 = clock64();
  if (lane == 0) cyc[warp] = t1 - t0;
  out[threadIdx.x] = p0 + p1 + q0 + q1;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 64 * 8);
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 2, 4, 8}) {
      const int iters = 200;
      k<<<1, 32 * warps, 104 * 64 * 8>>>(out, cyc, iters, mode);
      cudaDeviceSynchronize();
      long long h[64];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("mode %d warps %d: %.1f cycles per iteration (2 DMMA + 4 LDS)  [%s]\n", mode, warps, (double)h[0] / (12.0 * iters), cudaGetErrorString(cudaGetLastError()));
    }
}
