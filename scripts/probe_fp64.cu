// Microbenchmarks that decide how the element-tile contraction is issued on B200 (sm_100a):
//   DFMA stream vs DMMA (mma.sync m8n8k4 f64) throughput, whether the two overlap, and the dependent-issue
//   latencies (DFMA, DMMA, LDS.64/LDS.128) that bound the short phases of the element kernels.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/probe_fp64 scripts/probe_fp64.cu
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// mode 0: all warps DFMA; 1: all warps DMMA; 2: even warps DFMA, odd warps DMMA
template <int NACC>
__global__ void __launch_bounds__(1024) mix_kernel(double *out, int iters, int mode, unsigned long long *ops) {
  const int warp = threadIdx.x / 32;
  const bool use_mma = (mode == 1) || (mode == 2 && (warp & 1));
  double a[2 * NACC];
  for (int k = 0; k < 2 * NACC; k++) a[k] = 1.0 + k + threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  if (use_mma) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < NACC; k++) dmma(a[2 * k], a[2 * k + 1], m, c);
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 2 * NACC; k++) a[k] = fma(a[k], m, c);
    }
  }
  double s = 0.0;
  for (int k = 0; k < 2 * NACC; k++) s += a[k];
  if (s == 12345.678) out[0] = s;
}

__global__ void lat_kernel(double *out, long long *cyc, int iters) {
  __shared__ __align__(16) double sh[512];
  for (int k = threadIdx.x; k < 512; k += blockDim.x) sh[k] = (double)((k * 8 + 8) % 4096);  // pointer chase in bytes
  __syncthreads();
  double x = 1.0 + threadIdx.x, m = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) x = fma(x, m, c);
  long long t1 = clock64();
  double c0 = x, c1 = x;
  for (int it = 0; it < iters; it++) dmma(c0, c1, m, c);
  long long t2 = clock64();
  // LDS.64 dependent chain
  int idx = threadIdx.x % 64;
  for (int it = 0; it < iters; it++) idx = (int)sh[idx] / 8;
  long long t3 = clock64();
  // DMUL -> DADD alternating (same pipe) chain
  double y = x;
  for (int it = 0; it < iters; it++) y = y * m + c0 * 0.0;
  long long t4 = clock64();
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0;
    cyc[1] = t2 - t1;
    cyc[2] = t3 - t2;
    cyc[3] = t4 - t3;
  }
  out[threadIdx.x] = x + c0 + c1 + idx + y;
}

int main() {
  double *out;
  long long *cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 64);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  const int warps_list[] = {4, 8, 12, 16, 32};
  for (int mode = 0; mode < 3; mode++) {
    for (int wi = 0; wi < 5; wi++) {
      const int warps = warps_list[wi];
      const int threads = warps * 32;
      mix_kernel<8><<<p.multiProcessorCount, threads>>>(out, 100, mode, nullptr);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      mix_kernel<8><<<p.multiProcessorCount, threads>>>(out, iters, mode, nullptr);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      // flops: DFMA thread op = 2 flop; DMMA warp op = 2*8*8*4 = 512 flop
      double dfma_warps = (mode == 0) ? warps : (mode == 2 ? warps / 2 : 0);
      double dmma_warps = (mode == 1) ? warps : (mode == 2 ? warps / 2 : 0);
      double fl = p.multiProcessorCount * (double)iters * (dfma_warps * 32 * 16 * 2.0 + dmma_warps * 8 * 512.0);
      printf("mode %d (%s) warps/SM %2d: %.3f ms  %.2f TFLOP/s\n", mode,
             mode == 0 ? "DFMA" : (mode == 1 ? "DMMA" : "DFMA+DMMA"), warps, ms, fl / ms * 1e-9);
    }
  }
  lat_kernel<<<1, 32>>>(out, cyc, 4096);
  cudaDeviceSynchronize();
  long long h[4];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("dependent latency (cycles/op, 1 warp): DFMA %.2f  DMMA %.2f  LDS.64 chase(+cvt) %.2f  DFMA-chain2 %.2f\n",
         h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / 4096.0);
  lat_kernel<<<1, 32 * 4>>>(out, cyc, 4096);
  cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("same with 4 warps (1 per SMSP): DFMA %.2f  DMMA %.2f  LDS %.2f\n", h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0);
  lat_kernel<<<1, 32 * 12>>>(out, cyc, 4096);
  cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("same with 12 warps (3 per SMSP): DFMA %.2f  DMMA %.2f  LDS %.2f\n", h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
