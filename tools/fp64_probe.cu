// Micro-probe: fp64 vector (DFMA) and tensor (DMMA m8n8k4) throughput on the current GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double x = 1.0000001, y = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c[16][2];
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int wpsm : {4, 8, 16, 32}) {
    int threads = 128, blocks = p.multiProcessorCount * wpsm / 4;
    int iters = 20000;
    dfma_kernel<<<blocks, threads>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * iters * (double)blocks * threads;
    printf("DFMA warps/SM %2d : %.2f TFLOP/s (%.3f ms)\n", wpsm, fl / ms * 1e-9, ms);
    dmma_kernel<<<blocks, threads>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dmma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 2.0 * 256 * 16 * iters * (double)blocks * (threads / 32);
    printf("DMMA warps/SM %2d : %.2f TFLOP/s (%.3f ms)\n", wpsm, fl / ms * 1e-9, ms);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
