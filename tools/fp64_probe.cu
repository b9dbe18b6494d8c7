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

// Same DMMA operand pattern as gram_s_kernel<1,5,10,2>: 3 A values, 5x2 B values, 10 accumulator pairs, 4 passes.
__global__ void dmma_gram_pattern(double* out, int iters) {
  double cre[5][2], cim[5][2];
  for (int i = 0; i < 5; ++i) cre[i][0] = cre[i][1] = cim[i][0] = cim[i][1] = 0.0;
  double ar = threadIdx.x * 1e-3, ai = 0.5 + threadIdx.x * 1e-3, nai = -ai;
  double br[5], bi[5];
  for (int i = 0; i < 5; ++i) { br[i] = 1.0 + i + threadIdx.x * 1e-4; bi[i] = 2.0 + i - threadIdx.x * 1e-4; }
#define DM(c, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" \
                                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b))
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 5; ++i) DM(cre[i], ar, br[i]);
#pragma unroll
    for (int i = 0; i < 5; ++i) DM(cim[i], ar, bi[i]);
#pragma unroll
    for (int i = 0; i < 5; ++i) DM(cre[i], ai, bi[i]);
#pragma unroll
    for (int i = 0; i < 5; ++i) DM(cim[i], nai, br[i]);
  }
  double s = 0;
  for (int i = 0; i < 5; ++i) s += cre[i][0] + cre[i][1] + cim[i][0] + cim[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent-issue probe: NACC accumulators used round-robin -> dependent DMMAs are NACC instructions apart
template <int NACC>
__global__ void dmma_dep_kernel(double* out, int iters) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16 / NACC; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run_dep(double* out, int sms, int wpsm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int threads = 32 * wpsm, blocks = sms, iters = 5000;
  dmma_dep_kernel<NACC><<<blocks, threads>>>(out, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  dmma_dep_kernel<NACC><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 2.0 * 256 * 16 * iters * (double)blocks * wpsm;
  printf("DMMA dep-distance %2d, warps/SM %2d : %.2f TFLOP/s\n", NACC, wpsm, fl / ms * 1e-9);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int wpsm : {4, 8, 16, 20, 32}) {
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
    dmma_gram_pattern<<<blocks, threads>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dmma_gram_pattern<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 2.0 * 256 * 20 * iters * (double)blocks * (threads / 32);
    printf("DMMA gram-pattern warps/SM %2d : %.2f TFLOP/s (%.3f ms)\n", wpsm, fl / ms * 1e-9, ms);
  }
  for (int w : {4, 8, 20}) {
    run_dep<1>(out, p.multiProcessorCount, w);
    run_dep<2>(out, p.multiProcessorCount, w);
    run_dep<4>(out, p.multiProcessorCount, w);
    run_dep<8>(out, p.multiProcessorCount, w);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
