// Times cuSOLVER eigen-solvers on random complex Hermitian matrices: Xsyevd (divide & conquer, what jvmc_eigh calls)
// vs Zheevj (Jacobi).   nvcc -O2 -o build/heev_probe tools/heev_probe.cu -lcusolver -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <cuComplex.h>

static void fill(std::vector<cuDoubleComplex>& A, int n) {
  srand(1);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double re = rand() / (double)RAND_MAX - 0.5, im = (i == j) ? 0.0 : rand() / (double)RAND_MAX - 0.5;
      A[(size_t)j * n + i] = make_cuDoubleComplex(re, im);
      A[(size_t)i * n + j] = make_cuDoubleComplex(re, -im);
    }
}

int main(int argc, char** argv) {
  cusolverDnHandle_t h;
  cusolverDnCreate(&h);
  for (int a = 1; a < argc; ++a) {
    const int n = atoi(argv[a]);
    std::vector<cuDoubleComplex> A((size_t)n * n);
    fill(A, n);
    cuDoubleComplex* dA; double* dW; int* dInfo;
    cudaMalloc(&dA, sizeof(cuDoubleComplex) * n * n); cudaMalloc(&dW, sizeof(double) * n); cudaMalloc(&dInfo, sizeof(int));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // Xsyevd
    {
      cusolverDnParams_t p; cusolverDnCreateParams(&p);
      size_t wd = 0, wh = 0;
      cusolverDnXsyevd_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_C_64F, dA, n, CUDA_R_64F, dW,
                                  CUDA_C_64F, &wd, &wh);
      void* dwork; cudaMalloc(&dwork, wd); std::vector<char> hwork(wh + 1);
      float best = 1e30f;
      for (int r = 0; r < 3; ++r) {
        cudaMemcpy(dA, A.data(), sizeof(cuDoubleComplex) * n * n, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        cusolverDnXsyevd(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_C_64F, dA, n, CUDA_R_64F, dW, CUDA_C_64F,
                         dwork, wd, hwork.data(), wh, dInfo);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("n=%d Xsyevd (heevd): %.2f ms\n", n, best);
      cudaFree(dwork);
    }
    // Zheevj
    {
      syevjInfo_t params; cusolverDnCreateSyevjInfo(&params);
      cusolverDnXsyevjSetTolerance(params, 1e-14); cusolverDnXsyevjSetMaxSweeps(params, 100); cusolverDnXsyevjSetSortEig(params, 1);
      int lwork = 0;
      cusolverDnZheevj_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, &lwork, params);
      cuDoubleComplex* dwork; cudaMalloc(&dwork, sizeof(cuDoubleComplex) * (size_t)lwork);
      float best = 1e30f;
      for (int r = 0; r < 3; ++r) {
        cudaMemcpy(dA, A.data(), sizeof(cuDoubleComplex) * n * n, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        cusolverDnZheevj(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dW, dwork, lwork, dInfo, params);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      int sweeps = 0; double res = 0; cusolverDnXsyevjGetSweeps(h, params, &sweeps); cusolverDnXsyevjGetResidual(h, params, &res);
      printf("n=%d Zheevj (Jacobi): %.2f ms, %d sweeps, residual %.2e\n", n, best, sweeps, res);
      cudaFree(dwork);
    }
    cudaFree(dA); cudaFree(dW); cudaFree(dInfo);
  }
  return 0;
}
