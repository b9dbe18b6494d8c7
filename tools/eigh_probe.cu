// Probe of the dense Hermitian eigensolvers available for the config-2 TDVP solve (P_c = 40 000 complex):
// cuSOLVER 64-bit API (Xsyevd), legacy Zheevd, cusolverMg on 1..G devices.  Prints status codes, workspace sizes
// and timings at sizes that finish in seconds; `--run N` runs one solver at size N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/eigh_probe tools/eigh_probe.cu -lcusolver -lcusolverMg -lcublas
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <cusolverMg.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>

__global__ void fill_hermitian(double2* A, long long n, unsigned seed) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // row
  long long j = blockIdx.y;                                        // column (column-major)
  if (i >= n || i < j) return;
  unsigned long long h = (unsigned long long)(i * 1315423911ull) ^ (unsigned long long)(j * 2654435761ull) ^ seed;
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
  double re = (double)(h & 0xFFFFFF) / 16777216.0 - 0.5, im = (double)((h >> 24) & 0xFFFFFF) / 16777216.0 - 0.5;
  if (i == j) { A[j * n + i] = make_double2(re + (double)n * 1e-3 * (double)(i % 97), 0.0); return; }
  A[j * n + i] = make_double2(re, im);
  A[i * n + j] = make_double2(re, -im);
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void probe_sizes(cusolverDnHandle_t h, cusolverDnParams_t p) {
  const int ns[] = {8192, 16384, 20000, 23000, 23170, 23171, 30000, 32768, 40000, 46340, 46341};
  for (int n : ns) {
    size_t d = 0, hh = 0;
    cusolverStatus_t s = cusolverDnXsyevd_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_C_64F,
                                                     nullptr, n, CUDA_R_64F, nullptr, CUDA_C_64F, &d, &hh);
    int lw = -1;
    cusolverStatus_t s2 = cusolverDnZheevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, nullptr, n,
                                                      nullptr, &lw);
    size_t dr = 0, hr = 0;
    cusolverStatus_t s3 = cusolverDnXsyevd_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F,
                                                      nullptr, n, CUDA_R_64F, nullptr, CUDA_R_64F, &dr, &hr);
    printf("n=%6d  Xsyevd C64: status %d dev %.2f GB host %.3f GB | Zheevd: status %d lwork %d | Xsyevd R64: status %d dev %.2f GB\n",
           n, (int)s, d / 1e9, hh / 1e9, (int)s2, lw, (int)s3, dr / 1e9);
  }
}

static void probe_other(cusolverDnHandle_t h, cusolverDnParams_t p) {
  const int ns[] = {16384, 32768, 36000, 40000};
  for (int n : ns) {
    size_t d = 0, hh = 0; int64_t meig = 0;
    cusolverStatus_t s1 = cusolverDnXsyevdx_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_LOWER, n,
                                                       CUDA_C_64F, nullptr, n, nullptr, nullptr, 1, n / 2, &meig, CUDA_R_64F, nullptr,
                                                       CUDA_C_64F, &d, &hh);
    printf("n=%6d Xsyevdx(I, half): status %d dev %.2f GB", n, (int)s1, d / 1e9);
    d = hh = 0;
    cusolverStatus_t s2 = cusolverDnXsyevBatched_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_C_64F, nullptr,
                                                            n, CUDA_R_64F, nullptr, CUDA_C_64F, &d, &hh, 1);
    printf(" | XsyevBatched: status %d dev %.2f GB", (int)s2, d / 1e9);
    d = hh = 0;
    cusolverStatus_t s3 = cusolverDnXgesvdp_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, 0, n, n, CUDA_C_64F, nullptr, n, CUDA_R_64F, nullptr,
                                                       CUDA_C_64F, nullptr, n, CUDA_C_64F, nullptr, n, CUDA_C_64F, &d, &hh);
    printf(" | Xgesvdp: status %d dev %.2f GB host %.2f GB", (int)s3, d / 1e9, hh / 1e9);
    d = hh = 0;
    cusolverStatus_t s4 = cusolverDnXgesvd_bufferSize(h, p, 'A', 'A', n, n, CUDA_C_64F, nullptr, n, CUDA_R_64F, nullptr, CUDA_C_64F, nullptr,
                                                      n, CUDA_C_64F, nullptr, n, CUDA_C_64F, &d, &hh);
    printf(" | Xgesvd: status %d dev %.2f GB", (int)s4, d / 1e9);
    int lw = -1;
    cusolverStatus_t s5 = cusolverDnZhetrd_bufferSize(h, CUBLAS_FILL_MODE_LOWER, n, nullptr, n, nullptr, nullptr, nullptr, &lw);
    printf(" | Zhetrd: status %d lwork %d", (int)s5, lw);
    lw = -1;
    cusolverStatus_t s6 = cusolverDnZunmtr_bufferSize(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, nullptr, n, nullptr,
                                                      nullptr, n, &lw);
    printf(" | Zunmtr: status %d lwork %d\n", (int)s6, lw);
  }
}

// SVD of a Hermitian positive semi-definite matrix = its eigen-decomposition (polar-decomposition based gesvdp)
static int run_svdp(cusolverDnHandle_t h, cusolverDnParams_t p, int n) {
  double2 *A, *U, *V; double* S; int* info;
  if (cudaMalloc(&A, sizeof(double2) * (size_t)n * n) != cudaSuccess) { printf("alloc A failed\n"); return 1; }
  cudaMalloc(&U, sizeof(double2) * (size_t)n * n); cudaMalloc(&V, sizeof(double2) * (size_t)n * n);
  cudaMalloc(&S, sizeof(double) * n); cudaMalloc(&info, sizeof(int));
  fill_hermitian<<<dim3((n + 255) / 256, n), 256>>>(A, n, 17u);
  cudaDeviceSynchronize();
  size_t d = 0, hh = 0;
  cusolverStatus_t s = cusolverDnXgesvdp_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, 0, n, n, CUDA_C_64F, A, n, CUDA_R_64F, S, CUDA_C_64F, U, n,
                                                    CUDA_C_64F, V, n, CUDA_C_64F, &d, &hh);
  if (s != CUSOLVER_STATUS_SUCCESS) { printf("n=%d gesvdp bufferSize status %d\n", n, (int)s); return 1; }
  void* work; if (cudaMalloc(&work, d) != cudaSuccess) { printf("alloc work %.1f GB failed\n", d / 1e9); return 1; }
  void* hbuf = hh ? malloc(hh) : nullptr;
  double herr = 0;
  double t0 = now();
  s = cusolverDnXgesvdp(h, p, CUSOLVER_EIG_MODE_VECTOR, 0, n, n, CUDA_C_64F, A, n, CUDA_R_64F, S, CUDA_C_64F, U, n, CUDA_C_64F, V, n, CUDA_C_64F,
                        work, d, hbuf, hh, info, &herr);
  cudaError_t ce = cudaDeviceSynchronize();
  double t1 = now();
  int hi = -7; cudaMemcpy(&hi, info, sizeof(int), cudaMemcpyDeviceToHost);
  double w0 = 0, w1 = 0; cudaMemcpy(&w0, S, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&w1, S + n - 1, 8, cudaMemcpyDeviceToHost);
  printf("n=%d Xgesvdp C64: status %d cuda %d info %d  %.2f s  workspace %.2f GB  s[0]=%.4f s[-1]=%.4f h_err %.2e\n", n, (int)s, (int)ce, hi,
         t1 - t0, d / 1e9, w0, w1, herr);
  fflush(stdout);
  cudaFree(A); cudaFree(U); cudaFree(V); cudaFree(S); cudaFree(info); cudaFree(work); free(hbuf);
  return 0;
}

static int run_dn(cusolverDnHandle_t h, cusolverDnParams_t p, int n, bool legacy) {
  double2* A; double* W; int* info;
  if (cudaMalloc(&A, sizeof(double2) * (size_t)n * n) != cudaSuccess) { printf("alloc A failed\n"); return 1; }
  cudaMalloc(&W, sizeof(double) * n); cudaMalloc(&info, sizeof(int));
  fill_hermitian<<<dim3((n + 255) / 256, n), 256>>>(A, n, 17u);
  cudaDeviceSynchronize();
  void* work = nullptr; size_t d = 0, hh = 0; int lw = 0;
  cusolverStatus_t s;
  if (!legacy) {
    s = cusolverDnXsyevd_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_C_64F, A, n, CUDA_R_64F, W,
                                    CUDA_C_64F, &d, &hh);
    if (s != CUSOLVER_STATUS_SUCCESS) { printf("n=%d Xsyevd bufferSize status %d\n", n, (int)s); return 1; }
  } else {
    s = cusolverDnZheevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, (cuDoubleComplex*)A, n, W, &lw);
    if (s != CUSOLVER_STATUS_SUCCESS || lw <= 0) { printf("n=%d Zheevd bufferSize status %d lwork %d\n", n, (int)s, lw); return 1; }
    d = sizeof(double2) * (size_t)lw;
  }
  if (cudaMalloc(&work, d) != cudaSuccess) { printf("alloc work %.1f GB failed\n", d / 1e9); return 1; }
  void* hbuf = hh ? malloc(hh) : nullptr;
  double t0 = now();
  if (!legacy)
    s = cusolverDnXsyevd(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_C_64F, A, n, CUDA_R_64F, W, CUDA_C_64F,
                         work, d, hbuf, hh, info);
  else
    s = cusolverDnZheevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, (cuDoubleComplex*)A, n, W,
                         (cuDoubleComplex*)work, lw, info);
  cudaError_t ce = cudaDeviceSynchronize();
  double t1 = now();
  int hi = -7; cudaMemcpy(&hi, info, sizeof(int), cudaMemcpyDeviceToHost);
  double w0 = 0, w1 = 0; cudaMemcpy(&w0, W, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&w1, W + n - 1, 8, cudaMemcpyDeviceToHost);
  printf("n=%d %s: status %d cuda %d info %d  %.2f s  workspace %.2f GB  ev[0]=%.4f ev[-1]=%.4f\n", n, legacy ? "Zheevd" : "Xsyevd C64",
         (int)s, (int)ce, hi, t1 - t0, d / 1e9, w0, w1);
  fflush(stdout);
  cudaFree(A); cudaFree(W); cudaFree(info); cudaFree(work); free(hbuf);
  return 0;
}

// cusolverMg on G devices, 1-D column block-cyclic layout with block size TA
static int run_mg(int n, int G, int TA) {
  cusolverMgHandle_t mh; cusolverMgCreate(&mh);
  std::vector<int> devs(G); for (int g = 0; g < G; ++g) devs[g] = g;
  cusolverStatus_t s = cusolverMgDeviceSelect(mh, G, devs.data());
  if (s != CUSOLVER_STATUS_SUCCESS) { printf("Mg DeviceSelect status %d\n", (int)s); return 1; }
  for (int a = 0; a < G; ++a) { cudaSetDevice(a); for (int b = 0; b < G; ++b) if (a != b) cudaDeviceEnablePeerAccess(b, 0); }
  cudaGetLastError();
  cudaLibMgGrid_t grid; cudaLibMgMatrixDesc_t desc;
  cusolverMgCreateDeviceGrid(&grid, 1, G, devs.data(), CUDALIBMG_GRID_MAPPING_COL_MAJOR);
  cusolverMgCreateMatrixDesc(&desc, n, n, n, TA, CUDA_C_64F, grid);
  const long long nblk = (n + TA - 1) / TA;
  std::vector<void*> dA(G), dWork(G);
  std::vector<long long> cols(G, 0);
  for (long long b = 0; b < nblk; ++b) cols[b % G] += TA;      // padded to whole blocks
  // fill on device 0 then scatter column blocks
  cudaSetDevice(0);
  double2* full; if (cudaMalloc(&full, sizeof(double2) * (size_t)n * n) != cudaSuccess) { printf("alloc full failed\n"); return 1; }
  fill_hermitian<<<dim3((n + 255) / 256, n), 256>>>(full, n, 17u);
  cudaDeviceSynchronize();
  for (int g = 0; g < G; ++g) { cudaSetDevice(g); cudaMalloc(&dA[g], sizeof(double2) * (size_t)n * cols[g]); cudaMemset(dA[g], 0, sizeof(double2) * (size_t)n * cols[g]); }
  for (long long b = 0; b < nblk; ++b) {
    int g = (int)(b % G); long long lb = b / G;
    long long w = (b + 1) * TA <= n ? TA : n - b * TA;
    cudaMemcpy((double2*)dA[g] + (size_t)lb * TA * n, full + (size_t)b * TA * n, sizeof(double2) * (size_t)n * w, cudaMemcpyDefault);
  }
  cudaSetDevice(0); cudaFree(full);
  std::vector<double> W(n);
  int64_t lwork = 0;
  s = cusolverMgSyevd_bufferSize(mh, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA.data(), 1, 1, desc, W.data(), CUDA_R_64F,
                                 CUDA_C_64F, &lwork);
  printf("Mg n=%d G=%d TA=%d: bufferSize status %d lwork %lld elements (%.2f GB per device)\n", n, G, TA, (int)s, (long long)lwork,
         lwork * 16.0 / 1e9);
  if (s != CUSOLVER_STATUS_SUCCESS) return 1;
  for (int g = 0; g < G; ++g) { cudaSetDevice(g); if (cudaMalloc(&dWork[g], sizeof(double2) * (size_t)lwork) != cudaSuccess) { printf("alloc work failed\n"); return 1; } }
  int info = -7;
  double t0 = now();
  s = cusolverMgSyevd(mh, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA.data(), 1, 1, desc, W.data(), CUDA_R_64F, CUDA_C_64F,
                      dWork.data(), lwork, &info);
  for (int g = 0; g < G; ++g) { cudaSetDevice(g); cudaDeviceSynchronize(); }
  double t1 = now();
  printf("Mg n=%d G=%d: status %d info %d  %.2f s  ev[0]=%.4f ev[-1]=%.4f\n", n, G, (int)s, info, t1 - t0, W[0], W[n - 1]);
  fflush(stdout);
  for (int g = 0; g < G; ++g) { cudaSetDevice(g); cudaFree(dA[g]); cudaFree(dWork[g]); }
  cusolverMgDestroyMatrixDesc(desc); cusolverMgDestroyGrid(grid); cusolverMgDestroy(mh);
  return 0;
}

int main(int argc, char** argv) {
  cusolverDnHandle_t h; cusolverDnParams_t p;
  cusolverDnCreate(&h); cusolverDnCreateParams(&p);
  int ndev = 0; cudaGetDeviceCount(&ndev);
  if (argc >= 3 && !strcmp(argv[1], "--run")) return run_dn(h, p, atoi(argv[2]), false);
  if (argc >= 3 && !strcmp(argv[1], "--run-legacy")) return run_dn(h, p, atoi(argv[2]), true);
  if (argc >= 3 && !strcmp(argv[1], "--run-mg")) return run_mg(atoi(argv[2]), argc >= 4 ? atoi(argv[3]) : ndev, argc >= 5 ? atoi(argv[4]) : 256);
  if (argc >= 2 && !strcmp(argv[1], "--unm")) {
    const int n = argc >= 3 ? atoi(argv[2]) : 40000;
    const int cols[] = {128, 256, 1024, 4096, 16384};
    const int ks[] = {128, 256, 1024, 4096, 16384, n - 1};
    for (int nc : cols) {
      int lw = -1;
      cusolverStatus_t s1 = cusolverDnZunmtr_bufferSize(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, nc, nullptr, n,
                                                        nullptr, nullptr, n, &lw);
      printf("Zunmtr n=%d ncols=%d: status %d lwork %d\n", n, nc, (int)s1, lw);
      for (int k : ks) {
        lw = -1;
        cusolverStatus_t s2 = cusolverDnZunmqr_bufferSize(h, CUBLAS_SIDE_LEFT, CUBLAS_OP_N, n - 1, nc, k, nullptr, n, nullptr, nullptr, n, &lw);
        printf("   Zunmqr m=%d ncols=%d k=%d: status %d lwork %d\n", n - 1, nc, k, (int)s2, lw);
      }
      lw = -1;
      cusolverStatus_t s3 = cusolverDnDormtr_bufferSize(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, nc, nullptr, n,
                                                        nullptr, nullptr, n, &lw);
      printf("   Dormtr n=%d ncols=%d: status %d lwork %d\n", n, nc, (int)s3, lw);
    }
    return 0;
  }
  if (argc >= 3 && !strcmp(argv[1], "--run-svdp")) return run_svdp(h, p, atoi(argv[2]));
  if (argc >= 2 && !strcmp(argv[1], "--other")) {
    probe_other(h, p);
    run_svdp(h, p, 8192);
    run_svdp(h, p, 16384);
    return 0;
  }
  probe_sizes(h, p);
  run_dn(h, p, 4096, false);
  run_dn(h, p, 8192, false);
  run_dn(h, p, 16384, false);
  run_dn(h, p, 8192, true);
  run_mg(8192, 1, 256);
  if (ndev > 1) run_mg(8192, ndev, 256);
  return 0;
}
