// jvmc_rbm_logpsi / jvmc_rbm_tables -- (Cpx)RBM forward pass and flip-ratio tables.
//
// Replaces NQS.__call__/_eval for CpxRBM/RBM (reference jVMC/vqs.py:223-251,
// jVMC/nets/rbm.py:30-38,82-88, jVMC/nets/activation_functions.py:19-22) and the sampler's final
// re-evaluation (jVMC/sampler.py:293-296).
//
// theta[b,j] = sum_i (2 s[b,i]-1) W[i,j] + bias[j];  logpsi[b] = sum_j lncosh(theta[b,j]);
// tau[b,j] = tanh(theta[b,j]) is kept for the local-energy / gradient / Gram kernels.
#include "common.cuh"

namespace {

constexpr int LP_THREADS = 128;
constexpr int LP_TB = 8;  // samples per CTA: each W element loaded once per 8 samples

__global__ void __launch_bounds__(LP_THREADS)
rbm_logpsi_kernel(const int32_t* __restrict__ s, long long B, int N, int M,
                  const cplx* __restrict__ W, const cplx* __restrict__ bias,
                  cplx* __restrict__ logpsi, cplx* __restrict__ tau) {
  extern __shared__ unsigned char smem_raw[];
  double* sig = reinterpret_cast<double*>(smem_raw);          // [N][LP_TB] (+-1, 0 for padding rows)
  __shared__ double red[LP_TB][2][LP_THREADS / 32];
  const long long b0 = (long long)blockIdx.x * LP_TB;
  const int nb = (int)min((long long)LP_TB, B - b0);
  for (int t = threadIdx.x; t < LP_TB * N; t += LP_THREADS) {
    int bb = t / N, i = t - bb * N;
    sig[i * LP_TB + bb] = (bb < nb) ? (double)(2 * s[(b0 + bb) * N + i] - 1) : 0.0;
  }
  __syncthreads();
  cplx part[LP_TB];
#pragma unroll
  for (int bb = 0; bb < LP_TB; ++bb) part[bb] = cmk(0.0, 0.0);
  for (int j = threadIdx.x; j < M; j += LP_THREADS) {
    cplx acc[LP_TB];
    cplx bj = bias ? bias[j] : cmk(0.0, 0.0);
#pragma unroll
    for (int bb = 0; bb < LP_TB; ++bb) acc[bb] = bj;
    for (int i = 0; i < N; ++i) {
      cplx w = W[(size_t)i * M + j];
      const double2* sg2 = reinterpret_cast<const double2*>(sig + i * LP_TB);
#pragma unroll
      for (int bb = 0; bb < LP_TB; bb += 2) {
        double2 sg = sg2[bb >> 1];
        acc[bb].x = fma(sg.x, w.x, acc[bb].x);
        acc[bb].y = fma(sg.x, w.y, acc[bb].y);
        acc[bb + 1].x = fma(sg.y, w.x, acc[bb + 1].x);
        acc[bb + 1].y = fma(sg.y, w.y, acc[bb + 1].y);
      }
    }
#pragma unroll
    for (int bb = 0; bb < LP_TB; ++bb) {
      if (bb < nb) {
        cplx lc, th;
        lncosh_tanh(acc[bb], lc, th);
        part[bb] = cadd(part[bb], lc);
        if (tau) tau[(size_t)(b0 + bb) * M + j] = th;
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int bb = 0; bb < LP_TB; ++bb) {
    cplx v = warp_csum(part[bb]);
    if (lane == 0) { red[bb][0][wid] = v.x; red[bb][1][wid] = v.y; }
  }
  __syncthreads();
  if (threadIdx.x < nb) {
    double re = 0.0, im = 0.0;
    for (int w = 0; w < LP_THREADS / 32; ++w) { re += red[threadIdx.x][0][w]; im += red[threadIdx.x][1][w]; }
    logpsi[b0 + threadIdx.x] = cmk(re, im);
  }
}

// T[i,j] = tanh(2 W[i,j]);  lc[i] = sum_j lncosh(2 W[i,j]);  tb2[j] = tanh(2 b[j]);
// lcb = sum_j lncosh(2 b[j]).  One CTA per site (+1 for the bias row).
__global__ void __launch_bounds__(128)
rbm_tables_kernel(int N, int M, const cplx* __restrict__ W, const cplx* __restrict__ bias,
                  cplx* __restrict__ T, cplx* __restrict__ lc, cplx* __restrict__ tb2, cplx* __restrict__ lcb) {
  __shared__ double red[2][4];
  const int i = blockIdx.x;
  const bool biasRow = (i == N);
  if (biasRow && !bias) {
    for (int j = threadIdx.x; j < M; j += blockDim.x) tb2[j] = cmk(0.0, 0.0);
    if (threadIdx.x == 0) lcb[0] = cmk(0.0, 0.0);
    return;
  }
  cplx sum = cmk(0.0, 0.0);
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    cplx w = biasRow ? bias[j] : W[(size_t)i * M + j];
    cplx l, t;
    lncosh_tanh(cmk(2.0 * w.x, 2.0 * w.y), l, t);
    sum = cadd(sum, l);   // l = log cosh(2w) (branch irrelevant: only exp(sum) is used)
    if (biasRow) tb2[j] = t; else T[(size_t)i * M + j] = t;
  }
  cplx v = warp_csum(sum);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = v.x; red[1][threadIdx.x >> 5] = v.y; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double re = 0.0, im = 0.0;
    for (int w = 0; w < 4; ++w) { re += red[0][w]; im += red[1][w]; }
    cplx out = cmk(re, im);
    if (biasRow) lcb[0] = out; else lc[i] = out;
  }
}

}  // namespace

extern "C" int jvmc_rbm_logpsi(const int32_t* s, long long B, int N, int M, const double* W, const double* bias,
                               double* logpsi, double* tau, void* stream) {
  if (B == 0) return JVMC_OK;
  if (B < 0 || N <= 0 || M <= 0 || !s || !W || !logpsi) return JVMC_ERR_ARG;
  size_t smem = (size_t)LP_TB * N * sizeof(double);
  if (smem > 200 * 1024) return JVMC_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(rbm_logpsi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long grid = (B + LP_TB - 1) / LP_TB;
  rbm_logpsi_kernel<<<(unsigned)grid, LP_THREADS, smem, (cudaStream_t)stream>>>(
      s, B, N, M, (const cplx*)W, (const cplx*)bias, (cplx*)logpsi, (cplx*)tau);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

// Table buffer layout (complex128 elements): T[N*M] | lc[N] | tb2[M] | lcb[1]
extern "C" long long jvmc_rbm_tables_elems(int N, int M) { return (long long)N * M + N + M + 1; }

extern "C" int jvmc_rbm_tables(int N, int M, const double* W, const double* bias, double* tables, void* stream) {
  if (N <= 0 || M <= 0 || !W || !tables) return JVMC_ERR_ARG;
  cplx* T = (cplx*)tables;
  cplx* lc = T + (size_t)N * M;
  cplx* tb2 = lc + N;
  cplx* lcb = tb2 + M;
  rbm_tables_kernel<<<N + 1, 128, 0, (cudaStream_t)stream>>>(N, M, (const cplx*)W, (const cplx*)bias, T, lc, tb2, lcb);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
