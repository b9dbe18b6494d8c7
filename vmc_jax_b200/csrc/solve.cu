// TDVP / SR linear-system pieces.
//
//  jvmc_expand_S       A (P_c x P_c complex Hermitian, Khatri-Rao order) -> S = q(S0) (+ diagonal
//                      shift) in the reference's flat parameter layout   <- jVMC/util/tdvp.py:140-146,
//                      gradient layout of jVMC/vqs.py:66-69
//  jvmc_eigh(+_workspace)  cuSOLVER syevd/heevd (library call)            <- jnp.linalg.eigh, tdvp.py:153-171
//  jvmc_tdvp_regularize    eigenvalue cut, SNR soft cut-off and the cutoff-halving loop, entirely on the
//                      device (the reference syncs with the host once per iteration) <- tdvp.py:193-209
#include <cusolverDn.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

// flat reference index -> (complex parameter index, multiplied-by-i flag), holomorphic layout
__device__ __forceinline__ void flat_to_cpx(long long f, long long Mb, long long NM, long long& c, int& ph) {
  if (f < 2 * Mb) { ph = f >= Mb; c = ph ? f - Mb : f; }
  else {
    long long g = f - 2 * Mb;
    ph = g >= NM;
    c = Mb + (ph ? g - NM : g);
  }
}

// out is column-major S (out[fb*P + fa] = S[fa][fb]); mode 0: S = Re(S0) as double, mode 1: S = i Im(S0)
// as complex128.  diagonal shift: S[f][f] *= (1 + shift) when shift > 1e-10 (tdvp.py:145-146).
__global__ void expand_S_kernel(const cplx* __restrict__ A, long long Mb, long long NM, int mode, double shift,
                                double* __restrict__ out) {
  const long long P = 2 * (Mb + NM), Pc = Mb + NM;
  const long long fa = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // fastest: coalesced col-major write
  const long long fb = blockIdx.y;
  if (fa >= P) return;
  long long ca, cb;
  int pa, pb;
  flat_to_cpx(fa, Mb, NM, ca, pa);
  flat_to_cpx(fb, Mb, NM, cb, pb);
  cplx a = A[ca * Pc + cb];
  // conj(ph_a) ph_b : (0,0)->1, (0,1)->i, (1,0)->-i, (1,1)->1
  cplx v = a;
  if (pa == 0 && pb == 1) v = cmk(-a.y, a.x);
  else if (pa == 1 && pb == 0) v = cmk(a.y, -a.x);
  if (mode == 0) {
    double x = v.x;
    if (fa == fb && shift > 1e-10) x = x + shift * x;
    out[fb * P + fa] = x;
  } else {
    double y = v.y;
    if (fa == fb && shift > 1e-10) y = y + shift * y;
    reinterpret_cast<cplx*>(out)[fb * P + fa] = cmk(0.0, y);
  }
}

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

__device__ __forceinline__ double pow6(double x) { double x2 = x * x; return x2 * x2 * x2; }

// scal: [0]=residual, [1]=final cutoff (max(cutoff, pinvCutoff))
__global__ void __launch_bounds__(1024)
tdvp_regularize_kernel(int n, const double* __restrict__ ev, const cplx* __restrict__ VtF,
                       const double* __restrict__ snr, const cplx* __restrict__ F, double pinvTol, double pinvCutoff,
                       double snrTol, double* __restrict__ pinvEv, double* __restrict__ scal) {
  __shared__ double red[32];
  const double evmax = ev[n - 1];
  double fpart = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) fpart += cabs2(F[k]);
  const double Fnorm = sqrt(block_sum(fpart, red));
  double residual = 1.0, cutoff = 1e-2;
  bool wrote = false;
  while (residual > pinvTol && cutoff > pinvCutoff) {
    cutoff *= 0.8;
    const double c = fmax(cutoff, pinvCutoff);
    double part = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const double e = ev[k];
      const double rel = fabs(e / evmax);
      const double inv = (rel > 1e-14) ? 1.0 / e : 0.0;
      double reg = 1.0 / (1.0 + pow6(c / rel));
      if (snr) reg *= 1.0 / (1.0 + pow6(snrTol / snr[k]));
      const double p = inv * reg;
      pinvEv[k] = p;
      const double d = p * e - 1.0;
      part += d * d * cabs2(VtF[k]);
    }
    residual = sqrt(block_sum(part, red)) / Fnorm;
    wrote = true;
  }
  if (!wrote)
    for (int k = threadIdx.x; k < n; k += blockDim.x) pinvEv[k] = 0.0;
  if (threadIdx.x == 0) { scal[0] = residual; scal[1] = fmax(cutoff, pinvCutoff); }
}

cusolverDnHandle_t g_handle = nullptr;
cusolverDnParams_t g_params = nullptr;
int ensure_handle() {
  if (!g_handle) {
    if (cusolverDnCreate(&g_handle) != CUSOLVER_STATUS_SUCCESS) return JVMC_ERR_SOLVER;
    if (cusolverDnCreateParams(&g_params) != CUSOLVER_STATUS_SUCCESS) return JVMC_ERR_SOLVER;
  }
  return JVMC_OK;
}

}  // namespace

extern "C" int jvmc_expand_S(const double* A, int M, int N, int hasBias, int mode, double shift, double* out,
                             void* stream) {
  if (!A || !out || M <= 0 || N <= 0 || mode < 0 || mode > 1) return JVMC_ERR_ARG;
  long long Mb = hasBias ? M : 0, NM = (long long)N * M;
  long long P = 2 * (Mb + NM);
  if (P > 65535LL * 32) return JVMC_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((P + 255) / 256), (unsigned)P);
  if (P > 65535) return JVMC_ERR_UNSUPPORTED;
  expand_S_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const cplx*)A, Mb, NM, mode, shift, out);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_eigh_workspace(int n, int isComplex, long long* deviceBytes, long long* hostBytes) {
  if (n <= 0 || !deviceBytes || !hostBytes) return JVMC_ERR_ARG;
  int rc = ensure_handle();
  if (rc) return rc;
  size_t d = 0, h = 0;
  cudaDataType ta = isComplex ? CUDA_C_64F : CUDA_R_64F;
  if (cusolverDnXsyevd_bufferSize(g_handle, g_params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, ta, nullptr,
                                  n, CUDA_R_64F, nullptr, ta, &d, &h) != CUSOLVER_STATUS_SUCCESS)
    return JVMC_ERR_SOLVER;
  *deviceBytes = (long long)d;
  *hostBytes = (long long)h;
  return JVMC_OK;
}

// A: column-major n x n (lower triangle referenced), overwritten by the eigenvectors (column k =
// eigenvector k); w: n ascending eigenvalues; info: device int.
extern "C" int jvmc_eigh(int n, int isComplex, double* A, double* w, void* work, long long deviceBytes, int* info,
                         void* stream) {
  if (n <= 0 || !A || !w || !info) return JVMC_ERR_ARG;
  int rc = ensure_handle();
  if (rc) return rc;
  if (cusolverDnSetStream(g_handle, (cudaStream_t)stream) != CUSOLVER_STATUS_SUCCESS) return JVMC_ERR_SOLVER;
  size_t d = 0, h = 0;
  cudaDataType ta = isComplex ? CUDA_C_64F : CUDA_R_64F;
  if (cusolverDnXsyevd_bufferSize(g_handle, g_params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, ta, A, n,
                                  CUDA_R_64F, w, ta, &d, &h) != CUSOLVER_STATUS_SUCCESS)
    return JVMC_ERR_SOLVER;
  if ((long long)d > deviceBytes) return JVMC_ERR_ARG;
  void* hbuf = h ? malloc(h) : nullptr;
  cusolverStatus_t stt = cusolverDnXsyevd(g_handle, g_params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, ta,
                                          A, n, CUDA_R_64F, w, ta, work, d, hbuf, h, info);
  if (hbuf) { cudaStreamSynchronize((cudaStream_t)stream); free(hbuf); }
  return stt == CUSOLVER_STATUS_SUCCESS ? JVMC_OK : JVMC_ERR_SOLVER;
}

extern "C" int jvmc_tdvp_regularize(int n, const double* ev, const double* VtF, const double* snr, const double* F,
                                    double pinvTol, double pinvCutoff, double snrTol, double* pinvEv, double* scal,
                                    void* stream) {
  if (n <= 0 || !ev || !VtF || !F || !pinvEv || !scal) return JVMC_ERR_ARG;
  tdvp_regularize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n, ev, (const cplx*)VtF, snr, (const cplx*)F, pinvTol,
                                                               pinvCutoff, snrTol, pinvEv, scal);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
