// jvmc_tdvp_solve / jvmc_minsr_solve -- the regularised solves of one VMC step as single C-ABI entry points, so that a
// binding that cannot run Python between kernels (XLA-FFI custom call) gets the whole solve:
//
//  jvmc_tdvp_solve   TDVP.solve, reference jVMC/util/tdvp.py:153-213:
//        ev, V = eigh(S)                       (:153-171, cuSOLVER through jvmc_eigh)
//        VtF = V^dagger F                      (:171)
//        rho_n = V^dagger q(-x conj(dO_n) dE_n),  rhoVar = Var_w rho,  snr = sqrt(|N |VtF|^2 / rhoVar|)   (:173-181,
//                                               jVMC/stats.py:255-265 covar_data, :282-292 transform, :248-252 var)
//        cutoff loop -> pinvEv, residual       (:193-209, jvmc_tdvp_regularize)
//        update = Re(V (pinvEv . VtF))         (:211)
//     The per-sample projections are one dense contraction [B, n] x [n, n] (cuBLAS gemm on sample chunks, library call);
//     with a communicator the first/second moments of rho are summed over ranks in-stream (jvmc_comm_allreduce_sum_f64).
//  jvmc_minsr_solve  MinSR.solve, reference jVMC/util/minsr.py:59-65 (and :67-78 with the stacked real matrix):
//        x = pinv(T, rtol, hermitian) e  through eigh(T): V (w . V^dagger e), w_k = 1/ev_k for |ev_k| > rtol max|ev|
#include <cublas_v2.h>

#include "common.cuh"

extern "C" int jvmc_eigh_workspace(int n, int isComplex, long long* deviceBytes, long long* hostBytes);
extern "C" int jvmc_eigh(int n, int isComplex, double* A, double* w, void* work, long long deviceBytes, int* info, void* stream);
extern "C" int jvmc_tdvp_regularize(int n, const double* ev, const double* VtF, const double* snr, const double* F,
                                    double pinvTol, double pinvCutoff, double snrTol, double* pinvEv, double* scal, void* stream);
extern "C" int jvmc_comm_allreduce_sum_f64(void* comm, double* buf, long long count, void* stream);

namespace {

cublasHandle_t g_blas = nullptr;
int ensure_blas(cudaStream_t st) {
  if (!g_blas && cublasCreate(&g_blas) != CUBLAS_STATUS_SUCCESS) return JVMC_ERR_SOLVER;
  return cublasSetStream(g_blas, st) == CUBLAS_STATUS_SUCCESS ? JVMC_OK : JVMC_ERR_SOLVER;
}

// V column-major n x n (column k = eigenvector k): out[k] = sum_j conj(V[j,k]) f[j]   -- one warp per k
template <bool CPLXV>
__global__ void vdagger_vec_kernel(int n, const double* __restrict__ V, const cplx* __restrict__ f, cplx* __restrict__ out) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= n) return;
  double ar = 0.0, ai = 0.0;
  for (int j = lane; j < n; j += 32) {
    const cplx x = f[j];
    if (CPLXV) {
      const cplx v = reinterpret_cast<const cplx*>(V)[(size_t)k * n + j];
      ar += v.x * x.x + v.y * x.y;       // conj(v) x
      ai += v.x * x.y - v.y * x.x;
    } else {
      const double v = V[(size_t)k * n + j];
      ar += v * x.x; ai += v * x.y;
    }
  }
  ar = warp_sum(ar); ai = warp_sum(ai);
  if (lane == 0) out[k] = cmk(ar, ai);
}

// out[j] = Re sum_k V[j,k] c[k]  (update, real part only) or the complex sum (MinSR) -- one thread per row j, coalesced over j
template <bool CPLXV, bool REALOUT>
__global__ void v_vec_kernel(int n, const double* __restrict__ V, const double* __restrict__ wgt, const cplx* __restrict__ c,
                             double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double ar = 0.0, ai = 0.0;
  for (int k = 0; k < n; ++k) {
    cplx ck = c[k];
    if (wgt) { const double w = wgt[k]; ck.x *= w; ck.y *= w; }
    if (CPLXV) {
      const cplx v = reinterpret_cast<const cplx*>(V)[(size_t)k * n + j];
      ar += v.x * ck.x - v.y * ck.y;
      ai += v.x * ck.y + v.y * ck.x;
    } else {
      const double v = V[(size_t)k * n + j];
      ar += v * ck.x; ai += v * ck.y;
    }
  }
  if (REALOUT) out[j] = ar;
  else reinterpret_cast<cplx*>(out)[j] = cmk(ar, ai);
}

// X[c, k] = q(-x conj(D[c,k]) e[c] / w[c]) for the samples of one chunk; q = Re (mode 0, written as double) or
// i Im (mode 1, written as complex (0, y))
__global__ void snr_x_kernel(long long rows, int n, const cplx* __restrict__ D, const cplx* __restrict__ e,
                             const double* __restrict__ w, double xre, double xim, int mode, double* __restrict__ X) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * n) return;
  const long long c = idx / n;
  const cplx d = D[idx];
  const cplx y = cscale(cmul(cconj(d), e[c]), 1.0 / w[c]);
  const cplx z = cmul(cmk(-xre, -xim), y);
  if (mode == 0) X[idx] = z.x;
  else reinterpret_cast<cplx*>(X)[idx] = cmk(0.0, z.y);
}

// mom[0..2n) += sum_c w_c rho[c,k] (complex), mom[2n..3n) += sum_c w_c |rho[c,k]|^2   (rho row-major [rows, n])
template <bool CPLX>
__global__ void snr_moments_kernel(long long rows, int n, const double* __restrict__ rho, const double* __restrict__ w,
                                   double* __restrict__ mom) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const long long r0 = (long long)blockIdx.y * 256, r1 = min(rows, r0 + 256);
  double s1r = 0.0, s1i = 0.0, s2 = 0.0;
  for (long long c = r0; c < r1; ++c) {
    const double wc = w[c];
    if (CPLX) {
      const cplx v = reinterpret_cast<const cplx*>(rho)[c * n + k];
      s1r += wc * v.x; s1i += wc * v.y; s2 += wc * cabs2(v);
    } else {
      const double v = rho[c * n + k];
      s1r += wc * v; s2 += wc * v * v;
    }
  }
  atomicAdd(mom + 2 * k, s1r);
  atomicAdd(mom + 2 * k + 1, s1i);
  atomicAdd(mom + 2 * (size_t)n + k, s2);
}

__global__ void snr_finish_kernel(int n, const double* __restrict__ mom, const cplx* __restrict__ VtF, double numSamples,
                                  double* __restrict__ rhoVar, double* __restrict__ snr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double var = mom[2 * (size_t)n + k] - (mom[2 * k] * mom[2 * k] + mom[2 * k + 1] * mom[2 * k + 1]);
  rhoVar[k] = var;
  snr[k] = sqrt(fabs(numSamples * cabs2(VtF[k]) / var));
}

__global__ void pinv_weights_kernel(int n, const double* __restrict__ ev, double rtol, double* __restrict__ wgt) {
  __shared__ double red[32];
  double m = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) m = fmax(m, fabs(ev[k]));
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
  const double cut = rtol * m;
  for (int k = threadIdx.x; k < n; k += blockDim.x) wgt[k] = (fabs(ev[k]) > cut) ? 1.0 / ev[k] : 0.0;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline long long snr_chunk(int n, long long B) {
  long long c = (1LL << 27) / (n > 0 ? n : 1);     // <= 2 GB of complex X per chunk
  if (c < 256) c = 256;
  return c < B ? c : B;
}

}  // namespace

extern "C" int jvmc_tdvp_solve_workspace(int n, int mode, long long B, long long* bytes) {
  if (n <= 0 || mode < 0 || mode > 1 || B < 0 || !bytes) return JVMC_ERR_ARG;
  long long d = 0, h = 0;
  int rc = jvmc_eigh_workspace(n, mode, &d, &h);
  if (rc) return rc;
  const size_t el = mode ? 16 : 8;
  const long long ch = B > 0 ? snr_chunk(n, B) : 0;
  size_t snr = 2 * align256((size_t)ch * n * el) + align256(sizeof(double) * 3 * (size_t)n);
  size_t tot = align256((size_t)d) > snr ? align256((size_t)d) : snr;       // the eigensolver's scratch is reused
  *bytes = (long long)(tot + align256(16 * (size_t)n));
  return JVMC_OK;
}

// S: column-major n x n (double for mode 0 = 'real', complex for mode 1 = 'imag'), overwritten by the eigenvectors.
// useSnr = 0: the SNR is computed (when D is given) but not applied in the cutoff loop (ExactSampler, tdvp.py:203).
// F: complex[n] = q(F0).  D: complex[B, n] centred data sqrt(w)(O - <O>) or NULL (no SNR, ExactSampler semantics,
// tdvp.py:203); e: complex[B] centred sqrt(w)(E - <E>); w: [B].  x = rhsPrefactor.  comm: jvmc_comm handle or NULL.
// Outputs (device): ev[n], VtF complex[n], rhoVar[n], snr[n] (untouched without D), pinvEv[n], update[n] (real),
// scal[2] = (residual, cutoff), info (cuSOLVER devInfo).
extern "C" int jvmc_tdvp_solve(int n, int mode, double* S, const double* F, long long B, const double* D, const double* e,
                               const double* w, double xre, double xim, double numSamplesGlobal, int useSnr, double snrTol,
                               double pinvTol, double pinvCutoff, void* comm, double* ev, double* VtF, double* rhoVar,
                               double* snr, double* pinvEv, double* update, double* scal, int* info, void* work,
                               long long workBytes, void* stream) {
  if (n <= 0 || mode < 0 || mode > 1 || !S || !F || !ev || !VtF || !pinvEv || !update || !scal || !info || !work)
    return JVMC_ERR_ARG;
  if (D && (!e || !w || !rhoVar || !snr || B <= 0)) return JVMC_ERR_ARG;
  long long need = 0;
  int rc = jvmc_tdvp_solve_workspace(n, mode, D ? B : 0, &need);
  if (rc) return rc;
  if (workBytes < need) return JVMC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  long long d = 0, h = 0;
  jvmc_eigh_workspace(n, mode, &d, &h);
  rc = jvmc_eigh(n, mode, S, ev, work, d, info, stream);
  if (rc) return rc;
  const int warpsPerBlock = 8;
  if (mode) vdagger_vec_kernel<true><<<(n + warpsPerBlock - 1) / warpsPerBlock, 32 * warpsPerBlock, 0, st>>>(n, S, (const cplx*)F, (cplx*)VtF);
  else vdagger_vec_kernel<false><<<(n + warpsPerBlock - 1) / warpsPerBlock, 32 * warpsPerBlock, 0, st>>>(n, S, (const cplx*)F, (cplx*)VtF);
  JVMC_CHECK_LAUNCH();
  if (D) {
    if ((rc = ensure_blas(st))) return rc;
    const size_t el = mode ? 16 : 8;
    const long long ch = snr_chunk(n, B);
    unsigned char* base = (unsigned char*)work;
    double* X = (double*)base;
    double* rho = (double*)(base + align256((size_t)ch * n * el));
    double* mom = (double*)(base + 2 * align256((size_t)ch * n * el));
    cudaMemsetAsync(mom, 0, sizeof(double) * 3 * (size_t)n, st);
    for (long long lo = 0; lo < B; lo += ch) {
      const long long rows = (lo + ch < B) ? ch : B - lo;
      const long long tot = rows * n;
      snr_x_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(rows, n, (const cplx*)D + lo * n, (const cplx*)e + lo, w + lo, xre, xim,
                                                                 mode, X);
      JVMC_CHECK_LAUNCH();
      // rho^T [n x rows] (column-major) = V^dagger [n x n] X^T [n x rows]; X row-major [rows, n] IS X^T column-major
      cublasStatus_t bs;
      if (mode) {
        const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0), zero = make_cuDoubleComplex(0.0, 0.0);
        bs = cublasZgemm(g_blas, CUBLAS_OP_C, CUBLAS_OP_N, n, (int)rows, n, &one, (const cuDoubleComplex*)S, n,
                         (const cuDoubleComplex*)X, n, &zero, (cuDoubleComplex*)rho, n);
      } else {
        const double one = 1.0, zero = 0.0;
        bs = cublasDgemm(g_blas, CUBLAS_OP_T, CUBLAS_OP_N, n, (int)rows, n, &one, S, n, X, n, &zero, rho, n);
      }
      if (bs != CUBLAS_STATUS_SUCCESS) return JVMC_ERR_SOLVER;
      dim3 g((n + 127) / 128, (unsigned)((rows + 255) / 256));
      if (mode) snr_moments_kernel<true><<<g, 128, 0, st>>>(rows, n, rho, w + lo, mom);
      else snr_moments_kernel<false><<<g, 128, 0, st>>>(rows, n, rho, w + lo, mom);
      JVMC_CHECK_LAUNCH();
    }
    if (comm && (rc = jvmc_comm_allreduce_sum_f64(comm, mom, 3LL * n, stream))) return rc;
    snr_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, mom, (const cplx*)VtF, numSamplesGlobal, rhoVar, snr);
    JVMC_CHECK_LAUNCH();
  }
  rc = jvmc_tdvp_regularize(n, ev, VtF, (D && useSnr) ? snr : nullptr, F, pinvTol, pinvCutoff, snrTol, pinvEv, scal, stream);
  if (rc) return rc;
  if (mode) v_vec_kernel<true, true><<<(n + 127) / 128, 128, 0, st>>>(n, S, pinvEv, (const cplx*)VtF, update);
  else v_vec_kernel<false, true><<<(n + 127) / 128, 128, 0, st>>>(n, S, pinvEv, (const cplx*)VtF, update);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_minsr_solve_workspace(int n, int isComplex, long long* bytes) {
  if (n <= 0 || !bytes) return JVMC_ERR_ARG;
  long long d = 0, h = 0;
  int rc = jvmc_eigh_workspace(n, isComplex, &d, &h);
  if (rc) return rc;
  *bytes = (long long)(align256((size_t)d) + align256(16 * (size_t)n) + align256(8 * (size_t)n));
  return JVMC_OK;
}

// T: column-major n x n Hermitian (complex) or symmetric (double), overwritten by its eigenvectors; e: complex[n]
// (imaginary parts zero in the real case); x: complex[n] <- pinv(T, rtol) e; ev[n]; info: cuSOLVER devInfo.
extern "C" int jvmc_minsr_solve(int n, int isComplex, double* T, const double* e, double rtol, double* x, double* ev, int* info,
                                void* work, long long workBytes, void* stream) {
  if (n <= 0 || !T || !e || !x || !ev || !info || !work) return JVMC_ERR_ARG;
  long long need = 0, d = 0, h = 0;
  int rc = jvmc_minsr_solve_workspace(n, isComplex, &need);
  if (rc) return rc;
  if (workBytes < need) return JVMC_ERR_ARG;
  jvmc_eigh_workspace(n, isComplex, &d, &h);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* base = (unsigned char*)work;
  cplx* proj = (cplx*)(base + align256((size_t)d));
  double* wgt = (double*)(base + align256((size_t)d) + align256(16 * (size_t)n));
  rc = jvmc_eigh(n, isComplex, T, ev, work, d, info, stream);
  if (rc) return rc;
  pinv_weights_kernel<<<1, 1024, 0, st>>>(n, ev, rtol, wgt);
  JVMC_CHECK_LAUNCH();
  if (isComplex) {
    vdagger_vec_kernel<true><<<(n + 7) / 8, 256, 0, st>>>(n, T, (const cplx*)e, proj);
    JVMC_CHECK_LAUNCH();
    v_vec_kernel<true, false><<<(n + 127) / 128, 128, 0, st>>>(n, T, wgt, proj, x);
  } else {
    vdagger_vec_kernel<false><<<(n + 7) / 8, 256, 0, st>>>(n, T, (const cplx*)e, proj);
    JVMC_CHECK_LAUNCH();
    v_vec_kernel<false, false><<<(n + 127) / 128, 128, 0, st>>>(n, T, wgt, proj, x);
  }
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
