// Building blocks of the dense Hermitian eigen-decomposition BEYOND cuSOLVER's size limit.
//
// Every dense eigensolver of cuSOLVER 11.7 (Xsyevd, Xsyevdx, XsyevBatched, Zheevd, Xgesvdp) returns
// CUSOLVER_STATUS_INVALID_VALUE for n > 32768 (probed on B200: profiles/r2_eigh_probe.txt), and the config-2 TDVP solve
// (reference jVMC/util/tdvp.py:153-171, jnp.linalg.eigh) needs n = P_c = 40 000.  The tridiagonalisation (hetrd) and
// the back-transformation (unmtr) do work at that size, so the solver is split by ONE level of Cuppen's divide and
// conquer on the tridiagonal matrix (kernels.py:eigh_large):
//   A = H T H^dagger (hetrd)  ->  T = diag(T1', T2') + rho v v^T  ->  T1' = Q1 D1 Q1^T, T2' = Q2 D2 Q2^T (Xsyevd, n/2 each)
//   ->  D + rho z z^T, z = (last row of Q1, +- first row of Q2): deflation (host scan, O(n)), secular equation
//   (jvmc_secular_roots), Gu-Eisenstat eigenvectors (jvmc_secular_vectors)  ->  Z = diag(Q1, Q2) U (cuBLAS)  ->  V = H Z (unmtr).
// The merge follows LAPACK's dlaed2 / dlaed3 / dlaed4 (published algorithm; bracketed bisection instead of the rational
// interpolation of dlaed4: robust, and O(100 k^2) divisions are milliseconds on a B200).
#include <cusolverDn.h>

#include "common.cuh"

namespace {

cusolverDnHandle_t g_h = nullptr;
int ensure(cudaStream_t st) {
  if (!g_h && cusolverDnCreate(&g_h) != CUSOLVER_STATUS_SUCCESS) return JVMC_ERR_SOLVER;
  return cusolverDnSetStream(g_h, st) == CUSOLVER_STATUS_SUCCESS ? JVMC_OK : JVMC_ERR_SOLVER;
}

__global__ void tridiag_dense_kernel(int m, const double* __restrict__ d, const double* __restrict__ e, double shiftFirst,
                                     double shiftLast, double* __restrict__ T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double di = d[i];
  if (i == 0) di += shiftFirst;
  if (i == m - 1) di += shiftLast;
  T[(size_t)i * m + i] = di;
  if (i + 1 < m) { T[(size_t)i * m + i + 1] = e[i]; T[(size_t)(i + 1) * m + i] = e[i]; }
}

// f(lam) = 1 + rho sum_j z2_j / (d_j - lam) evaluated in coordinates shifted to the pole d[orig]: lam = d[orig] + mu,
// d_j - lam = (d_j - d[orig]) - mu (no cancellation when mu is tiny).  One CTA per root, fixed bisection count.
__global__ void __launch_bounds__(256)
secular_roots_kernel(int k, const double* __restrict__ d, const double* __restrict__ z2, double rho, double z2sum,
                     int* __restrict__ origOut, double* __restrict__ muOut) {
  __shared__ double red[8];
  __shared__ double bc;
  const int i = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double di = d[i];
  const double upper = (i + 1 < k) ? d[i + 1] : d[k - 1] + rho * z2sum;
  const double gap = upper - di;
  auto fval = [&](int orig, double mu) {
    const double dorig = d[orig];
    double acc = 0.0;
    for (int j = threadIdx.x; j < k; j += blockDim.x) acc += z2[j] / ((d[j] - dorig) - mu);
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      bc = 1.0 + rho * t;
    }
    __syncthreads();
    return bc;
  };
  // origin: the pole closer to the root (sign of f at the midpoint; f increases on the interval)
  const double mid = 0.5 * gap;
  const double fm = fval(i, mid);
  int orig;
  double lo, hi;
  if (fm > 0.0) { orig = i; lo = 0.0; hi = mid; }
  else if (i + 1 < k) { orig = i + 1; lo = -mid; hi = 0.0; }
  else { orig = i; lo = mid; hi = gap; }
  for (int it = 0; it < 110; ++it) {
    const double mu = 0.5 * (lo + hi);
    if (fval(orig, mu) > 0.0) hi = mu; else lo = mu;
  }
  if (threadIdx.x == 0) { origOut[i] = orig; muOut[i] = 0.5 * (lo + hi); }
}

// Gu-Eisenstat: zhat_i^2 = prod_j (lam_j - d_i) / prod_{j != i} (d_j - d_i) / rho with lam_j - d_i = (d[orig_j] - d_i) + mu_j
__global__ void __launch_bounds__(256)
loewner_kernel(int k, const double* __restrict__ d, const double* __restrict__ z, double rho, const int* __restrict__ orig,
               const double* __restrict__ mu, double* __restrict__ zhat) {
  __shared__ double red[8];
  const int i = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double di = d[i];
  double p = 1.0;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const double num = (d[orig[j]] - di) + mu[j];
    const double den = (j == i) ? 1.0 : d[j] - di;
    p *= num / den;
  }
  for (int o = 16; o; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
  if (lane == 0) red[wid] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 1.0;
    for (int w = 0; w < 8; ++w) t *= red[w];
    const double v = sqrt(fabs(t / rho));
    zhat[i] = z[i] < 0.0 ? -v : v;
  }
}

// column j of the k x k eigenvector block: u_j[i] = zhat_i / (d_i - lam_j), normalised, scattered to
// U[rowIdx[i] + colIdx[j] * ldu]; one CTA per column
__global__ void __launch_bounds__(256)
secular_vectors_kernel(int k, const double* __restrict__ d, const double* __restrict__ zhat, const int* __restrict__ orig,
                       const double* __restrict__ mu, const int* __restrict__ rowIdx, const int* __restrict__ colIdx,
                       double* __restrict__ U, long long ldu) {
  __shared__ double red[8];
  __shared__ double bc;
  const int j = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double dorig = d[orig[j]], muj = mu[j];
  double nrm = 0.0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const double v = zhat[i] / -((dorig - d[i]) + muj);
    nrm += v * v;
  }
  nrm = warp_sum(nrm);
  if (lane == 0) red[wid] = nrm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    bc = rsqrt(t);
  }
  __syncthreads();
  const double inv = bc;
  double* col = U + (size_t)colIdx[j] * ldu;
  for (int i = threadIdx.x; i < k; i += blockDim.x) col[rowIdx[i]] = zhat[i] / -((dorig - d[i]) + muj) * inv;
}

// Givens rotations on ROWS (ri, rj) of the column-major matrix U, applied in order; one thread per column
__global__ void apply_rotations_kernel(int ncols, long long ldu, double* __restrict__ U, int nrot, const int* __restrict__ ri,
                                       const int* __restrict__ rj, const double* __restrict__ c, const double* __restrict__ s) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  double* u = U + (size_t)col * ldu;
  for (int r = 0; r < nrot; ++r) {
    const double a = u[ri[r]], b = u[rj[r]];
    u[ri[r]] = c[r] * a - s[r] * b;
    u[rj[r]] = s[r] * a + c[r] * b;
  }
}

__global__ void real_to_complex_kernel(long long count, const double* __restrict__ src, cplx* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = cmk(src[i], 0.0);
}

}  // namespace

// T (m x m, zero-initialised by the caller) <- tridiagonal matrix (d, e) with d[0] += shiftFirst, d[m-1] += shiftLast
extern "C" int jvmc_tridiag_dense(int m, const double* d, const double* e, double shiftFirst, double shiftLast, double* T,
                                  void* stream) {
  if (m <= 0 || !d || (!e && m > 1) || !T) return JVMC_ERR_ARG;
  tridiag_dense_kernel<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(m, d, e, shiftFirst, shiftLast, T);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_hetrd_workspace(int n, int isComplex, long long* bytes) {
  if (n <= 0 || !bytes) return JVMC_ERR_ARG;
  int rc = ensure(0);
  if (rc) return rc;
  int lw = 0;
  cusolverStatus_t s = isComplex
      ? cusolverDnZhetrd_bufferSize(g_h, CUBLAS_FILL_MODE_LOWER, n, nullptr, n, nullptr, nullptr, nullptr, &lw)
      : cusolverDnDsytrd_bufferSize(g_h, CUBLAS_FILL_MODE_LOWER, n, nullptr, n, nullptr, nullptr, nullptr, &lw);
  if (s != CUSOLVER_STATUS_SUCCESS || lw <= 0) return JVMC_ERR_SOLVER;
  *bytes = (long long)lw * (isComplex ? 16 : 8);
  return JVMC_OK;
}

// A: column-major n x n, lower triangle referenced; overwritten by the Householder vectors.  d[n], e[n-1] real;
// tau[n-1] (complex when isComplex).  info: device int.
extern "C" int jvmc_hetrd(int n, int isComplex, double* A, double* d, double* e, double* tau, void* work, long long bytes,
                          int* info, void* stream) {
  if (n <= 0 || !A || !d || !e || !tau || !work || !info) return JVMC_ERR_ARG;
  int rc = ensure((cudaStream_t)stream);
  if (rc) return rc;
  cusolverStatus_t s;
  if (isComplex)
    s = cusolverDnZhetrd(g_h, CUBLAS_FILL_MODE_LOWER, n, (cuDoubleComplex*)A, n, d, e, (cuDoubleComplex*)tau,
                         (cuDoubleComplex*)work, (int)(bytes / 16), info);
  else
    s = cusolverDnDsytrd(g_h, CUBLAS_FILL_MODE_LOWER, n, A, n, d, e, tau, (double*)work, (int)(bytes / 8), info);
  return s == CUSOLVER_STATUS_SUCCESS ? JVMC_OK : JVMC_ERR_SOLVER;
}

extern "C" int jvmc_unmtr_workspace(int n, int ncols, int isComplex, long long* bytes) {
  if (n <= 0 || ncols <= 0 || !bytes) return JVMC_ERR_ARG;
  int rc = ensure(0);
  if (rc) return rc;
  int lw = 0;
  cusolverStatus_t s = isComplex
      ? cusolverDnZunmtr_bufferSize(g_h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, ncols, nullptr, n, nullptr,
                                    nullptr, n, &lw)
      : cusolverDnDormtr_bufferSize(g_h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, ncols, nullptr, n, nullptr,
                                    nullptr, n, &lw);
  if (s != CUSOLVER_STATUS_SUCCESS || lw <= 0) return JVMC_ERR_SOLVER;
  *bytes = (long long)lw * (isComplex ? 16 : 8);
  return JVMC_OK;
}

// C (n x ncols, column-major, leading dimension n) <- H C with H the unitary factor of jvmc_hetrd (A, tau)
extern "C" int jvmc_unmtr(int n, int ncols, int isComplex, double* A, double* tau, double* C, void* work, long long bytes,
                          int* info, void* stream) {
  if (n <= 0 || ncols <= 0 || !A || !tau || !C || !work || !info) return JVMC_ERR_ARG;
  int rc = ensure((cudaStream_t)stream);
  if (rc) return rc;
  cusolverStatus_t s;
  if (isComplex)
    s = cusolverDnZunmtr(g_h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, ncols, (cuDoubleComplex*)A, n,
                         (cuDoubleComplex*)tau, (cuDoubleComplex*)C, n, (cuDoubleComplex*)work, (int)(bytes / 16), info);
  else
    s = cusolverDnDormtr(g_h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, ncols, A, n, tau, C, n, (double*)work,
                         (int)(bytes / 8), info);
  return s == CUSOLVER_STATUS_SUCCESS ? JVMC_OK : JVMC_ERR_SOLVER;
}

// The same back-transformation in chunks of reflectors (cusolverDnZunmtr / Dormtr refuse n > 32768, Zunmqr / Dormqr accept
// m = 40 000 with k <= 16384 reflectors per call, probed: profiles/r2_eigh_probe.txt): C (m x ncols, leading dimension ldc)
// <- H_0 .. H_{k-1} C with the k reflectors stored geqrf-style in the columns of A (m x k, leading dimension lda).
extern "C" int jvmc_unmqr_workspace(int m, int ncols, int k, int isComplex, int lda, int ldc, long long* bytes) {
  if (m <= 0 || ncols <= 0 || k <= 0 || k > m || !bytes) return JVMC_ERR_ARG;
  int rc = ensure(0);
  if (rc) return rc;
  int lw = 0;
  cusolverStatus_t s = isComplex
      ? cusolverDnZunmqr_bufferSize(g_h, CUBLAS_SIDE_LEFT, CUBLAS_OP_N, m, ncols, k, nullptr, lda, nullptr, nullptr, ldc, &lw)
      : cusolverDnDormqr_bufferSize(g_h, CUBLAS_SIDE_LEFT, CUBLAS_OP_N, m, ncols, k, nullptr, lda, nullptr, nullptr, ldc, &lw);
  if (s != CUSOLVER_STATUS_SUCCESS || lw <= 0) return JVMC_ERR_SOLVER;
  *bytes = (long long)lw * (isComplex ? 16 : 8);
  return JVMC_OK;
}

extern "C" int jvmc_unmqr(int m, int ncols, int k, int isComplex, double* A, int lda, double* tau, double* C, int ldc, void* work,
                          long long bytes, int* info, void* stream) {
  if (m <= 0 || ncols <= 0 || k <= 0 || k > m || !A || !tau || !C || !work || !info) return JVMC_ERR_ARG;
  int rc = ensure((cudaStream_t)stream);
  if (rc) return rc;
  cusolverStatus_t s;
  if (isComplex)
    s = cusolverDnZunmqr(g_h, CUBLAS_SIDE_LEFT, CUBLAS_OP_N, m, ncols, k, (cuDoubleComplex*)A, lda, (cuDoubleComplex*)tau,
                         (cuDoubleComplex*)C, ldc, (cuDoubleComplex*)work, (int)(bytes / 16), info);
  else
    s = cusolverDnDormqr(g_h, CUBLAS_SIDE_LEFT, CUBLAS_OP_N, m, ncols, k, A, lda, tau, C, ldc, (double*)work, (int)(bytes / 8), info);
  return s == CUSOLVER_STATUS_SUCCESS ? JVMC_OK : JVMC_ERR_SOLVER;
}

extern "C" int jvmc_real_to_complex(long long count, const double* src, double* dst, void* stream) {
  if (count < 0 || !src || !dst) return JVMC_ERR_ARG;
  if (count == 0) return JVMC_OK;
  real_to_complex_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(count, src, (cplx*)dst);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

// roots of 1 + rho sum_j z2_j / (d_j - lam) = 0 for ascending, distinct poles d[k] and z2 > 0, rho > 0:
// lam_i = d[orig[i]] + mu[i] in (d_i, d_{i+1}) (last one: (d_k-1, d_k-1 + rho sum z2)).
extern "C" int jvmc_secular_roots(int k, const double* d, const double* z2, double rho, double z2sum, int* orig, double* mu,
                                  void* stream) {
  if (k <= 0 || !d || !z2 || !orig || !mu || !(rho > 0.0)) return JVMC_ERR_ARG;
  secular_roots_kernel<<<k, 256, 0, (cudaStream_t)stream>>>(k, d, z2, rho, z2sum, orig, mu);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

// eigenvectors of diag(d) + rho z z^T for the roots of jvmc_secular_roots: column j goes to U[rowIdx[.], colIdx[j]]
// (U column-major, leading dimension ldu); zhat: k doubles of scratch
extern "C" int jvmc_secular_vectors(int k, const double* d, const double* z, double rho, const int* orig, const double* mu,
                                    const int* rowIdx, const int* colIdx, double* U, long long ldu, double* zhat,
                                    void* stream) {
  if (k <= 0 || !d || !z || !orig || !mu || !rowIdx || !colIdx || !U || !zhat || ldu < k) return JVMC_ERR_ARG;
  loewner_kernel<<<k, 256, 0, (cudaStream_t)stream>>>(k, d, z, rho, orig, mu, zhat);
  JVMC_CHECK_LAUNCH();
  secular_vectors_kernel<<<k, 256, 0, (cudaStream_t)stream>>>(k, d, zhat, orig, mu, rowIdx, colIdx, U, ldu);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

// Givens rotations (ri[r], rj[r], c[r], s[r]) applied in order to the rows of the column-major matrix U [*, ncols]:
// (u_i, u_j) <- (c u_i - s u_j, s u_i + c u_j)
extern "C" int jvmc_apply_row_rotations(int ncols, long long ldu, double* U, int nrot, const int* ri, const int* rj,
                                        const double* c, const double* s, void* stream) {
  if (ncols <= 0 || !U || nrot < 0) return JVMC_ERR_ARG;
  if (nrot == 0) return JVMC_OK;
  if (!ri || !rj || !c || !s) return JVMC_ERR_ARG;
  apply_rotations_kernel<<<(ncols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ncols, ldu, U, nrot, ri, rj, c, s);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
