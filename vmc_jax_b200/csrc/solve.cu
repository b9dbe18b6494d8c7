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


// packed <-> A: the suffix [r M, Pc) of every matrix row i = (r, j), rows back to back (direction 0: pack, 1: unpack)
__global__ void hermitian_pack_kernel(cplx* __restrict__ A, int Pc, int M, cplx* __restrict__ packed, int unpack, int row0) {
  const int i = row0 + blockIdx.y;
  const int r = i / M, j = i - r * M;
  const long long c0 = (long long)r * M;
  const long long off = (long long)M * ((long long)r * Pc - (long long)M * r * (r - 1) / 2) + (long long)j * (Pc - c0);
  for (long long k = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < Pc; k += (long long)gridDim.x * blockDim.x) {
    if (unpack) A[(size_t)i * Pc + k] = packed[off + k - c0];
    else packed[off + k - c0] = A[(size_t)i * Pc + k];
  }
}

// A[i][k] = conj(A[k][i]) for every element below the block diagonal (block = M x M): restores the full Hermitian
// matrix after only the block rows' suffixes [r0 M, Pc) were all-reduced.  32 x 32 tiles through shared memory.
__global__ void hermitian_mirror_kernel(cplx* __restrict__ A, int Pc, int M) {
  __shared__ cplx tile[32][33];
  const int bi = blockIdx.y, bk = blockIdx.x;          // destination tile (rows bi, cols bk), bk <= bi
  if (bk > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;        // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int srow = bk * 32 + r, scol = bi * 32 + tx; // source A[k][i]
    if (srow < Pc && scol < Pc) tile[r][tx] = A[(size_t)srow * Pc + scol];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, k = bk * 32 + tx;
    if (i < Pc && k < Pc && k < (i / M) * M) {         // strictly below the block diagonal
      const cplx v = tile[tx][r];
      A[(size_t)i * Pc + k] = cmk(v.x, -v.y);
    }
  }
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

extern "C" int jvmc_hermitian_mirror_blocks(double* A, int Pc, int M, void* stream) {
  if (!A || Pc <= 0 || M <= 0) return JVMC_ERR_ARG;
  const unsigned nt = (unsigned)((Pc + 31) / 32);
  hermitian_mirror_kernel<<<dim3(nt, nt), dim3(32, 8), 0, (cudaStream_t)stream>>>((cplx*)A, Pc, M);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

/* number of complex elements of the packed upper block triangle */
extern "C" long long jvmc_hermitian_packed_elems(int Pc, int M) {
  if (Pc <= 0 || M <= 0 || Pc % M != 0) return -1;
  const long long R = Pc / M;
  return (long long)M * (R * Pc - (long long)M * R * (R - 1) / 2);
}

extern "C" int jvmc_hermitian_pack_blocks(double* A, int Pc, int M, double* packed, int unpack, void* stream) {
  if (!A || !packed || Pc <= 0 || M <= 0 || Pc % M != 0 || Pc > 65535) return JVMC_ERR_ARG;
  unsigned gx = (unsigned)((Pc + 1023) / 1024);
  hermitian_pack_kernel<<<dim3(gx, (unsigned)Pc), 256, 0, (cudaStream_t)stream>>>((cplx*)A, Pc, M, (cplx*)packed, unpack, 0);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

/* the same for matrix rows [row0, row0 + nrows) only (packed: base of the WHOLE packed buffer; block rows are contiguous in
 * it, elements [jvmc_hermitian_packed_offset(row0 / M), ...)): lets a finished range of block rows be reduced over ranks
 * while the Gram kernel still works on the others */
extern "C" int jvmc_hermitian_pack_rows(double* A, int Pc, int M, int row0, int nrows, double* packed, int unpack, void* stream) {
  if (!A || !packed || Pc <= 0 || M <= 0 || Pc % M != 0 || Pc > 65535 || row0 < 0 || nrows <= 0 || row0 + nrows > Pc)
    return JVMC_ERR_ARG;
  unsigned gx = (unsigned)((Pc + 1023) / 1024);
  hermitian_pack_kernel<<<dim3(gx, (unsigned)nrows), 256, 0, (cudaStream_t)stream>>>((cplx*)A, Pc, M, (cplx*)packed, unpack, row0);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
