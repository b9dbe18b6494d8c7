// Per-sample gradients and Khatri-Rao moment reductions for (Cpx)RBM.
//
// For an RBM, O_k(s_n) = d logpsi / d param is Khatri-Rao structured: with the "site" index
// r (r = 0 is the bias pseudo-site with sigma = +1 when the net has a bias, then the N lattice
// sites) and hidden index j,   O_n[r*M + j] = sigma_{n,r} * tau_{n,j},  tau = tanh(theta).
//
//  jvmc_rbm_grad     materialises O in the reference's flat layout (small P only)
//                      <- NQS.gradients / flat_gradient_holo / flat_gradient (jVMC/vqs.py:46-69,256-287)
//  jvmc_rbm_moments  out[r,j] = sum_n wgt_n sigma_{n,r} tau~_{n,j}  (tau~ = tau or conj tau)
//                      <- SampledObs mean of the gradients (jVMC/stats.py:50,204), the force
//                         covar(grads, Eloc) (stats.py:52-58,245; tdvp.py:139) and MinSR's
//                         -O^dagger x contraction (minsr.py:65) without ever forming O.
//  jvmc_pack_sigma   bit-packed, transposed spins for the Gram kernels.
#include "common.cuh"

namespace {

// layout 0: holomorphic  [b, i b, W, i W]   (P = 2*(Mb + N*M))
// layout 1: real params  [b, W]             (P = Mb + N*M)
__global__ void __launch_bounds__(256)
rbm_grad_kernel(const int32_t* __restrict__ s, const cplx* __restrict__ tau, long long B, int N, int M,
                int hasBias, int layout, cplx* __restrict__ out) {
  const long long b = blockIdx.x;
  const int Mb = hasBias ? M : 0;
  const int NM = N * M;
  const long long P = (layout == 0) ? 2LL * (Mb + NM) : (long long)(Mb + NM);
  const cplx* t = tau + b * M;
  const int32_t* c = s + b * N;
  cplx* o = out + b * P;
  for (long long k = threadIdx.x; k < P; k += blockDim.x) {
    long long q = k;
    bool timesI = false;
    cplx v;
    if (layout == 0) {
      if (q < 2LL * Mb) { timesI = q >= Mb; v = t[q % M]; }
      else {
        q -= 2LL * Mb;
        timesI = q >= NM;
        if (timesI) q -= NM;
        int i = (int)(q / M), j = (int)(q - (long long)i * M);
        double sg = (double)(2 * c[i] - 1);
        v = cmk(sg * t[j].x, sg * t[j].y);
      }
    } else {
      if (q < Mb) v = t[q];
      else {
        q -= Mb;
        int i = (int)(q / M), j = (int)(q - (long long)i * M);
        double sg = (double)(2 * c[i] - 1);
        v = cmk(sg * t[j].x, sg * t[j].y);
      }
    }
    o[k] = timesI ? cmk(-v.y, v.x) : v;
  }
}

constexpr int MO_RG = 8;      // site rows per thread
constexpr int MO_TS = 64;     // samples staged per smem tile
constexpr int MO_THREADS = 128;

// grid: (ceil(M/128), ceil(R/8), chunks);  partial[chunk][r][j]
__global__ void __launch_bounds__(MO_THREADS)
rbm_moments_kernel(const int32_t* __restrict__ s, const cplx* __restrict__ tau, const cplx* __restrict__ wgt,
                   long long B, int N, int M, int hasBias, int conjTau, long long chunkLen,
                   cplx* __restrict__ partial) {
  __shared__ double sg[MO_TS][MO_RG];
  __shared__ cplx sw[MO_TS];
  const int R = N + (hasBias ? 1 : 0);
  const int j = blockIdx.x * MO_THREADS + threadIdx.x;
  const int r0 = blockIdx.y * MO_RG;
  const long long n0 = (long long)blockIdx.z * chunkLen;
  const long long n1 = min(B, n0 + chunkLen);
  cplx acc[MO_RG];
#pragma unroll
  for (int q = 0; q < MO_RG; ++q) acc[q] = cmk(0.0, 0.0);
  for (long long nt = n0; nt < n1; nt += MO_TS) {
    const int cnt = (int)min((long long)MO_TS, n1 - nt);
    __syncthreads();
    for (int t = threadIdx.x; t < MO_TS * MO_RG; t += MO_THREADS) {
      int k = t / MO_RG, q = t - k * MO_RG;
      int r = r0 + q;
      double v = 0.0;
      if (k < cnt && r < R) {
        if (hasBias && r == 0) v = 1.0;
        else v = (double)(2 * s[(nt + k) * N + (r - (hasBias ? 1 : 0))] - 1);
      }
      sg[k][q] = v;
    }
    for (int k = threadIdx.x; k < MO_TS; k += MO_THREADS) sw[k] = (k < cnt) ? wgt[nt + k] : cmk(0.0, 0.0);
    __syncthreads();
    if (j < M) {
      for (int k = 0; k < cnt; ++k) {
        cplx t = tau[(nt + k) * M + j];
        if (conjTau) t.y = -t.y;
        cplx wt = cmul(sw[k], t);
#pragma unroll
        for (int q = 0; q < MO_RG; ++q) {
          double g = sg[k][q];
          acc[q].x = fma(g, wt.x, acc[q].x);
          acc[q].y = fma(g, wt.y, acc[q].y);
        }
      }
    }
  }
  if (j < M) {
#pragma unroll
    for (int q = 0; q < MO_RG; ++q) {
      int r = r0 + q;
      if (r < R) partial[((size_t)blockIdx.z * R + r) * M + j] = acc[q];
    }
  }
}

__global__ void moments_finish_kernel(const cplx* __restrict__ partial, int chunks, long long RM, cplx* __restrict__ out) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= RM) return;
  cplx a = cmk(0.0, 0.0);
  for (int c = 0; c < chunks; ++c) a = cadd(a, partial[(size_t)c * RM + k]);
  out[k] = a;
}

// out[n] = sum_{r,j} sigma_{n,r} tau~_{n,j} x[r,j]  (O_n . x, Khatri-Rao mat-vec; one warp per sample).
// x is staged through shared memory in site blocks; lanes stride over hidden units.
__global__ void __launch_bounds__(256)
rbm_krmatvec_kernel(const int32_t* __restrict__ s, const cplx* __restrict__ tau, const cplx* __restrict__ x,
                    long long B, int N, int M, int hasBias, int conjTau, cplx* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n = (long long)blockIdx.x * 8 + warp;
  if (n >= B) return;
  const int R = N + (hasBias ? 1 : 0);
  cplx acc = cmk(0.0, 0.0);
  for (int j = lane; j < M; j += 32) {
    cplx t = tau[n * M + j];
    if (conjTau) t.y = -t.y;
    cplx col = cmk(0.0, 0.0);
    for (int r = 0; r < R; ++r) {
      double sg = (hasBias && r == 0) ? 1.0 : (double)(2 * s[n * N + (r - (hasBias ? 1 : 0))] - 1);
      cplx xv = x[(size_t)r * M + j];
      col.x = fma(sg, xv.x, col.x);
      col.y = fma(sg, xv.y, col.y);
    }
    acc = cadd(acc, cmul(t, col));
  }
  acc = warp_csum(acc);
  if (lane == 0) out[n] = acc;
}

// sigT[r][w] bit k of word w = (sigma_{32w+k, r} == +1); bias pseudo-site row (all ones) first.
__global__ void pack_sigma_kernel(const int32_t* __restrict__ s, long long B, int N, int hasBias, long long words,
                                  uint32_t* __restrict__ sigT) {
  const int lane = threadIdx.x & 31;
  long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per (site, word)
  const int R = N + (hasBias ? 1 : 0);
  if (gw >= (long long)R * words) return;
  int r = (int)(gw / words);
  long long w = gw - (long long)r * words;
  long long n = w * 32 + lane;
  bool bit = false;
  if (n < B) bit = (hasBias && r == 0) ? true : (s[n * N + (r - (hasBias ? 1 : 0))] != 0);
  uint32_t word = __ballot_sync(0xffffffffu, bit);
  if (lane == 0) sigT[(size_t)r * words + w] = word;
}

}  // namespace

extern "C" int jvmc_rbm_grad(const int32_t* s, const double* tau, long long B, int N, int M, int hasBias,
                             int layout, double* out, void* stream) {
  if (B == 0) return JVMC_OK;
  if (!s || !tau || !out || B < 0 || N <= 0 || M <= 0 || layout < 0 || layout > 1) return JVMC_ERR_ARG;
  rbm_grad_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(s, (const cplx*)tau, B, N, M, hasBias, layout, (cplx*)out);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_rbm_moments_chunks(long long B) {
  long long c = (B + 4095) / 4096;
  if (c < 1) c = 1;
  if (c > 64) c = 64;
  return (int)c;
}

// workspace: chunks * R * M complex128
extern "C" int jvmc_rbm_moments(const int32_t* s, const double* tau, const double* wgt, long long B, int N, int M,
                                int hasBias, int conjTau, double* workspace, double* out, void* stream) {
  if (!workspace || !out || B < 0 || N <= 0 || M <= 0) return JVMC_ERR_ARG;
  if (B > 0 && (!s || !tau || !wgt)) return JVMC_ERR_ARG;
  const int R = N + (hasBias ? 1 : 0);
  const int chunks = jvmc_rbm_moments_chunks(B);
  long long chunkLen = (B + chunks - 1) / chunks;
  chunkLen = ((chunkLen + MO_TS - 1) / MO_TS) * MO_TS;
  if (chunkLen == 0) chunkLen = MO_TS;
  dim3 grid((M + MO_THREADS - 1) / MO_THREADS, (R + MO_RG - 1) / MO_RG, chunks);
  rbm_moments_kernel<<<grid, MO_THREADS, 0, (cudaStream_t)stream>>>(
      s, (const cplx*)tau, (const cplx*)wgt, B, N, M, hasBias, conjTau, chunkLen, (cplx*)workspace);
  JVMC_CHECK_LAUNCH();
  long long RM = (long long)R * M;
  moments_finish_kernel<<<(unsigned)((RM + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const cplx*)workspace, chunks, RM, (cplx*)out);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_rbm_krmatvec(const int32_t* s, const double* tau, const double* x, long long B, int N, int M,
                                 int hasBias, int conjTau, double* out, void* stream) {
  if (B == 0) return JVMC_OK;
  if (!s || !tau || !x || !out || B < 0 || N <= 0 || M <= 0) return JVMC_ERR_ARG;
  rbm_krmatvec_kernel<<<(unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      s, (const cplx*)tau, (const cplx*)x, B, N, M, hasBias, conjTau, (cplx*)out);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_pack_sigma(const int32_t* s, long long B, int N, int hasBias, unsigned int* sigT, void* stream) {
  if (B == 0) return JVMC_OK;
  if (!s || !sigT || B < 0 || N <= 0) return JVMC_ERR_ARG;
  long long words = (B + 31) / 32;
  const int R = N + (hasBias ? 1 : 0);
  long long warps = (long long)R * words;
  pack_sigma_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(s, B, N, hasBias, words, sigT);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
