// Real-parameter convolutional ansatz (reference jVMC/nets/cnn.py:18-81, class CNN): periodic ("wrap") padding by
// F-1 at the end of every axis, cross-correlation with stride, activation, repeated per layer, then
// sum over positions and channels / sqrt(#positions * #channels of the last layer).  1-D and 2-D lattices.
//
//  jvmc_cnn_logpsi  <- NQS.__call__ / _eval (vqs.py:223-251) for nets.CNN
//  jvmc_cnn_grad    <- NQS.gradients with flat_gradient (vqs.py:46-51, 256-287): per-sample back-propagation,
//                      output in the reference's flat parameter order (Conv_l: bias, kernel[F..., Cin, Cout])
//  jvmc_cnn_mcmc    <- MCSampler._get_samples / _sweep (sampler.py:301-356) with a full forward pass per proposal
//                      (no low-rank update exists for a CNN), proposers spin_flip / _Z2 / _zeroMag (:15-64)
//
// The flat parameter vector theta (NQS.get_parameters) is the kernels' weight format: for every layer first the bias
// (if the layer has one), then the kernel in C order [Fx, Fy, Cin, Cout].  One CTA per sample / chain; activations
// live in shared memory.  Correctness-first kernels for the "next" row of the scope table (config 4).
#include "common.cuh"
#include "cnn_common.cuh"

namespace {

// pre-activation of output element (px, py, co) of layer l from the previous layer's output `in`
__device__ __forceinline__ double conv_at(const CnnDesc& d, int l, const double* __restrict__ theta,
                                          const double* __restrict__ in, int px, int py, int co) {
  const int ci_n = d.ch[l], co_n = d.ch[l + 1];
  const int ix = d.ox[l], iy = d.oy[l];
  const double* Kp = theta + d.offK[l];
  double acc = d.hasBias[l] ? theta[d.offB[l] + co] : 0.0;
  // wrap without integer division: the window start is inside the lattice, a tap wraps at most a few times
  int qx = px * d.sx;
  for (int fx = 0; fx < d.Fx; ++fx, ++qx) {
    while (qx >= ix) qx -= ix;
    int qy = py * d.sy;
    for (int fy = 0; fy < d.Fy; ++fy, ++qy) {
      while (qy >= iy) qy -= iy;
      const double* row = in + (size_t)(qx * iy + qy) * ci_n;
      const double* kr = Kp + (size_t)((fx * d.Fy + fy) * ci_n) * co_n + co;
      for (int ci = 0; ci < ci_n; ++ci) acc = fma(row[ci], kr[(size_t)ci * co_n], acc);
    }
  }
  return acc;
}

// CTA-cooperative forward pass.  act: scratch of d.totA doubles (input at offA[0] must be filled, +-1);
// zbuf (optional, same layout): pre-activations.  Returns the network output in every thread.
__device__ double cnn_forward(const CnnDesc& d, const double* __restrict__ theta, double* act, double* zbuf,
                              double* red) {
  for (int l = 0; l < d.nl; ++l) {
    const int co_n = d.ch[l + 1];
    const int n = d.ox[l + 1] * d.oy[l + 1] * co_n;
    const double* in = act + d.offA[l];
    double* out = act + d.offA[l + 1];
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const int co = e % co_n, p = e / co_n;
      const int py = p % d.oy[l + 1], px = p / d.oy[l + 1];
      const double z = conv_at(d, l, theta, in, px, py, co);
      if (zbuf) zbuf[d.offA[l + 1] + e] = z;
      out[e] = actf(d.act[l], z);
    }
    __syncthreads();
  }
  // block sum of the last layer
  const int n = d.ox[d.nl] * d.oy[d.nl] * d.ch[d.nl];
  const double* last = act + d.offA[d.nl];
  double s = 0.0;
  for (int e = threadIdx.x; e < n; e += blockDim.x) s += last[e];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  __syncthreads();
  return tot / d.nrm;
}

__global__ void cnn_logpsi_kernel(CnnDesc d, const double* __restrict__ theta, const int32_t* __restrict__ s, int N,
                                  cplx* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* act = reinterpret_cast<double*>(smem_raw);
  double* red = act + d.totA;
  const long long b = blockIdx.x;
  for (int i = threadIdx.x; i < N; i += blockDim.x) act[i] = (double)(2 * s[b * N + i] - 1);
  __syncthreads();
  const double v = cnn_forward(d, theta, act, nullptr, red);
  if (threadIdx.x == 0) out[b] = cmk(v, 0.0);
}

// per-sample gradient by back-propagation; delta buffers share the layout of the activations
__global__ void cnn_grad_kernel(CnnDesc d, const double* __restrict__ theta, const int32_t* __restrict__ s, int N,
                                cplx* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* act = reinterpret_cast<double*>(smem_raw);
  double* zb = act + d.totA;
  double* dl = zb + d.totA;
  double* red = dl + d.totA;
  const long long b = blockIdx.x;
  for (int i = threadIdx.x; i < N; i += blockDim.x) act[i] = (double)(2 * s[b * N + i] - 1);
  __syncthreads();
  (void)cnn_forward(d, theta, act, zb, red);
  // delta of the last layer: d out / d z = f'(z) / nrm
  {
    const int n = d.ox[d.nl] * d.oy[d.nl] * d.ch[d.nl];
    for (int e = threadIdx.x; e < n; e += blockDim.x)
      dl[d.offA[d.nl] + e] = dactf(d.act[d.nl - 1], zb[d.offA[d.nl] + e]) / d.nrm;
  }
  __syncthreads();
  cplx* row = out + b * (long long)d.P;
  for (int l = d.nl - 1; l >= 0; --l) {
    const int ci_n = d.ch[l], co_n = d.ch[l + 1];
    const int ix = d.ox[l], iy = d.oy[l], ox = d.ox[l + 1], oy = d.oy[l + 1];
    const double* in = act + d.offA[l];
    const double* del = dl + d.offA[l + 1];
    // parameter gradients of layer l
    if (d.hasBias[l])
      for (int co = threadIdx.x; co < co_n; co += blockDim.x) {
        double acc = 0.0;
        for (int p = 0; p < ox * oy; ++p) acc += del[(size_t)p * co_n + co];
        row[d.offB[l] + co] = cmk(acc, 0.0);
      }
    const int nk = d.Fx * d.Fy * ci_n * co_n;
    for (int e = threadIdx.x; e < nk; e += blockDim.x) {
      const int co = e % co_n, ci = (e / co_n) % ci_n, f = e / (co_n * ci_n);
      const int fy = f % d.Fy, fx = f / d.Fy;
      double acc = 0.0;
      for (int px = 0; px < ox; ++px) {
        const int qx = (px * d.sx + fx) % ix;
        for (int py = 0; py < oy; ++py) {
          const int qy = (py * d.sy + fy) % iy;
          acc = fma(in[(size_t)(qx * iy + qy) * ci_n + ci], del[(size_t)(px * oy + py) * co_n + co], acc);
        }
      }
      row[d.offK[l] + e] = cmk(acc, 0.0);
    }
    // delta of the previous layer (not needed below the first one)
    if (l > 0) {
      const double* Kp = theta + d.offK[l];
      const int n = ix * iy * ci_n;
      for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int ci = e % ci_n, q = e / ci_n;
        const int qy = q % iy, qx = q / iy;
        double acc = 0.0;
        // outputs (px, py) whose window covers (qx, qy): px sx + fx = qx + w ix for a wrap count w >= 0, same for y
        for (int fx = 0; fx < d.Fx; ++fx)
          for (int vx = qx - fx; vx <= (ox - 1) * d.sx; vx += ix) {
            if (vx < 0 || vx % d.sx != 0) continue;
            const int px = vx / d.sx;
            for (int fy = 0; fy < d.Fy; ++fy)
              for (int vy = qy - fy; vy <= (oy - 1) * d.sy; vy += iy) {
                if (vy < 0 || vy % d.sy != 0) continue;
                const int py = vy / d.sy;
                const double* kr = Kp + (size_t)((fx * d.Fy + fy) * ci_n + ci) * co_n;
                const double* dr = del + (size_t)(px * oy + py) * co_n;
                for (int co = 0; co < co_n; ++co) acc = fma(kr[co], dr[co], acc);
              }
          }
        dl[d.offA[l] + e] = acc * dactf(d.act[l - 1], zb[d.offA[l] + e]);
      }
    }
    __syncthreads();
  }
}

struct CnnMcmcArgs {
  int32_t* states;
  long long C;
  int N;
  unsigned long long seed, step0;
  long long chain0;
  int proposer;
  double mu;
  int K;
  long long thermSteps;
  int numSamples;
  int32_t* out;
  unsigned long long* counters;
  int thetaInSmem;
};

// one CTA per chain, full forward pass per proposal
__global__ void cnn_mcmc_kernel(CnnDesc d, const double* theta, CnnMcmcArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* act = reinterpret_cast<double*>(smem_raw);
  double* red = act + d.totA;
  int32_t* cfg = reinterpret_cast<int32_t*>(red + 32);     // current configuration
  int32_t* prop = cfg + a.N;                                 // proposal
  // the chain evaluates the net at every step: keep the parameters in shared memory when they fit
  if (a.thetaInSmem) {
    double* th = reinterpret_cast<double*>(prop + a.N + (a.N & 1));
    for (int i = threadIdx.x; i < d.P; i += blockDim.x) th[i] = theta[i];
    theta = th;
  }
  __shared__ int sh_accept;
  const long long chain = blockIdx.x;
  const int N = a.N;
  const unsigned long long gchain = (unsigned long long)(a.chain0 + chain);
  const Philox rng(a.seed);
  for (int i = threadIdx.x; i < N; i += blockDim.x) cfg[i] = a.states[chain * N + i] != 0;
  __syncthreads();
  auto eval = [&](const int32_t* c) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) act[i] = (double)(2 * c[i] - 1);
    __syncthreads();
    return cnn_forward(d, theta, act, nullptr, red);
  };
  double cur = eval(cfg);
  unsigned long long nAcc = 0, nProp = 0;
  const long long total = a.thermSteps + (long long)a.numSamples * a.K;
  long long nextEmit = a.thermSteps + a.K;
  int emitted = 0;
  for (long long st = 0; st < total; ++st) {
    const unsigned long long gs = a.step0 + (unsigned long long)st;
    const uint4 r = rng((uint32_t)gs, (uint32_t)(gs >> 32), (uint32_t)gchain, (uint32_t)(gchain >> 32) << 8);
    for (int i = threadIdx.x; i < N; i += blockDim.x) prop[i] = cfg[i];
    __syncthreads();
    if (threadIdx.x == 0) {
      bool g = false;
      if (a.proposer == 2) {
        // exchange the ru-th up spin with the rd-th down spin (sampler.py:42-64); both counts are N/2 in the
        // zero-magnetisation sector the chains start in
        const int half = N / 2;
        int ru = 1 + (int)__umulhi(r.x, (uint32_t)half), rd = 1 + (int)__umulhi(r.y, (uint32_t)half);
        int iu = -1, id = -1;
        for (int i = 0; i < N; ++i) {
          if (cfg[i]) { if (--ru == 0) iu = i; } else { if (--rd == 0) id = i; }
        }
        if (iu >= 0 && id >= 0) { prop[iu] = 0; prop[id] = 1; }
        else if (iu >= 0) prop[iu] = 0;
        else if (id >= 0) prop[id] = 1;
        const uint4 r2 = rng((uint32_t)gs, (uint32_t)(gs >> 32), (uint32_t)gchain, ((uint32_t)(gchain >> 32) << 8) | 1u);
        g = __umulhi(r2.x, 5u) == 0u;
      } else {
        const int k = (int)__umulhi(r.x, (uint32_t)N);
        prop[k] ^= 1;
        g = (a.proposer == 1) && (__umulhi(r.y, 5u) == 0u);
      }
      if (g) for (int i = 0; i < N; ++i) prop[i] ^= 1;
    }
    __syncthreads();
    const double nxt = eval(prop);
    if (threadIdx.x == 0) {
      const double P = exp(a.mu * (nxt - cur));
      sh_accept = u01_from_bits(r.z, r.w) < P;
    }
    __syncthreads();
    nProp += 1;
    if (sh_accept) {
      nAcc += 1;
      cur = nxt;
      for (int i = threadIdx.x; i < N; i += blockDim.x) cfg[i] = prop[i];
    }
    __syncthreads();
    if (st + 1 == nextEmit) {
      const long long row = (long long)emitted * a.C + chain;   // time-major, chain-minor (sampler.py:323)
      for (int i = threadIdx.x; i < N; i += blockDim.x) a.out[row * N + i] = cfg[i];
      ++emitted;
      nextEmit += a.K;
    }
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) a.states[chain * N + i] = cfg[i];
  if (threadIdx.x == 0) {
    atomicAdd(a.counters + 0, nProp);
    atomicAdd(a.counters + 1, nAcc);
  }
}

// threads per CTA: enough to give every thread a few output elements of the widest layer
int cnn_threads(const CnnDesc& d) {
  int widest = 0;
  for (int l = 1; l <= d.nl; ++l) widest = max(widest, d.ox[l] * d.oy[l] * d.ch[l]);
  return widest >= 1024 ? 512 : (widest >= 384 ? 256 : 128);
}

}  // namespace

extern "C" int jvmc_cnn_num_parameters(const int* desc, int ndesc, int* P) {
  CnnDesc d;
  int rc = make_desc(desc, ndesc, d);
  if (rc != JVMC_OK || !P) return rc != JVMC_OK ? rc : JVMC_ERR_ARG;
  *P = d.P;
  return JVMC_OK;
}

extern "C" int jvmc_cnn_logpsi(const int* desc, int ndesc, const double* theta, const int32_t* s, long long B,
                               double* logpsi, void* stream) {
  CnnDesc d;
  int rc = make_desc(desc, ndesc, d);
  if (rc != JVMC_OK) return rc;
  if (B == 0) return JVMC_OK;
  if (!theta || !s || !logpsi || B < 0) return JVMC_ERR_ARG;
  size_t smem = (size_t)(d.totA + 32) * sizeof(double);
  if (smem > 227 * 1024) return JVMC_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) cudaFuncSetAttribute(cnn_logpsi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cnn_logpsi_kernel<<<(unsigned)B, cnn_threads(d), smem, (cudaStream_t)stream>>>(d, theta, s, d.Lx * d.Ly, (cplx*)logpsi);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_cnn_grad(const int* desc, int ndesc, const double* theta, const int32_t* s, long long B, double* out,
                             void* stream) {
  CnnDesc d;
  int rc = make_desc(desc, ndesc, d);
  if (rc != JVMC_OK) return rc;
  if (B == 0) return JVMC_OK;
  if (!theta || !s || !out || B < 0) return JVMC_ERR_ARG;
  size_t smem = (size_t)(3 * d.totA + 32) * sizeof(double);
  if (smem > 227 * 1024) return JVMC_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) cudaFuncSetAttribute(cnn_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cnn_grad_kernel<<<(unsigned)B, cnn_threads(d), smem, (cudaStream_t)stream>>>(d, theta, s, d.Lx * d.Ly, (cplx*)out);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_cnn_mcmc_inc(const int* desc, int ndesc, const double* theta, int32_t* states, long long C,
                                 unsigned long long seed, unsigned long long step0, long long chain0, int proposer,
                                 double mu, int sweepSteps, long long thermSteps, int numSamplesPerChain, int32_t* out,
                                 unsigned long long* counters, void* stream);

extern "C" int jvmc_cnn_mcmc(const int* desc, int ndesc, const double* theta, int32_t* states, long long C,
                             unsigned long long seed, unsigned long long step0, long long chain0, int proposer, double mu,
                             int sweepSteps, long long thermSteps, int numSamplesPerChain, int32_t* out,
                             unsigned long long* counters, void* stream) {
  // stride-1 nets: one warp per chain with incremental updates of the cached activations (cnn_inc.cu)
  int rc = jvmc_cnn_mcmc_inc(desc, ndesc, theta, states, C, seed, step0, chain0, proposer, mu, sweepSteps, thermSteps,
                             numSamplesPerChain, out, counters, stream);
  if (rc != JVMC_ERR_UNSUPPORTED) return rc;
  CnnDesc d;
  rc = make_desc(desc, ndesc, d);
  if (rc != JVMC_OK) return rc;
  if (!theta || !states || !counters || C < 0) return JVMC_ERR_ARG;
  if (numSamplesPerChain > 0 && !out) return JVMC_ERR_ARG;
  if (proposer < 0 || proposer > 2 || sweepSteps <= 0 || thermSteps < 0 || numSamplesPerChain < 0) return JVMC_ERR_ARG;
  const int N = d.Lx * d.Ly;
  if (proposer == 2 && (N % 2 != 0)) return JVMC_ERR_ARG;
  if (C == 0) return JVMC_OK;
  CnnMcmcArgs a;
  a.states = states; a.C = C; a.N = N; a.seed = seed; a.step0 = step0; a.chain0 = chain0; a.proposer = proposer;
  a.mu = mu; a.K = sweepSteps; a.thermSteps = thermSteps; a.numSamples = numSamplesPerChain; a.out = out;
  a.counters = counters;
  size_t smem = (size_t)(d.totA + 32) * sizeof(double) + (size_t)(2 * N + (N & 1)) * sizeof(int32_t);
  if (smem > 227 * 1024) return JVMC_ERR_UNSUPPORTED;
  a.thetaInSmem = (smem + (size_t)d.P * sizeof(double) <= 100 * 1024) ? 1 : 0;
  if (a.thetaInSmem) smem += (size_t)d.P * sizeof(double);
  if (smem > 48 * 1024) cudaFuncSetAttribute(cnn_mcmc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cnn_mcmc_kernel<<<(unsigned)C, cnn_threads(d), smem, (cudaStream_t)stream>>>(d, theta, a);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
