// Orbit-averaged (Cpx)RBM: the reference's SymNet wrapper around an RBM (jVMC/nets/sym_wrapper.py:8-66 with
// avgFun_Coefficients_Exp, orbits from jVMC/util/symmetries.py),
//
//     log Psi(s) = log sum_g f_g exp( sum_j logcosh( sum_i x^g_i W_ij + b_j ) ),   x^g = O_g sigma,
//
// where every O_g is a signed permutation: x^g_i = e_g(i) sigma_{p_g(i)}.  Because the symmetry acts on the INPUT,
// group element g is an RBM with permuted / sign-flipped weight rows, so the flip-ratio identity of the plain RBM
// (rbm_mcmc.cu) holds per group element and the sampler keeps tanh(theta^g) for all g instead of doing |G| forward
// passes per proposal.
//
//  jvmc_symrbm_logpsi  <- SymNet.__call__ under NQS.__call__ (vqs.py:223-251)
//  jvmc_symrbm_grad    <- NQS.gradients (vqs.py:46-69, 256-287) for the wrapped net, reference flat layout
//  jvmc_symrbm_mcmc    <- MCSampler._get_samples / _sweep (sampler.py:301-356) with propose_spin_flip
#include "common.cuh"

namespace {

struct SymArgs {
  int N, M, G;
  const cplx* W;        // [N, M]
  const cplx* bias;     // [M] or null
  const int32_t* pmap;  // [G, N]  x^g_i = esgn[g,i] * sigma[pmap[g,i]]
  const int32_t* esgn;  // [G, N]  +-1
  const cplx* fac;      // [G]
};

// a_g = sum_j logcosh(theta^g_j) for one group element; optionally stores tanh(theta^g_j)
__device__ __forceinline__ cplx group_amplitude(const SymArgs& a, const int32_t* __restrict__ cfg, int g, cplx* tauOut) {
  cplx acc = cmk(0.0, 0.0);
  const int32_t* pm = a.pmap + (size_t)g * a.N;
  const int32_t* es = a.esgn + (size_t)g * a.N;
  for (int j = 0; j < a.M; ++j) {
    cplx th = a.bias ? a.bias[j] : cmk(0.0, 0.0);
    for (int i = 0; i < a.N; ++i) {
      const double x = (double)(es[i] * (2 * cfg[pm[i]] - 1));
      const cplx w = a.W[(size_t)i * a.M + j];
      th.x = fma(x, w.x, th.x);
      th.y = fma(x, w.y, th.y);
    }
    cplx l, t;
    lncosh_tanh(th, l, t);
    acc = cadd(acc, l);
    if (tauOut) tauOut[j] = t;
  }
  return acc;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// one warp per sample, lanes over group elements
__global__ void symrbm_logpsi_kernel(SymArgs a, const int32_t* __restrict__ s, long long B, cplx* __restrict__ logpsi,
                                     cplx* __restrict__ wout) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const int32_t* cfg = s + b * a.N;
  // pass 1: the shift m = max_g Re a_g over the elements with a non-zero factor (jax.scipy.special.logsumexp)
  double m = -1e300;
  for (int g = lane; g < a.G; g += 32) {
    const cplx f = a.fac[g];
    if (f.x != 0.0 || f.y != 0.0) m = fmax(m, group_amplitude(a, cfg, g, nullptr).x);
  }
  m = warp_max(m);
  if (m == -1e300) m = 0.0;
  cplx S = cmk(0.0, 0.0);
  for (int g = lane; g < a.G; g += 32) {
    const cplx ag = group_amplitude(a, cfg, g, nullptr);
    const cplx t = cmul(a.fac[g], cexp(cmk(ag.x - m, ag.y)));
    S = cadd(S, t);
    if (wout) wout[b * a.G + g] = t;
  }
  S = warp_csum(S);
  if (wout) {
    __syncwarp();
    for (int g = lane; g < a.G; g += 32) wout[b * a.G + g] = cdiv(wout[b * a.G + g], S);
  }
  if (lane == 0) logpsi[b] = cmk(log(hypot(S.x, S.y)) + m, atan2(S.y, S.x));
}

// one CTA per sample: tau^g_j into shared memory, then every parameter sums over the group elements
__global__ void symrbm_grad_kernel(SymArgs a, const int32_t* __restrict__ s, long long B, const cplx* __restrict__ wts,
                                   int layout, cplx* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* tau = reinterpret_cast<cplx*>(smem_raw);                         // [G][M]
  int32_t* cfg = reinterpret_cast<int32_t*>(tau + (size_t)a.G * a.M);    // [N]
  const long long b = blockIdx.x;
  for (int i = threadIdx.x; i < a.N; i += blockDim.x) cfg[i] = s[b * a.N + i];
  __syncthreads();
  for (int u = threadIdx.x; u < a.G * a.M; u += blockDim.x) {
    const int g = u / a.M, j = u - g * a.M;
    const int32_t* pm = a.pmap + (size_t)g * a.N;
    const int32_t* es = a.esgn + (size_t)g * a.N;
    cplx th = a.bias ? a.bias[j] : cmk(0.0, 0.0);
    for (int i = 0; i < a.N; ++i) {
      const double x = (double)(es[i] * (2 * cfg[pm[i]] - 1));
      const cplx w = a.W[(size_t)i * a.M + j];
      th.x = fma(x, w.x, th.x);
      th.y = fma(x, w.y, th.y);
    }
    cplx l, t;
    lncosh_tanh(th, l, t);
    tau[u] = t;
  }
  __syncthreads();
  const int Mb = a.bias ? a.M : 0;
  const int Pc = Mb + a.N * a.M;
  cplx* row = out + b * (long long)(layout == 0 ? 2 * Pc : Pc);
  const cplx* w = wts + b * a.G;
  for (int c = threadIdx.x; c < Pc; c += blockDim.x) {
    cplx acc = cmk(0.0, 0.0);
    if (c < Mb) {
      for (int g = 0; g < a.G; ++g) acc = cadd(acc, cmul(w[g], tau[g * a.M + c]));
    } else {
      const int i = (c - Mb) / a.M, j = (c - Mb) - i * a.M;
      for (int g = 0; g < a.G; ++g) {
        const double x = (double)(a.esgn[(size_t)g * a.N + i] * (2 * cfg[a.pmap[(size_t)g * a.N + i]] - 1));
        acc = cadd(acc, cscale(cmul(w[g], tau[g * a.M + j]), x));
      }
    }
    if (layout == 0) {
      // holomorphic: per leaf [g, i g] (vqs.py:66-69); leaves in sorted-key order bias, kernel
      const bool isB = c < Mb;
      const int local = isB ? c : c - Mb;
      const int size = isB ? Mb : a.N * a.M;
      const long long base = isB ? 0 : 2LL * Mb;
      row[base + local] = acc;
      row[base + size + local] = cmk(-acc.y, acc.x);
    } else {
      row[c] = acc;
    }
  }
}

struct SymMcmcArgs {
  SymArgs n;
  const int32_t* qmap;  // [G, N]  row i of group element g that reads site k: i = qmap[g, k]
  const cplx* T;        // tanh(2W) [N, M]
  const cplx* lc;       // sum_j ln cosh(2 W_ij) [N]
  int32_t* states;
  long long C;
  unsigned long long seed, step0;
  long long chain0;
  double mu;
  int K;
  long long thermSteps;
  int numSamples;
  int refreshEvery;
  int32_t* out;
  unsigned long long* counters;
  int wpc;
};

constexpr int SYM_GQ = 8;   // group elements per lane: G <= 256

__global__ void symrbm_mcmc_kernel(SymMcmcArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long chain = (long long)blockIdx.x * a.wpc + warp;
  if (chain >= a.C) return;
  const int N = a.n.N, M = a.n.M, G = a.n.G;
  cplx* tau = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * G * M;                               // [G][M]
  int32_t* cfg = reinterpret_cast<int32_t*>(reinterpret_cast<cplx*>(smem_raw) + (size_t)a.wpc * G * M) + (size_t)warp * N;
  const unsigned long long gchain = (unsigned long long)(a.chain0 + chain);
  const Philox rng(a.seed);
  for (int i = lane; i < N; i += 32) cfg[i] = a.states[chain * N + i] != 0;
  __syncwarp();

  cplx cg[SYM_GQ];      // f_g psi_g(s) up to a common scale, for g = lane + 32 q
  cplx Ssum = cmk(0.0, 0.0);
  auto refresh = [&]() {
    cplx ag[SYM_GQ];
    double m = -1e300;
#pragma unroll
    for (int q = 0; q < SYM_GQ; ++q) {
      const int g = lane + 32 * q;
      ag[q] = cmk(0.0, 0.0);
      if (g < G) {
        ag[q] = group_amplitude(a.n, cfg, g, tau + (size_t)g * M);
        const cplx f = a.n.fac[g];
        if (f.x != 0.0 || f.y != 0.0) m = fmax(m, ag[q].x);
      }
    }
    m = warp_max(m);
    if (m == -1e300) m = 0.0;
    cplx S = cmk(0.0, 0.0);
#pragma unroll
    for (int q = 0; q < SYM_GQ; ++q) {
      const int g = lane + 32 * q;
      cg[q] = (g < G) ? cmul(a.n.fac[g], cexp(cmk(ag[q].x - m, ag[q].y))) : cmk(0.0, 0.0);
      S = cadd(S, cg[q]);
    }
    Ssum = warp_csum(S);
    __syncwarp();
  };

  unsigned long long nAcc = 0, nProp = 0;
  const long long total = a.thermSteps + (long long)a.numSamples * a.K;
  long long nextEmit = a.thermSteps + a.K;
  int emitted = 0, sweepCtr = 0;
  refresh();
  for (long long st = 0; st < total; ++st) {
    const unsigned long long gs = a.step0 + (unsigned long long)st;
    const uint4 r = rng((uint32_t)gs, (uint32_t)(gs >> 32), (uint32_t)gchain, (uint32_t)(gchain >> 32) << 8);
    const int k = (int)__umulhi(r.x, (uint32_t)N);
    const double sigk = cfg[k] ? 1.0 : -1.0;
    cplx rg[SYM_GQ];
    cplx Snew = cmk(0.0, 0.0);
#pragma unroll
    for (int q = 0; q < SYM_GQ; ++q) {
      const int g = lane + 32 * q;
      rg[q] = cmk(1.0, 0.0);
      if (g < G) {
        const int i = a.qmap[(size_t)g * N + k];
        const double sga = -sigk * (double)a.n.esgn[(size_t)g * N + i];   // -x^g_i
        const cplx* Ti = a.T + (size_t)i * M;
        const cplx* tg = tau + (size_t)g * M;
        cplx p = cmk(1.0, 0.0);
        for (int j = 0; j < M; ++j) {
          const cplx tt = Ti[j], tj = tg[j];
          p = cmul(p, cmk(fma(sga, fma(tj.x, tt.x, -tj.y * tt.y), 1.0), sga * fma(tj.x, tt.y, tj.y * tt.x)));
        }
        rg[q] = cmul(cexp(a.lc[i]), p);
        Snew = cadd(Snew, cmul(cg[q], rg[q]));
      }
    }
    Snew = warp_csum(Snew);
    const double ratio2 = cabs2(Snew) / cabs2(Ssum);               // |Psi(s')/Psi(s)|^2
    const double P = (a.mu == 2.0) ? ratio2 : pow(ratio2, 0.5 * a.mu);
    const double u = u01_from_bits(r.z, r.w);
    nProp += 1;
    if (u < P) {
      nAcc += 1;
#pragma unroll
      for (int q = 0; q < SYM_GQ; ++q) {
        const int g = lane + 32 * q;
        if (g < G) {
          const int i = a.qmap[(size_t)g * N + k];
          const double sga = -sigk * (double)a.n.esgn[(size_t)g * N + i];
          const cplx* Ti = a.T + (size_t)i * M;
          cplx* tg = tau + (size_t)g * M;
          for (int j = 0; j < M; ++j) {
            const cplx n = cscale(Ti[j], sga), tj = tg[j];
            tg[j] = cdiv(cadd(tj, n), cadd(cmk(1.0, 0.0), cmul(tj, n)));
          }
          cg[q] = cmul(cg[q], rg[q]);
        }
      }
      Ssum = Snew;
      __syncwarp();
      if (lane == 0) cfg[k] ^= 1;
      __syncwarp();
    }
    if (st + 1 == nextEmit) {
      const long long row = (long long)emitted * a.C + chain;   // time-major, chain-minor (sampler.py:323)
      for (int i = lane; i < N; i += 32) a.out[row * N + i] = cfg[i];
      ++emitted;
      nextEmit += a.K;
    }
    if ((st + 1) % a.K == 0) {
      if (++sweepCtr >= a.refreshEvery && st + 1 < total) { sweepCtr = 0; refresh(); }
    }
  }
  for (int i = lane; i < N; i += 32) a.states[chain * N + i] = cfg[i];
  if (lane == 0) {
    atomicAdd(a.counters + 0, nProp);
    atomicAdd(a.counters + 1, nAcc);
  }
}

SymArgs make_sym(int N, int M, int G, const double* W, const double* bias, const int32_t* pmap, const int32_t* esgn,
                 const double* fac) {
  SymArgs a;
  a.N = N; a.M = M; a.G = G;
  a.W = (const cplx*)W; a.bias = (const cplx*)bias; a.pmap = pmap; a.esgn = esgn; a.fac = (const cplx*)fac;
  return a;
}

}  // namespace

extern "C" int jvmc_symrbm_logpsi(const int32_t* s, long long B, int N, int M, int G, const double* W, const double* bias,
                                  const int32_t* pmap, const int32_t* esgn, const double* fac, double* logpsi,
                                  double* weights, void* stream) {
  if (B == 0) return JVMC_OK;
  if (!s || !W || !pmap || !esgn || !fac || !logpsi || B < 0 || N <= 0 || M <= 0 || G <= 0) return JVMC_ERR_ARG;
  SymArgs a = make_sym(N, M, G, W, bias, pmap, esgn, fac);
  symrbm_logpsi_kernel<<<(unsigned)((B * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, s, B, (cplx*)logpsi,
                                                                                           (cplx*)weights);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_symrbm_grad(const int32_t* s, long long B, int N, int M, int G, const double* W, const double* bias,
                                const int32_t* pmap, const int32_t* esgn, const double* fac, const double* weights,
                                int layout, double* out, void* stream) {
  if (B == 0) return JVMC_OK;
  if (!s || !W || !pmap || !esgn || !fac || !weights || !out || B < 0 || N <= 0 || M <= 0 || G <= 0) return JVMC_ERR_ARG;
  if (layout != 0 && layout != 1) return JVMC_ERR_ARG;
  SymArgs a = make_sym(N, M, G, W, bias, pmap, esgn, fac);
  size_t smem = (size_t)G * M * sizeof(cplx) + (size_t)N * sizeof(int32_t);
  if (smem > 227 * 1024) return JVMC_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) cudaFuncSetAttribute(symrbm_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  symrbm_grad_kernel<<<(unsigned)B, 256, smem, (cudaStream_t)stream>>>(a, s, B, (const cplx*)weights, layout, (cplx*)out);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_symrbm_mcmc(int32_t* states, long long C, int N, int M, int G, const double* W, const double* bias,
                                const int32_t* pmap, const int32_t* esgn, const int32_t* qmap, const double* fac,
                                const double* tables, unsigned long long seed, unsigned long long step0,
                                long long chain0, double mu, int sweepSteps, long long thermSteps,
                                int numSamplesPerChain, int refreshEvery, int32_t* out, unsigned long long* counters,
                                void* stream) {
  if (!states || !W || !pmap || !esgn || !qmap || !fac || !tables || !counters || C < 0 || N <= 0 || M <= 0 || G <= 0)
    return JVMC_ERR_ARG;
  if (numSamplesPerChain > 0 && !out) return JVMC_ERR_ARG;
  if (sweepSteps <= 0 || thermSteps < 0 || numSamplesPerChain < 0) return JVMC_ERR_ARG;
  if (G > 32 * SYM_GQ) return JVMC_ERR_UNSUPPORTED;
  if (C == 0) return JVMC_OK;
  SymMcmcArgs a;
  a.n = make_sym(N, M, G, W, bias, pmap, esgn, fac);
  a.qmap = qmap;
  a.T = (const cplx*)tables; a.lc = a.T + (size_t)N * M;
  a.states = states; a.C = C; a.seed = seed; a.step0 = step0; a.chain0 = chain0; a.mu = mu; a.K = sweepSteps;
  a.thermSteps = thermSteps; a.numSamples = numSamplesPerChain; a.refreshEvery = refreshEvery > 0 ? refreshEvery : 1;
  a.out = out; a.counters = counters;
  const size_t perWarp = (size_t)G * M * sizeof(cplx) + (size_t)N * sizeof(int32_t);
  int wpc = (int)((200 * 1024) / perWarp);
  if (wpc < 1) return JVMC_ERR_UNSUPPORTED;
  if (wpc > 4) wpc = 4;
  a.wpc = wpc;
  const size_t smem = (size_t)wpc * perWarp;
  if (smem > 48 * 1024) cudaFuncSetAttribute(symrbm_mcmc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  symrbm_mcmc_kernel<<<(unsigned)((C + wpc - 1) / wpc), wpc * 32, smem, (cudaStream_t)stream>>>(a);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
