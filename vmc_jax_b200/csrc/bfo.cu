// Branch-free operator kernels.
//
//  jvmc_bfo_matels   matrix elements of every operator string + nonzero bookkeeping
//  jvmc_bfo_emit     compacted s' / matEl in the reference's order and padding
//        <- Operator.get_s_primes, _find_nonzero, set_zero_to_zero (reference
//           jVMC/operator/base.py:91-160) and BranchFreeOperator._get_s_primes
//           (jVMC/operator/branch_free.py:443-487).  Integer outputs and matrix elements are
//           bit-exact w.r.t. the oracle (products rounded separately, sequential diagonal merge).
//  jvmc_oloc_reduce  O_loc = sum_k matEl_k exp(logpsi(s'_k) - logpsi(s))   <- base.py:162-164
//  jvmc_rbm_eloc_bfo fused local energy for (Cpx)RBM: psi(s')/psi(s) from tanh/cosh addition
//           theorems on the cached tau (no s' materialisation, no forward passes)
//        <- Operator.get_O_loc fast path (base.py:166-192) for strings flipping <= 2 sites.
#include "common.cuh"

namespace {

constexpr int BFO_MAXLEN = 16;

struct BfoTables {
  int numOps, len, lDim;
  const int32_t* idx;     // [numOps, len]
  const int32_t* map;     // [numOps, len, lDim]
  const cplx* matEls;     // [numOps, len, lDim]
  const int32_t* fermi;   // [numOps, len]
  const uint8_t* isDiag;  // [numOps]
};

// walk one operator string on configuration c (read-only) ; returns matEl, records modified sites
__device__ __forceinline__ cplx walk_string(const BfoTables& t, int o, const int32_t* __restrict__ c, int N,
                                            cplx m, int* msite, int* mval, int& nmod) {
  nmod = 0;
  for (int k = 0; k < t.len; ++k) {
    const int i = t.idx[o * t.len + k];
    int cur = c[i];
    for (int q = 0; q < nmod; ++q) if (msite[q] == i) cur = mval[q];
    m = cmul_exact(m, t.matEls[((size_t)o * t.len + k) * t.lDim + cur]);
    if (t.fermi[o * t.len + k]) {
      // Jordan-Wigner sign prod_{j>i} (1 - 2 c_j) on the current (modified) configuration
      int sgn = 1;
      for (int j = i + 1; j < N; ++j) {
        int cj = c[j];
        for (int q = 0; q < nmod; ++q) if (msite[q] == j) cj = mval[q];
        sgn *= (1 - 2 * cj);
      }
      m = cmk(m.x * (double)sgn, m.y * (double)sgn);
    }
    const int nv = t.map[((size_t)o * t.len + k) * t.lDim + cur];
    bool found = false;
    for (int q = 0; q < nmod; ++q) if (msite[q] == i) { mval[q] = nv; found = true; }
    if (!found) { msite[nmod] = i; mval[nmod] = nv; ++nmod; }
  }
  return m;
}

__global__ void bfo_matels_kernel(BfoTables t, const int32_t* __restrict__ s, long long B, int N,
                                  const cplx* __restrict__ pref, cplx* __restrict__ mAll) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * t.numOps) return;
  long long b = gid / t.numOps;
  int o = (int)(gid - b * t.numOps);
  int msite[BFO_MAXLEN], mval[BFO_MAXLEN], nmod;
  mAll[gid] = walk_string(t, o, s + b * N, N, pref[o], msite, mval, nmod);
}

// one thread per sample: sequential diagonal merge (branch_free.py:483-485), nonzero choice list
// (base.py:91-103: index if |m|>1e-6 else numOps-1, sorted ascending), count and running max.
__global__ void bfo_nonzero_kernel(BfoTables t, long long B, int numDiag, cplx* __restrict__ mAll,
                                   int32_t* __restrict__ choice, int32_t* __restrict__ count, int32_t* maxCount) {
  long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  cplx* m = mAll + b * t.numOps;
  if (numDiag > 1) {
    cplx acc = cmk(0.0, 0.0);
    int first = -1;
    for (int o = 0; o < t.numOps; ++o)
      if (t.isDiag[o]) {
        acc = cmk(__dadd_rn(acc.x, m[o].x), __dadd_rn(acc.y, m[o].y));
        if (first < 0) first = o;
      }
    for (int o = 0; o < t.numOps; ++o) if (t.isDiag[o]) m[o] = cmk(0.0, 0.0);
    m[first] = acc;
  }
  int n = 0;
  int32_t* ch = choice + b * t.numOps;
  for (int o = 0; o < t.numOps; ++o) {
    double ab = hypot(m[o].x, m[o].y);
    if (ab > 1e-6) ch[n++] = o;
  }
  for (int k = n; k < t.numOps; ++k) ch[k] = t.numOps - 1;
  count[b] = n;
  atomicMax(maxCount, n);
}

// one warp per output row (b, k): copy the configuration, apply string choice[b,k]
__global__ void bfo_emit_kernel(BfoTables t, const int32_t* __restrict__ s, long long B, int N,
                                const cplx* __restrict__ mAll, const int32_t* __restrict__ choice,
                                const int32_t* __restrict__ count, int Kmax, int32_t* __restrict__ sp,
                                cplx* __restrict__ matEl) {
  const int lane = threadIdx.x & 31;
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= B * Kmax) return;
  long long b = row / Kmax;
  int k = (int)(row - b * Kmax);
  int o = choice[b * t.numOps + k];
  const int32_t* c = s + b * N;
  int32_t* dst = sp + row * N;
  // every lane replays the (short) string redundantly and patches the sites it owns
  int msite[BFO_MAXLEN], mval[BFO_MAXLEN], nmod;
  (void)walk_string(t, o, c, N, cmk(1.0, 0.0), msite, mval, nmod);
  for (int i = lane; i < N; i += 32) {
    int v = c[i];
    for (int q = 0; q < nmod; ++q) if (msite[q] == i) v = mval[q];
    dst[i] = v;
  }
  if (lane == 0) matEl[row] = (k < count[b]) ? mAll[b * t.numOps + o] : cmk(0.0, 0.0);
}

__global__ void oloc_reduce_kernel(const cplx* __restrict__ matEl, const cplx* __restrict__ lpS,
                                   const cplx* __restrict__ lpSP, long long B, int K, cplx* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  cplx l0 = lpS[b];
  cplx acc = cmk(0.0, 0.0);
  for (int k = lane; k < K; k += 32) {
    cplx m = matEl[b * K + k];
    cplx e = cexp(csub(lpSP[b * K + k], l0));
    acc = cadd(acc, cmul(m, e));
  }
  acc = warp_csum(acc);
  if (lane == 0) out[b] = acc;
}

// ---------------------------------------------------------------------------------------------
// Fused RBM local energy.  One warp per sample; tau row in shared memory.  Lanes first walk 32
// strings in parallel (matrix element, flipped sites), then the warp evaluates the amplitude
// ratio of every off-diagonal string with lanes striding over hidden units.
constexpr int EL_WPC = 4;

__global__ void __launch_bounds__(EL_WPC * 32)
rbm_eloc_kernel(BfoTables t, const int32_t* __restrict__ s, const cplx* __restrict__ tauG, long long B, int N, int M,
                const cplx* __restrict__ T, const cplx* __restrict__ lc, const cplx* __restrict__ pref,
                int numDiag, cplx* __restrict__ out, int* __restrict__ errFlag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long b = (long long)blockIdx.x * EL_WPC + warp;
  if (b >= B) return;
  cplx* tau = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * M;
  int32_t* cfg = reinterpret_cast<int32_t*>(reinterpret_cast<cplx*>(smem_raw) + (size_t)EL_WPC * M) + (size_t)warp * N;
  for (int j = lane; j < M; j += 32) tau[j] = tauG[b * M + j];
  for (int i = lane; i < N; i += 32) cfg[i] = s[b * N + i];
  __syncwarp();

  cplx diagAcc = cmk(0.0, 0.0);   // merged diagonal strings
  cplx eloc = cmk(0.0, 0.0);
  for (int o0 = 0; o0 < t.numOps; o0 += 32) {
    const int o = o0 + lane;
    cplx m = cmk(0.0, 0.0);
    int f0 = -1, f1 = -1;
    bool isd = true;
    if (o < t.numOps) {
      int msite[BFO_MAXLEN], mval[BFO_MAXLEN], nmod;
      m = walk_string(t, o, cfg, N, pref[o], msite, mval, nmod);
      isd = t.isDiag[o] != 0;
      int nf = 0;
      for (int q = 0; q < nmod; ++q)
        if (mval[q] != cfg[msite[q]]) { if (nf == 0) f0 = msite[q]; else if (nf == 1) f1 = msite[q]; ++nf; }
      if (nf > 2) atomicExch(errFlag, 1);
    }
    // diagonal strings are merged into one entry (branch_free.py:483-485) ...
    diagAcc = cadd(diagAcc, warp_csum((isd && o < t.numOps) ? m : cmk(0.0, 0.0)));
    // off-diagonal strings: warp-cooperative ratio
    unsigned live = __ballot_sync(0xffffffffu, (o < t.numOps) && !isd && (hypot(m.x, m.y) > 1e-6));
    while (live) {
      const int src = __ffs(live) - 1;
      live &= live - 1;
      const cplx mm = cmk(__shfl_sync(0xffffffffu, m.x, src), __shfl_sync(0xffffffffu, m.y, src));
      const int a0 = __shfl_sync(0xffffffffu, f0, src);
      const int a1 = __shfl_sync(0xffffffffu, f1, src);
      cplx ratio = cmk(1.0, 0.0);
      if (a0 >= 0) {
        const double sg0 = cfg[a0] ? -1.0 : 1.0;   // -sigma
        const cplx* T0 = T + (size_t)a0 * M;
        cplx lcs = lc[a0];
        cplx p = cmk(1.0, 0.0);
        if (a1 < 0) {
          for (int j = lane; j < M; j += 32) {
            cplx tt = T0[j], tj = tau[j];
            cplx f = cmk(fma(sg0, fma(tj.x, tt.x, -tj.y * tt.y), 1.0), sg0 * fma(tj.x, tt.y, tj.y * tt.x));
            p = cmul(p, f);
          }
        } else {
          const double sg1 = cfg[a1] ? -1.0 : 1.0;
          const cplx* T1 = T + (size_t)a1 * M;
          lcs = cadd(lcs, lc[a1]);
          for (int j = lane; j < M; j += 32) {
            cplx ta = cscale(T0[j], sg0), tb = cscale(T1[j], sg1), tj = tau[j];
            cplx n = cadd(ta, tb);
            cplx d = cadd(cmk(1.0, 0.0), cmul(ta, tb));
            p = cmul(p, cadd(d, cmul(tj, n)));
          }
        }
        p = warp_cprod(p);
        ratio = cmul(cexp(lcs), p);
      }
      eloc = cadd(eloc, cmul(mm, ratio));
    }
  }
  // ... and that entry passes the same |m| > 1e-6 filter as every other one (base.py:93-103)
  if (hypot(diagAcc.x, diagAcc.y) <= 1e-6) diagAcc = cmk(0.0, 0.0);
  if (lane == 0) out[b] = cadd(eloc, diagAcc);
}

BfoTables make_tables(int numOps, int len, int lDim, const int32_t* idx, const int32_t* map, const double* matEls,
                      const int32_t* fermi, const uint8_t* isDiag) {
  BfoTables t;
  t.numOps = numOps; t.len = len; t.lDim = lDim;
  t.idx = idx; t.map = map; t.matEls = (const cplx*)matEls; t.fermi = fermi; t.isDiag = isDiag;
  return t;
}

}  // namespace

extern "C" int jvmc_bfo_matels(const int32_t* s, long long B, int N, int numOps, int len, int lDim,
                               const int32_t* idx, const int32_t* map, const double* matEls, const int32_t* fermi,
                               const uint8_t* isDiag, int numDiag, const double* pref, double* mAll,
                               int32_t* choice, int32_t* count, int32_t* maxCount, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!maxCount) return JVMC_ERR_ARG;
  cudaMemsetAsync(maxCount, 0, sizeof(int32_t), st);
  if (B == 0) return JVMC_OK;
  if (!s || !idx || !map || !matEls || !fermi || !isDiag || !pref || !mAll || !choice || !count) return JVMC_ERR_ARG;
  if (len > BFO_MAXLEN || len <= 0 || numOps <= 0 || lDim < 2) return JVMC_ERR_UNSUPPORTED;
  BfoTables t = make_tables(numOps, len, lDim, idx, map, matEls, fermi, isDiag);
  long long tot = B * numOps;
  bfo_matels_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(t, s, B, N, (const cplx*)pref, (cplx*)mAll);
  JVMC_CHECK_LAUNCH();
  bfo_nonzero_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(t, B, numDiag, (cplx*)mAll, choice, count, maxCount);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_bfo_emit(const int32_t* s, long long B, int N, int numOps, int len, int lDim,
                             const int32_t* idx, const int32_t* map, const double* matEls, const int32_t* fermi,
                             const uint8_t* isDiag, const double* mAll, const int32_t* choice, const int32_t* count,
                             int Kmax, int32_t* sp, double* matEl, void* stream) {
  if (B == 0 || Kmax == 0) return JVMC_OK;
  if (!s || !idx || !map || !matEls || !fermi || !mAll || !choice || !count) return JVMC_ERR_ARG;
  if (len > BFO_MAXLEN || len <= 0 || numOps <= 0 || Kmax < 0 || Kmax > numOps) return JVMC_ERR_ARG;
  if (!sp || !matEl) return JVMC_ERR_ARG;
  BfoTables t = make_tables(numOps, len, lDim, idx, map, matEls, fermi, isDiag);
  long long rows = B * Kmax;
  bfo_emit_kernel<<<(unsigned)((rows * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      t, s, B, N, (const cplx*)mAll, choice, count, Kmax, sp, (cplx*)matEl);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_oloc_reduce(const double* matEl, const double* logPsiS, const double* logPsiSP, long long B,
                                int K, double* out, void* stream) {
  if (B == 0) return JVMC_OK;
  if (!logPsiS || !out || B < 0 || K < 0) return JVMC_ERR_ARG;
  if (K > 0 && (!matEl || !logPsiSP)) return JVMC_ERR_ARG;
  oloc_reduce_kernel<<<(unsigned)((B * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      (const cplx*)matEl, (const cplx*)logPsiS, (const cplx*)logPsiSP, B, K, (cplx*)out);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

extern "C" int jvmc_rbm_eloc_bfo(const int32_t* s, const double* tau, long long B, int N, int M,
                                 const double* tables, int numOps, int len, int lDim, const int32_t* idx,
                                 const int32_t* map, const double* matEls, const int32_t* fermi,
                                 const uint8_t* isDiag, int numDiag, const double* pref, double* out,
                                 int* errFlag, void* stream) {
  if (B == 0) return JVMC_OK;
  if (!s || !tau || !tables || !idx || !map || !matEls || !fermi || !isDiag || !pref || !out || !errFlag)
    return JVMC_ERR_ARG;
  if (len > BFO_MAXLEN || len <= 0 || numOps <= 0 || lDim != 2) return JVMC_ERR_UNSUPPORTED;
  BfoTables t = make_tables(numOps, len, lDim, idx, map, matEls, fermi, isDiag);
  const cplx* T = (const cplx*)tables;
  const cplx* lc = T + (size_t)N * M;
  size_t smem = (size_t)EL_WPC * M * sizeof(cplx) + (size_t)EL_WPC * N * sizeof(int32_t);
  if (smem > 227 * 1024) return JVMC_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(rbm_eloc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rbm_eloc_kernel<<<(unsigned)((B + EL_WPC - 1) / EL_WPC), EL_WPC * 32, smem, (cudaStream_t)stream>>>(
      t, s, (const cplx*)tau, B, N, M, T, lc, (const cplx*)pref, numDiag, (cplx*)out, errFlag);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
