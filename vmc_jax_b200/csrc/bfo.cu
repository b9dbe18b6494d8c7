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
#include "bfo_common.cuh"

namespace {

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
// Fused RBM local energy, tiled over samples.  A CTA holds the tau rows of WPC x SPW samples in shared memory;
// every warp owns SPW samples.  Lanes first walk 32 operator strings in parallel (matrix element, flipped sites),
// then the warp evaluates the amplitude ratio of every off-diagonal string with lanes striding over the hidden
// units.  The kernel is bound by the L2 -> SM stream of the weight rows T_a (16 M bytes per string and sample if
// nothing is shared), so
//  * a row loaded into registers is applied to the warp's SPW samples whenever they flip the same sites (always
//    the case for TFIM/Heisenberg-type strings), and
//  * the warps of a CTA are re-aligned at every block of 32 strings, so that they request the same rows at about
//    the same time and share them through L1 (ld.global.nc);
//  * exp(lc_a) is tabulated once per CTA instead of one complex exponential per string and sample.
// tunables (tools/build_variants.py compiles alternatives for A/B timing)
#ifndef JVMC_EL_MAXWPC
#define JVMC_EL_MAXWPC 8        // warps per CTA, one sample per warp
#endif
#ifndef JVMC_EL_MAXW_SPLIT
#define JVMC_EL_MAXW_SPLIT 16   // warps per CTA when WPS warps share a sample
#endif
#ifndef JVMC_EL_MINB
#define JVMC_EL_MINB 2          // minimum CTAs per SM promised to the compiler: 128 registers, 16 warps per SM
#endif                          // (with 1 the compiler takes 212 registers and the kernel is 30 % slower)
#ifndef JVMC_EL_MINB_SPLIT
#define JVMC_EL_MINB_SPLIT 1    // the same for the kernels with several warps per sample (512-thread CTAs)
#endif
#ifndef JVMC_EL_WPS1_MAXM
#define JVMC_EL_WPS1_MAXM 512   // largest M handled by one warp per sample
#endif
constexpr int EL_MAXWPC = JVMC_EL_MAXWPC;

template <bool TWO, int SPW>
__device__ __forceinline__ void flip_ratio(const cplx* __restrict__ T0, const cplx* __restrict__ T1,
                                           const cplx* const* tau, const double* sg0, const double* sg1, int M,
                                           int lane, cplx* p) {
  cplx pa[SPW], pb[SPW];   // two independent product chains per sample
#pragma unroll
  for (int q = 0; q < SPW; ++q) { pa[q] = cmk(1.0, 0.0); pb[q] = cmk(1.0, 0.0); }
  auto step = [&](int j, cplx* acc) {
    const cplx t0 = __ldg(reinterpret_cast<const double2*>(T0 + j));
    cplx t1 = cmk(0.0, 0.0);
    if (TWO) t1 = __ldg(reinterpret_cast<const double2*>(T1 + j));
#pragma unroll
    for (int q = 0; q < SPW; ++q) {
      const cplx tj = tau[q][j];
      cplx f;
      if (!TWO) {
        f = cmk(fma(sg0[q], fma(tj.x, t0.x, -tj.y * t0.y), 1.0), sg0[q] * fma(tj.x, t0.y, tj.y * t0.x));
      } else {
        const cplx ta = cscale(t0, sg0[q]), tb = cscale(t1, sg1[q]);
        f = cadd(cadd(cmk(1.0, 0.0), cmul(ta, tb)), cmul(tj, cadd(ta, tb)));
      }
      acc[q] = cmul(acc[q], f);
    }
  };
  int j = lane;
  for (; j + 32 < M; j += 64) { step(j, pa); step(j + 32, pb); }
  if (j < M) step(j, pa);
#pragma unroll
  for (int q = 0; q < SPW; ++q) p[q] = warp_cprod(cmul(pa[q], pb[q]));
}

// sign * x with sign = +-1 given as a sign-bit mask (0 / 0x80000000): one integer-pipe LOP3 instead of a DMUL
__device__ __forceinline__ double xor_sign(double x, int mask) {
  return __hiloint2double(__double2hiint(x) ^ mask, __double2loint(x));
}

// Same product with the sample's tau row held in registers (lane l owns the hidden units l + 32 k): no shared-memory
// traffic at all in the inner loop, which otherwise binds before the fp64 pipe does (16 B of tau + 16 B of T per
// ten DFMA).  JT = ceil(M / 32) exactly, so only the last k can run past the row: its index is clamped (jl) and its
// tau register is zero, which makes the factor exactly 1 -- no predicates or selects in the unrolled loop.
// With WPS warps per sample the 32-unit chunks are dealt round-robin (chunk c -> warp c % WPS), which keeps the
// "only the last k can be ragged" property; j0 = 32 sw + lane is the lane's first unit, jl the clamped last one.
template <bool TWO, int JT, int WPS>
__device__ __forceinline__ cplx flip_ratio_reg(const cplx* __restrict__ T0, const cplx* __restrict__ T1,
                                               const cplx (&tr)[JT], double sg0, double sg1, int lane, int j0, int jl) {
  cplx pa = cmk(1.0, 0.0), pb = cmk(1.0, 0.0);
  const int m0 = sg0 < 0.0 ? (int)0x80000000 : 0, m1 = sg1 < 0.0 ? (int)0x80000000 : 0;
  const cplx* p0 = T0 + j0;
  const cplx* p1 = TWO ? T1 + j0 : nullptr;
#pragma unroll
  for (int k = 0; k < JT; ++k) {
    const cplx t0 = __ldg(reinterpret_cast<const double2*>(k < JT - 1 ? p0 + 32 * WPS * k : T0 + jl));
    const cplx tj = tr[k];
    const cplx ta = cmk(xor_sign(t0.x, m0), xor_sign(t0.y, m0));   // sg0 * t0
    cplx f;
    if (!TWO) {
      f = cmk(fma(tj.x, ta.x, fma(-tj.y, ta.y, 1.0)), fma(tj.x, ta.y, tj.y * ta.x));
    } else {
      const cplx t1 = __ldg(reinterpret_cast<const double2*>(k < JT - 1 ? p1 + 32 * WPS * k : T1 + jl));
      const cplx tb = cmk(xor_sign(t1.x, m1), xor_sign(t1.y, m1));
      f = cadd(cadd(cmk(1.0, 0.0), cmul(ta, tb)), cmul(tj, cadd(ta, tb)));
      if (k == JT - 1 && j0 + 32 * WPS * k != jl) f = cmk(1.0, 0.0);   // padding lane of the ragged tail
    }
    if (k & 1) pb = cmul(pb, f); else pa = cmul(pa, f);
  }
  return warp_cprod(cmul(pa, pb));
}

// Two single-flip strings (sites a and b) of the same sample at once: four independent product chains hide the
// fp64 latency, and the two warp reductions are folded into one butterfly -- after the first exchange lanes 0-15
// carry string a and lanes 16-31 string b.  Returns, in every lane, the product of the string its half belongs to.
template <int JT, int WPS>
__device__ __forceinline__ cplx flip_ratio_reg_dual(const cplx* __restrict__ Ta, const cplx* __restrict__ Tb,
                                                    const cplx (&tr)[JT], double sga, double sgb, int lane, int j0,
                                                    int jl) {
  cplx pa0 = cmk(1.0, 0.0), pa1 = cmk(1.0, 0.0), pb0 = cmk(1.0, 0.0), pb1 = cmk(1.0, 0.0);
  const int ma = sga < 0.0 ? (int)0x80000000 : 0, mb = sgb < 0.0 ? (int)0x80000000 : 0;
  const cplx* qa = Ta + j0;
  const cplx* qb = Tb + j0;
#pragma unroll
  for (int k = 0; k < JT; ++k) {
    const cplx ta = __ldg(reinterpret_cast<const double2*>(k < JT - 1 ? qa + 32 * WPS * k : Ta + jl));
    const cplx tb = __ldg(reinterpret_cast<const double2*>(k < JT - 1 ? qb + 32 * WPS * k : Tb + jl));
    const cplx tj = tr[k];
    // f = 1 + (sg t) tau_j: the sign goes onto the row entry by an XOR of the sign bits (integer pipe), which leaves
    // 4 fp64 instructions per factor (3 DFMA + 1 DMUL) instead of 6
    const cplx sa = cmk(xor_sign(ta.x, ma), xor_sign(ta.y, ma));
    const cplx sb = cmk(xor_sign(tb.x, mb), xor_sign(tb.y, mb));
    const cplx fa = cmk(fma(tj.x, sa.x, fma(-tj.y, sa.y, 1.0)), fma(tj.x, sa.y, tj.y * sa.x));
    const cplx fb = cmk(fma(tj.x, sb.x, fma(-tj.y, sb.y, 1.0)), fma(tj.x, sb.y, tj.y * sb.x));
    if (k & 1) { pa1 = cmul(pa1, fa); pb1 = cmul(pb1, fb); } else { pa0 = cmul(pa0, fa); pb0 = cmul(pb0, fb); }
  }
  const cplx A = cmul(pa0, pa1), Bv = cmul(pb0, pb1);
  const bool hi = lane >= 16;
  const cplx send = hi ? A : Bv;          // give away the string of the other half ...
  cplx v = hi ? Bv : A;                   // ... keep the own one
  v = cmul(v, cmk(__shfl_xor_sync(0xffffffffu, send.x, 16), __shfl_xor_sync(0xffffffffu, send.y, 16)));
#pragma unroll
  for (int o = 8; o > 0; o >>= 1)
    v = cmul(v, cmk(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o)));
  return v;
}

template <int SPW, int JT, int WPS>
__global__ void __launch_bounds__(WPS > 1 ? JVMC_EL_MAXW_SPLIT * 32 : EL_MAXWPC * 32, WPS > 1 ? JVMC_EL_MINB_SPLIT : JVMC_EL_MINB)
rbm_eloc_kernel(BfoTables t, const int32_t* __restrict__ s, const cplx* __restrict__ tauG, long long B, int N, int M,
                const cplx* __restrict__ T, const cplx* __restrict__ lc, const cplx* __restrict__ pref,
                int numDiag, cplx* __restrict__ out, int* __restrict__ errFlag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  static_assert(JT == 0 || SPW == 1, "register-resident tau: one sample per warp (group)");
  static_assert(WPS == 1 || JT > 0, "several warps per sample only on the register path");
  cplx* tauAll = reinterpret_cast<cplx*>(smem_raw);                 // [wpc * SPW][M], absent when JT > 0
  cplx* elc = tauAll + (JT > 0 ? 0 : (size_t)wpc * SPW * M);
  cplx* xch = elc + N;                                              // [wpc / WPS][2][WPS][2] partial products (WPS > 1)
  int32_t* cfgAll = reinterpret_cast<int32_t*>(xch + (WPS > 1 ? 4 * wpc : 0));
  constexpr int JR = (JT > 0) ? JT : 1;
  cplx tr[JR];
  const int sw = warp % WPS, grp = warp / WPS;                      // slice / sample group of this warp
  const int j0 = 32 * sw + lane;                                    // first hidden unit of the lane (chunks round-robin)
  const int jlast = j0 + 32 * WPS * (JT > 0 ? JT - 1 : 0);
  const int jl = min(jlast, M - 1);                                 // clamped index of the (possibly ragged) last chunk
  unsigned xpar = 0;                                                // parity of the exchange slots
  for (int i = threadIdx.x; i < N; i += blockDim.x) elc[i] = cexp(lc[i]);
  const long long b0 = ((long long)blockIdx.x * (wpc / WPS) + grp) * SPW;
  const cplx* tau[SPW];
  const int32_t* cfg[SPW];
  bool ok[SPW];
#pragma unroll
  for (int q = 0; q < SPW; ++q) {
    ok[q] = b0 + q < B;
    const long long b = ok[q] ? b0 + q : B - 1;   // padding samples repeat the last one and are not written
    cplx* tq = tauAll + (size_t)(warp * SPW + q) * M;
    int32_t* cq = cfgAll + (size_t)(warp * SPW + q) * N;
    if (JT > 0) {
#pragma unroll
      for (int k = 0; k < (JT > 0 ? JT : 1); ++k)
        tr[k] = (j0 + 32 * WPS * k < M) ? tauG[b * M + j0 + 32 * WPS * k] : cmk(0.0, 0.0);
    } else {
      for (int j = lane; j < M; j += 32) tq[j] = tauG[b * M + j];
    }
    for (int i = lane; i < N; i += 32) cq[i] = s[b * N + i];
    tau[q] = tq; cfg[q] = cq;
  }
  __syncthreads();

  // product over the WPS warps of a sample group: slot [parity][warp][which]; one named barrier per exchange
  // (all warps of a group follow the same control flow: same sample, same strings)
  auto group_prod = [&](cplx v, int which) {
    cplx* slot = xch + ((size_t)(grp * 2 + (xpar & 1)) * WPS) * 2;
    xpar ^= 1;
    if ((lane & 15) == 0) slot[sw * 2 + (lane >> 4)] = v;   // lane 0: string a (or the only one), lane 16: string b
    asm volatile("bar.sync %0, %1;\n" ::"r"(1 + grp), "r"(WPS * 32) : "memory");
    cplx r = slot[which];
#pragma unroll
    for (int w = 1; w < WPS; ++w) r = cmul(r, slot[w * 2 + which]);
    return r;
  };
  cplx diagAcc[SPW], eloc[SPW];
  cplx eloc2 = cmk(0.0, 0.0);   // register path: contributions accumulated by lanes 0 and 16 (string pairs)
#pragma unroll
  for (int q = 0; q < SPW; ++q) { diagAcc[q] = cmk(0.0, 0.0); eloc[q] = cmk(0.0, 0.0); }
  for (int o0 = 0; o0 < t.numOps; o0 += 32) {
    if (o0 > 0) __syncthreads();   // re-align the CTA's warps: same rows at the same time -> L1 sharing
    const int o = o0 + lane;
    cplx m[SPW];
    int f0[SPW], f1[SPW];
    unsigned live[SPW], any = 0;
#pragma unroll
    for (int q = 0; q < SPW; ++q) {
      m[q] = cmk(0.0, 0.0); f0[q] = -1; f1[q] = -1;
      bool isd = true;
      if (o < t.numOps) {
        int msite[BFO_MAXLEN], mval[BFO_MAXLEN], nmod;
        m[q] = walk_string(t, o, cfg[q], N, pref[o], msite, mval, nmod);
        isd = t.isDiag[o] != 0;
        int nf = 0;
        for (int k = 0; k < nmod; ++k)
          if (mval[k] != cfg[q][msite[k]]) { if (nf == 0) f0[q] = msite[k]; else if (nf == 1) f1[q] = msite[k]; ++nf; }
        if (nf > 2) atomicExch(errFlag, 1);
      }
      // diagonal strings are merged into one entry (branch_free.py:483-485) ...
      diagAcc[q] = cadd(diagAcc[q], warp_csum((isd && o < t.numOps) ? m[q] : cmk(0.0, 0.0)));
      live[q] = __ballot_sync(0xffffffffu, (o < t.numOps) && !isd && (hypot(m[q].x, m[q].y) > 1e-6));
      any |= live[q];
    }
    // off-diagonal strings: warp-cooperative ratios
    while (any) {
      const int src = __ffs(any) - 1;
      any &= any - 1;
      if (JT > 0 && any) {
        // register path: take two single-flip strings at a time
        const int src2 = __ffs(any) - 1;
        const int aa = __shfl_sync(0xffffffffu, f0[0], src), ab = __shfl_sync(0xffffffffu, f0[0], src2);
        const int a1a = __shfl_sync(0xffffffffu, f1[0], src), a1b = __shfl_sync(0xffffffffu, f1[0], src2);
        if (aa >= 0 && ab >= 0 && a1a < 0 && a1b < 0) {
          any &= any - 1;
          const bool hi = lane >= 16;
          const int mysrc = hi ? src2 : src;
          const cplx mmh = cmk(__shfl_sync(0xffffffffu, m[0].x, mysrc), __shfl_sync(0xffffffffu, m[0].y, mysrc));
          const double sga = cfg[0][aa] ? -1.0 : 1.0, sgb = cfg[0][ab] ? -1.0 : 1.0;
          cplx pr = flip_ratio_reg_dual<JR, WPS>(T + (size_t)aa * M, T + (size_t)ab * M, tr, sga, sgb, lane,
                                                              j0, jl);
          if (WPS > 1) pr = group_prod(pr, hi ? 1 : 0);
          // lanes 0 and 16 (of the group's first warp) carry the two contributions; they are summed at the end
          if ((lane & 15) == 0 && sw == 0) eloc2 = cadd(eloc2, cmul(mmh, cmul(elc[hi ? ab : aa], pr)));
          continue;
        }
      }
      cplx mm[SPW];
      int a0[SPW], a1[SPW];
      bool lv[SPW];
      bool joint = true;
#pragma unroll
      for (int q = 0; q < SPW; ++q) {
        mm[q] = cmk(__shfl_sync(0xffffffffu, m[q].x, src), __shfl_sync(0xffffffffu, m[q].y, src));
        a0[q] = __shfl_sync(0xffffffffu, f0[q], src);
        a1[q] = __shfl_sync(0xffffffffu, f1[q], src);
        lv[q] = (live[q] >> src) & 1u;
        joint = joint && lv[q] && a0[q] == a0[0] && a1[q] == a1[0];
      }
      if (joint && a0[0] >= 0) {
        // all samples of the warp flip the same sites: one pass over the row(s) serves them all
        double sg0[SPW], sg1[SPW];
#pragma unroll
        for (int q = 0; q < SPW; ++q) {
          sg0[q] = cfg[q][a0[0]] ? -1.0 : 1.0;   // -sigma
          sg1[q] = (a1[0] >= 0 && cfg[q][a1[0]]) ? -1.0 : 1.0;
        }
        cplx p[SPW];
        cplx e = elc[a0[0]];
        if (a1[0] < 0) {
          if (JT > 0) {
            p[0] = flip_ratio_reg<false, JR, WPS>(T + (size_t)a0[0] * M, nullptr, tr, sg0[0], sg1[0], lane, j0, jl);
            if (WPS > 1) p[0] = group_prod(p[0], 0);
          }
          else flip_ratio<false, SPW>(T + (size_t)a0[0] * M, nullptr, tau, sg0, sg1, M, lane, p);
        } else {
          if (JT > 0) {
            p[0] = flip_ratio_reg<true, JR, WPS>(T + (size_t)a0[0] * M, T + (size_t)a1[0] * M, tr, sg0[0], sg1[0],
                                                         lane, j0, jl);
            if (WPS > 1) p[0] = group_prod(p[0], 0);
          }
          else flip_ratio<true, SPW>(T + (size_t)a0[0] * M, T + (size_t)a1[0] * M, tau, sg0, sg1, M, lane, p);
          e = cmul(e, elc[a1[0]]);
        }
#pragma unroll
        for (int q = 0; q < SPW; ++q) eloc[q] = cadd(eloc[q], cmul(mm[q], cmul(e, p[q])));
      } else {
#pragma unroll
        for (int q = 0; q < SPW; ++q) {
          if (!lv[q]) continue;
          cplx ratio = cmk(1.0, 0.0);
          if (a0[q] >= 0) {
            const double g0 = cfg[q][a0[q]] ? -1.0 : 1.0;
            const double g1 = (a1[q] >= 0 && cfg[q][a1[q]]) ? -1.0 : 1.0;
            const cplx* tq = tau[q];
            cplx p1;
            cplx e = elc[a0[q]];
            if (JT > 0) {
              p1 = cmk(1.0, 0.0);   // unreachable: with one sample per warp a live flipping string takes the joint path
            } else if (a1[q] < 0) {
              flip_ratio<false, 1>(T + (size_t)a0[q] * M, nullptr, &tq, &g0, &g1, M, lane, &p1);
            } else {
              flip_ratio<true, 1>(T + (size_t)a0[q] * M, T + (size_t)a1[q] * M, &tq, &g0, &g1, M, lane, &p1);
              e = cmul(e, elc[a1[q]]);
            }
            ratio = cmul(e, p1);
          }
          eloc[q] = cadd(eloc[q], cmul(mm[q], ratio));
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < SPW; ++q) {
    // ... and that entry passes the same |m| > 1e-6 filter as every other one (base.py:93-103)
    cplx d = diagAcc[q];
    if (hypot(d.x, d.y) <= 1e-6) d = cmk(0.0, 0.0);
    if (JT > 0) {
      const cplx e16 = cmk(__shfl_sync(0xffffffffu, eloc2.x, 16), __shfl_sync(0xffffffffu, eloc2.y, 16));
      eloc[q] = cadd(eloc[q], cadd(eloc2, e16));
    }
    if (lane == 0 && sw == 0 && ok[q]) out[b0 + q] = cadd(eloc[q], d);
  }
}

template <int SPW, int JT, int WPS>
int launch_eloc(BfoTables t, const int32_t* s, const cplx* tau, long long B, int N, int M, const cplx* T, const cplx* lc,
                const cplx* pref, int numDiag, cplx* out, int* errFlag, cudaStream_t st) {
  int wpc;
  size_t smem;
  if (JT > 0) {
    // register path: shared memory only holds the configurations (per warp), exp(lc) and the exchange slots
    const int maxw = WPS > 1 ? JVMC_EL_MAXW_SPLIT : EL_MAXWPC;
    long long groups = maxw / WPS;
    if (B < groups) groups = B;
    wpc = (int)groups * WPS;
    smem = (size_t)N * sizeof(cplx) + (WPS > 1 ? (size_t)4 * wpc * sizeof(cplx) : 0) + (size_t)wpc * N * sizeof(int32_t);
    if (smem > 227 * 1024) return JVMC_ERR_UNSUPPORTED;
  } else {
    const size_t perSample = (size_t)M * sizeof(cplx) + (size_t)N * sizeof(int32_t);
    const size_t fixed = (size_t)N * sizeof(cplx);
    // as many samples per CTA as fit ~110 KB (two CTAs per SM); one big CTA when that leaves fewer than 4 warps
    auto warpsFor = [&](size_t budget) {
      long long w = ((long long)budget - (long long)fixed) / (long long)(SPW * perSample);
      return (int)(w < 0 ? 0 : (w > EL_MAXWPC ? EL_MAXWPC : w));
    };
    wpc = warpsFor(110 * 1024);
    if (wpc < 4) wpc = warpsFor(227 * 1024);
    if (wpc < 1) return JVMC_ERR_UNSUPPORTED;
    if (B < (long long)wpc * SPW) wpc = (int)((B + SPW - 1) / SPW);
    smem = fixed + (size_t)wpc * SPW * perSample;
  }
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(rbm_eloc_kernel<SPW, JT, WPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const long long perCta = (long long)(wpc / WPS) * SPW;
  const long long ctas = (B + perCta - 1) / perCta;
  rbm_eloc_kernel<SPW, JT, WPS><<<(unsigned)ctas, wpc * 32, smem, st>>>(t, s, tau, B, N, M, T, lc, pref, numDiag, out,
                                                                     errFlag);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
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
#define JVMC_ELOC(SPW, JT, WPS) launch_eloc<SPW, JT, WPS>(t, s, (const cplx*)tau, B, N, M, T, lc, (const cplx*)pref, \
                                                         numDiag, (cplx*)out, errFlag, (cudaStream_t)stream)
#define JVMC_ELOC_LO(WPS)                                                                                          \
  switch ((M + 32 * WPS - 1) / (32 * WPS)) {                                                                       \
    case 1: return JVMC_ELOC(1, 1, WPS);   case 2: return JVMC_ELOC(1, 2, WPS);   case 3: return JVMC_ELOC(1, 3, WPS);   \
    case 4: return JVMC_ELOC(1, 4, WPS);   case 5: return JVMC_ELOC(1, 5, WPS);   case 6: return JVMC_ELOC(1, 6, WPS);   \
    case 7: return JVMC_ELOC(1, 7, WPS);   case 8: return JVMC_ELOC(1, 8, WPS);                                          \
    default: break;                                                                                                \
  }
#define JVMC_ELOC_JT(WPS)                                                                                          \
  switch ((M + 32 * WPS - 1) / (32 * WPS)) {                                                                       \
    case 9: return JVMC_ELOC(1, 9, WPS);                                                                           \
    case 10: return JVMC_ELOC(1, 10, WPS); case 11: return JVMC_ELOC(1, 11, WPS); case 12: return JVMC_ELOC(1, 12, WPS); \
    case 13: return JVMC_ELOC(1, 13, WPS); case 14: return JVMC_ELOC(1, 14, WPS); case 15: return JVMC_ELOC(1, 15, WPS); \
    case 16: return JVMC_ELOC(1, 16, WPS);                                                                         \
    default: break;                                                                                                \
  }
  // tau in registers: one warp per sample up to M = 512, then 2 / 4 / 8 warps per sample (M <= 4096); the weight
  // rows are shared through L1 by the warps of a CTA
  if (M <= 256) { JVMC_ELOC_LO(1) }
  else if (M <= JVMC_EL_WPS1_MAXM) { JVMC_ELOC_JT(1) }
  else if (M <= 512) { JVMC_ELOC_LO(2) }
  else if (M <= 1024) { JVMC_ELOC_JT(2) }
  else if (M <= 2048) { JVMC_ELOC_JT(4) }
  else if (M <= 4096) { JVMC_ELOC_JT(8) }
  // larger M: tau in shared memory, two samples per warp share every weight row in registers
  int rc = JVMC_ELOC(2, 0, 1);
  if (rc == JVMC_ERR_UNSUPPORTED) rc = JVMC_ELOC(1, 0, 1);
  return rc;
#undef JVMC_ELOC_JT
#undef JVMC_ELOC_LO
#undef JVMC_ELOC
}
