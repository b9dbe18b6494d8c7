// jvmc_rbm_mcmc -- parallel-chain Metropolis sampler for (Cpx)RBM wave functions.
//
// Replaces MCSampler._get_samples / _sweep.perform_mc_update (reference jVMC/sampler.py:301-356)
// and the proposers propose_spin_flip / _Z2 / _zeroMag (jVMC/sampler.py:15-19,30-39,42-64).
//
// Two kernels: rbm_mcmc_flip_kernel (below; single-flip proposers, the hot path) and the generic
// rbm_mcmc_kernel (exchange proposer, anything else).  Generic kernel: one warp per Markov chain, the chain
// keeps tau_j = tanh(theta_j) (theta = sigma W + b) in shared
// memory and the configuration as a bit mask in registers (lane l owns sites 32l..32l+31).  A
// proposal that flips sites a (and b) changes theta by Delta = -2 sigma_a W_a (- 2 sigma_b W_b) and
//    psi(s')/psi(s) = prod_a cosh-prod(a) * prod_j [ d_j + tau_j n_j ],
//    one flip : n = t_a, d = 1;   two flips: n = t_a + t_b, d = 1 + t_a t_b,  t_a = -sigma_a tanh(2W_aj)
// (tanh/cosh addition theorems), so a proposal costs M complex multiply-adds instead of a full
// N*M forward pass; tanh(2W) and log prod_j cosh(2W_ij) are per-parameter tables (jvmc_rbm_tables).
// Accepted moves update tau rationally, tau' = (tau d + n)/(d + tau n); tau is re-derived from
// theta = sigma W + b every `refreshEvery` sweeps to stop round-off drift.
// Global Z2 flip (prob. 1/5 in the _Z2/_zeroMag proposers): theta -> -theta + 2b.
//
// RNG: Philox4x32-10 keyed by the seed, counter = (Metropolis step, global chain id, purpose):
// results are independent of how chains are distributed over GPUs.  The reference's threefry
// stream is not reproduced (parity is distributional).
#include "common.cuh"

namespace {

constexpr int MC_WPC = 4;  // warps (= chains) per CTA
// largest M handled by 1 / 2 / 4 warps per chain (tunables, tools/build_variants.py)
#ifndef JVMC_MC_MAXM1
#define JVMC_MC_MAXM1 512
#endif
#ifndef JVMC_MC_MAXM2
#define JVMC_MC_MAXM2 1024
#endif
#ifndef JVMC_MC_MAXM4
#define JVMC_MC_MAXM4 2048
#endif

struct McmcArgs {
  int32_t* states;
  long long C;
  int N, M;
  const cplx* W;
  const cplx* bias;
  const cplx* T;
  const cplx* lc;
  const cplx* tb2;
  const cplx* lcb;
  unsigned long long seed, step0;
  long long chain0;
  int proposer;
  double mu;
  int K;
  long long thermSteps;
  int numSamples;
  int refreshEvery;
  int32_t* out;
  unsigned long long* counters;
};

__device__ __forceinline__ int nth_set_bit(uint32_t w, int n) {  // n >= 1, position of n-th set bit
  for (int k = 1; k < n; ++k) w &= w - 1;
  return __ffs(w) - 1;
}

// site index of the r-th (1-based) set bit of the warp-distributed mask; -1 if fewer than r bits set
__device__ __forceinline__ int warp_select(uint32_t word, int r, int lane) {
  int cnt = __popc(word);
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  int exc = inc - cnt;
  bool mine = (r > exc) && (r <= inc);
  int pos = mine ? (lane * 32 + nth_set_bit(word, r - exc)) : -1;
  unsigned who = __ballot_sync(0xffffffffu, mine);
  if (who == 0) return -1;
  return __shfl_sync(0xffffffffu, pos, __ffs(who) - 1);
}

__global__ void __launch_bounds__(MC_WPC * 32)
rbm_mcmc_kernel(McmcArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long chain = (long long)blockIdx.x * MC_WPC + warp;
  if (chain >= a.C) return;  // whole warp exits together
  const int N = a.N, M = a.M;
  cplx* tau = reinterpret_cast<cplx*>(smem_raw) + (size_t)warp * M;
  uint32_t* sbits = reinterpret_cast<uint32_t*>(reinterpret_cast<cplx*>(smem_raw) + (size_t)MC_WPC * M) + warp * 32;
  const bool hasBias = a.bias != nullptr;
  const unsigned long long gchain = (unsigned long long)(a.chain0 + chain);
  const Philox rng(a.seed);

  // load configuration bits
  uint32_t valid = 0, bits = 0;
  {
    int base = lane * 32;
    for (int k = 0; k < 32; ++k) {
      int i = base + k;
      if (i < N) {
        valid |= 1u << k;
        if (a.states[chain * N + i] != 0) bits |= 1u << k;
      }
    }
  }

  auto refresh = [&]() {
    sbits[lane] = bits;
    __syncwarp();
    for (int j0 = 0; j0 < M; j0 += 32) {
      int j = j0 + lane;
      if (j < M) {
        cplx acc = hasBias ? a.bias[j] : cmk(0.0, 0.0);
        for (int i = 0; i < N; ++i) {
          double sg = ((sbits[i >> 5] >> (i & 31)) & 1u) ? 1.0 : -1.0;
          cplx w = a.W[(size_t)i * M + j];
          acc.x = fma(sg, w.x, acc.x);
          acc.y = fma(sg, w.y, acc.y);
        }
        cplx l, t;
        lncosh_tanh(acc, l, t);
        tau[j] = t;
      }
    }
    __syncwarp();
  };

  unsigned long long nAcc = 0, nProp = 0;
  const long long total = a.thermSteps + (long long)a.numSamples * a.K;
  long long nextEmit = a.thermSteps + a.K;
  int emitted = 0;
  int sweepCtr = 0;
  refresh();

  for (long long st = 0; st < total; ++st) {
    const unsigned long long gs = a.step0 + (unsigned long long)st;
    uint4 r = rng((uint32_t)gs, (uint32_t)(gs >> 32), (uint32_t)gchain, (uint32_t)(gchain >> 32) << 8);
    int sa = -1, sb = -1;  // flipped sites
    bool g = false;
    if (a.proposer == 2) {
      int half = N / 2;
      int ru = 1 + (int)__umulhi(r.x, (uint32_t)half);
      int rd = 1 + (int)__umulhi(r.y, (uint32_t)half);
      sa = warp_select(bits, ru, lane);
      sb = warp_select((~bits) & valid, rd, lane);
      if (sa < 0) { sa = sb; sb = -1; }
      uint4 r2 = rng((uint32_t)gs, (uint32_t)(gs >> 32), (uint32_t)gchain, ((uint32_t)(gchain >> 32) << 8) | 1u);
      g = __umulhi(r2.x, 5u) == 0u;
    } else {
      sa = (int)__umulhi(r.x, (uint32_t)N);
      g = (a.proposer == 1) && (__umulhi(r.y, 5u) == 0u);
    }
    const bool gb = g && hasBias;
    double sga = 0.0, sgb = 0.0;  // -sigma of the flipped sites
    double lre = 0.0;
    const cplx* Ta = a.T;
    const cplx* Tb = a.T;
    if (sa >= 0) {
      uint32_t w = __shfl_sync(0xffffffffu, bits, sa >> 5);
      sga = ((w >> (sa & 31)) & 1u) ? -1.0 : 1.0;
      Ta = a.T + (size_t)sa * M;
      lre += a.lc[sa].x;
    }
    if (sb >= 0) {
      uint32_t w = __shfl_sync(0xffffffffu, bits, sb >> 5);
      sgb = ((w >> (sb & 31)) & 1u) ? -1.0 : 1.0;
      Tb = a.T + (size_t)sb * M;
      lre += a.lc[sb].x;
    }
    if (gb) lre += a.lcb[0].x;

    double prod = 1.0;
    if (sb < 0 && !gb) {
      // hot path: single flip, f = 1 + tau * t_a
      if (sa >= 0) {
        for (int j = lane; j < M; j += 32) {
          cplx t = Ta[j];
          cplx tj = tau[j];
          double fr = fma(sga, fma(tj.x, t.x, -tj.y * t.y), 1.0);
          double fi = sga * fma(tj.x, t.y, tj.y * t.x);
          prod *= fma(fr, fr, fi * fi);
        }
      }
    } else {
      for (int j = lane; j < M; j += 32) {
        cplx ta = cscale(Ta[j], sga);
        cplx n = ta, d = cmk(1.0, 0.0);
        if (sb >= 0) {
          cplx tb = cscale(Tb[j], sgb);
          n = cadd(ta, tb);
          d = cadd(d, cmul(ta, tb));
        }
        cplx tj = tau[j];
        cplx f = cadd(d, cmul(tj, n));
        if (gb) {
          cplx t1 = cdiv(cadd(cmul(tj, d), n), f);
          cplx f2 = csub(cmk(1.0, 0.0), cmul(t1, a.tb2[j]));
          f = cmul(f, f2);
        }
        prod *= cabs2(f);
      }
    }
    prod = warp_prod(prod);
    double P;
    if (a.mu == 2.0) P = exp(2.0 * lre) * prod;
    else P = exp(a.mu * (lre + 0.5 * log(prod)));
    const double u = u01_from_bits(r.z, r.w);
    const bool accept = u < P;
    nProp += 1;
    if (accept) {
      nAcc += 1;
      for (int j = lane; j < M; j += 32) {
        cplx tj = tau[j];
        cplx tn = tj;
        if (sa >= 0) {
          cplx ta = cscale(Ta[j], sga);
          cplx n = ta, d = cmk(1.0, 0.0);
          if (sb >= 0) {
            cplx tb = cscale(Tb[j], sgb);
            n = cadd(ta, tb);
            d = cadd(d, cmul(ta, tb));
          }
          tn = cdiv(cadd(cmul(tj, d), n), cadd(d, cmul(tj, n)));
        }
        if (g) {
          if (hasBias) {
            cplx b2 = a.tb2[j];
            tn = cdiv(csub(b2, tn), csub(cmk(1.0, 0.0), cmul(b2, tn)));
          } else {
            tn = cmk(-tn.x, -tn.y);
          }
        }
        tau[j] = tn;
      }
      if (sa >= 0 && (sa >> 5) == lane) bits ^= 1u << (sa & 31);
      if (sb >= 0 && (sb >> 5) == lane) bits ^= 1u << (sb & 31);
      if (g) bits = (~bits) & valid;
      __syncwarp();
    }
    // sweep bookkeeping
    if (st + 1 == nextEmit) {
      const long long row = (long long)emitted * a.C + chain;  // time-major, chain-minor (sampler.py:323)
      int32_t* dst = a.out + row * N;
      for (int w = 0; w * 32 < N; ++w) {
        uint32_t word = __shfl_sync(0xffffffffu, bits, w);
        int i = w * 32 + lane;
        if (i < N) dst[i] = (int32_t)((word >> lane) & 1u);
      }
      ++emitted;
      nextEmit += a.K;
    }
    if ((st + 1) % a.K == 0) {
      if (++sweepCtr >= a.refreshEvery && st + 1 < total) { sweepCtr = 0; refresh(); }
    }
  }
  // persist chain state and counters
  for (int w = 0; w * 32 < N; ++w) {
    uint32_t word = __shfl_sync(0xffffffffu, bits, w);
    int i = w * 32 + lane;
    if (i < N) a.states[chain * N + i] = (int32_t)((word >> lane) & 1u);
  }
  if (lane == 0) {
    atomicAdd(a.counters + 0, nProp);
    atomicAdd(a.counters + 1, nAcc);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Fast path for the single-spin-flip proposers (propose_spin_flip, propose_spin_flip_Z2) and M <= 32 JT:
//  * the row T_a of the NEXT proposal is copied global -> shared with cp.async while the current proposal is
//    evaluated (the site of step st+1 only depends on the counter RNG, not on the accept decision), so the L2
//    latency of the weight-row stream is off the per-step critical path;
//  * tau lives in registers (JT complex numbers per lane), the rows in two ping-pong shared buffers, the loop over
//    hidden units is fully unrolled (JT = ceil(M / 32 / WPCH) exactly, zero padding instead of predicates);
//  * Philox is evaluated once per 32 steps (lane l computes step base + l) and broadcast by shuffles;
//  * exp(mu Re lc_a) comes from a per-CTA shared table instead of an exp per proposal.
// Same RNG counters as the generic kernel: the accept/reject sequence is identical up to floating-point rounding.
// WPCH > 1 (large M): one chain per CTA, its hidden units split over WPCH warps (slices of 32 JT units); the
// partial products meet in shared memory (one __syncthreads per Metropolis step, double-buffered slots) and every
// warp takes the same accept decision from the same RNG counters.
// Code size is what this kernel has to watch: its Metropolis loop must stay inside the 32 KB L1.5 instruction cache
// (ncu: 14 % of the warp samples were "no instruction" stalls with two inlined copies of the 13-times unrolled
// refresh and IEEE divisions in the accept path -- 17.7 k SASS instructions).
//  * the log-cosh / tanh evaluation of the (rare) refresh is an out-of-line call;
//  * the accept path divides by Newton iterations on the hardware reciprocal seed (MUFU.RCP64H): 1 + tau n is
//    O(1) there, so the IEEE slow paths (denormals, overflow) that `1.0 / x` carries along are dead weight.
// tau_j = tanh(b_j + sum_i sigma_i W_ij) of one hidden unit from the bit-packed configuration in shared memory
__device__ __noinline__ void refresh_unit(const cplx* __restrict__ W, const cplx* __restrict__ bias,
                                          const uint32_t* sbits, int N, int M, int j, double* out2) {
  cplx acc = bias ? bias[j] : cmk(0.0, 0.0);
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    const double sg = ((sbits[i >> 5] >> (i & 31)) & 1u) ? 1.0 : -1.0;
    const cplx w = W[(size_t)i * M + j];
    acc.x = fma(sg, w.x, acc.x);
    acc.y = fma(sg, w.y, acc.y);
  }
  cplx l, t;
  lncosh_tanh(acc, l, t);
  out2[0] = t.x; out2[1] = t.y;
}
__device__ __forceinline__ double rcp_newton(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));   // ~2^-20 relative
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);                                    // third step: below 1 ulp whatever the seed quality
  return fma(r, e, r);
}
__device__ __forceinline__ cplx cdiv_fast(cplx a, cplx b) {
  const double inv = rcp_newton(cabs2(b));
  return cmk(fma(a.x, b.x, a.y * b.y) * inv, fma(a.y, b.x, -a.x * b.y) * inv);
}

template <int JT, int WPCH>
__global__ void __launch_bounds__(WPCH == 1 ? MC_WPC * 32 : WPCH * 32)
rbm_mcmc_flip_kernel(McmcArgs a) {
  constexpr int CPB = (WPCH == 1) ? MC_WPC : 1;    // chains per CTA
  constexpr int NW = CPB * WPCH;                   // warps per CTA
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cw = (WPCH == 1) ? warp : 0;           // chain of this warp within the CTA
  const int sw = (WPCH == 1) ? 0 : warp;           // slice of the hidden units this warp owns
  const long long chain = (long long)blockIdx.x * CPB + cw;
  const int N = a.N, M = a.M;
  const int jbase = sw * 32 * JT;                  // first hidden unit of the slice
  // tau lives in REGISTERS (lane l owns the units jbase + l + 32 k); the weight rows are staged in two shared
  // buffers per warp (ping-pong: the row of step st+1 lands while step st is evaluated).  Buffers are padded to
  // 32 JT entries and the padding is kept at zero, tau of a padding unit is zero as well: its factor is exactly 1
  // and its update 0 -> 0, so the unrolled loops need no predicates.  Shared-memory traffic per hidden unit and
  // proposal: 16 B written by cp.async + 16 B read -- the L1/shared pipe is what binds this kernel.
  cplx* bufAll = reinterpret_cast<cplx*>(smem_raw);            // [NW][2][32 JT]
  double* elc = reinterpret_cast<double*>(bufAll + (size_t)NW * 2 * JT * 32);
  double* part = elc + N;                          // [2][WPCH] partial products (WPCH > 1)
  uint32_t* sbitsAll = reinterpret_cast<uint32_t*>(part + 2 * WPCH);
  // two mbarriers per warp: "row of the even / odd step has landed" (TMA bulk copy, complete_tx)
  uint64_t* barAll = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sbitsAll + NW * 32) + 7) & ~(uintptr_t)7);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((unsigned)__cvta_generic_to_shared(barAll + 2 * warp)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((unsigned)__cvta_generic_to_shared(barAll + 2 * warp + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = threadIdx.x; i < N; i += NW * 32) elc[i] = exp(a.mu * a.lc[i].x);
  for (int i = threadIdx.x; i < 2 * NW * JT * 32; i += NW * 32) bufAll[i] = cmk(0.0, 0.0);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // zero fill (generic proxy) before the TMA writes
  __syncthreads();
  if (chain >= a.C) return;  // WPCH == 1 only (whole warp exits together; no further block-wide barriers there)
  cplx* buf0 = bufAll + (size_t)warp * 2 * JT * 32;
  uint32_t* sbits = sbitsAll + warp * 32;
  const bool hasBias = a.bias != nullptr;
  const unsigned long long gchain = (unsigned long long)(a.chain0 + chain);
  const Philox rng(a.seed);
  const uint32_t c2 = (uint32_t)gchain, c3 = (uint32_t)(gchain >> 32) << 8;

  uint32_t valid = 0, bits = 0;
  {
    int base = lane * 32;
    for (int k = 0; k < 32; ++k) {
      int i = base + k;
      if (i < N) {
        valid |= 1u << k;
        if (a.states[chain * N + i] != 0) bits |= 1u << k;
      }
    }
  }
  cplx tau[JT];
  auto refresh = [&]() {
    sbits[lane] = bits;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < JT; ++k) {
      const int j = jbase + lane + 32 * k;
      cplx t = cmk(0.0, 0.0);
      if (j < M) {
        double o2[2];
        refresh_unit(a.W, a.bias, sbits, N, M, j, o2);
        t = cmk(o2[0], o2[1]);
      }
      tau[k] = t;
    }
    __syncwarp();
  };
  // stage this warp's slice of row T_site into buffer `which`: ONE bulk copy by the TMA engine (cp.async.bulk, SASS
  // UBLKCP) that completes on the buffer's mbarrier -- no per-lane copy instructions and no L1 involvement (the
  // per-lane cp.async version kept the L1/shared pipe at 82 %).  The elements beyond M stay zero.
  const unsigned rowBytes = (unsigned)(min(32 * JT, M - jbase) * (int)sizeof(cplx));
  const unsigned barAddr = (unsigned)__cvta_generic_to_shared(barAll + 2 * warp);
  auto prefetch_row = [&](int site, int which) {
    __syncwarp();                                    // every lane is done reading this buffer (step st - 1)
    if (lane == 0) {
      const cplx* src = a.T + (size_t)site * M + jbase;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(buf0 + which * JT * 32);
      const unsigned bar = barAddr + 8u * (unsigned)which;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(rowBytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                   "l"(src), "r"(rowBytes), "r"(bar) : "memory");
    }
  };
  auto wait_row = [&](int which, unsigned parity) {
    const unsigned bar = barAddr + 8u * (unsigned)which;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MC_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni MC_WAIT_DONE;\n"
        "bra.uni MC_WAIT_LOOP;\n"
        "MC_WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
  };
  // product over all hidden units of the chain from this warp's partial product
  auto chain_prod = [&](double p, long long st) {
    p = warp_prod(p);
    if (WPCH > 1) {
      double* slot = part + (st & 1) * WPCH;
      if (lane == 0) slot[sw] = p;
      __syncthreads();
      p = 1.0;
#pragma unroll
      for (int w = 0; w < WPCH; ++w) p *= slot[w];   // same order in every warp: bitwise identical decisions
    }
    return p;
  };

  unsigned long long nAcc = 0, nProp = 0;
  const long long total = a.thermSteps + (long long)a.numSamples * a.K;
  long long nextEmit = a.thermSteps + a.K;
  int emitted = 0, sweepCtr = 0, untilSweep = a.K;
  refresh();

  // RNG batches: q holds the Philox output of step (batch base + lane); batches are aligned on the LOCAL step
  // index, the counter of step st is step0 + st exactly as in the generic kernel.
  uint4 q = make_uint4(0, 0, 0, 0);
  auto draw = [&](long long st, uint4& r) {   // r of step st (warp-uniform); refills the batch every 32 steps
    const int idx = (int)(st & 31);
    if (idx == 0) {
      const unsigned long long gs = a.step0 + (unsigned long long)st + (unsigned long long)lane;
      q = rng((uint32_t)gs, (uint32_t)(gs >> 32), c2, c3);
    }
    r.x = __shfl_sync(0xffffffffu, q.x, idx);
    r.y = __shfl_sync(0xffffffffu, q.y, idx);
    r.z = __shfl_sync(0xffffffffu, q.z, idx);
    r.w = __shfl_sync(0xffffffffu, q.w, idx);
  };
  uint4 rc;
  if (total > 0) {
    draw(0, rc);
    prefetch_row((int)__umulhi(rc.x, (uint32_t)N), 0);
  }
  for (long long st = 0; st < total; ++st) {
    const int sa = (int)__umulhi(rc.x, (uint32_t)N);
    const bool g = (a.proposer == 1) && (__umulhi(rc.y, 5u) == 0u);
    const double u = u01_from_bits(rc.z, rc.w);
    const cplx* tv = buf0 + (int)(st & 1) * JT * 32 + lane;   // row of this step; the other buffer receives the next one
    wait_row((int)(st & 1), (unsigned)((st >> 1) & 1));   // buffer st & 1 is used every second step: phase st / 2
    uint4 rn = rc;
    if (st + 1 < total) {
      draw(st + 1, rn);
      prefetch_row((int)__umulhi(rn.x, (uint32_t)N), (int)((st + 1) & 1));
    }
    const uint32_t wsa = __shfl_sync(0xffffffffu, bits, sa >> 5);
    const double sga = ((wsa >> (sa & 31)) & 1u) ? -1.0 : 1.0;   // -sigma_a
    const bool gb = g && hasBias;
    bool accept;
    if (!gb) {
      // |1 + sga tau t|^2 = (1 + sga Re)^2 + Im^2, two independent partial products
      double p0 = 1.0, p1 = 1.0;
      const int smask = sga < 0.0 ? (int)0x80000000 : 0;
#pragma unroll
      for (int k = 0; k < JT; ++k) {
        const cplx tj = tau[k];
        const cplx t = tv[32 * k];
        // 1 + sga Re(tau t): the sign is XORed into the row entry (integer pipe), 2 DFMA instead of DMUL + 2 DFMA
        const double tsx = __hiloint2double(__double2hiint(t.x) ^ smask, __double2loint(t.x));
        const double tsy = __hiloint2double(__double2hiint(t.y) ^ smask, __double2loint(t.y));
        const double fr = fma(tj.x, tsx, fma(-tj.y, tsy, 1.0));
        const double im = fma(tj.x, t.y, tj.y * t.x);
        const double f2 = fma(fr, fr, im * im);
        if (k & 1) p1 *= f2; else p0 *= f2;
      }
      const double prod = chain_prod(p0 * p1, st);
      const double P = (a.mu == 2.0) ? elc[sa] * prod : elc[sa] * pow(prod, 0.5 * a.mu);
      accept = u < P;
      nProp += 1;
      if (accept) {
        nAcc += 1;
        const double gsn = g ? -1.0 : 1.0;   // Z2 flip without bias: tau -> -tau
#pragma unroll
        for (int k = 0; k < JT; ++k) {
          const cplx tj = tau[k];
          const cplx n = cscale(tv[32 * k], sga);
          const cplx tn = cdiv_fast(cadd(tj, n), cadd(cmk(1.0, 0.0), cmul(tj, n)));
          tau[k] = cscale(tn, gsn);
        }
      }
    } else {
      // global flip with a bias (probability 1/5 of the Z2 proposer's steps): theta -> -theta' + 2b
      double prod = 1.0;
#pragma unroll
      for (int k = 0; k < JT; ++k) {
        const int j = jbase + lane + 32 * k;
        if (j < M) {
          const cplx tj = tau[k];
          const cplx n = cscale(tv[32 * k], sga);
          cplx f = cadd(cmk(1.0, 0.0), cmul(tj, n));
          const cplx t1 = cdiv_fast(cadd(tj, n), f);
          f = cmul(f, csub(cmk(1.0, 0.0), cmul(t1, a.tb2[j])));
          prod *= cabs2(f);
        }
      }
      prod = chain_prod(prod, st);
      const double lre = a.lc[sa].x + a.lcb[0].x;
      const double P = (a.mu == 2.0) ? exp(2.0 * lre) * prod : exp(a.mu * (lre + 0.5 * log(prod)));
      accept = u < P;
      nProp += 1;
      if (accept) {
        nAcc += 1;
#pragma unroll
        for (int k = 0; k < JT; ++k) {
          const int j = jbase + lane + 32 * k;
          if (j < M) {
            const cplx tj = tau[k];
            const cplx n = cscale(tv[32 * k], sga);
            cplx tn = cdiv_fast(cadd(tj, n), cadd(cmk(1.0, 0.0), cmul(tj, n)));
            const cplx b2 = a.tb2[j];
            tn = cdiv_fast(csub(b2, tn), csub(cmk(1.0, 0.0), cmul(b2, tn)));
            tau[k] = tn;
          }
        }
      }
    }
    if (accept) {
      if ((sa >> 5) == lane) bits ^= 1u << (sa & 31);
      if (g) bits = (~bits) & valid;
    }
    if (st + 1 == nextEmit) {
      if (sw == 0) {
        const long long row = (long long)emitted * a.C + chain;  // time-major, chain-minor (sampler.py:323)
        int32_t* dst = a.out + row * N;
        for (int w = 0; w * 32 < N; ++w) {
          uint32_t word = __shfl_sync(0xffffffffu, bits, w);
          int i = w * 32 + lane;
          if (i < N) dst[i] = (int32_t)((word >> lane) & 1u);
        }
      }
      ++emitted;
      nextEmit += a.K;
    }
    if (--untilSweep == 0) {
      untilSweep = a.K;
      if (++sweepCtr >= a.refreshEvery && st + 1 < total) { sweepCtr = 0; refresh(); }
    }
    rc = rn;
  }
  if (sw == 0) {
    for (int w = 0; w * 32 < N; ++w) {
      uint32_t word = __shfl_sync(0xffffffffu, bits, w);
      int i = w * 32 + lane;
      if (i < N) a.states[chain * N + i] = (int32_t)((word >> lane) & 1u);
    }
    if (lane == 0) {
      atomicAdd(a.counters + 0, nProp);
      atomicAdd(a.counters + 1, nAcc);
    }
  }
}

template <int JT, int WPCH>
int launch_flip(const McmcArgs& a, cudaStream_t stream) {
  constexpr int CPB = (WPCH == 1) ? MC_WPC : 1;
  constexpr int NW = CPB * WPCH;
  size_t smem = (size_t)2 * NW * JT * 32 * sizeof(cplx) + (size_t)(a.N + 2 * WPCH) * sizeof(double) + NW * 32 * sizeof(uint32_t) +
                8 + (size_t)2 * NW * sizeof(uint64_t);
  if (smem > 227 * 1024) return JVMC_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(rbm_mcmc_flip_kernel<JT, WPCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  unsigned grid = (unsigned)((a.C + CPB - 1) / CPB);
  rbm_mcmc_flip_kernel<JT, WPCH><<<grid, NW * 32, smem, stream>>>(a);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

// JT = ceil(M / (32 WPCH)) in 1..16; with several warps per chain only 9..16 occur (M > 256 WPCH)
template <int WPCH>
int dispatch_flip(const McmcArgs& a, cudaStream_t stream) {
  const int jt = (a.M + 32 * WPCH - 1) / (32 * WPCH);
#define JVMC_FLIP(J) case J: return launch_flip<J, WPCH>(a, stream)
  if constexpr (WPCH == 1) {
    switch (jt) {
      JVMC_FLIP(1); JVMC_FLIP(2); JVMC_FLIP(3); JVMC_FLIP(4); JVMC_FLIP(5); JVMC_FLIP(6); JVMC_FLIP(7); JVMC_FLIP(8);
      default: break;
    }
  }
#ifdef JVMC_MC_LOWJT   // variants that split smaller M over several warps (tools/build_variants.py)
  else {
    switch (jt) {
      JVMC_FLIP(5); JVMC_FLIP(6); JVMC_FLIP(7); JVMC_FLIP(8);
      default: break;
    }
  }
#endif
  switch (jt) {
    JVMC_FLIP(9); JVMC_FLIP(10); JVMC_FLIP(11); JVMC_FLIP(12); JVMC_FLIP(13); JVMC_FLIP(14); JVMC_FLIP(15); JVMC_FLIP(16);
    default: return JVMC_ERR_UNSUPPORTED;
  }
#undef JVMC_FLIP
}

}  // namespace

static int g_mcmc_generic = 0;
// development knob: 1 forces the generic kernel (A/B comparisons of the fast single-flip path)
extern "C" int jvmc_mcmc_set_generic(int on) { g_mcmc_generic = on; return JVMC_OK; }

extern "C" int jvmc_rbm_mcmc(int32_t* states, long long C, int N, int M, const double* W, const double* bias,
                             const double* tables, unsigned long long seed, unsigned long long step0,
                             long long chain0, int proposer, double mu, int sweepSteps, long long thermSteps,
                             int numSamplesPerChain, int refreshEvery, int32_t* out, unsigned long long* counters,
                             void* stream) {
  if (!states || !W || !tables || !counters || C < 0 || N <= 0 || M <= 0) return JVMC_ERR_ARG;
  if (numSamplesPerChain > 0 && !out) return JVMC_ERR_ARG;
  if (N > 1024) return JVMC_ERR_UNSUPPORTED;
  if (proposer < 0 || proposer > 2 || sweepSteps <= 0 || thermSteps < 0 || numSamplesPerChain < 0) return JVMC_ERR_ARG;
  if (proposer == 2 && (N % 2 != 0)) return JVMC_ERR_ARG;
  if (C == 0) return JVMC_OK;
  McmcArgs a;
  a.states = states; a.C = C; a.N = N; a.M = M;
  a.W = (const cplx*)W; a.bias = (const cplx*)bias;
  a.T = (const cplx*)tables; a.lc = a.T + (size_t)N * M; a.tb2 = a.lc + N; a.lcb = a.tb2 + M;
  a.seed = seed; a.step0 = step0; a.chain0 = chain0; a.proposer = proposer; a.mu = mu;
  a.K = sweepSteps; a.thermSteps = thermSteps; a.numSamples = numSamplesPerChain;
  a.refreshEvery = refreshEvery > 0 ? refreshEvery : 1;
  a.out = out; a.counters = counters;
  if (proposer != 2 && !g_mcmc_generic) {
    // single-flip fast path: one warp per chain up to M = 512, then 2 / 4 / 8 warps per chain (M <= 4096)
    int rc = JVMC_ERR_UNSUPPORTED;
    if (M <= JVMC_MC_MAXM1) rc = dispatch_flip<1>(a, (cudaStream_t)stream);
    else if (M <= JVMC_MC_MAXM2) rc = dispatch_flip<2>(a, (cudaStream_t)stream);
    else if (M <= JVMC_MC_MAXM4) rc = dispatch_flip<4>(a, (cudaStream_t)stream);
    else if (M <= 4096) rc = dispatch_flip<8>(a, (cudaStream_t)stream);
    if (rc != JVMC_ERR_UNSUPPORTED) return rc;
  }
  size_t smem = (size_t)MC_WPC * M * sizeof(cplx) + MC_WPC * 32 * sizeof(uint32_t);
  if (smem > 227 * 1024) return JVMC_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(rbm_mcmc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  unsigned grid = (unsigned)((C + MC_WPC - 1) / MC_WPC);
  rbm_mcmc_kernel<<<grid, MC_WPC * 32, smem, (cudaStream_t)stream>>>(a);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
