// Incremental evaluation of the real-parameter CNN (reference jVMC/nets/cnn.py:18-81) for local moves.
//
// The reference evaluates the whole net for every Metropolis proposal (jVMC/sampler.py:327-356) and for every
// connected configuration s' of the local energy (jVMC/operator/base.py:166-192).  A move that changes one or two
// sites only changes the pre-activations of layer l inside a box of (l (F-1) + 1)^d positions around each site, so
//   z_1'(p)  = z_1(p) + sum_k K_1[site_k - p] (s'_k - s_k)                      (F^d positions per site)
//   z_l'(p)  = z_l(p) + sum_{q in window(p), q affected} K_l[q - p] (a_{l-1}'(q) - a_{l-1}(q))
//   log psi' = log psi + sum_{p affected} (act(z_L'(p)) - act(z_L(p))) / nrm
// with the pre-activations z_l and activations a_l of the current configuration cached in shared memory.
//
//  jvmc_cnn_mcmc_inc  <- MCSampler._sweep for nets.CNN with stride 1: one or two warps per Markov chain (two when shared
//                        memory limits the chains per SM), cached state updated in place on acceptance; global flips (Z2 / zero-magnetisation proposers, sampler.py:28-64) and
//                        the end of every sweep re-evaluate the net in full (bounds the drift of the cached sums)
//  jvmc_cnn_eloc_bfo  <- Operator.get_O_loc for nets.CNN: one CTA per sample, full forward once, then every warp takes
//                        off-diagonal strings and evaluates psi(s')/psi(s) from the delta (strings changing <= 2 sites)
// Same Philox counters and proposal rules as the generic kernel of cnn.cu: identical chains up to rounding of the
// acceptance probability.  Eligibility (else JVMC_ERR_UNSUPPORTED and the caller takes the generic path): stride 1,
// <= 4 layers, F <= L per axis, N <= 1024, state + scratch within the shared memory of an SM.
#include "common.cuh"
#include "cnn_common.cuh"
#include "bfo_common.cuh"

namespace {

constexpr int INC_MAXL = 4;
constexpr size_t INC_SMEM_SM = 227 * 1024;     // dynamic shared memory of an SM; 1 KB per CTA is reserved on top

struct IncLayout {
  int nl, Lx, Ly, N, Fx, Fy, P;
  int C[INC_MAXL + 1], C4[INC_MAXL + 1];        // channels, and rounded up to a multiple of 4
  int act[INC_MAXL], hasBias[INC_MAXL], offB[INC_MAXL], offK[INC_MAXL];   // offsets in theta (reference order)
  int wB[INC_MAXL], wK[INC_MAXL], wSize;        // shared-memory weight image: bias [C4], kernel [f][ci][C4], zero padded
  int offZ[INC_MAXL + 1], offA[INC_MAXL + 1];   // state (doubles): a_0 = +-1 input, then z_l and (l < nl) a_l, [channel][position]
  int stateSize;
  int Wx[INC_MAXL + 1], Wy[INC_MAXL + 1];       // box of affected positions per changed site
  int maxAff[INC_MAXL + 1];                     // list capacity (two changed sites)
  int offV[INC_MAXL + 1];                       // scratch (doubles): z' [maxAff C_l]; for l < nl also a' and a' - a
  int valSize;                                  // + 12 doubles at the end: 4 reduction slots, list counts, proposal exchange
  int offLst[INC_MAXL + 1], offMap[INC_MAXL + 1];   // scratch (16-bit units): affected (x, y) lists, position -> list index (l < nl)
  int shortSize;                                // 16-bit units, multiple of 4
  double nrm;
};

inline bool make_inc_layout(const CnnDesc& d, IncLayout& L) {
  if (d.nl > INC_MAXL || d.sx != 1 || d.sy != 1 || d.Fx > d.Lx || d.Fy > d.Ly) return false;
  L.nl = d.nl; L.Lx = d.Lx; L.Ly = d.Ly; L.N = d.Lx * d.Ly; L.Fx = d.Fx; L.Fy = d.Fy; L.P = d.P; L.nrm = d.nrm;
  if (L.N > 1024) return false;
  int off = L.N, offV = 0, offS = 0, offW = 0;
  L.C[0] = 1; L.C4[0] = 1; L.offA[0] = 0; L.offZ[0] = 0;
  L.Wx[0] = L.Wy[0] = 1; L.maxAff[0] = 2; L.offV[0] = 0; L.offLst[0] = L.offMap[0] = 0;
  for (int l = 1; l <= d.nl; ++l) {
    L.C[l] = d.ch[l];
    L.C4[l] = (d.ch[l] + 3) & ~3;
    L.act[l - 1] = d.act[l - 1]; L.hasBias[l - 1] = d.hasBias[l - 1]; L.offB[l - 1] = d.offB[l - 1]; L.offK[l - 1] = d.offK[l - 1];
    L.wB[l - 1] = offW; offW += L.C4[l];
    L.wK[l - 1] = offW; offW += d.Fx * d.Fy * L.C[l - 1] * L.C4[l];
    L.offZ[l] = off; off += L.C[l] * L.N;
    L.offA[l] = off; if (l < d.nl) off += L.C[l] * L.N;
    L.Wx[l] = min(d.Lx, l * (d.Fx - 1) + 1);
    L.Wy[l] = min(d.Ly, l * (d.Fy - 1) + 1);
    L.maxAff[l] = min(L.N, 2 * L.Wx[l] * L.Wy[l]);
    L.offV[l] = offV; offV += (l < d.nl ? 3 : 1) * L.maxAff[l] * L.C[l];
    L.offLst[l] = offS; offS += 2 * L.maxAff[l];
    L.offMap[l] = offS; if (l < d.nl) offS += L.N;
  }
  L.wSize = offW;
  L.stateSize = off; L.valSize = offV + 12; L.shortSize = (offS + 3) & ~3;
  return true;
}

inline size_t inc_scratch_bytes(const IncLayout& L) {
  return (size_t)L.valSize * sizeof(double) + (size_t)L.shortSize * sizeof(int16_t);
}

__device__ __forceinline__ int wrap_up(int v, int n) { return v >= n ? v - n : v; }     // v in [0, 2n)
__device__ __forceinline__ int wrap_dn(int v, int n) { return v < 0 ? v + n : v; }      // v in [-n, n)
// a / d for 0 <= a < 2^23 (float reciprocal + one correction step instead of the ~20-instruction integer division)
__device__ __forceinline__ int fast_div(int a, int d, float inv) {
  int q = (int)(__int2float_rn(a) * inv);
  const int r = a - q * d;
  if (r >= d) ++q; else if (r < 0) --q;
  return q;
}

// Shared-memory image of the parameters: per layer the bias and the kernel [f][ci][C4] with the output channels padded
// by zeros to a multiple of 4 (16-byte rows: vector loads, no channel predicates in the inner loops).
__device__ void inc_load_weights(const IncLayout& L, const double* __restrict__ theta, double* w) {
  for (int i = threadIdx.x; i < L.wSize; i += blockDim.x) w[i] = 0.0;
  __syncthreads();
  for (int l = 0; l < L.nl; ++l) {
    const int C = L.C[l + 1], C4 = L.C4[l + 1], rows = L.Fx * L.Fy * L.C[l];
    if (L.hasBias[l])
      for (int i = threadIdx.x; i < C; i += blockDim.x) w[L.wB[l] + i] = theta[L.offB[l] + i];
    for (int i = threadIdx.x; i < rows * C; i += blockDim.x) {
      const int r = i / C, co = i - r * C;
      w[L.wK[l] + r * C4 + co] = theta[L.offK[l] + i];
    }
  }
  __syncthreads();
}

template <int NT>
__device__ __forceinline__ void group_sync() {
  if (NT == 32) __syncwarp(); else __syncthreads();
}

// sum over the NT threads of a group (a warp, or the whole CTA); red: the 12 trailing doubles of the group's scratch
template <int NT>
__device__ __forceinline__ double group_sum(double v, double* red) {
  v = warp_sum(v);
  if (NT == 32) return v;
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int k = 0; k < NT / 32; ++k) tot += red[k];
  __syncthreads();
  return tot;
}

// Compile-time shape of a net.  SpecRuntime: everything is read from the layout (any eligible net).  Spec2: a two-layer
// net with known filter, channel counts and activation -- the loops over taps, channels and layers unroll, the
// activation switch and the channel predicates fold, the weight offsets become immediates (about half the instructions
// of the runtime version).  Instantiated for the BASELINE configs[3] net; same arithmetic in the same order, so both
// versions (and the generic full-forward kernels of cnn.cu) produce identical chains.
struct SpecRuntime {
  static constexpr bool fixed = false;
  static constexpr int NL = 0, FX = 0, FY = 0, ACT = 0;
  __host__ __device__ static constexpr int C(int) { return 0; }
};
template <int FX_, int FY_, int C1_, int C2_, int ACT_>
struct Spec2 {
  static constexpr bool fixed = true;
  static constexpr int NL = 2, FX = FX_, FY = FY_, ACT = ACT_;
  __host__ __device__ static constexpr int C(int l) { return l == 0 ? 1 : (l == 1 ? C1_ : C2_); }
};
#define S_NL (S::fixed ? S::NL : L.nl)
#define S_FX (S::fixed ? S::FX : L.Fx)
#define S_FY (S::fixed ? S::FY : L.Fy)
#define S_C(l) (S::fixed ? S::C(l) : L.C[l])
#define S_C4(l) (S::fixed ? ((S::C(l) + 3) & ~3) : L.C4[l])
#define S_ACT(l) (S::fixed ? S::ACT : L.act[l])

// z[0..3] += sum_ci K[ci][co0..co0+3] x[ci]; kr: row of the tap (ci = 0) at column co0, 16-byte aligned, row stride C4
template <int CP>
__device__ __forceinline__ void tap4_n(const double* __restrict__ kr, int C4, const double* __restrict__ x, int xs, int Cp,
                                       double (&z)[4]) {
#pragma unroll
  for (int ci = 0; ci < (CP > 0 ? CP : Cp); ++ci) {
    const double v = x[ci * xs];
    const double2 k01 = *reinterpret_cast<const double2*>(kr + ci * C4);
    const double2 k23 = *reinterpret_cast<const double2*>(kr + ci * C4 + 2);
    z[0] = fma(v, k01.x, z[0]); z[1] = fma(v, k01.y, z[1]);
    z[2] = fma(v, k23.x, z[2]); z[3] = fma(v, k23.y, z[3]);
  }
}
// the usual input-channel counts get an unrolled body (the loop bookkeeping is otherwise as long as the arithmetic)
__device__ __forceinline__ void tap4(const double* __restrict__ kr, int C4, const double* __restrict__ x, int xs, int Cp,
                                     double (&z)[4]) {
  switch (Cp) {
    case 1: tap4_n<1>(kr, C4, x, xs, Cp, z); break;
    case 2: tap4_n<2>(kr, C4, x, xs, Cp, z); break;
    case 3: tap4_n<3>(kr, C4, x, xs, Cp, z); break;
    case 4: tap4_n<4>(kr, C4, x, xs, Cp, z); break;
    case 6: tap4_n<6>(kr, C4, x, xs, Cp, z); break;
    case 8: tap4_n<8>(kr, C4, x, xs, Cp, z); break;
    default: tap4_n<0>(kr, C4, x, xs, Cp, z); break;
  }
}

// Full evaluation into the state buffer (a_0 must hold the configuration as +-1) by the NT threads of the group.
// Returns log psi in every thread.  Out of line: the sampler calls it from four places and its loop must stay inside
// the instruction cache.
template <int NT, class S>
__device__ __noinline__ double inc_forward(const IncLayout& L, const double* __restrict__ w, double* st, double* red, int tid) {
  const int N = L.N;
  const float invLy = 1.0f / (float)L.Ly;
  double sum = 0.0;
#pragma unroll
  for (int l = 1; l <= S_NL; ++l) {
    const int ci_n = S_C(l - 1), co_n = S_C(l), C4 = S_C4(l), a = S_ACT(l - 1);
    const double* in = st + L.offA[l - 1];
    double* z = st + L.offZ[l];
    double* ao = st + L.offA[l];
    const bool last = (l == S_NL);
    const double* Kp = w + L.wK[l - 1];
    const double* bp = w + L.wB[l - 1];
    for (int p = tid; p < N; p += NT) {
      const int px = fast_div(p, L.Ly, invLy), py = p - px * L.Ly;
#pragma unroll
      for (int co0 = 0; co0 < co_n; co0 += 4) {
        double acc[4] = {bp[co0], bp[co0 + 1], bp[co0 + 2], bp[co0 + 3]};
#pragma unroll 1
        for (int fx = 0; fx < S_FX; ++fx) {
          const int qx = wrap_up(px + fx, L.Lx);
#pragma unroll 1
          for (int fy = 0; fy < S_FY; ++fy) {
            const int q = qx * L.Ly + wrap_up(py + fy, L.Ly);
            tap4(Kp + ((fx * S_FY + fy) * ci_n) * C4 + co0, C4, in + q, N, ci_n, acc);
          }
        }
        double av4[4];
        actf4(a, acc, av4);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (co0 + j < co_n) {
            z[(co0 + j) * N + p] = acc[j];
            const double av = av4[j];
            if (last) sum += av; else ao[(co0 + j) * N + p] = av;
          }
      }
    }
    group_sync<NT>();
  }
  return group_sum<NT>(sum, red) / L.nrm;
}

struct IncChange {
  int n;            // changed sites (0..2)
  int site[2];
  double delta[2];  // new - old input value (+-2)
};

// log psi(s') - log psi(s) for the change c, evaluated by the NT threads of the group (tid: index in the group).
// The new pre-activations / activations of the affected entries stay in the group's scratch (val, sh; cnt[l] positions
// per layer) for inc_commit.  The position maps are left clean (-1 everywhere).
template <int NT, class S>
__device__ double inc_delta(const IncLayout& L, const double* __restrict__ w, const double* __restrict__ st, double* val,
                            int16_t* sh, const IncChange& c, int* cnt, int tid) {
  const int N = L.N, Ly = L.Ly, Lx = L.Lx, lane = tid & 31;
  const float invLy = 1.0f / (float)Ly;
  const int sx0 = fast_div(c.site[0], Ly, invLy), sy0 = c.site[0] - sx0 * Ly;
  const int sx1 = fast_div(c.site[1], Ly, invLy), sy1 = c.site[1] - sx1 * Ly;
  double dsum = 0.0;
  // ---- affected positions of every layer: union of the boxes of the changed sites (ballot compaction by one warp per
  // layer; with several warps in the group the layers are built concurrently by different warps)
  {
    const int gw = tid >> 5;                      // warp within the group
#pragma unroll
    for (int l = 1; l <= S_NL; ++l) {
      if ((l - 1) % (NT / 32) != gw) continue;
      const int Wx = L.Wx[l], Wy = L.Wy[l], per = Wx * Wy, ncand = c.n * per;
      uint16_t* lx = reinterpret_cast<uint16_t*>(sh) + L.offLst[l];
      uint16_t* ly = lx + L.maxAff[l];
      int16_t* map = sh + L.offMap[l];
      const bool last = (l == S_NL);
      const float invWy = 1.0f / (float)Wy;
      int n = 0;
      for (int base = 0; base < ncand; base += 32) {
        const int cand = base + lane;
        bool keep = false;
        int px = 0, py = 0;
        if (cand < ncand) {
          const int k = cand >= per ? 1 : 0;
          const int r = cand - k * per;
          const int dx = fast_div(r, Wy, invWy), dy = r - dx * Wy;
          px = wrap_dn((k ? sx1 : sx0) - dx, Lx);
          py = wrap_dn((k ? sy1 : sy0) - dy, Ly);
          keep = !(k == 1 && wrap_dn(sx0 - px, Lx) < Wx && wrap_dn(sy0 - py, Ly) < Wy);   // not already in box 0
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const int i = n + __popc(m & ((1u << lane) - 1u));
          lx[i] = (uint16_t)px; ly[i] = (uint16_t)py;
          if (!last) map[px * Ly + py] = (int16_t)i;
        }
        n += __popc(m);
      }
      if (NT == 32) cnt[l] = n;
      else if (lane == 0) reinterpret_cast<int*>(val + L.valSize - 12)[8 + l] = n;   // behind the 4 reduction slots
    }
    group_sync<NT>();
    if (NT > 32) {
#pragma unroll
      for (int l = 1; l <= S_NL; ++l) cnt[l] = reinterpret_cast<const int*>(val + L.valSize - 12)[8 + l];
    }
  }
#pragma unroll
  for (int l = 1; l <= S_NL; ++l) {
    const uint16_t* lx = reinterpret_cast<const uint16_t*>(sh) + L.offLst[l];
    const uint16_t* ly = lx + L.maxAff[l];
    const bool last = (l == S_NL);
    const int n = cnt[l];
    // ---- new values: one thread per affected position, output channels in chunks of four
    const int C = S_C(l), C4 = S_C4(l), a = S_ACT(l - 1);
    double* zn = val + L.offV[l];
    double* an = zn + L.maxAff[l] * C;
    double* da = an + L.maxAff[l] * C;
    const double* zo = st + L.offZ[l];
    // work items: (affected position, chunk of four output channels)
    const int nch = (C + 3) >> 2;
    const float invNch = 1.0f / (float)nch;
    for (int it = tid; it < n * nch; it += NT) {
      const int i = nch == 1 ? it : (nch == 2 ? (it >> 1) : fast_div(it, nch, invNch));
      const int co0 = 4 * (it - i * nch);
      const int px = lx[i], py = ly[i], p = px * Ly + py;
      {
        double z[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = (co0 + j < C) ? zo[(co0 + j) * N + p] : 0.0;
        if (l == 1) {
          const double* K1 = w + L.wK[0];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (k < c.n) {
              const int fx = wrap_dn((k ? sx1 : sx0) - px, Lx), fy = wrap_dn((k ? sy1 : sy0) - py, Ly);
              if (fx < S_FX && fy < S_FY) {
                const double* kr = K1 + (fx * S_FY + fy) * C4 + co0;
                const double dk = k ? c.delta[1] : c.delta[0];
#pragma unroll
                for (int j = 0; j < 4; ++j) z[j] = fma(kr[j], dk, z[j]);
              }
            }
          }
        } else {
          const int Cp = S_C(l - 1);
          const int16_t* pmap = sh + L.offMap[l - 1];
          const double* pda = val + L.offV[l - 1] + 2 * L.maxAff[l - 1] * Cp;
          const double* Kp = w + L.wK[l - 1];
#pragma unroll 1
          for (int fx = 0; fx < S_FX; ++fx) {
            const int qx = wrap_up(px + fx, Lx);
#pragma unroll
            for (int fy = 0; fy < S_FY; ++fy) {
              const int j = pmap[qx * Ly + wrap_up(py + fy, Ly)];
              if (j >= 0) tap4(Kp + ((fx * S_FY + fy) * Cp) * C4 + co0, C4, pda + j * Cp, 1, Cp, z);
            }
          }
        }
        double av4[4], ao4[4] = {0.0, 0.0, 0.0, 0.0};
        actf4(a, z, av4);
        if (last) {
          double zo4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) zo4[j] = (co0 + j < C) ? zo[(co0 + j) * N + p] : 0.0;
          actf4(a, zo4, ao4);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (co0 + j < C) {
            const int e = i * C + co0 + j;
            zn[e] = z[j];
            const double av = av4[j];
            if (last) {
              dsum += av - ao4[j];
            } else {
              an[e] = av;
              da[e] = av - st[L.offA[l] + (co0 + j) * N + p];
            }
          }
      }
    }
    if (l < S_NL) group_sync<NT>();             // (the reduction below synchronises after the last layer)
  }
  // ---- sum over the group; the maps are cleaned between the two barriers of the reduction (every thread is past its
  // last map read at the first one)
  dsum = warp_sum(dsum);
  double* red = val + L.valSize - 12;
  if (NT > 32) {
    if (lane == 0) red[tid >> 5] = dsum;
    __syncthreads();
    dsum = 0.0;
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) dsum += red[k];
  } else {
    __syncwarp();
  }
  for (int l = 1; l < S_NL; ++l) {
    const uint16_t* lx = reinterpret_cast<const uint16_t*>(sh) + L.offLst[l];
    const uint16_t* ly = lx + L.maxAff[l];
    int16_t* map = sh + L.offMap[l];
    for (int i = tid; i < cnt[l]; i += NT) map[lx[i] * Ly + ly[i]] = -1;
  }
  group_sync<NT>();
  return dsum / L.nrm;
}

// Make the change of the last inc_delta the cached state.
template <int NT, class S>
__device__ void inc_commit(const IncLayout& L, double* st, const double* val, const int16_t* sh, const IncChange& c,
                           const int* cnt, int tid) {
  const int N = L.N;
  if (tid < c.n) st[c.site[tid]] += c.delta[tid];
#pragma unroll
  for (int l = 1; l <= S_NL; ++l) {
    const int C = S_C(l);
    const uint16_t* lx = reinterpret_cast<const uint16_t*>(sh) + L.offLst[l];
    const uint16_t* ly = lx + L.maxAff[l];
    const double* zn = val + L.offV[l];
    const double* an = zn + L.maxAff[l] * C;
    for (int i = tid; i < cnt[l]; i += NT) {
      const int p = lx[i] * L.Ly + ly[i];
#pragma unroll
      for (int co = 0; co < C; ++co) {
        st[L.offZ[l] + co * N + p] = zn[i * C + co];
        if (l < S_NL) st[L.offA[l] + co * N + p] = an[i * C + co];
      }
    }
  }
  group_sync<NT>();
}

template <int NT>
__device__ __forceinline__ void inc_clear_maps(const IncLayout& L, int16_t* sh, int tid) {
  for (int l = 1; l < L.nl; ++l)
    for (int i = tid; i < L.N; i += NT) sh[L.offMap[l] + i] = -1;
  group_sync<NT>();
}

struct IncMcmcArgs {
  int32_t* states;
  long long C;
  unsigned long long seed, step0;
  long long chain0;
  int proposer;
  double mu;
  int K;
  long long thermSteps;
  int numSamples;
  int32_t* out;
  unsigned long long* counters;
};

// site of the r-th (1-based) set bit over the lanes' words (site order), -1 if there are fewer
__device__ __forceinline__ int nth_set_site(unsigned word, int r, int lane) {
  const int mine = __popc(word);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const bool here = (incl >= r) && (incl - mine < r);
  const unsigned m = __ballot_sync(0xffffffffu, here);
  if (m == 0u) return -1;
  const int src = __ffs(m) - 1;
  int bit = 0;
  if (here) {
    unsigned w = word;
    int n = r - (incl - mine);                   // the n-th set bit of this word, by halving
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) {
      const int low = __popc(w & ((1u << sh) - 1u));
      if (n > low) { n -= low; w >>= sh; bit += sh; }
    }
  }
  return 32 * src + __shfl_sync(0xffffffffu, bit, src);
}

// NT = 32: one warp per chain, several chains per CTA; NT = 64 / 128: the two / four warps of a CTA share one chain (the
// layout when shared memory, not threads, limits the chains per SM: more warps to issue from)
template <int NT, class S>
__global__ void __launch_bounds__(NT == 32 ? 256 : NT, NT == 32 ? 1 : 8)
cnn_inc_mcmc_kernel(const __grid_constant__ IncLayout L, const double* __restrict__ theta, IncMcmcArgs a, int perChainBytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* w = reinterpret_cast<double*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tid = NT == 32 ? lane : (int)threadIdx.x;
  const int N = L.N;
  inc_load_weights(L, theta, w);
  const long long chain = NT == 32 ? (long long)blockIdx.x * (blockDim.x >> 5) + warp : (long long)blockIdx.x;
  if (chain >= a.C) return;                      // NT = 32 only: no block-wide barrier below
  unsigned char* mine = smem_raw + (size_t)L.wSize * sizeof(double) + (size_t)(NT == 32 ? warp : 0) * perChainBytes;
  double* st = reinterpret_cast<double*>(mine);
  double* val = st + L.stateSize;
  double* red = val + L.valSize - 12;
  int16_t* sh = reinterpret_cast<int16_t*>(val + L.valSize);
  inc_clear_maps<NT>(L, sh, tid);
  const unsigned long long gchain = (unsigned long long)(a.chain0 + chain);
  const Philox rng(a.seed);
  // lane l of every warp owns the sites [32 l, 32 l + 32)
  unsigned valid = 0, bits = 0;
  for (int k = 0; k < 32; ++k) {
    const int i = 32 * lane + k;
    if (i < N) {
      valid |= 1u << k;
      if (a.states[chain * N + i] != 0) bits |= 1u << k;
    }
  }
  auto load_input = [&](unsigned wbits) {
    for (int w0 = 0; w0 < N; w0 += 32) {          // word by word: consecutive lanes write consecutive sites
      const unsigned wb = __shfl_sync(0xffffffffu, wbits, w0 >> 5);
      if ((NT == 32 || warp == 0) && w0 + lane < N) st[w0 + lane] = ((wb >> lane) & 1u) ? 1.0 : -1.0;
    }
    group_sync<NT>();
  };
  load_input(bits);
  double cur = inc_forward<NT, S>(L, w, st, red, tid);
  unsigned long long nAcc = 0, nProp = 0;
  const long long total = a.thermSteps + (long long)a.numSamples * a.K;
  long long nextEmit = a.thermSteps + a.K;
  int emitted = 0;
  int cnt[INC_MAXL + 1];
  uint4 qa = make_uint4(0, 0, 0, 0);
  uint32_t qb = 0;
  for (long long stp = 0; stp < total; ++stp) {
    const unsigned long long gs = a.step0 + (unsigned long long)stp;
    IncChange c;
    c.n = 0; c.site[0] = c.site[1] = 0; c.delta[0] = c.delta[1] = 0.0;
    bool g = false;
    double u = 0.0;
    int* px = reinterpret_cast<int*>(red + 8);         // proposal exchange slots (behind the list counts)
    if (NT == 32 || warp == 0) {
    // Philox once per 32 steps: lane l evaluates the counters of step (batch base + l), the step's values come by shuffle
    const int bidx = (int)(stp & 31);
    if (bidx == 0) {
      const unsigned long long gl = gs + (unsigned long long)lane;
      qa = rng((uint32_t)gl, (uint32_t)(gl >> 32), (uint32_t)gchain, (uint32_t)(gchain >> 32) << 8);
      if (a.proposer == 2)
        qb = rng((uint32_t)gl, (uint32_t)(gl >> 32), (uint32_t)gchain, ((uint32_t)(gchain >> 32) << 8) | 1u).x;
    }
    uint4 r;
    r.x = __shfl_sync(0xffffffffu, qa.x, bidx); r.y = __shfl_sync(0xffffffffu, qa.y, bidx);
    r.z = __shfl_sync(0xffffffffu, qa.z, bidx); r.w = __shfl_sync(0xffffffffu, qa.w, bidx);
    u = u01_from_bits(r.z, r.w);
    if (a.proposer == 2) {
      // exchange the ru-th up spin with the rd-th down spin (sampler.py:42-64)
      const int half = N / 2;
      const int ru = 1 + (int)__umulhi(r.x, (uint32_t)half), rd = 1 + (int)__umulhi(r.y, (uint32_t)half);
      const int iu = nth_set_site(bits & valid, ru, lane), id = nth_set_site(~bits & valid, rd, lane);
      if (iu >= 0) { c.site[0] = iu; c.delta[0] = -2.0; c.n = 1; }
      if (id >= 0) {
        if (c.n == 0) { c.site[0] = id; c.delta[0] = 2.0; } else { c.site[1] = id; c.delta[1] = 2.0; }
        ++c.n;
      }
      g = __umulhi(__shfl_sync(0xffffffffu, qb, bidx), 5u) == 0u;
    } else {
      const int k = (int)__umulhi(r.x, (uint32_t)N);
      const unsigned wk = __shfl_sync(0xffffffffu, bits, k >> 5);
      c.site[0] = k; c.delta[0] = ((wk >> (k & 31)) & 1u) ? -2.0 : 2.0; c.n = 1;
      g = (a.proposer == 1) && (__umulhi(r.y, 5u) == 0u);
    }
    }
    if (NT > 32) {
      // the first warp draws the proposal, the others read it
      double* pu = red + 10;
      if (threadIdx.x == 0) {
        px[0] = c.n | (g ? 4 : 0) | (c.delta[0] > 0.0 ? 8 : 0) | (c.delta[1] > 0.0 ? 16 : 0);
        px[1] = c.site[0]; px[2] = c.site[1]; *pu = u;
      }
      __syncthreads();
      const int w0 = px[0];
      c.n = w0 & 3; g = (w0 & 4) != 0; c.site[0] = px[1]; c.site[1] = px[2]; u = *pu;
      c.delta[0] = c.n > 0 ? ((w0 & 8) ? 2.0 : -2.0) : 0.0;
      c.delta[1] = c.n > 1 ? ((w0 & 16) ? 2.0 : -2.0) : 0.0;
    }
    unsigned nbits = bits;
    if (c.n > 0 && (c.site[0] >> 5) == lane) nbits ^= 1u << (c.site[0] & 31);
    if (c.n > 1 && (c.site[1] >> 5) == lane) nbits ^= 1u << (c.site[1] & 31);
    nProp += 1;
    if (!g) {
      const double dlt = inc_delta<NT, S>(L, w, st, val, sh, c, cnt, tid);
      if (u < exp(a.mu * dlt)) {
        nAcc += 1;
        inc_commit<NT, S>(L, st, val, sh, c, cnt, tid);
        cur += dlt;
        bits = nbits;
      }
    } else {
      // local change followed by the flip of every spin: nothing of the cached state survives -> full evaluation;
      // a rejection evaluates the old configuration again
      nbits = ~nbits & valid;
      load_input(nbits);
      const double nxt = inc_forward<NT, S>(L, w, st, red, tid);
      if (u < exp(a.mu * (nxt - cur))) {
        nAcc += 1;
        cur = nxt;
        bits = nbits;
      } else {
        load_input(bits);
        cur = inc_forward<NT, S>(L, w, st, red, tid);
      }
    }
    if (stp + 1 == nextEmit) {
      const long long row = (long long)emitted * a.C + chain;   // time-major, chain-minor (sampler.py:323)
      for (int w0 = 0; w0 < N; w0 += 32) {                       // warp-uniform trip count (shuffles inside)
        const unsigned wb = __shfl_sync(0xffffffffu, bits, w0 >> 5);
        if ((NT == 32 || warp == 0) && w0 + lane < N) a.out[row * N + w0 + lane] = (int32_t)((wb >> lane) & 1u);
      }
      ++emitted;
      nextEmit += a.K;
      cur = inc_forward<NT, S>(L, w, st, red, tid);   // once per sweep: the cached sums never drift far
    }
  }
  for (int w0 = 0; w0 < N; w0 += 32) {
    const unsigned wb = __shfl_sync(0xffffffffu, bits, w0 >> 5);
    if ((NT == 32 || warp == 0) && w0 + lane < N) a.states[chain * N + w0 + lane] = (int32_t)((wb >> lane) & 1u);
  }
  if ((NT == 32 ? lane : (int)threadIdx.x) == 0) {
    atomicAdd(a.counters + 0, nProp);
    atomicAdd(a.counters + 1, nAcc);
  }
}

// one CTA per sample: the four warps evaluate the net once together, then take the off-diagonal strings in turn
template <class S, int NW>
__global__ void __launch_bounds__(32 * NW)
cnn_inc_eloc_kernel(const __grid_constant__ IncLayout L, const double* __restrict__ theta, BfoTables t, const int32_t* __restrict__ s, long long B,
                    const cplx* __restrict__ pref, int numDiag, cplx* __restrict__ out, int* __restrict__ errFlag,
                    int perWarpBytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int N = L.N;
  double* w = reinterpret_cast<double*>(smem_raw);
  double* st = w + L.wSize;
  double* red = st + L.stateSize;                                   // [16]
  int32_t* cfg = reinterpret_cast<int32_t*>(red + 16);
  unsigned char* scratch = reinterpret_cast<unsigned char*>(cfg + ((N + 3) & ~3)) + (size_t)warp * perWarpBytes;
  double* val = reinterpret_cast<double*>(scratch);
  int16_t* sh = reinterpret_cast<int16_t*>(val + L.valSize);
  const long long b = blockIdx.x;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int v = s[b * N + i];
    cfg[i] = v;
    st[i] = v ? 1.0 : -1.0;
  }
  inc_clear_maps<32>(L, sh, lane);
  inc_load_weights(L, theta, w);                                    // (block barriers inside)
  (void)inc_forward<32 * NW, S>(L, w, st, red, (int)threadIdx.x);
  cplx eloc = cmk(0.0, 0.0), diagAcc = cmk(0.0, 0.0);
  int cnt[INC_MAXL + 1];
  int rank = 0;
  for (int o0 = 0; o0 < t.numOps; o0 += 32) {
    const int o = o0 + lane;
    cplx m = cmk(0.0, 0.0);
    int f0 = -1, f1 = -1, nf = 0;
    bool isd = true;
    if (o < t.numOps) {
      int msite[BFO_MAXLEN], mval[BFO_MAXLEN], nmod;
      m = walk_string(t, o, cfg, N, pref[o], msite, mval, nmod);
      isd = t.isDiag[o] != 0;
      for (int k = 0; k < nmod; ++k)
        if (mval[k] != cfg[msite[k]]) { if (nf == 0) f0 = msite[k]; else if (nf == 1) f1 = msite[k]; ++nf; }
      if (nf == 2 && f1 < f0) { const int tmp = f0; f0 = f1; f1 = tmp; }
    }
    // diagonal strings are merged into one entry (branch_free.py:483-485) ...
    diagAcc = cadd(diagAcc, warp_csum((isd && o < t.numOps) ? m : cmk(0.0, 0.0)));
    unsigned live = __ballot_sync(0xffffffffu, (o < t.numOps) && !isd && (hypot(m.x, m.y) > 1e-6));
    while (live) {
      const int src = __ffs(live) - 1;
      live &= live - 1;
      const int a0 = __shfl_sync(0xffffffffu, f0, src), a1 = __shfl_sync(0xffffffffu, f1, src);
      const int nfl = __shfl_sync(0xffffffffu, nf, src);
      cplx mm = cmk(__shfl_sync(0xffffffffu, m.x, src), __shfl_sync(0xffffffffu, m.y, src));
      // the live strings that follow with the same changed sites lead to the same s' (e.g. Sx Sx and Sy Sy on a bond):
      // their matrix elements share one amplitude ratio
      while (live) {
        const int nx = __ffs(live) - 1;
        if (__shfl_sync(0xffffffffu, f0, nx) != a0 || __shfl_sync(0xffffffffu, f1, nx) != a1 ||
            __shfl_sync(0xffffffffu, nf, nx) != nfl)
          break;
        live &= live - 1;
        mm = cadd(mm, cmk(__shfl_sync(0xffffffffu, m.x, nx), __shfl_sync(0xffffffffu, m.y, nx)));
      }
      // matrix elements that cancel exactly (S^xS^x + S^yS^y on a bond of parallel spins) need no amplitude ratio:
      // the reference adds m r and -m r with the same r
      if (mm.x == 0.0 && mm.y == 0.0) continue;
      if ((rank++ % nw) != warp) continue;       // the CTA's warps take the distinct s' in turn
      if (nfl > 2) { if (lane == 0) atomicExch(errFlag, 1); continue; }
      IncChange c;
      c.n = nfl;
      c.site[0] = a0 >= 0 ? a0 : 0; c.site[1] = a1 >= 0 ? a1 : 0;
      c.delta[0] = a0 >= 0 ? -2.0 * st[c.site[0]] : 0.0;
      c.delta[1] = a1 >= 0 ? -2.0 * st[c.site[1]] : 0.0;
      double ratio = 1.0;
      if (nfl > 0) ratio = exp(inc_delta<32, S>(L, w, st, val, sh, c, cnt, lane));
      eloc = cadd(eloc, cmk(mm.x * ratio, mm.y * ratio));
    }
  }
  __syncthreads();
  if (lane == 0) { red[2 * warp] = eloc.x; red[2 * warp + 1] = eloc.y; }
  __syncthreads();
  if (threadIdx.x == 0) {
    cplx e = cmk(0.0, 0.0);
    for (int k = 0; k < nw; ++k) e = cadd(e, cmk(red[2 * k], red[2 * k + 1]));
    // ... and that entry passes the same |m| > 1e-6 filter as every other one (base.py:93-103)
    cplx d = diagAcc;
    if (hypot(d.x, d.y) <= 1e-6) d = cmk(0.0, 0.0);
    out[b] = cadd(e, d);
  }
}

int g_inc_nt = 64;

template <class S>
int launch_inc_mcmc(const IncLayout& L, const double* theta, const IncMcmcArgs& a, long long C, size_t wBytes, size_t perChain,
                    void* stream) {
  // chains an SM can hold with one chain per CTA
  const size_t smem1 = wBytes + perChain;
  if (smem1 > INC_SMEM_SM) return JVMC_ERR_UNSUPPORTED;
  int perSm1 = (int)(INC_SMEM_SM / (smem1 + 1024));
  if (perSm1 > 32) perSm1 = 32;
  if (perSm1 <= 16 && g_inc_nt == 64) {
    // shared memory limits the residency: two warps per chain (measured at config 4: 28.6 ms against 31.2 ms with four
    // and 65 ms with one; jvmc_cnn_set_generic(2) selects four for A/B runs)
    if (smem1 > 48 * 1024)
      cudaFuncSetAttribute(cnn_inc_mcmc_kernel<64, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    cudaFuncSetAttribute(cnn_inc_mcmc_kernel<64, S>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cnn_inc_mcmc_kernel<64, S><<<(unsigned)C, 64, smem1, (cudaStream_t)stream>>>(L, theta, a, (int)perChain);
  } else if (perSm1 <= 16) {
    if (smem1 > 48 * 1024)
      cudaFuncSetAttribute(cnn_inc_mcmc_kernel<128, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    // the chains per SM follow from shared memory: ask for the largest carve-out
    cudaFuncSetAttribute(cnn_inc_mcmc_kernel<128, S>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cnn_inc_mcmc_kernel<128, S><<<(unsigned)C, 128, smem1, (cudaStream_t)stream>>>(L, theta, a, (int)perChain);
  } else {
    // small nets: one warp per chain, the chains of a CTA share the weight image
    int cpc = 1;
    while (cpc < 8 && wBytes + 2 * cpc * perChain <= 48 * 1024) cpc *= 2;
    const size_t smem = wBytes + cpc * perChain;
    cnn_inc_mcmc_kernel<32, S><<<(unsigned)((C + cpc - 1) / cpc), 32 * cpc, smem, (cudaStream_t)stream>>>(L, theta, a, (int)perChain);
  }
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

template <class S>
int launch_inc_eloc(const IncLayout& L, const double* theta, const BfoTables& t, const int32_t* s, long long B,
                    const cplx* pref, int numDiag, cplx* out, int* errFlag, int nw, size_t smem, size_t perWarp, void* stream) {
  if (nw == 8) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(cnn_inc_eloc_kernel<S, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(cnn_inc_eloc_kernel<S, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cnn_inc_eloc_kernel<S, 8><<<(unsigned)B, 256, smem, (cudaStream_t)stream>>>(L, theta, t, s, B, pref, numDiag, out, errFlag,
                                                                             (int)perWarp);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(cnn_inc_eloc_kernel<S, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(cnn_inc_eloc_kernel<S, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cnn_inc_eloc_kernel<S, 4><<<(unsigned)B, 128, smem, (cudaStream_t)stream>>>(L, theta, t, s, B, pref, numDiag, out, errFlag,
                                                                             (int)perWarp);
  }
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

// the net of BASELINE configs[3]: F = (3, 3), channels (6, 4), elu
typedef Spec2<3, 3, 6, 4, 0> SpecCfg4;
inline bool is_cfg4(const IncLayout& L) {
  return L.nl == 2 && L.Fx == 3 && L.Fy == 3 && L.C[1] == 6 && L.C[2] == 4 && L.act[0] == 0 && L.act[1] == 0;
}

}  // namespace

static int g_cnn_generic = 0;
// development knob: 1 makes jvmc_cnn_mcmc_inc / jvmc_cnn_eloc_bfo answer JVMC_ERR_UNSUPPORTED (A/B against the generic path)
extern "C" int jvmc_cnn_set_generic(int on) {
  g_cnn_generic = (on == 1);
  g_inc_nt = (on == 2) ? 128 : 64;      // 2: incremental sampler with four instead of two warps per chain (A/B)
  return JVMC_OK;
}

extern "C" int jvmc_cnn_mcmc_inc(const int* desc, int ndesc, const double* theta, int32_t* states, long long C,
                                 unsigned long long seed, unsigned long long step0, long long chain0, int proposer,
                                 double mu, int sweepSteps, long long thermSteps, int numSamplesPerChain, int32_t* out,
                                 unsigned long long* counters, void* stream) {
  CnnDesc d;
  int rc = make_desc(desc, ndesc, d);
  if (rc != JVMC_OK) return rc;
  if (!theta || !states || !counters || C < 0) return JVMC_ERR_ARG;
  if (numSamplesPerChain > 0 && !out) return JVMC_ERR_ARG;
  if (proposer < 0 || proposer > 2 || sweepSteps <= 0 || thermSteps < 0 || numSamplesPerChain < 0) return JVMC_ERR_ARG;
  IncLayout L;
  if (g_cnn_generic || !make_inc_layout(d, L)) return JVMC_ERR_UNSUPPORTED;
  if (proposer == 2 && (L.N % 2 != 0)) return JVMC_ERR_ARG;
  if (C == 0) return JVMC_OK;
  const size_t perChain = (((size_t)L.stateSize * sizeof(double) + inc_scratch_bytes(L)) + 15) & ~(size_t)15;
  const size_t wBytes = (size_t)L.wSize * sizeof(double);
  IncMcmcArgs a;
  a.states = states; a.C = C; a.seed = seed; a.step0 = step0; a.chain0 = chain0; a.proposer = proposer; a.mu = mu;
  a.K = sweepSteps; a.thermSteps = thermSteps; a.numSamples = numSamplesPerChain; a.out = out; a.counters = counters;
  return is_cfg4(L) ? launch_inc_mcmc<SpecCfg4>(L, theta, a, C, wBytes, perChain, stream)
                    : launch_inc_mcmc<SpecRuntime>(L, theta, a, C, wBytes, perChain, stream);
}

extern "C" int jvmc_cnn_eloc_bfo(const int* desc, int ndesc, const double* theta, const int32_t* s, long long B,
                                 int numOps, int len, int lDim, const int32_t* idx, const int32_t* map,
                                 const double* matEls, const int32_t* fermi, const uint8_t* isDiag, int numDiag,
                                 const double* pref, double* out, int* errFlag, void* stream) {
  CnnDesc d;
  int rc = make_desc(desc, ndesc, d);
  if (rc != JVMC_OK) return rc;
  if (B == 0) return JVMC_OK;
  if (!theta || !s || !idx || !map || !matEls || !fermi || !isDiag || !pref || !out || !errFlag || B < 0) return JVMC_ERR_ARG;
  IncLayout L;
  if (g_cnn_generic || len > BFO_MAXLEN || len <= 0 || numOps <= 0 || lDim != 2 || !make_inc_layout(d, L))
    return JVMC_ERR_UNSUPPORTED;
  BfoTables t = make_tables(numOps, len, lDim, idx, map, matEls, fermi, isDiag);
  const int nw = g_inc_nt == 128 ? 8 : 4;      // warps per sample (8: A/B knob jvmc_cnn_set_generic(2))
  const size_t perWarp = (inc_scratch_bytes(L) + 15) & ~(size_t)15;
  const size_t smem = (size_t)L.wSize * sizeof(double) + (size_t)L.stateSize * sizeof(double) + 16 * sizeof(double) +
                      (size_t)((L.N + 3) & ~3) * sizeof(int32_t) + nw * perWarp;
  if (smem > INC_SMEM_SM) return JVMC_ERR_UNSUPPORTED;
  return is_cfg4(L) ? launch_inc_eloc<SpecCfg4>(L, theta, t, s, B, (const cplx*)pref, numDiag, (cplx*)out, errFlag, nw, smem,
                                                perWarp, stream)
                    : launch_inc_eloc<SpecRuntime>(L, theta, t, s, B, (const cplx*)pref, numDiag, (cplx*)out, errFlag, nw, smem,
                                                   perWarp, stream);
}
