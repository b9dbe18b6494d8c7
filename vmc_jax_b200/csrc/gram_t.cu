// jvmc_rbm_gram_T -- centred tangent kernel  T = Obar Obar^dagger  of MinSR for (Cpx)RBM on fp64 tensor cores.
//
// Replaces SampledObs.tangent_kernel (reference jVMC/stats.py:332-336, used by MinSR.solve,
// jVMC/util/minsr.py:59-60).  The reference gathers the dense centred gradients Obar [N_T x P] to every
// rank and runs an N_T x N_T x P zgemm.  With O_n[(r,j)] = sigma_nr tau_nj (Khatri-Rao) the kernel factorises:
//
//   T_nm = scale * sqrt(p_n p_m) * [ D_nm * Q_nm - v_n - conj(v_m) + c ],
//   D_nm = sum_r sigma_nr sigma_mr = R - 2 popc(bits_n xor bits_m)         (exact integer, bit-packed spins)
//   Q_nm = sum_j tau_nj conj(tau_mj)                                        (fp64 DMMA, K = M instead of P = R*M)
//   v_n  = O_n . conj(mu)  (jvmc_rbm_krmatvec),   c = |mu|^2
//
// i.e. N (= number of sites) times fewer tensor flops than the dense product and no O in memory.
// Pipeline as in gram.cu: TMA-engine bulk copies by a producer warp into an mbarrier-guarded smem ring, 16
// DMMA consumer warps (CTA tile 64 x 64), Hermitian: only tile pairs nb >= mb, both images written.
#include <type_traits>
#include "common.cuh"

namespace {

constexpr int T_KC = 32;      // hidden units per stage (512-byte bulk copies: the 256-byte ones of T_KC = 16 starve the consumers)
constexpr int T_STAGES = 3;
constexpr int T_TS = 64;      // CTA tile (samples)
constexpr int T_LD = T_KC + 4;  // row pitch in complex elements, == 4 (mod 8): conflict-free LDS.128
constexpr int T_NCW = 16;

__device__ double2 g_zero_rowT[T_KC];

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void tri_decode(long long p, int& hi, int& lo) {
  long long h = (long long)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
  while ((h + 1) * (h + 2) / 2 <= p) ++h;
  while (h * (h + 1) / 2 > p) --h;
  hi = (int)h;
  lo = (int)(p - h * (h + 1) / 2);
}

// sigR[n][w]: bit k of word w = (sigma_{n, 32w+k} == +1) over the R Khatri-Rao sites (bias pseudo-site = bit 0).
__global__ void pack_sigma_rows_kernel(const int32_t* __restrict__ s, long long B, int N, int hasBias, int wordsR,
                                       uint32_t* __restrict__ sigR) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * wordsR) return;
  long long n = gid / wordsR;
  int w = (int)(gid - n * wordsR);
  const int R = N + (hasBias ? 1 : 0);
  uint32_t word = 0;
  for (int k = 0; k < 32; ++k) {
    int r = w * 32 + k;
    if (r >= R) break;
    bool up = (hasBias && r == 0) ? true : (s[n * N + (r - (hasBias ? 1 : 0))] != 0);
    if (up) word |= 1u << k;
  }
  sigR[gid] = word;
}

__global__ void __launch_bounds__(32 * (T_NCW + 1), 1)
gram_t_kernel(const cplx* __restrict__ Y, long long B, int M, int R, const uint32_t* __restrict__ sigR, int wordsR,
              const double* __restrict__ p, const cplx* __restrict__ v, const double* __restrict__ cptr, double scale,
              cplx* __restrict__ T, long long tile0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cplx* tiles = reinterpret_cast<cplx*>(smem_raw);  // [stage][2][T_TS][T_LD]
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)T_STAGES * 2 * T_TS * T_LD);
  uint64_t* empty = full + T_STAGES;
  int nb, mb;
  tri_decode(tile0 + blockIdx.x, nb, mb);   // mb <= nb
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rowA0 = (long long)nb * T_TS, rowB0 = (long long)mb * T_TS;
  const int KT = (M + T_KC - 1) / T_KC;

  for (int e = threadIdx.x; e < T_STAGES * 2 * T_TS * T_LD; e += blockDim.x) tiles[e] = cmk(0.0, 0.0);
  if (threadIdx.x == 0) {
    for (int sidx = 0; sidx < T_STAGES; ++sidx) { mbar_init(full + sidx, 1); mbar_init(empty + sidx, T_NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();

  if (warp == T_NCW) {
    // producer: per stage one bulk copy per (tile, sample row) of min(T_KC, M - k0) hidden units
    for (int kt = 0; kt < KT; ++kt) {
      const int slot = kt % T_STAGES;
      if (kt >= T_STAGES) mbar_wait(empty + slot, (unsigned)((kt / T_STAGES - 1) & 1));
      cplx* st = tiles + (size_t)slot * 2 * T_TS * T_LD;
      const int k0 = kt * T_KC;
      const unsigned bytes = (unsigned)(min(T_KC, M - k0) * (int)sizeof(cplx));
      if (lane == 0) mbar_expect_tx(full + slot, 2u * T_TS * bytes);
      __syncwarp();
      for (int e = lane; e < 2 * T_TS; e += 32) {
        const int which = e / T_TS, rr = e - which * T_TS;
        const long long n = (which ? rowB0 : rowA0) + rr;
        const cplx* src = (n < B) ? (Y + n * M + k0) : reinterpret_cast<const cplx*>(g_zero_rowT);
        bulk_g2s(st + ((size_t)which * T_TS + rr) * T_LD, src, bytes, full + slot);
      }
    }
    return;
  }

  // consumers: 8 x 2 warps, warp tile 8 (n) x 32 (m)
  const int wm = warp >> 1, wn = warp & 1;
  const int q4 = lane & 3, q8 = lane >> 2;
  double cre[4][2], cim[4][2];
#pragma unroll
  for (int b = 0; b < 4; ++b) { cre[b][0] = cre[b][1] = 0.0; cim[b][0] = cim[b][1] = 0.0; }
  for (int kt = 0; kt < KT; ++kt) {
    const int slot = kt % T_STAGES;
    const cplx* As = tiles + (size_t)slot * 2 * T_TS * T_LD;
    const cplx* Bs = As + (size_t)T_TS * T_LD;
    mbar_wait(full + slot, (unsigned)((kt / T_STAGES) & 1));
    // hidden units beyond M in the last stage hold stale data from earlier stages -> mask them there (only there:
    // the selects cost issue slots the DMMA stream needs)
    const int kvalid = min(T_KC, M - kt * T_KC);
    auto stage = [&](auto masked) {
      constexpr bool MASK = decltype(masked)::value;
#pragma unroll
      for (int ks = 0; ks < T_KC / 4; ++ks) {
        const int kk = ks * 4 + q4;
        const bool ok = !MASK || kk < kvalid;
        cplx av = As[(size_t)(wm * 8 + q8) * T_LD + kk];
        double ar = ok ? av.x : 0.0, ai = ok ? av.y : 0.0;
        double nar = -ar;
        double br[4], bi[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          cplx bv = Bs[(size_t)(wn * 32 + b * 8 + q8) * T_LD + kk];
          br[b] = ok ? bv.x : 0.0;
          bi[b] = ok ? bv.y : 0.0;
        }
        // a conj(b) = (ar br + ai bi) + i (ai br - ar bi)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          dmma884(cre[b][0], cre[b][1], ar, br[b]);
          dmma884(cim[b][0], cim[b][1], ai, br[b]);
          dmma884(cre[b][0], cre[b][1], ai, bi[b]);
          dmma884(cim[b][0], cim[b][1], nar, bi[b]);
        }
      }
    };
    if (kvalid == T_KC) stage(std::false_type{}); else stage(std::true_type{});
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + slot);
  }

  // epilogue: Hadamard with the integer spin overlap, centring, weights; write both Hermitian images
  const double c = cptr[0];
  const bool diagTile = (nb == mb);
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const long long n = rowA0 + wm * 8 + q8;
      const long long m = rowB0 + wn * 32 + b * 8 + 2 * q4 + e;
      if (n >= B || m >= B) continue;
      if (diagTile && m > n) continue;
      int diff = 0;
      for (int w = 0; w < wordsR; ++w) diff += __popc(sigR[n * wordsR + w] ^ sigR[m * wordsR + w]);
      const double D = (double)(R - 2 * diff);
      const cplx vn = v[n], vm = v[m];
      const double w8 = scale * sqrt(p[n] * p[m]);
      double tr = w8 * (D * cre[b][e] - vn.x - vm.x + c);
      double ti = w8 * (D * cim[b][e] - vn.y + vm.y);
      if (n == m) ti = 0.0;
      T[n * B + m] = cmk(tr, ti);
      if (n != m) T[m * B + n] = cmk(tr, -ti);
    }
}

}  // namespace

extern "C" int jvmc_pack_sigma_rows(const int32_t* s, long long B, int N, int hasBias, unsigned int* sigR,
                                    void* stream) {
  if (B == 0) return JVMC_OK;
  if (!s || !sigR || B < 0 || N <= 0) return JVMC_ERR_ARG;
  const int R = N + (hasBias ? 1 : 0);
  const int wordsR = (R + 31) / 32;
  long long tot = B * wordsR;
  pack_sigma_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(s, B, N, hasBias, wordsR, sigR);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

// T: [B, B] complex128 row-major, fully populated Hermitian.  scale = 2 reproduces the reference's doubled
// holomorphic layout (tangent_kernel of [g, i g] data), scale = 1 the plain complex-parameter kernel.
extern "C" long long jvmc_rbm_gram_T_tiles(long long B) {
  const long long nT = (B + T_TS - 1) / T_TS;
  return nT * (nT + 1) / 2;
}

// [tile0, tile0 + ntiles) of the jvmc_rbm_gram_T_tiles(B) tile pairs of the Hermitian half (ntiles <= 0: all): every
// element of T is written by exactly one tile pair (both images), so ranks that take disjoint ranges into zero-initialised
// buffers obtain the full kernel by one SUM all-reduce -- T is formed once over the ranks, not once per rank.
extern "C" int jvmc_rbm_gram_T(const double* Y, long long B, int M, int R, const unsigned int* sigR, const double* p,
                               const double* v, const double* c, double scale, long long tile0, long long ntiles, double* T,
                               void* stream) {
  if (B == 0) return JVMC_OK;
  if (!Y || !sigR || !p || !v || !c || !T || B < 0 || M <= 0 || R <= 0) return JVMC_ERR_ARG;
  const int wordsR = (R + 31) / 32;
  size_t smem = (size_t)T_STAGES * 2 * T_TS * T_LD * sizeof(cplx) + 2 * T_STAGES * sizeof(uint64_t);
  cudaFuncSetAttribute(gram_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long nT = (B + T_TS - 1) / T_TS;
  long long tiles = nT * (nT + 1) / 2;
  if (ntiles <= 0) { tile0 = 0; ntiles = tiles; }
  if (tile0 < 0 || tile0 + ntiles > tiles) return JVMC_ERR_ARG;
  if (ntiles > 2147483647LL) return JVMC_ERR_UNSUPPORTED;
  gram_t_kernel<<<(unsigned)ntiles, 32 * (T_NCW + 1), smem, (cudaStream_t)stream>>>(
      (const cplx*)Y, B, M, R, sigR, wordsR, p, (const cplx*)v, c, scale, (cplx*)T, tile0);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}
