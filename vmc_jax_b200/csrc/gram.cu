// jvmc_rbm_gram_S -- quantum Fisher matrix S = <O^dagger O>_c for (Cpx)RBM on fp64 tensor cores.
//
// Replaces SampledObs.covar() on the per-sample gradients (reference jVMC/stats.py:52-58,235-245
// called from jVMC/util/tdvp.py:142) -- the one dense contraction of the VMC step.  The reference
// materialises O [N_s x P] (complex128) and runs a P x P x N_s zgemm on the doubled [g, i g]
// layout; here O is never formed.  With the Khatri-Rao structure O_n[(r,j)] = sigma_{n,r} tau_{n,j}
//
//   A[(r,j),(r',l)] = alpha * sum_n s^{rr'}_n conj(Y_nj) Y_nl  -  kappa * conj(mu_rj) mu_r'l ,
//   s^{rr'}_n = sigma_nr sigma_nr' = +-1,   Y = sqrt(p) (.) tau  (or tau and alpha = p for uniform p),
//
// so every site pair (r,r') shares the same operand tiles of Y and differs only by a sign stream
// (bit-packed, transposed spins).  Symmetries used: G^{rr'} = G^{r'r} and G^{rr'} Hermitian in
// (j,l): only r <= r' and tile pairs jb >= lb are computed (1/4 of the P_c^2 complex products,
// 2 N_s P_c^2 real flop); the epilogue scatters each value to its 2 or 4 images of the full
// Hermitian A (row-major, A[a][b] = <conj(O_a) O_b>_c).
//
// Kernel: CTA tile TS x TS complex (TS = 64: 8x2 warps of 8x32; TS = 80: 10x2 warps of 8x40 -> 16/20 consumer
// warps, a multiple of the 4 SM sub-partitions so that the DMMA pipe of every sub-partition is equally
// loaded) + one producer warp.  K (= samples) is streamed in stages of 16 rows through a 4-slot shared-
// memory ring: the producer issues one TMA-engine bulk copy (cp.async.bulk, SASS UBLKCP) per tile row and
// arms a "full" mbarrier with the byte count; consumers wait on it, and release the slot through an
// "empty" mbarrier -- no CTA-wide barrier and no per-thread address arithmetic in the K loop.  Per k4 step
// a consumer warp loads interleaved (re,im) fragments with LDS.128, applies the sign stream by XOR on the
// sign bit (ALU pipe, not the tensor pipe) and issues 4 DMMA m8n8k4 per 8x8 complex block
// (rr, ii -> Re;  ri, -ir -> Im), accumulating in registers.
#include "common.cuh"

namespace {

constexpr int G_KC = 16;     // samples per pipeline stage
constexpr int G_STAGES = 4;  // smem ring depth

__device__ double2 g_zero_row[128];   // zero source for out-of-range sample rows (static storage is zero-initialised)

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
// TMA engine, 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
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
__device__ __forceinline__ double xor_sign(double v, unsigned mask) {
  return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
}

__device__ __forceinline__ void tri_decode(long long p, int& hi, int& lo) {  // p = hi(hi+1)/2 + lo, lo <= hi
  long long h = (long long)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
  while ((h + 1) * (h + 2) / 2 <= p) ++h;
  while (h * (h + 1) / 2 > p) --h;
  hi = (int)h;
  lo = (int)(p - h * (h + 1) / 2);
}

// WM x WN consumer warps, each owning WBM x WBN complex 8x8 blocks, plus ONE producer warp that feeds
// the smem ring with bulk copies; consumers never meet at a CTA-wide barrier inside the K loop.
template <int WBM, int WBN, int WM, int WN>
__global__ void __launch_bounds__(32 * (WM * WN + 1), 1)
gram_s_kernel(const cplx* __restrict__ Y, long long B, int M, int R, const uint32_t* __restrict__ sigT,
              long long words, const cplx* __restrict__ mu, double alpha, double kappa, cplx* __restrict__ A) {
  constexpr int TS = 8 * WBM * WM;
  static_assert(TS == 8 * WBN * WN, "square CTA tile");
  static_assert(TS <= 128, "g_zero_row");
  constexpr int LD = TS + 2;  // row pitch in complex elements, == 2 (mod 8): conflict-free LDS.128
  constexpr int NCW = WM * WN;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cplx* tiles = reinterpret_cast<cplx*>(smem_raw);  // [stage][2][G_KC][LD]
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)G_STAGES * 2 * G_KC * LD);
  uint64_t* empty = full + G_STAGES;

  int r1, r0, jb, lb;
  tri_decode(blockIdx.x, r1, r0);   // r0 <= r1
  tri_decode(blockIdx.y, jb, lb);   // lb <= jb
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int colA0 = jb * TS, colB0 = lb * TS;
  const int KT = (int)((B + G_KC - 1) / G_KC);

  // zero the ring once (columns beyond M are never overwritten), then set up the barriers
  for (int e = threadIdx.x; e < G_STAGES * 2 * G_KC * LD; e += blockDim.x) tiles[e] = cmk(0.0, 0.0);
  if (threadIdx.x == 0) {
    for (int sidx = 0; sidx < G_STAGES; ++sidx) { mbar_init(full + sidx, 1); mbar_init(empty + sidx, NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy zero fill vs async-proxy bulk writes
  __syncthreads();

  if (warp == NCW) {
    // ===== producer warp: one bulk copy per (tile, sample row), lanes in parallel =====
    const unsigned bytesA = (unsigned)(min(TS, M - colA0) * (int)sizeof(cplx));
    const unsigned bytesB = (unsigned)(min(TS, M - colB0) * (int)sizeof(cplx));
    for (int kt = 0; kt < KT; ++kt) {
      const int slot = kt % G_STAGES;
      if (kt >= G_STAGES) mbar_wait(empty + slot, (unsigned)((kt / G_STAGES - 1) & 1));
      cplx* st = tiles + (size_t)slot * 2 * G_KC * LD;
      if (lane == 0) mbar_expect_tx(full + slot, (unsigned)G_KC * (bytesA + bytesB));
      __syncwarp();
      const long long n0 = (long long)kt * G_KC;
      for (int e = lane; e < 2 * G_KC; e += 32) {
        const int which = e / G_KC, k = e - which * G_KC;
        const long long n = n0 + k;
        const cplx* src = (n < B) ? (Y + n * M + (which ? colB0 : colA0)) : reinterpret_cast<const cplx*>(g_zero_row);
        bulk_g2s(st + ((size_t)which * G_KC + k) * LD, src, which ? bytesB : bytesA, full + slot);
      }
    }
    return;
  }

  // ===== consumer warps =====
  const int wm = warp / WN, wn = warp % WN;
  const int q4 = lane & 3, q8 = lane >> 2;
  const uint32_t* sg0 = sigT + (size_t)r0 * words;
  const uint32_t* sg1 = sigT + (size_t)r1 * words;
  double cre[WBM][WBN][2], cim[WBM][WBN][2];
#pragma unroll
  for (int a = 0; a < WBM; ++a)
#pragma unroll
    for (int b = 0; b < WBN; ++b) { cre[a][b][0] = cre[a][b][1] = 0.0; cim[a][b][0] = cim[a][b][1] = 0.0; }

  // sign words are fetched one stage ahead so that their L2 latency never sits in front of the DMMAs
  uint32_t xnext = (KT > 0) ? (sg0[0] ^ sg1[0]) : 0u;
  for (int kt = 0; kt < KT; ++kt) {
    const int slot = kt % G_STAGES;
    const cplx* As = tiles + (size_t)slot * 2 * G_KC * LD;
    const cplx* Bs = As + (size_t)G_KC * LD;
    const long long n0 = (long long)kt * G_KC;
    const uint32_t x = xnext;
    if (kt + 1 < KT) xnext = sg0[(n0 + G_KC) >> 5] ^ sg1[(n0 + G_KC) >> 5];
    const int sh = (int)(n0 & 31);
    mbar_wait(full + slot, (unsigned)((kt / G_STAGES) & 1));
#pragma unroll
    for (int ks = 0; ks < G_KC / 4; ++ks) {
      const int kk = ks * 4 + q4;
      const unsigned mask = ((x >> (sh + kk)) & 1u) << 31;
      double ar[WBM], ai[WBM], nai[WBM], br[WBN], bi[WBN];
#pragma unroll
      for (int a = 0; a < WBM; ++a) {
        cplx v = As[(size_t)kk * LD + (wm * WBM + a) * 8 + q8];
        ar[a] = xor_sign(v.x, mask);
        ai[a] = xor_sign(v.y, mask);
        nai[a] = xor_sign(v.y, mask ^ 0x80000000u);
      }
#pragma unroll
      for (int b = 0; b < WBN; ++b) {
        cplx v = Bs[(size_t)kk * LD + (wn * WBN + b) * 8 + q8];
        br[b] = v.x;
        bi[b] = v.y;
      }
#pragma unroll
      for (int a = 0; a < WBM; ++a)
#pragma unroll
        for (int b = 0; b < WBN; ++b) {
          dmma884(cre[a][b][0], cre[a][b][1], ar[a], br[b]);
          dmma884(cim[a][b][0], cim[a][b][1], ar[a], bi[b]);
          dmma884(cre[a][b][0], cre[a][b][1], ai[a], bi[b]);
          dmma884(cim[a][b][0], cim[a][b][1], nai[a], br[b]);
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + slot);   // this warp is done with the slot
  }

  // epilogue: scatter to the Hermitian images
  const long long Pc = (long long)R * M;
  const bool diagTile = (jb == lb);
  const bool samePair = (r0 == r1);
#pragma unroll
  for (int a = 0; a < WBM; ++a)
#pragma unroll
    for (int b = 0; b < WBN; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = colA0 + (wm * WBM + a) * 8 + q8;
        const int l = colB0 + (wn * WBN + b) * 8 + 2 * q4 + e;
        if (j >= M || l >= M) continue;
        if (diagTile && l > j) continue;
        double gr = alpha * cre[a][b][e];
        double gi = (diagTile && l == j) ? 0.0 : alpha * cim[a][b][e];
        const long long a0 = (long long)r0 * M + j, b1 = (long long)r1 * M + l;
        cplx m0j = mu ? mu[a0] : cmk(0.0, 0.0);
        cplx m1l = mu ? mu[b1] : cmk(0.0, 0.0);
        // A[(r0,j),(r1,l)] = G - kappa conj(mu_r0j) mu_r1l
        cplx v = cmk(gr - kappa * (m0j.x * m1l.x + m0j.y * m1l.y), gi - kappa * (m0j.x * m1l.y - m0j.y * m1l.x));
        if (samePair && l == j) v.y = 0.0;   // diagonal of a Hermitian matrix (x*y - y*x need not cancel under FMA)
        A[a0 * Pc + b1] = v;
        if (!(samePair && l == j)) A[b1 * Pc + a0] = cconj(v);
        if (!samePair && l != j) {   // (l == j: the r0<->r1 images coincide with the two above)
          const long long a1 = (long long)r1 * M + j, b0 = (long long)r0 * M + l;
          cplx m1j = mu ? mu[a1] : cmk(0.0, 0.0);
          cplx m0l = mu ? mu[b0] : cmk(0.0, 0.0);
          cplx w = cmk(gr - kappa * (m1j.x * m0l.x + m1j.y * m0l.y), gi - kappa * (m1j.x * m0l.y - m1j.y * m0l.x));
          A[a1 * Pc + b0] = w;
          A[b0 * Pc + a1] = cconj(w);
        }
      }
}

template <int WBM, int WBN, int WM, int WN>
int launch_gram(const cplx* Y, long long B, int M, int R, const uint32_t* sigT, long long words, const cplx* mu,
                double alpha, double kappa, cplx* A, cudaStream_t st) {
  constexpr int TS = 8 * WBM * WM;
  constexpr int LD = TS + 2;
  size_t smem = (size_t)G_STAGES * 2 * G_KC * LD * sizeof(cplx) + 2 * G_STAGES * sizeof(uint64_t);
  auto kern = gram_s_kernel<WBM, WBN, WM, WN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int nT = (M + TS - 1) / TS;
  long long pairs = (long long)R * (R + 1) / 2;
  long long tpairs = (long long)nT * (nT + 1) / 2;
  if (tpairs > 65535 || pairs > 2147483647LL) return JVMC_ERR_UNSUPPORTED;
  dim3 grid((unsigned)pairs, (unsigned)tpairs);
  kern<<<grid, 32 * (WM * WN + 1), smem, st>>>(Y, B, M, R, sigT, words, mu, alpha, kappa, A);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

}  // namespace

// A: [R*M, R*M] complex128 row-major, fully populated Hermitian.  tile = 0 auto, 64 or 80.
extern "C" int jvmc_rbm_gram_S(const double* Y, long long B, int M, int R, const unsigned int* sigT,
                               const double* mu, double alpha, double kappa, double* A, int tile, void* stream) {
  if (!sigT || !A || B < 0 || M <= 0 || R <= 0 || (B > 0 && !Y)) return JVMC_ERR_ARG;
  long long words = (B + 31) / 32;
  if (tile == 0) tile = (M % 80 == 0 || M == 40) ? 80 : 64;
  cudaStream_t st = (cudaStream_t)stream;
  if (tile == 80)
    return launch_gram<1, 5, 10, 2>((const cplx*)Y, B, M, R, sigT, words, (const cplx*)mu, alpha, kappa, (cplx*)A, st);
  if (tile == 64)
    return launch_gram<1, 4, 8, 2>((const cplx*)Y, B, M, R, sigT, words, (const cplx*)mu, alpha, kappa, (cplx*)A, st);
  return JVMC_ERR_ARG;
}
