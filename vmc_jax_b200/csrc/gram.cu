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
// Kernel: CTA tile TS x TS complex (TS = 64: 2x2 warps of 32x32; TS = 80: 5x2 warps of 16x40),
// K (= samples) streamed in stages of 16 rows through a 3-stage cp.async ring; per k4 step each
// warp loads interleaved (re,im) fragments with LDS.128, applies the sign stream by XOR on the
// sign bit (ALU pipe, not the fp64 pipe) and issues 4 DMMA m8n8k4 per 8x8 complex block
// (rr, ii -> Re;  ri, -ir -> Im), accumulating in registers.
#include "common.cuh"

namespace {

constexpr int G_KC = 16;
constexpr int G_STAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int srcBytes) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(srcBytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

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

template <int WBM, int WBN, int WM, int WN>
__global__ void __launch_bounds__(32 * WM * WN, 1)
gram_s_kernel(const cplx* __restrict__ Y, long long B, int M, int R, const uint32_t* __restrict__ sigT,
              long long words, const cplx* __restrict__ mu, double alpha, double kappa, cplx* __restrict__ A) {
  constexpr int TS = 8 * WBM * WM;
  static_assert(TS == 8 * WBN * WN, "square CTA tile");
  constexpr int LD = TS + 2;  // row pitch in complex elements, == 2 (mod 8): conflict-free LDS.128
  constexpr int THREADS = 32 * WM * WN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* tiles = reinterpret_cast<cplx*>(smem_raw);  // [stage][2][G_KC][LD]

  int r1, r0, jb, lb;
  tri_decode(blockIdx.x, r1, r0);   // r0 <= r1
  tri_decode(blockIdx.y, jb, lb);   // lb <= jb
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / WN, wn = warp % WN;
  const int q4 = lane & 3, q8 = lane >> 2;
  const int colA0 = jb * TS, colB0 = lb * TS;
  const uint32_t* sg0 = sigT + (size_t)r0 * words;
  const uint32_t* sg1 = sigT + (size_t)r1 * words;

  double cre[WBM][WBN][2], cim[WBM][WBN][2];
#pragma unroll
  for (int a = 0; a < WBM; ++a)
#pragma unroll
    for (int b = 0; b < WBN; ++b) { cre[a][b][0] = cre[a][b][1] = 0.0; cim[a][b][0] = cim[a][b][1] = 0.0; }

  const int KT = (int)((B + G_KC - 1) / G_KC);
  auto issue = [&](int kt) {
    if (kt < KT) {
      cplx* st = tiles + (size_t)(kt % G_STAGES) * 2 * G_KC * LD;
      const long long n0 = (long long)kt * G_KC;
      for (int e = threadIdx.x; e < 2 * G_KC * TS; e += THREADS) {
        int which = e / (G_KC * TS);
        int rem = e - which * (G_KC * TS);
        int k = rem / TS, c = rem - k * TS;
        int col = (which ? colB0 : colA0) + c;
        long long n = n0 + k;
        bool ok = (n < B) && (col < M);
        const cplx* src = ok ? (Y + n * M + col) : Y;
        cp_async16(st + ((size_t)which * G_KC + k) * LD + c, src, ok ? 16 : 0);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int sidx = 0; sidx < G_STAGES - 1; ++sidx) issue(sidx);

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();
    issue(kt + G_STAGES - 1);
    const cplx* As = tiles + (size_t)(kt % G_STAGES) * 2 * G_KC * LD;
    const cplx* Bs = As + (size_t)G_KC * LD;
    const long long n0 = (long long)kt * G_KC;
    const uint32_t x = sg0[n0 >> 5] ^ sg1[n0 >> 5];
    const int sh = (int)(n0 & 31);
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
          dmma884(cre[a][b][0], cre[a][b][1], ai[a], bi[b]);
          dmma884(cim[a][b][0], cim[a][b][1], ar[a], bi[b]);
          dmma884(cim[a][b][0], cim[a][b][1], nai[a], br[b]);
        }
    }
  }
  cp_async_wait<0>();

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
  size_t smem = (size_t)G_STAGES * 2 * G_KC * LD * sizeof(cplx);
  auto kern = gram_s_kernel<WBM, WBN, WM, WN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int nT = (M + TS - 1) / TS;
  long long pairs = (long long)R * (R + 1) / 2;
  long long tpairs = (long long)nT * (nT + 1) / 2;
  if (tpairs > 65535 || pairs > 2147483647LL) return JVMC_ERR_UNSUPPORTED;
  dim3 grid((unsigned)pairs, (unsigned)tpairs);
  kern<<<grid, 32 * WM * WN, smem, st>>>(Y, B, M, R, sigT, words, mu, alpha, kappa, A);
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
    return launch_gram<2, 5, 5, 2>((const cplx*)Y, B, M, R, sigT, words, (const cplx*)mu, alpha, kappa, (cplx*)A, st);
  if (tile == 64)
    return launch_gram<4, 4, 2, 2>((const cplx*)Y, B, M, R, sigT, words, (const cplx*)mu, alpha, kappa, (cplx*)A, st);
  return JVMC_ERR_ARG;
}
