// jvmc_rbm_gram_S_i8 -- the quantum Fisher matrix S on the 5th-generation (tcgen05) INT8 tensor cores with an
// Ozaki-style error-free splitting: fp64-equivalent accuracy (<= 1e-11 of the column scales) at several times the
// fp64 DMMA rate.  Same contract as jvmc_rbm_gram_S (gram.cu); replaces SampledObs.covar() on the gradients
// (reference jVMC/stats.py:52-58,235-245 via jVMC/util/tdvp.py:142).
//
//   A[(r,j),(r',l)] = alpha sum_n s_n conj(Y_nj) Y_nl - kappa conj(mu_rj) mu_r'l ,   s_n = sigma_nr sigma_nr'
//
// 1. jvmc_i8_slice: every real column z of Z = [Re Y_j, Im Y_j] (interleaved, 2M columns) gets the scale
//    c_z = 2 max_n |Z_nz| (1 + 1/128) and each entry is split into 5 base-256 digits, TWICE:
//      B encoding (column operand, never negated): Z_nz = c_z sum_k b_k 256^-k,        b_k in [-128, 127] integers;
//      A encoding (row operand, gets the signs):   Z_nz = c_z sum_k (a_k + 1/2) 256^-k, a_k in [-128, 127], i.e.
//      half-integer digits -127.5 .. 127.5: the negative of the digit stored as byte a is stored as ~a (one's
//      complement), so applying s_n = -1 to 4 samples of a row is ONE xor of a 32-bit word with a byte mask -- no
//      carries, no SWAR arithmetic (the two's-complement digits of round 1 needed ~6 ALU instructions per word and
//      made the eight sign warps, not the tensor pipe, the bottleneck: tools/gram_trace.py).
//    Both are exact to 2^-41 of the column scale.  Digits are stored in the K-major SWIZZLE_NONE UMMA canonical layout
//    [n/32][digit][z/8][(n/16)%2][z%8][n%16], so that a (tile, 32-sample stage, digit) is one contiguous block (one TMA
//    bulk copy each).  Padding samples are all-zero bytes in both encodings (a' = 0, b = 0: no contribution).
// 2. gram_s_i8_kernel: CTA = (site pair r<=r', tile): 128 real rows (64 complex j) x NC <= 80 real columns (the tile list
//    comes from the host: kernels.py:i8_tile_list; tiles that end at the diagonal are narrower).  Clusters of two CTAs
//    work on consecutive site pairs of the same tile and share every operand copy by TMA multicast.  Per 32-sample
//    stage the producer warp bulk-copies (TMA engine) the 5 A-digit and 5 B-digit tiles into an mbarrier-guarded
//    6-slot smem ring; eight "sign" warps xor the A digits with the byte masks of s_n (256-entry lookup table in shared
//    memory) and store them into a double-buffered TMEM A operand (tcgen05.st); warp 1 issues 7 tcgen05.mma kind::i8
//    (M=128, N = NC n, K=32; SASS UTCIMMA; A from TMEM, B from smem): digit pair (k,k') accumulates exactly in int32 into
//    the TMEM accumulator of level t = k+k', and because the B digit tiles are contiguous in smem and the level
//    accumulators adjacent TMEM column blocks, ONE instruction covers n consecutive digit pairs (5 levels x 80 columns
//    + 2 x 40 A-operand columns = 480 of the 512 TMEM columns).
//    sum_n s_n (a_n + 1/2) b_n = sum_n a'_n b_n + 1/2 sum_n b_n: the second term does not depend on the site pair; it is
//    formed once per launch from the column sums of the B digits (i8_colsum / i8_corr kernels) and added per level in
//    the epilogue.  One launch covers at most 25600 samples (int32 head-room 5 * 128^2 * 25600 < 2^31); at its end four
//    epilogue warps read TMEM (tcgen05.ld), weight level t by 256^-t in fp64, combine real/imag parts with a lane
//    shuffle, apply scales, alpha and the mean correction and scatter the Hermitian images (later launches add into A).
//    Dropped digit pairs (k+k' > 6) are below 4 * 256^-5 / 4 of c_z c_z' per sample.
//    Compile-time knobs for experiments: JVMC_I8_TN (full tile width), JVMC_I8_NB (TMEM A buffers).
#include "common.cuh"

namespace {

constexpr int I8_S = 5;                 // digits
constexpr int I8_LEV = 5;               // levels t = 2..6
#ifndef JVMC_I8_TN
#define JVMC_I8_TN 80
#endif
#ifndef JVMC_I8_NB
#define JVMC_I8_NB 2
#endif
constexpr int I8_NB = JVMC_I8_NB;       // TMEM A-operand buffers
#ifndef JVMC_I8_SIGNWARPS
#define JVMC_I8_SIGNWARPS 8
#endif
constexpr int I8_SW = JVMC_I8_SIGNWARPS;  // 8: two warps per TMEM lane quarter (digits 1-3 / 4-5); 4: one warp, all digits
constexpr int I8_KW = (I8_SW == 8) ? 3 : 5;   // digits per sign warp (max)
constexpr int I8_TM = 128, I8_TN = JVMC_I8_TN;  // full tile in real columns
constexpr int I8_KS = 32;               // samples per stage (one MMA K)
constexpr int I8_SLOTS = 6;
constexpr int I8_A_BYTES = I8_TM * I8_KS;   // per digit
constexpr int I8_B_BYTES = I8_TN * I8_KS;
constexpr int I8_STAGE_BYTES = I8_S * (I8_A_BYTES + I8_B_BYTES);
constexpr int I8_MAXSTAGES = 800;       // stages per launch: 800*32*5*128^2 < 2^31 (int32 head-room in TMEM)
constexpr int I8_DEFAULT_STAGES = 800;  // default: as many as the head-room allows (see jvmc_rbm_gram_S_i8)
constexpr int I8_THREADS = 96 + 32 * I8_SW;   // warp 0 A producer, 1 MMA issuer, 2-9 sign warps (2-5 also epilogue), 10 B producer
constexpr int I8_BPROD_WARP = 2 + I8_SW;
constexpr int I8_ACOL = I8_LEV * I8_TN;   // first TMEM column of the A operand buffers (I8_NB x 5 digits x 8 columns)
static_assert(I8_LEV * JVMC_I8_TN + JVMC_I8_NB * 5 * 8 <= 512, "TMEM columns");

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
// arrive on the barrier at the same shared-memory offset of CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, unsigned rank) {
  unsigned raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  // relaxed: the shared-memory loads this arrive stands for have returned their data already (it was consumed), so
  // no release fence over the cluster is needed (a release.cluster arrive cost ~1500 cycles here)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(raddr) : "memory");
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
// non-blocking probe (the potentially-blocking try_wait costs ~150-200 cycles even when the phase completed long ago)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// multicast variant: the bytes land at the same shared-memory offset of every CTA in ctaMask and complete_tx the
// mbarrier at the same offset of each of them
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, unsigned bytes, uint64_t* bar, unsigned short mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// K-major SWIZZLE_NONE UMMA shared-memory descriptor (canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units)
__device__ __forceinline__ uint64_t umma_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// A operand from TMEM (lane = row, 8 columns = 32 K-bytes), B from shared memory
__device__ __forceinline__ void umma_i8_ts(uint32_t tacc, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tacc), "r"(ta), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// arrive (once the preceding MMAs retire) on the mbarrier at this offset in every CTA of ctaMask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, unsigned short mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tri_decode(long long p, int& hi, int& lo) {
  long long h = (long long)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
  while ((h + 1) * (h + 2) / 2 <= p) ++h;
  while (h * (h + 1) / 2 > p) --h;
  hi = (int)h;
  lo = (int)(p - h * (h + 1) / 2);
}

// ------------------------------------------------------------------------------------------ slicing
// column maxima of |Re|, |Im| : colmax[2j], colmax[2j+1] as bit patterns (non-negative doubles order like uint64)
__global__ void i8_colmax_kernel(const cplx* __restrict__ Y, long long B, int M, unsigned long long* __restrict__ colmax) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const long long n0 = (long long)blockIdx.y * 1024;
  const long long n1 = min(B, n0 + 1024);
  double mr = 0.0, mi = 0.0;
  for (long long n = n0; n < n1; ++n) {
    cplx v = Y[n * M + j];
    mr = fmax(mr, fabs(v.x));
    mi = fmax(mi, fabs(v.y));
  }
  atomicMax(colmax + 2 * j, (unsigned long long)__double_as_longlong(mr));
  atomicMax(colmax + 2 * j + 1, (unsigned long long)__double_as_longlong(mi));
}

// scale[z] = 2 max_n |Z_nz| (1 + 1/128): |Z/scale| <= 0.4962 keeps both digit encodings inside [-128, 127] (the
// half-integer encoding is shifted by (256^5 - 1)/510 ~ 0.002 * 256^5); 1 for empty columns
__global__ void i8_scale_kernel(const unsigned long long* __restrict__ colmax, int twoM, double* __restrict__ scale) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= twoM) return;
  double m = __longlong_as_double((long long)colmax[z]);
  scale[z] = (m > 0.0) ? 2.0 * m * (1.0 + 0.0078125) : 1.0;
}

// Heavy-tail diagnostic: per real column sum_n Z_nz^2 and max_n |Z_nz| (scratch zeroed by the caller), then
// ratio[z] = max_n|Z_nz| / rms_n(Z_nz) (0 for an all-zero column).  The digit resolution is relative to the column
// maximum: the error of A_zz' relative to its natural size sqrt(A_zz A_z'z') grows like ratio_z ratio_z' / sqrt(B).
__global__ void i8_colsq_kernel(const cplx* __restrict__ Y, long long B, int M, double* __restrict__ colsq,
                                unsigned long long* __restrict__ colmax) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const long long n0 = (long long)blockIdx.y * 1024;
  const long long n1 = min(B, n0 + 1024);
  double sr = 0.0, si = 0.0, mr = 0.0, mi = 0.0;
  for (long long n = n0; n < n1; ++n) {
    cplx v = Y[n * M + j];
    sr = fma(v.x, v.x, sr); si = fma(v.y, v.y, si);
    mr = fmax(mr, fabs(v.x)); mi = fmax(mi, fabs(v.y));
  }
  atomicAdd(colsq + 2 * j, sr);
  atomicAdd(colsq + 2 * j + 1, si);
  atomicMax(colmax + 2 * j, (unsigned long long)__double_as_longlong(mr));
  atomicMax(colmax + 2 * j + 1, (unsigned long long)__double_as_longlong(mi));
}
__global__ void i8_tail_ratio_kernel(const double* __restrict__ colsq, const unsigned long long* __restrict__ colmax,
                                     int twoM, double invB, double* __restrict__ ratio) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= twoM) return;
  const double ms = colsq[z] * invB;
  ratio[z] = ms > 0.0 ? __longlong_as_double((long long)colmax[z]) / sqrt(ms) : 0.0;
}

// flag[n] = 1 if any entry of sample n exceeds T times the rms of its real column (one warp per sample)
__global__ void i8_rowflag_kernel(const cplx* __restrict__ Y, long long B, int M, const double* __restrict__ colsq, double invB,
                                  double T, unsigned char* __restrict__ flag) {
  const long long n = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= B) return;
  const double* Yd = reinterpret_cast<const double*>(Y) + n * 2 * M;
  const double T2 = T * T * invB;
  bool hit = false;
  for (int z = lane; z < 2 * M; z += 32) {
    const double v = Yd[z];
    hit |= v * v > T2 * colsq[z];
  }
  const unsigned any = __ballot_sync(0xffffffffu, hit);
  if (lane == 0) flag[n] = any ? 1 : 0;
}

// one thread per (16-sample chunk, real column z): 16 strided loads, 2 x 5 x 16 B stores (A and B encodings)
__global__ void __launch_bounds__(256)
i8_slice_kernel(const cplx* __restrict__ Y, long long B, int M, const double* __restrict__ scale, long long numChunks,
                int numZGroups, int8_t* __restrict__ digA, int8_t* __restrict__ digB) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;     // fastest: coalesced over columns
  const long long ch = blockIdx.y;
  if (z >= 2 * M) return;
  const double inv = 1.0 / scale[z];
  const double* Yd = reinterpret_cast<const double*>(Y);
  int8_t da[I8_S][16], db[I8_S][16];
  for (int q = 0; q < 16; ++q) {
    const long long n = ch * 16 + q;
    if (n < B) {
      const double X = Yd[n * 2 * M + z] * inv * 1099511627776.0;     // 256^5 = 2^40: exact scaling
      long long NB = llrint(X);                                      // integer digits
      long long NA = llrint(X - 2155905152.5);                       // half-integer digits: minus (256^5 - 1) / 510
#pragma unroll
      for (int k = I8_S - 1; k >= 0; --k) {
        const int8_t b8 = (int8_t)(NB & 0xFF), a8 = (int8_t)(NA & 0xFF);   // balanced remainders in [-128, 127]
        db[k][q] = b8; da[k][q] = a8;
        NB = (NB - b8) >> 8; NA = (NA - a8) >> 8;
      }
    } else {
#pragma unroll
      for (int k = 0; k < I8_S; ++k) { da[k][q] = 0; db[k][q] = 0; }  // padding sample: contributes nothing
    }
  }
  // [stage][digit][zgroup][chunk of the stage][z%8][n%16]: a (tile, stage, digit) is one contiguous block
  const size_t off = ((size_t)(z >> 3) * 2 + (size_t)(ch & 1)) * 128 + (size_t)(z & 7) * 16;
#pragma unroll
  for (int k = 0; k < I8_S; ++k) {
    const size_t o = ((size_t)(ch >> 1) * I8_S + k) * numZGroups * 256 + off;
    *reinterpret_cast<int4*>(digA + o) = *reinterpret_cast<const int4*>(da[k]);
    *reinterpret_cast<int4*>(digB + o) = *reinterpret_cast<const int4*>(db[k]);
  }
}

// SB[k][z] += sum over the samples of stages [stage0, stage1) of the B digit k of column z (exact, int32):
// thread = (digit, column), blockIdx.y = slice of 32 stages
__global__ void i8_colsum_kernel(const int8_t* __restrict__ digB, int numZGroups, long long stage0, long long stage1,
                                 int* __restrict__ SB) {
  const int ZP = numZGroups * 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= I8_S * ZP) return;
  const int k = idx / ZP, z = idx - k * ZP;
  const long long s0 = stage0 + (long long)blockIdx.y * 32, s1 = min(stage1, s0 + 32);
  int acc = 0;
  for (long long st = s0; st < s1; ++st) {
    const int8_t* p = digB + ((size_t)st * I8_S + k) * numZGroups * 256 + (size_t)(z >> 3) * 256 + (size_t)(z & 7) * 16;
#pragma unroll
    for (int chunk = 0; chunk < 2; ++chunk) {
      const int4 v = *reinterpret_cast<const int4*>(p + chunk * 128);
      acc = __dp4a(v.x, 0x01010101, acc); acc = __dp4a(v.y, 0x01010101, acc);
      acc = __dp4a(v.z, 0x01010101, acc); acc = __dp4a(v.w, 0x01010101, acc);
    }
  }
  if (acc) atomicAdd(SB + idx, acc);
}
// corr[t][z] = 1/2 sum_{k' : 1 <= t - k' <= 5} SB[k'][z], levels t = 2..6 (digits k' = 1..5)
__global__ void i8_corr_kernel(const int* __restrict__ SB, int ZP, double* __restrict__ corr) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= I8_LEV * ZP) return;
  const int t = idx / ZP + 2, z = idx - (t - 2) * ZP;
  long long acc = 0;
  for (int kp = 1; kp <= I8_S; ++kp)
    if (t - kp >= 1 && t - kp <= I8_S) acc += SB[(kp - 1) * ZP + z];
  corr[idx] = 0.5 * (double)acc;
}

// ------------------------------------------------------------------------------------------ main kernel
struct I8Args {
  const int8_t* dig;       // A encoding (half-integer digits)
  const int8_t* digB;      // B encoding (integer digits)
  const double* corr;      // [5 levels][numZGroups * 8]: 1/2 sum_n b_n per level and column, this launch's samples
  long long numChunks;     // 16-sample chunks (even)
  int numZGroups;          // padded real columns / 8
  const double* scale;     // [2M]
  const uint32_t* sigT;    // [R][words]
  long long words;
  const int* tiles;        // 8 ints per tile: first row group, first column group, real columns NC (multiple of 16,
                           // <= I8_TN), jlo, jhi (rows written: jlo <= j < jhi), 3 unused
  const cplx* mu;
  double alpha, kappa;
  cplx* A;
  int M, R;
  long long stage0, stage1;   // this launch covers sample stages [stage0, stage1), at most I8_MAXSTAGES
  int accumulate;             // 0: A = alpha G - kappa mu^H mu ; 1: A += alpha G
  int dbg;                    // development ablations: 1 = MMA issue only, 2 = TMA + MMA (no sign pass), 4 = no TMEM store,
                              // 32 = no shared-memory reads in the sign pass, 64 = no sign arithmetic
  int cl;                     // thread-block cluster size along the pair axis (operand tiles are TMA-multicast)
  long long pairs;            // R (R + 1) / 2 site pairs in total
  long long pair0, npairs;    // this launch covers pairs [pair0, pair0 + npairs) in the order r0 ascending, r1 = r0 .. R-1
                              // (CTAs beyond it only pad the last cluster)
  long long* trace;           // development: clock64 time stamps of one CTA, 8 events per stage (jvmc_i8_set_trace)
  int traceTile;
};
#define I8_TRACE(ev) do { if (tracing) a.trace[(size_t)g * 16 + (ev)] = clock64(); } while (0)
// CTA-level phases go to the row of (unused) stage 831: 0 entry, 1 set-up done, 2 last MMA committed, 3 accumulators
// complete (seen by the epilogue), 4 epilogue done, 5 after the final CTA / cluster barriers
#define I8_TRACE_CTA(ev) do { if (tracing) a.trace[(size_t)831 * 16 + (ev)] = clock64(); } while (0)

__global__ void __launch_bounds__(I8_THREADS, 1) gram_s_i8_kernel(I8Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;                                                     // [slot][A digits | B digits]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)I8_SLOTS * I8_STAGE_BYTES);
  // The A and the B half of a ring slot are recycled independently: the A digits are free as soon as the sign warps
  // (of every CTA of the cluster: the copies are multicast) hold them in registers, ~2 stages before the MMAs that
  // read the B digits of the same stage retire.  With one barrier pair per slot the A copies were issued only ~2000
  // cycles before their use and every L2 miss (~1200 instead of ~900 cycles) stalled the sign warps: 20 % of the
  // stages took 1100 instead of 700 cycles (trace).
  uint64_t* fullA = bars;                   // A digits of the slot landed
  uint64_t* emptyA = bars + I8_SLOTS;       // sign warps of all CTAs of the cluster have loaded them
  uint64_t* full = bars + 2 * I8_SLOTS;     // B digits of the slot landed
  uint64_t* empty = bars + 3 * I8_SLOTS;    // MMAs reading them retired (in every CTA of the cluster)
  uint64_t* aready = bars + 4 * I8_SLOTS;   // [2] signed A digits of a stage are in TMEM buffer b
  uint64_t* afree = aready + I8_NB;         // [NB] MMAs reading TMEM buffer b retired
  uint64_t* accfull = afree + I8_NB;
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(accfull + 1);
  uint2* lut = reinterpret_cast<uint2*>(bars + 32);       // [256] sign bits of 8 samples -> two words of byte masks

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The CTAs of a cluster work on consecutive site pairs of the SAME tile: the raw digit tiles are identical for
  // all of them (only the signs differ), so every bulk copy is multicast to the whole cluster and the L2 -> SM
  // operand traffic (the binding resource of the single-CTA version) drops by the cluster size.
  const unsigned crank = (a.cl > 1) ? cluster_ctarank() : 0u;
  const unsigned short cmask = (unsigned short)((1u << a.cl) - 1u);
  const bool padCta = (long long)blockIdx.x >= a.npairs;
  int r1, r0;
  {
    // pair p <-> (r0 ascending, r1 >= r0): block row r0 of the upper block triangle of A is complete once the pairs of
    // r0 are done, so a launch over a range of pairs finishes a range of rows (overlapped reduction over ranks)
    int hi, lo;
    tri_decode(a.pairs - 1 - (padCta ? a.pair0 : a.pair0 + blockIdx.x), hi, lo);
    r0 = a.R - 1 - hi;
    r1 = a.R - 1 - lo;
  }
  const int* tile = a.tiles + 8 * (size_t)blockIdx.y;
  // tile rows start at real column 8 RG, its NC columns at real column 8 CG; tiles that end at the diagonal are
  // narrower than I8_TN so that less of the rectangle above the diagonal is computed
  const int RG = tile[0], CG = tile[1], NC = tile[2], jlo = tile[3], jhi = tile[4];
  const unsigned bBytes = (unsigned)(NC * I8_KS);       // bytes of one B digit tile of a stage
  const long long numStages = a.stage1 - a.stage0;
  const bool tracing = a.trace != nullptr && blockIdx.x == 0 && (int)blockIdx.y == a.traceTile && lane == 0;
  if (warp == 2) I8_TRACE_CTA(0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < I8_SLOTS; ++s) {
      mbar_init(full + s, 1); mbar_init(empty + s, (unsigned)a.cl);
      mbar_init(fullA + s, 1); mbar_init(emptyA + s, (unsigned)(a.cl * (I8_SW / 2)));
    }
    for (int b = 0; b < I8_NB; ++b) { mbar_init(aready + b, I8_SW / 2); mbar_init(afree + b, 1); }   // 4 warps per buffer
    mbar_init(accfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 256) {
    for (unsigned v = threadIdx.x - 64; v < 256; v += 256) {
      const unsigned lo = v & 0xFu, hi = v >> 4;
      lut[v] = make_uint2(((lo * 0x00204081u) & 0x01010101u) * 0xFFu, ((hi * 0x00204081u) & 0x01010101u) * 0xFFu);
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_base_p)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (a.cl > 1) cluster_sync_all();   // peers' barriers are initialised before anything is multicast to them
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_base_p;
  if (warp == 2) I8_TRACE_CTA(1);

  if (warp == 0 || warp == I8_BPROD_WARP) {
    // ===================== producers: one bulk copy per digit and stage; warp 0 the A tiles, warp 10 the B tiles =====
    const bool isA = warp == 0;
    uint64_t* fullBar = isA ? fullA : full;
    uint64_t* emptyBar = isA ? emptyA : empty;
    const unsigned bytes = isA ? (unsigned)I8_A_BYTES : bBytes;
    const int8_t* srcBase = isA ? a.dig + (size_t)RG * 256 : a.digB + (size_t)CG * 256;
    long long nProd = (a.dbg & 1) ? 0 : numStages;
    if (isA && (a.dbg & 3)) nProd = 0;               // ablation modes without a sign pass: nobody consumes the A tiles
    uint32_t slot = 0, par = 0;
    for (long long g = 0; g < nProd; ++g) {
      if (g >= I8_SLOTS) mbar_wait(emptyBar + slot, par ^ 1u);
      unsigned char* st = ring + (size_t)slot * I8_STAGE_BYTES + (isA ? 0 : I8_S * I8_A_BYTES);
      if (lane == 0) mbar_expect_tx(fullBar + slot, (unsigned)I8_S * bytes);
      __syncwarp();
      if (lane < I8_S) {
        const size_t base = ((size_t)(a.stage0 + g) * I8_S + lane) * a.numZGroups;
        unsigned char* dst = st + lane * bytes;                                       // digit tiles packed
        const int8_t* src = srcBase + base * 256;
        if (a.cl == 1) bulk_g2s(dst, src, bytes, fullBar + slot);
        else if ((unsigned)lane % (unsigned)a.cl == crank) bulk_g2s_mc(dst, src, bytes, fullBar + slot, cmask);   // my share
      }
      if (isA) I8_TRACE(0);
      if (++slot == I8_SLOTS) { slot = 0; par ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp 1, convergent; one elected lane issues =====================
    // Digit tiles of B are contiguous in smem ([digit][rowgroup][chunk]), and the level accumulators are adjacent
    // TMEM column blocks, so ONE instruction with N = NC n multiplies A_k with B_k'..B_k'+n-1 and accumulates into
    // levels k+k' .. k+k'+n-1: 7 instructions per stage instead of 15.
    // The warp stays convergent and every address / descriptor is computed from warp-uniform values, so that the
    // operands live in uniform registers: issuing from a divergent `lane == 0` branch (round 1) made the compiler
    // wrap every UTCIMMA into an ELECT / R2UR / BRA.U.ANY loop, ~60 cycles per instruction, and the single-thread
    // bookkeeping between stages (64-bit g % 6, two barrier polls) left the tensor pipe idle for ~365 of ~1080
    // cycles per stage (trace: tools/gram_trace.py).  The barriers of stage g+1 are polled between the MMAs of stage
    // g; with the sign pass on, `aready` alone is waited for (the sign warps observed `full` before arriving on it).
    {
      const uint32_t NCu = (uint32_t)__shfl_sync(0xffffffffu, NC, 0);
      const uint32_t tmemU = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t ibase = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_TM >> 4) << 24);
      const uint32_t idesc1 = ibase | (((1u * NCu) >> 3) << 17);
      const uint32_t idesc2 = ibase | (((2u * NCu) >> 3) << 17);
      const uint32_t idesc3 = ibase | (((3u * NCu) >> 3) << 17);
      const bool useFull = !(a.dbg & 1);                         // B digits of the stage landed
      const bool useReady = !(a.dbg & 3);
      const uint32_t ring0 = smem_u32(ring) + I8_S * I8_A_BYTES;
      const uint32_t bStep = (NCu * (uint32_t)I8_KS) >> 4;       // one B digit tile in 16-byte units
      auto wait_stage = [&](uint32_t sl, uint32_t fp, uint32_t bb, uint32_t bp) {
        // both probes are issued before either result is consumed: their ~100-cycle latencies overlap
        const bool okF = !useFull || mbar_test(full + sl, fp);
        const bool okR = !useReady || mbar_test(aready + bb, bp);
        if (!okF) mbar_wait(full + sl, fp);
        if (!okR) mbar_wait(aready + bb, bp);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      };
      if (numStages > 0) wait_stage(0u, 0u, 0u, 0u);
      const bool leader = elect_one();
      // The loop over stages is unrolled over the six ring slots (six is even: the TMEM A buffer is slot & 1), so the
      // slot's shared-memory descriptor and every TMEM address are loop-invariant values the compiler keeps in uniform
      // registers; only the barrier phases depend on the round q.
      static_assert(I8_SLOTS % I8_NB == 0, "ring slots cycle through the TMEM A buffers");
      for (long long q = 0; q * I8_SLOTS < numStages; ++q) {
#pragma unroll
        for (uint32_t slot = 0; slot < (uint32_t)I8_SLOTS; ++slot) {
          const long long g = q * I8_SLOTS + slot;
          if (g >= numStages) break;
          I8_TRACE(2);
          const uint32_t b = slot % (uint32_t)I8_NB;
          // K-major SWIZZLE_NONE descriptor of B digit tile k': low word = (address >> 4) | LBO(128 B) << 16,
          // high word = SBO(256 B) | version bit
          const uint32_t dlo = (((ring0 + slot * (uint32_t)I8_STAGE_BYTES) >> 4) & 0x3FFFu) | ((128u >> 4) << 16);
          const uint32_t ta = tmemU + (uint32_t)I8_ACOL + b * (uint32_t)(I8_S * 8);
          // (digit k, first k', count n): level column = (k + k' - 2) * NC
          auto mma = [&](uint32_t k, uint32_t kp, uint32_t idesc, uint32_t accumulate) {
            const uint64_t db = ((uint64_t)((256u >> 4) | (1u << 14)) << 32) | (uint64_t)(dlo + (kp - 1u) * bStep);
            umma_i8_ts(tmemU + (k + kp - 2u) * NCu, ta + (k - 1u) * 8u, db, idesc, accumulate);
          };
          if (leader) {
            const uint32_t acc = (g == 0) ? 0u : 1u;
            mma(1, 1, idesc3, acc);        // levels 2,3,4 (first writer at g = 0)
            mma(1, 4, idesc2, acc);        // levels 5,6   (first writer at g = 0)
            mma(2, 1, idesc3, 1u);         // levels 3,4,5
            mma(2, 4, idesc1, 1u);         // level 6
            mma(3, 1, idesc3, 1u);         // levels 4,5,6
          }
          __syncwarp();
          I8_TRACE(1);
          if (g + 1 < numStages) {           // barriers of the next stage, polled while the MMAs above are queued
            const uint32_t nslot = (slot + 1 == (uint32_t)I8_SLOTS) ? 0u : slot + 1;
            const uint32_t nfullPar = (uint32_t)((slot + 1 == (uint32_t)I8_SLOTS) ? (q + 1) : q) & 1u;
            wait_stage(nslot, nfullPar, nslot % (uint32_t)I8_NB, (uint32_t)((g + 1) / I8_NB) & 1u);
          }
          I8_TRACE(8);
          if (leader) {
            mma(4, 1, idesc2, 1u);         // levels 5,6
            mma(5, 1, idesc1, 1u);         // level 6
            if (a.cl == 1) umma_commit(empty + slot);        // smem slot reusable once these MMAs retire
            else umma_commit_mc(empty + slot, cmask);        // ... in every CTA of the cluster (they all write into it)
            umma_commit(afree + b);                          // ... and so is the TMEM A buffer
            if (g + 1 == numStages) umma_commit(accfull);
          }
          __syncwarp();
          I8_TRACE(3);
        }
      }
      I8_TRACE_CTA(2);
    }
  } else {
    // ===================== sign warps: A_k <- s_n * A_k, smem -> registers -> TMEM =====================
    // thread = tile row; per digit the row's 32 sample bytes are two 16-byte units of the canonical layout.  The digits
    // are half-integers stored in one's complement: the sign is one xor with the byte mask of the stage's sign bits.
    // Warps 2-5 serve the even stages (TMEM buffer 0), warps 6-9 the odd ones (buffer 1), one lane quarter each and all
    // five digits: a warp's chain per stage (poll `full`, shared-memory loads, poll `afree`, TMEM store, arrive) is
    // latency- not throughput-bound (~800 cycles, trace), so every warp gets two stage periods for it.
    const int ew = warp & 3;                              // TMEM lane quarter this warp may access
    const int row = ew * 32 + lane;                       // real row of the tile (z = 8 RG + row)
    {
      static_assert(I8_NB == 2 && I8_SW == 8, "sign warps: two groups of four, one TMEM A buffer each");
      const uint32_t grp = (uint32_t)(warp - 2) >> 2;     // 0: even stages, 1: odd stages; also the TMEM buffer
      const uint32_t* sg0 = a.sigT + (size_t)r0 * a.words + a.stage0;
      const uint32_t* sg1 = a.sigT + (size_t)r1 * a.words + a.stage0;
      const uint32_t rowOff = (uint32_t)(row >> 3) * 256u + (uint32_t)(row & 7) * 16u;
      const unsigned char* aBase = ring + rowOff;
      const uint32_t tBase = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(I8_ACOL + grp * (I8_S * 8));
      const long long nSign = (a.dbg & 3) ? 0 : numStages;
      uint32_t xn = (grp < nSign) ? (sg0[grp] ^ sg1[grp]) : 0u;
      uint32_t slot = grp, fullPar = 0, freePar = 0;
      for (long long g = grp; g < nSign; g += 2) {
        const uint32_t x = xn;                              // bit = 1 -> s_n = -1 (32 samples of the stage)
        if (g + 2 < numStages) xn = sg0[g + 2] ^ sg1[g + 2];
        uint32_t msk[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint2 m = lut[(x >> (8 * q)) & 0xFFu];
          msk[2 * q] = m.x; msk[2 * q + 1] = m.y;
        }
        // The signed digits are prepared in registers BEFORE waiting for the TMEM buffer, so that only the
        // TMEM store sits on the MMA(g-2) -> sign(g) -> MMA(g) dependency chain.
        mbar_wait(fullA + slot, fullPar);
        if (warp == 2) I8_TRACE(4);
        const unsigned char* st = aBase + slot * (uint32_t)I8_STAGE_BYTES;
        uint32_t w[I8_S][8];
#pragma unroll
        for (int k = 0; k < I8_S; ++k) {
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const uint4 v = (a.dbg & 32) ? make_uint4(0u, 0u, 0u, 0u)
                                         : *reinterpret_cast<const uint4*>(st + k * I8_A_BYTES + ch * 128);
            w[k][ch * 4 + 0] = v.x ^ msk[ch * 4 + 0]; w[k][ch * 4 + 1] = v.y ^ msk[ch * 4 + 1];
            w[k][ch * 4 + 2] = v.z ^ msk[ch * 4 + 2]; w[k][ch * 4 + 3] = v.w ^ msk[ch * 4 + 3];
          }
        }
        __syncwarp();
        if (lane == 0) {                                      // the A half of the slot may be refilled
          if (a.cl == 1) mbar_arrive(emptyA + slot);
          else for (unsigned rk = 0; rk < (unsigned)a.cl; ++rk) mbar_arrive_cluster(emptyA + slot, rk);
        }
        if (warp == 2) I8_TRACE(5);
        if (g >= I8_NB) { mbar_wait(afree + grp, freePar); freePar ^= 1u; }
        if (warp == 2) I8_TRACE(6);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int k = 0; k < I8_S; ++k) {
          if (a.dbg & 4) continue;
          const uint32_t taddr = tBase + (uint32_t)(k * 8);
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr),
                       "r"(w[k][0]), "r"(w[k][1]), "r"(w[k][2]), "r"(w[k][3]), "r"(w[k][4]), "r"(w[k][5]), "r"(w[k][6]),
                       "r"(w[k][7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        if (warp == 2) I8_TRACE(7);
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(aready + grp);
        slot += 2;
        if (slot >= I8_SLOTS) { slot -= I8_SLOTS; fullPar ^= 1u; }
      }
    }
    if (warp < 6) {
    // ===================== epilogue (warps 2-5): TMEM -> fp64, combine re/im, scale, scatter =====================
    // Everything the inner loop needs per column (scales, means, level corrections) is staged in shared memory (the
    // operand ring is free once the accumulators are complete) and the row-side values sit in registers; later
    // launches add into A with fire-and-forget reductions (RED.ADD.F64) instead of load-add-store round trips -- the
    // first version of this epilogue was a chain of dependent L2 accesses, ~44 000 cycles = 8 % of a CTA (trace).
    const double wl[I8_LEV] = {1.0 / 65536.0, 1.0 / 16777216.0, 1.0 / 4294967296.0, 1.0 / 1099511627776.0,
                               1.0 / 281474976710656.0};   // 256^-t, t = 2..6
    const int ZP = a.numZGroups * 8;
    const int zrow = RG * 8 + row;
    const int j = zrow >> 1;                              // complex row index; lane parity = re/im part
    const bool isIm = (zrow & 1) != 0;
    const bool jok = j < a.M;
    const double srow = jok ? a.scale[zrow] : 0.0;
    const long long Pc = (long long)a.R * a.M;
    const bool samePair = (r0 == r1);
    const bool useMu = a.mu != nullptr && !a.accumulate;
    const cplx m0j = (useMu && jok) ? a.mu[(long long)r0 * a.M + j] : cmk(0.0, 0.0);
    const cplx m1j = (useMu && jok) ? a.mu[(long long)r1 * a.M + j] : cmk(0.0, 0.0);
    mbar_wait(accfull, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (warp == 2) I8_TRACE_CTA(3);
    double* stg = reinterpret_cast<double*>(ring);        // [80] column scales | [40] mu(r0,l) | [40] mu(r1,l) | [5][80] corr
    double* sSc = stg;
    cplx* sMu0 = reinterpret_cast<cplx*>(stg + I8_TN);
    cplx* sMu1 = sMu0 + I8_TN / 2;
    double* sCorr = stg + I8_TN + 2 * I8_TN;
    {
      const int et = (warp - 2) * 32 + lane;              // 0..127
      for (int c = et; c < I8_TN; c += 128) {
        const int zc = CG * 8 + c;
        sSc[c] = (c < NC && (zc >> 1) < a.M) ? a.scale[zc] : 0.0;
      }
      for (int i = et; i < I8_TN; i += 128) {              // first half: mu(r0, l), second half: mu(r1, l)
        const int which = i >= I8_TN / 2, li = which ? i - I8_TN / 2 : i;
        const int l = CG * 4 + li;
        const cplx m = (useMu && li < NC / 2 && l < a.M) ? a.mu[(long long)(which ? r1 : r0) * a.M + l] : cmk(0.0, 0.0);
        (which ? sMu1 : sMu0)[li] = m;
      }
      for (int i = et; i < I8_LEV * I8_TN; i += 128) {
        const int t = i / I8_TN, c = i - t * I8_TN;
        sCorr[i] = (c < NC) ? a.corr[(size_t)t * ZP + CG * 8 + c] : 0.0;
      }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
    }
    for (int cb = 0; cb < NC / 16; ++cb) {
      double v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = 0.0;
#pragma unroll
      for (int t = 0; t < I8_LEV; ++t) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(t * NC + cb * 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; ++c)   // + 1/2 sum_n b_n of the level: the half of the half-integer A digits
          v[c] = fma((double)(int32_t)r[c] + sCorr[t * I8_TN + cb * 16 + c], wl[t], v[c]);
      }
      // C[zrow][zcol] with both column scales; the partner lane (lane ^ 1) holds the other part of the same j
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const int li = cb * 8 + cc;
        const int l = CG * 4 + li;
        const double c0 = v[2 * cc] * srow * sSc[2 * li];              // column Re Y_l (scale 0 beyond M)
        const double c1 = v[2 * cc + 1] * srow * sSc[2 * li + 1];      // column Im Y_l
        const double p1 = __shfl_xor_sync(0xffffffffu, c1, 1);
        // even lane (Re row): Re G = C[2j][2l] + C[2j+1][2l+1] = c0 + p1 ; odd lane (Im row): Im G = C[2j][2l+1] - C[2j+1][2l] = p1 - c0
        const double part = isIm ? (p1 - c0) : (c0 + p1);
        const double other = __shfl_xor_sync(0xffffffffu, part, 1);   // even lane receives Im G
        if (l >= a.M || !jok || j < jlo || j >= jhi || l > j || isIm || padCta) continue;
        const double gr = a.alpha * part;
        const double gi = (l == j) ? 0.0 : a.alpha * other;
        const long long a0 = (long long)r0 * a.M + j, b1 = (long long)r1 * a.M + l;
        cplx* pA = a.A + a0 * Pc + b1;
        cplx* pAt = a.A + b1 * Pc + a0;
        const bool diagEl = samePair && l == j;
        if (!a.accumulate) {
          const cplx m1l = sMu1[li];
          const cplx v0 = cmk(gr - a.kappa * (m0j.x * m1l.x + m0j.y * m1l.y),
                              diagEl ? 0.0 : gi - a.kappa * (m0j.x * m1l.y - m0j.y * m1l.x));
          *pA = v0;
          if (!diagEl) *pAt = cconj(v0);
        } else {
          atomicAdd(&pA->x, gr);
          if (!diagEl) { atomicAdd(&pA->y, gi); atomicAdd(&pAt->x, gr); atomicAdd(&pAt->y, -gi); }
        }
        if (!samePair && l != j) {
          const long long a1 = (long long)r1 * a.M + j, b0 = (long long)r0 * a.M + l;
          cplx* qA = a.A + a1 * Pc + b0;
          cplx* qAt = a.A + b0 * Pc + a1;
          if (!a.accumulate) {
            const cplx m0l = sMu0[li];
            const cplx w0 = cmk(gr - a.kappa * (m1j.x * m0l.x + m1j.y * m0l.y), gi - a.kappa * (m1j.x * m0l.y - m1j.y * m0l.x));
            *qA = w0;
            *qAt = cconj(w0);
          } else {
            atomicAdd(&qA->x, gr); atomicAdd(&qA->y, gi);
            atomicAdd(&qAt->x, gr); atomicAdd(&qAt->y, -gi);
          }
        }
      }
    }
    }
  }
  if (warp == 2) I8_TRACE_CTA(4);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (a.cl > 1) cluster_sync_all();   // no peer may still signal this CTA's barriers when it exits
  if (warp == 2) I8_TRACE_CTA(5);
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

}  // namespace

static int g_i8_dbg = 0;
static int g_i8_cluster = 2;
static int g_i8_max_stages = 0;   // development: stages per launch (0: I8_MAXSTAGES, the int32 head-room)
static long long* g_i8_trace = nullptr;
static int g_i8_trace_tile = 0;
// development: device buffer of 16 * 832 int64 that receives the clock64 time stamps of CTA (pair 0, tile `tile`) of the
// next launches (events per stage: 0 TMA issued, 2 MMA stage start, 1 six MMAs issued, 8 next stage's barriers seen, 3 stage committed, 4 sign warp saw full,
// 5 signs prepared, 6 sign warp saw TMEM buffer free, 7 TMEM stores done); nullptr switches tracing off
extern "C" int jvmc_i8_set_trace(void* buf, int tile) {
  g_i8_trace = (long long*)buf;
  g_i8_trace_tile = tile;
  return JVMC_OK;
}
// development knobs: bits 0-7 ablation flags (timing only, results invalid); bits 8-11, when non-zero, set the cluster size
extern "C" int jvmc_i8_set_debug(int flags) {
  g_i8_dbg = flags & 0xFF;
  const int cl = (flags >> 8) & 0xF;
  if (cl == 1 || cl == 2 || cl == 4 || cl == 8) g_i8_cluster = cl;
  const int ms = (flags >> 12) & 0xFFF;             // bits 12-23: stages per launch (L2 footprint experiments)
  g_i8_max_stages = (ms > 0 && ms <= I8_MAXSTAGES) ? ms : 0;
  return JVMC_OK;
}

// Tile shape in real columns (rows, cols); complex rows/columns are half of it.
extern "C" int jvmc_i8_tile_shape(int* rows, int* cols) {
  if (!rows || !cols) return JVMC_ERR_ARG;
  *rows = I8_TM; *cols = I8_TN;
  return JVMC_OK;
}

// Sizes of the scratch buffers the caller provides.
extern "C" int jvmc_i8_layout(long long B, int M, long long* numChunks, int* numZGroups, long long* digitBytes) {
  if (B < 0 || M <= 0 || !numChunks || !numZGroups || !digitBytes) return JVMC_ERR_ARG;
  long long stages = (B + I8_KS - 1) / I8_KS;
  if (stages < 1) stages = 1;
  *numChunks = stages * 2;
  int twoM = 2 * M;
  int padA = ((twoM + I8_TM - 1) / I8_TM) * I8_TM, padB = ((twoM + I8_TN - 1) / I8_TN) * I8_TN;
  int pad = padA > padB ? padA : padB;
  *numZGroups = pad / 8;
  // [A encoding | B encoding | scratch of the column-sum correction: int32 SB[5][ZP], double corr[5][ZP] (+ alignment)]
  const long long half = (long long)I8_S * (*numChunks) * (*numZGroups) * 128;
  *digitBytes = 2 * half + (long long)I8_S * pad * 4 + (long long)I8_LEV * pad * 8 + 256;
  return JVMC_OK;
}
static long long i8_half_bytes(long long numChunks, int numZGroups) { return (long long)I8_S * numChunks * numZGroups * 128; }

// ratios[z] <- max_n|Z_nz| / rms_n(Z_nz) for the 2M real columns Z = [Re Y_j, Im Y_j] (interleaved).  scratch: 4M doubles.
extern "C" int jvmc_i8_tail_ratios(const double* Y, long long B, int M, double* scratch, double* ratios, void* stream) {
  if (!Y || !scratch || !ratios || B <= 0 || M <= 0) return JVMC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(scratch, 0, sizeof(double) * 4 * M, st);
  dim3 g1((M + 127) / 128, (unsigned)((B + 1023) / 1024));
  i8_colsq_kernel<<<g1, 128, 0, st>>>((const cplx*)Y, B, M, scratch, (unsigned long long*)(scratch + 2 * M));
  JVMC_CHECK_LAUNCH();
  i8_tail_ratio_kernel<<<(2 * M + 255) / 256, 256, 0, st>>>(scratch, (const unsigned long long*)(scratch + 2 * M), 2 * M,
                                                            1.0 / (double)B, ratios);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

// flag[n] (device bytes) <- 1 where |Z_nz| > T rms_z for some real column z.  scratch: the buffer jvmc_i8_tail_ratios filled.
extern "C" int jvmc_i8_outlier_rows(const double* Y, long long B, int M, const double* scratch, double T, unsigned char* flag,
                                    void* stream) {
  if (!Y || !scratch || !flag || B <= 0 || M <= 0 || !(T > 0.0)) return JVMC_ERR_ARG;
  i8_rowflag_kernel<<<(unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const cplx*)Y, B, M, scratch, 1.0 / (double)B, T, flag);
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

// digits must be zero-initialised by the caller (padding rows/columns stay zero); colmax: 2M uint64 zeroed scratch.
extern "C" int jvmc_i8_slice(const double* Y, long long B, int M, unsigned long long* colmax, double* scale,
                             signed char* digits, void* stream) {
  if (!Y || !colmax || !scale || !digits || B <= 0 || M <= 0) return JVMC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  long long numChunks, digitBytes;
  int numZGroups;
  jvmc_i8_layout(B, M, &numChunks, &numZGroups, &digitBytes);
  cudaMemsetAsync(colmax, 0, sizeof(unsigned long long) * 2 * M, st);
  dim3 g1((M + 127) / 128, (unsigned)((B + 1023) / 1024));
  i8_colmax_kernel<<<g1, 128, 0, st>>>((const cplx*)Y, B, M, colmax);
  JVMC_CHECK_LAUNCH();
  i8_scale_kernel<<<(2 * M + 255) / 256, 256, 0, st>>>(colmax, 2 * M, scale);
  JVMC_CHECK_LAUNCH();
  dim3 g2((2 * M + 255) / 256, (unsigned)((B + 15) / 16));
  i8_slice_kernel<<<g2, 256, 0, st>>>((const cplx*)Y, B, M, scale, numChunks, numZGroups, (int8_t*)digits,
                                      (int8_t*)digits + i8_half_bytes(numChunks, numZGroups));
  JVMC_CHECK_LAUNCH();
  return JVMC_OK;
}

// tiles: device array of 8 ints per tile (rowGroup, colGroup, NC, jlo, jhi, 0, 0, 0): the tile's 128 rows start at real
// column 8 rowGroup (complex row 4 rowGroup), its NC real columns (multiple of 16, <= 80) at real column 8 colGroup; it
// writes the elements jlo <= j < jhi, l <= j.  The caller's list must cover every (j, l <= j) exactly once (later
// launches add into A); digit rows up to 8 colGroup + NC must lie inside the padded layout of jvmc_i8_layout.
extern "C" int jvmc_rbm_gram_S_i8(const signed char* digits, const double* scale, long long B, int M, int R,
                                  const unsigned int* sigT, const int* tiles, int numTiles, const double* mu,
                                  double alpha, double kappa, int accumulate, long long pair0, long long npairs, double* A,
                                  void* stream) {
  if (!digits || !scale || !sigT || !tiles || !A || B <= 0 || M <= 0 || R <= 0 || numTiles <= 0) return JVMC_ERR_ARG;
  I8Args a;
  long long digitBytes;
  jvmc_i8_layout(B, M, &a.numChunks, &a.numZGroups, &digitBytes);
  const long long half = i8_half_bytes(a.numChunks, a.numZGroups);
  const int ZP = a.numZGroups * 8;
  a.dig = (const int8_t*)digits; a.digB = a.dig + half;
  int* SB = (int*)((unsigned char*)digits + 2 * half);                                   // 16-byte aligned: half % 128 == 0
  double* corr = (double*)(((uintptr_t)(SB + (size_t)I8_S * ZP) + 255) & ~(uintptr_t)255);
  a.corr = corr;
  a.scale = scale; a.sigT = sigT; a.words = (B + 31) / 32;
  a.tiles = tiles; a.mu = (const cplx*)mu; a.alpha = alpha; a.kappa = kappa; a.A = (cplx*)A;
  a.M = M; a.R = R;
  size_t smem = (size_t)I8_SLOTS * I8_STAGE_BYTES + 32 * sizeof(uint64_t) + 256 * sizeof(uint2);   // ring | barriers | mask table
  cudaFuncSetAttribute(gram_s_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long pairs = (long long)R * (R + 1) / 2;
  if (numTiles > 65535 || pairs > 2147483647LL) return JVMC_ERR_UNSUPPORTED;
  // (tile descriptors are device data: NC in {16, 32, ..., I8_TN} and the padding bound are the caller's contract)
  a.cl = g_i8_cluster; a.pairs = pairs;
  if (npairs <= 0) { pair0 = 0; npairs = pairs; }
  if (pair0 < 0 || pair0 + npairs > pairs) return JVMC_ERR_ARG;
  a.pair0 = pair0; a.npairs = npairs;
  a.trace = g_i8_trace; a.traceTile = g_i8_trace_tile;
  dim3 grid((unsigned)((npairs + a.cl - 1) / a.cl * a.cl), (unsigned)numTiles);
  const long long numStages = a.numChunks / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(I8_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)a.cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  // one launch per <= 25600 samples (exact int32 accumulation), evenly split; later launches add into A
  const long long maxStages = g_i8_max_stages ? g_i8_max_stages : I8_DEFAULT_STAGES;
  const long long numLaunches = (numStages + maxStages - 1) / maxStages;
  const long long per = (numStages + numLaunches - 1) / numLaunches;
  for (long long s0 = 0; s0 < numStages; s0 += per) {
    a.stage0 = s0;
    a.stage1 = (s0 + per < numStages) ? s0 + per : numStages;
    a.accumulate = (s0 > 0 || accumulate) ? 1 : 0;   // accumulate != 0: A += alpha G (no mean correction), from the first launch on
    a.dbg = g_i8_dbg;
    // 1/2 sum_n b_n per level and column for the samples of this launch (the half of the half-integer A digits)
    cudaMemsetAsync(SB, 0, sizeof(int) * (size_t)I8_S * ZP, (cudaStream_t)stream);
    dim3 gc((unsigned)((I8_S * ZP + 127) / 128), (unsigned)((a.stage1 - a.stage0 + 31) / 32));
    i8_colsum_kernel<<<gc, 128, 0, (cudaStream_t)stream>>>(a.digB, a.numZGroups, a.stage0, a.stage1, SB);
    JVMC_CHECK_LAUNCH();
    i8_corr_kernel<<<(I8_LEV * ZP + 127) / 128, 128, 0, (cudaStream_t)stream>>>(SB, ZP, corr);
    JVMC_CHECK_LAUNCH();
    if (a.cl == 1) {
      gram_s_i8_kernel<<<grid, I8_THREADS, smem, (cudaStream_t)stream>>>(a);
      JVMC_CHECK_LAUNCH();
    } else {
      cudaError_t le = cudaLaunchKernelEx(&cfg, gram_s_i8_kernel, a);
      if (le != cudaSuccess) { jvmc_set_last_cuda_error(le); return JVMC_ERR_CUDA; }
    }
  }
  return JVMC_OK;
}
