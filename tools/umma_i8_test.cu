// Stage-0 probe for the int8 (Ozaki) Gram path: one CTA computes D[128 x N] = A[128 x K] * B[N x K]^T with
// tcgen05.mma kind::i8 (int8 x int8 -> int32 in TMEM), operands K-major in the SWIZZLE_NONE canonical layout
// [kchunk16][rowgroup8][8 rows][16 bytes], and checks the result on the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_i8_test tools/umma_i8_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int M_ = 128, N_ = 96, K_ = 64;   // K bytes (int8): 2 MMAs of K = 32

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (sm100)
  return d;                 // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}

__global__ void __launch_bounds__(128, 1) umma_test(const int8_t* __restrict__ Ag, const int8_t* __restrict__ Bg,
                                                    int32_t* __restrict__ D, int ts_mode) {
  extern __shared__ __align__(128) unsigned char smem[];
  int8_t* As = reinterpret_cast<int8_t*>(smem);                 // [K/16][M/8][8][16]
  int8_t* Bs = As + M_ * K_;                                    // [K/16][N/8][8][16]
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // operands are already in the canonical layout in global memory
  for (int i = tid; i < M_ * K_ / 16; i += 128) reinterpret_cast<int4*>(As)[i] = reinterpret_cast<const int4*>(Ag)[i];
  for (int i = tid; i < N_ * K_ / 16; i += 128) reinterpret_cast<int4*>(Bs)[i] = reinterpret_cast<const int4*>(Bg)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy smem writes -> async proxy (MMA)
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tacc = tmem_base;

  if (ts_mode) {
    // stage A into TMEM columns [128, 128 + K/4): lane = row, 8 columns (32 bytes) per MMA K-step
    const int row = warp * 32 + lane;
    for (int ks = 0; ks < K_ / 32; ++ks) {
      uint32_t w[8];
      for (int ch = 0; ch < 2; ++ch) {
        const int4 v = *reinterpret_cast<const int4*>(As + (((ks * 2 + ch) * (M_ / 8) + row / 8) * 8 + row % 8) * 16);
        w[ch * 4 + 0] = v.x; w[ch * 4 + 1] = v.y; w[ch * 4 + 2] = v.z; w[ch * 4 + 3] = v.w;
      }
      uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16) + 128u + (uint32_t)(ks * 8);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(w[0]),
                   "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }

  if (tid == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_ >> 3) << 17) | ((uint32_t)(M_ >> 4) << 24);
    for (int ks = 0; ks < K_ / 32; ++ks) {
      // one MMA consumes 2 k-chunks of 16 B: LBO = stride between chunks, SBO = stride between 8-row groups
      uint64_t da = make_desc(smem_u32(As) + ks * 2 * (M_ / 8) * 128, (M_ / 8) * 128, 128);
      uint64_t db = make_desc(smem_u32(Bs) + ks * 2 * (N_ / 8) * 128, (N_ / 8) * 128, 128);
      uint32_t acc = ks > 0 ? 1u : 0u;
      if (ts_mode) {
        uint32_t ta = tacc + 128u + (uint32_t)(ks * 8);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
            "}\n" ::"r"(tacc), "r"(ta), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
        continue;
      }
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "setp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
          "}\n" ::"r"(tacc), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMAs
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra.uni W_DONE;\n"
      "bra.uni W_LOOP;\n"
      "W_DONE:\n"
      "}\n" ::"r"(smem_u32(&mbar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31, 16 columns at a time
  for (int c0 = 0; c0 < N_; c0 += 16) {
    uint32_t r[16];
    uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int c = 0; c < 16; ++c) D[(warp * 32 + lane) * N_ + c0 + c] = (int32_t)r[c];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tacc));
}

int main() {
  std::vector<int8_t> A(M_ * K_), B(N_ * K_), Ac(M_ * K_), Bc(N_ * K_);
  srand(1);
  for (auto& v : A) v = (int8_t)(rand() % 255 - 127);
  for (auto& v : B) v = (int8_t)(rand() % 255 - 127);
  // canonical layout: [k/16][row/8][row%8][k%16]
  for (int r = 0; r < M_; ++r)
    for (int k = 0; k < K_; ++k) Ac[(((k / 16) * (M_ / 8) + r / 8) * 8 + r % 8) * 16 + k % 16] = A[r * K_ + k];
  for (int r = 0; r < N_; ++r)
    for (int k = 0; k < K_; ++k) Bc[(((k / 16) * (N_ / 8) + r / 8) * 8 + r % 8) * 16 + k % 16] = B[r * K_ + k];
  int8_t *dA, *dB;
  int32_t* dD;
  cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, M_ * N_ * 4);
  cudaMemcpy(dA, Ac.data(), A.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bc.data(), B.size(), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, M_ * N_ * 4);
  size_t smem = (M_ + N_) * K_;
  long badTotal = 0;
  for (int ts = 0; ts < 2; ++ts) {
  cudaMemset(dD, 0xFF, M_ * N_ * 4);
  umma_test<<<1, 128, smem>>>(dA, dB, dD, ts);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel (A from %s): %s\n", ts ? "TMEM" : "SMEM", cudaGetErrorString(e));
  std::vector<int32_t> D(M_ * N_);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < M_; ++m)
    for (int n = 0; n < N_; ++n) {
      int32_t ref = 0;
      for (int k = 0; k < K_; ++k) ref += (int32_t)A[m * K_ + k] * (int32_t)B[n * K_ + k];
      if (ref != D[m * N_ + n]) { if (bad < 5) printf("mismatch m=%d n=%d ref=%d got=%d\n", m, n, ref, D[m * N_ + n]); ++bad; }
    }
  printf("mismatches: %ld of %d\n", bad, M_ * N_);
  badTotal += bad;
  }
  return badTotal != 0;
}
