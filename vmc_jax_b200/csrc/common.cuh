// Shared device helpers for the jVMC hot-path kernels (sm_100a).
// fp64 complex arithmetic on double2, Philox4x32-10 counter RNG, warp reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define JVMC_OK 0
#define JVMC_ERR_ARG (-1)
#define JVMC_ERR_CUDA (-2)
#define JVMC_ERR_UNSUPPORTED (-3)
#define JVMC_ERR_SOLVER (-4)

#define JVMC_CHECK_LAUNCH()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) { jvmc_set_last_cuda_error(e__); return JVMC_ERR_CUDA; } \
  } while (0)

void jvmc_set_last_cuda_error(cudaError_t e);

typedef double2 cplx;  // (x = re, y = im); layout-compatible with complex128

__device__ __forceinline__ cplx cmk(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return cmk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cplx cconj(cplx a) { return cmk(a.x, -a.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return cmk(a.x * s, a.y * s); }
__device__ __forceinline__ double cabs2(cplx a) { return fma(a.x, a.x, a.y * a.y); }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  double inv = 1.0 / cabs2(b);
  return cmk(fma(a.x, b.x, a.y * b.y) * inv, fma(a.y, b.x, -a.x * b.y) * inv);
}
// complex multiply with separately rounded products (no FMA contraction): bit-compatible with
// the oracle's NumPy complex product; used only on the bit-exact matrix-element path.
__device__ __forceinline__ cplx cmul_exact(cplx a, cplx b) {
  return cmk(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
             __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}

// exp(z) - 1 without cancellation for small |z|
__device__ __forceinline__ cplx cexpm1(cplx z) {
  double s, c;
  sincos(z.y, &s, &c);
  double em1 = expm1(z.x);
  double sh = sin(0.5 * z.y);
  return cmk(fma(em1, c, -2.0 * sh * sh), (em1 + 1.0) * s);
}
__device__ __forceinline__ cplx cexp(cplx z) {
  double s, c;
  sincos(z.y, &s, &c);
  double e = exp(z.x);
  return cmk(e * c, e * s);
}
// principal-branch log(1+z)
__device__ __forceinline__ cplx clog1p(cplx z) {
  double re = 0.5 * log1p(fma(z.x, z.x + 2.0, z.y * z.y));
  double im = atan2(z.y, 1.0 + z.x);
  return cmk(re, im);
}

// log cosh(x) exactly as jVMC/nets/activation_functions.py:19-22 and, from the same
// exponential, tanh(x).  x' = sgn(Re x) x ; e = exp(-2x') ; lncosh = x' + log1p(e) - ln2 ;
// tanh x = sgn (1-e)/(1+e) with 1-e = -expm1(-2x').
__device__ __forceinline__ void lncosh_tanh(cplx x, cplx& lc, cplx& th) {
  double sgn = signbit(x.x) ? -1.0 : 1.0;
  cplx xp = cmk(x.x * sgn, x.y * sgn);
  cplx m2x = cmk(-2.0 * xp.x, -2.0 * xp.y);
  cplx em1 = cexpm1(m2x);               // e - 1
  cplx e = cmk(em1.x + 1.0, em1.y);
  cplx l1p = clog1p(e);
  lc = cmk(xp.x + l1p.x - 0.69314718055994530942, xp.y + l1p.y);
  cplx num = cmk(-em1.x, -em1.y);       // 1 - e
  cplx den = cmk(2.0 + em1.x, em1.y);   // 1 + e
  cplx t = cdiv(num, den);
  th = cmk(t.x * sgn, t.y * sgn);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ cplx warp_csum(cplx v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}
__device__ __forceinline__ cplx warp_cprod(cplx v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cplx w = cmk(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
    v = cmul(v, w);
  }
  return v;
}

// Philox4x32-10 (Salmon et al. 2011).  Counter-based: the sampler keys it with the user seed and
// counts (global chain id, Metropolis step, purpose), so streams do not depend on the GPU count.
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a;
      c1 = lo1;
      c2 = hi0 ^ c3 ^ b;
      c3 = lo0;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ double u01_from_bits(uint32_t hi, uint32_t lo) {
  uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}
