// Descriptor of the real-parameter CNN (reference jVMC/nets/cnn.py:18-81) shared by cnn.cu and cnn_inc.cu.
#pragma once
#include "common.cuh"

namespace {

constexpr int CNN_MAXL = 8;

struct CnnDesc {
  int nl;                 // layers
  int Lx, Ly;             // lattice (Ly = 1 for chains)
  int Fx, Fy, sx, sy;     // filter diameter and stride per axis
  int ch[CNN_MAXL + 1];   // channels, ch[0] = 1
  int ox[CNN_MAXL + 1], oy[CNN_MAXL + 1];   // spatial size after layer l (index 0: input)
  int act[CNN_MAXL];      // 0 elu, 1 relu, 2 tanh, 3 poly5, 4 poly6, 5 square
  int hasBias[CNN_MAXL];
  int offB[CNN_MAXL], offK[CNN_MAXL];       // offsets into theta
  int offA[CNN_MAXL + 1]; // offsets of the layer outputs in the activation scratch (index 0: input)
  int totA;               // total activation elements (input + all layers)
  int P;                  // number of parameters
  double nrm;             // sqrt(ox[nl] oy[nl] ch[nl])
};

// expm1(z) for z <= 0 (the negative branch of elu, evaluated for every affected entry of every proposal): argument
// reduction z = k ln2 + r, |r| <= ln2 / 2, Taylor polynomial of expm1(r) to r^14 (truncation 3e-19 relative),
// result 2^k expm1(r) + (2^k - 1) without cancellation; 2e-16 relative against libm, a quarter of its instructions.
__device__ __noinline__ double expm1_neg(double z) {
  z = fmax(z, -64.0);                                   // expm1 = -1 to the last bit below
  const double k = rint(z * 1.4426950408889634);
  double r = fma(-k, 6.93147180369123816490e-01, z);
  r = fma(-k, 1.90821492927058770002e-10, r);
  double q = 1.1470745597729725e-11;                    // 1/14!
  q = fma(q, r, 1.6059043836821613e-10);
  q = fma(q, r, 2.08767569878681e-09);
  q = fma(q, r, 2.505210838544172e-08);
  q = fma(q, r, 2.755731922398589e-07);
  q = fma(q, r, 2.7557319223985893e-06);
  q = fma(q, r, 2.48015873015873e-05);
  q = fma(q, r, 1.984126984126984e-04);
  q = fma(q, r, 1.388888888888889e-03);
  q = fma(q, r, 8.333333333333333e-03);
  q = fma(q, r, 4.1666666666666664e-02);
  q = fma(q, r, 1.6666666666666666e-01);
  q = fma(q, r, 0.5);
  const double p = fma(q, r * r, r);
  const double s = __hiloint2double((1023 + (int)k) << 20, 0);   // 2^k, k in [-93, 0]
  return fma(s, p, s - 1.0);
}

// The activation of four channels at once.  For elu one out-of-line call evaluates the four polynomials interleaved
// (an out-of-line call per value serialises 14-deep dependent chains; inlining them all costs instruction cache).
__device__ __forceinline__ double expm1_neg_inl(double z) {
  z = fmax(z, -64.0);
  const double k = rint(z * 1.4426950408889634);
  double r = fma(-k, 6.93147180369123816490e-01, z);
  r = fma(-k, 1.90821492927058770002e-10, r);
  double q = 1.1470745597729725e-11;
  q = fma(q, r, 1.6059043836821613e-10);
  q = fma(q, r, 2.08767569878681e-09);
  q = fma(q, r, 2.505210838544172e-08);
  q = fma(q, r, 2.755731922398589e-07);
  q = fma(q, r, 2.7557319223985893e-06);
  q = fma(q, r, 2.48015873015873e-05);
  q = fma(q, r, 1.984126984126984e-04);
  q = fma(q, r, 1.388888888888889e-03);
  q = fma(q, r, 8.333333333333333e-03);
  q = fma(q, r, 4.1666666666666664e-02);
  q = fma(q, r, 1.6666666666666666e-01);
  q = fma(q, r, 0.5);
  const double p = fma(q, r * r, r);
  const double s = __hiloint2double((1023 + (int)k) << 20, 0);
  return fma(s, p, s - 1.0);
}
__device__ __noinline__ double4 elu4(double a, double b, double c, double d) {
  const double ea = expm1_neg_inl(fmin(a, 0.0)), eb = expm1_neg_inl(fmin(b, 0.0));
  const double ec = expm1_neg_inl(fmin(c, 0.0)), ed = expm1_neg_inl(fmin(d, 0.0));
  return make_double4(a > 0.0 ? a : ea, b > 0.0 ? b : eb, c > 0.0 ? c : ec, d > 0.0 ? d : ed);
}

__device__ __forceinline__ double actf(int a, double z) {
  switch (a) {
    case 0: {   // branch-free: the evaluations of neighbouring channels interleave instead of serialising on divergent branches
      const double e = expm1_neg(fmin(z, 0.0));
      return z > 0.0 ? z : e;
    }
    case 1: return z > 0.0 ? z : 0.0;
    case 2: return tanh(z);
    case 3: { const double q = z * z; return ((0.133333333 * q - 0.333333333) * q + 1.) * z; }
    case 4: { const double q = z * z; return ((0.022222222 * q - 0.083333333) * q + 0.5) * q; }
    default: return z * z;
  }
}
__device__ __forceinline__ void actf4(int a, const double (&z)[4], double (&out)[4]) {
  if (a == 0) {
    const double4 r = elu4(z[0], z[1], z[2], z[3]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = actf(a, z[j]);
  }
}
__device__ __forceinline__ double dactf(int a, double z) {
  switch (a) {
    case 0: return z > 0.0 ? 1.0 : exp(z);
    case 1: return z > 0.0 ? 1.0 : 0.0;
    case 2: { const double t = tanh(z); return 1.0 - t * t; }
    case 3: { const double q = z * z; return (5.0 * 0.133333333 * q - 3.0 * 0.333333333) * q + 1.; }
    case 4: { const double q = z * z; return ((6.0 * 0.022222222 * q - 4.0 * 0.083333333) * q + 2.0 * 0.5) * z; }
    default: return 2.0 * z;
  }
}

// desc: int array [nl, Lx, Ly, Fx, Fy, sx, sy, firstLayerBias, bias, ch_1..ch_nl, act_1..act_nl]
inline int make_desc(const int* h, int n, CnnDesc& d) {
  if (n < 9) return JVMC_ERR_ARG;
  d.nl = h[0];
  if (d.nl < 1 || d.nl > CNN_MAXL || n != 9 + 2 * d.nl) return JVMC_ERR_ARG;
  d.Lx = h[1]; d.Ly = h[2]; d.Fx = h[3]; d.Fy = h[4]; d.sx = h[5]; d.sy = h[6];
  if (d.Lx < 1 || d.Ly < 1 || d.Fx < 1 || d.Fy < 1 || d.sx < 1 || d.sy < 1) return JVMC_ERR_ARG;
  d.ch[0] = 1; d.ox[0] = d.Lx; d.oy[0] = d.Ly;
  int off = 0, offA = d.Lx * d.Ly;
  d.offA[0] = 0;
  for (int l = 0; l < d.nl; ++l) {
    d.ch[l + 1] = h[9 + l];
    d.act[l] = h[9 + d.nl + l];
    if (d.ch[l + 1] < 1 || d.act[l] < 0 || d.act[l] > 5) return JVMC_ERR_ARG;
    d.hasBias[l] = (l == 0) ? h[7] : h[8];
    // wrap padding by F-1 then VALID convolution with stride s: floor((L - 1) / s) + 1 outputs
    d.ox[l + 1] = (d.ox[l] - 1) / d.sx + 1;
    d.oy[l + 1] = (d.oy[l] - 1) / d.sy + 1;
    d.offB[l] = off;
    if (d.hasBias[l]) off += d.ch[l + 1];
    d.offK[l] = off;
    off += d.Fx * d.Fy * d.ch[l] * d.ch[l + 1];
    d.offA[l + 1] = offA;
    offA += d.ox[l + 1] * d.oy[l + 1] * d.ch[l + 1];
  }
  d.P = off;
  d.totA = offA;
  d.nrm = sqrt((double)(d.ox[d.nl] * d.oy[d.nl] * d.ch[d.nl]));
  return JVMC_OK;
}

}  // namespace
