// Device-side tables of a BranchFreeOperator and the string walk shared by the local-energy kernels (bfo.cu, cnn_inc.cu).
#pragma once
#include "common.cuh"

namespace {

constexpr int BFO_MAXLEN = 16;

struct BfoTables {
  int numOps, len, lDim;
  const int32_t* idx;     // [numOps, len]
  const int32_t* map;     // [numOps, len, lDim]
  const cplx* matEls;     // [numOps, len, lDim]
  const int32_t* fermi;   // [numOps, len]
  const uint8_t* isDiag;  // [numOps]
};

// walk one operator string on configuration c (read-only) ; returns matEl, records modified sites
__device__ __forceinline__ cplx walk_string(const BfoTables& t, int o, const int32_t* __restrict__ c, int N,
                                            cplx m, int* msite, int* mval, int& nmod) {
  nmod = 0;
  for (int k = 0; k < t.len; ++k) {
    const int i = t.idx[o * t.len + k];
    int cur = c[i];
    for (int q = 0; q < nmod; ++q) if (msite[q] == i) cur = mval[q];
    m = cmul_exact(m, t.matEls[((size_t)o * t.len + k) * t.lDim + cur]);
    if (t.fermi[o * t.len + k]) {
      // Jordan-Wigner sign prod_{j>i} (1 - 2 c_j) on the current (modified) configuration
      int sgn = 1;
      for (int j = i + 1; j < N; ++j) {
        int cj = c[j];
        for (int q = 0; q < nmod; ++q) if (msite[q] == j) cj = mval[q];
        sgn *= (1 - 2 * cj);
      }
      m = cmk(m.x * (double)sgn, m.y * (double)sgn);
    }
    const int nv = t.map[((size_t)o * t.len + k) * t.lDim + cur];
    bool found = false;
    for (int q = 0; q < nmod; ++q) if (msite[q] == i) { mval[q] = nv; found = true; }
    if (!found) { msite[nmod] = i; mval[nmod] = nv; ++nmod; }
  }
  return m;
}

inline BfoTables make_tables(int numOps, int len, int lDim, const int32_t* idx, const int32_t* map, const double* matEls,
                      const int32_t* fermi, const uint8_t* isDiag) {
  BfoTables t;
  t.numOps = numOps; t.len = len; t.lDim = lDim;
  t.idx = idx; t.map = map; t.matEls = (const cplx*)matEls; t.fermi = fermi; t.isDiag = isDiag;
  return t;
}

}  // namespace
