/* jvmc_b200.h -- C ABI of the B200-native jVMC hot path (libjvmc_b200.so).
 *
 * The reference (markusschmitt/vmc_jax, jVMC 1.5.8) is pure Python on JAX and has no FFI of its
 * own; every entry point below replaces one jit/pmap'd Python function of the per-step VMC path
 * and is what an XLA-FFI / ctypes / torch binding for that function would bind.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless named host*; complex128 arrays are passed as
 *     `double*` with interleaved (re, im) pairs; configurations are int32 in {0,..,lDim-1};
 *   - `stream` is a cudaStream_t; functions enqueue work and return, they never allocate device
 *     memory and never synchronise (exception: jvmc_eigh waits if cuSOLVER needs a host buffer);
 *   - return value: 0 = ok, <0 = error code (jvmc_error_string).
 *   - "tables" is the buffer filled by jvmc_rbm_tables: T[N*M] | lc[N] | tb2[M] | lcb[1] (complex128).
 *   - Khatri-Rao "site" index r: r = 0 is the bias pseudo-site (sigma = +1) when hasBias, followed by
 *     the N lattice sites; complex parameter index c = r*M + j  (Flax leaf order: bias, kernel).
 */
#ifndef JVMC_B200_H
#define JVMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JVMC_OK 0
#define JVMC_ERR_ARG (-1)
#define JVMC_ERR_CUDA (-2)
#define JVMC_ERR_UNSUPPORTED (-3)
#define JVMC_ERR_SOLVER (-4)

int jvmc_version(void);
const char* jvmc_error_string(int code);

/* NQS.__call__/_eval for CpxRBM/RBM: jVMC/vqs.py:223-251, jVMC/nets/rbm.py:30-38,82-88,
 * jVMC/nets/activation_functions.py:19-22; also the sampler's final re-evaluation
 * jVMC/sampler.py:293-296.  s[B,N] -> logpsi[B] (complex), tau[B,M] = tanh(theta) (complex, may be NULL). */
int jvmc_rbm_logpsi(const int32_t* s, long long B, int N, int M, const double* W, const double* bias,
                    double* logpsi, double* tau, void* stream);

/* Per-parameter flip-ratio tables tanh(2W), log prod_j cosh(2W_ij), tanh(2b), log prod_j cosh(2 b_j). */
long long jvmc_rbm_tables_elems(int N, int M);
int jvmc_rbm_tables(int N, int M, const double* W, const double* bias, double* tables, void* stream);

/* MCSampler._get_samples/_sweep + proposers: jVMC/sampler.py:15-19,30-39,42-64,301-356.
 * proposer: 0 spin_flip, 1 spin_flip_Z2, 2 spin_flip_zeroMag.  states[C,N] in/out; out[numSamplesPerChain*C, N]
 * time-major / chain-minor (sampler.py:323); counters[2] += (proposed, accepted).  Philox stream:
 * (seed; step0 + step, chain0 + chain). */
int jvmc_rbm_mcmc(int32_t* states, long long C, int N, int M, const double* W, const double* bias,
                  const double* tables, unsigned long long seed, unsigned long long step0, long long chain0,
                  int proposer, double mu, int sweepSteps, long long thermSteps, int numSamplesPerChain,
                  int refreshEvery, int32_t* out, unsigned long long* counters, void* stream);
int jvmc_mcmc_set_generic(int on);   /* development knob: 1 = always use the generic sampler kernel (A/B timing) */

/* Operator.get_s_primes: jVMC/operator/base.py:91-160 + BranchFreeOperator._get_s_primes
 * jVMC/operator/branch_free.py:443-487.  Phase 1: all matrix elements mAll[B,numOps] (diagonal strings merged),
 * choice[B,numOps] (nonzero ops ascending, then numOps-1), count[B], maxCount[1].  Phase 2 (after the caller
 * read Kmax = *maxCount): sp[B*Kmax,N], matEl[B,Kmax] with the reference's padding. */
int jvmc_bfo_matels(const int32_t* s, long long B, int N, int numOps, int len, int lDim, const int32_t* idx,
                    const int32_t* map, const double* matEls, const int32_t* fermi, const uint8_t* isDiag,
                    int numDiag, const double* pref, double* mAll, int32_t* choice, int32_t* count,
                    int32_t* maxCount, void* stream);
int jvmc_bfo_emit(const int32_t* s, long long B, int N, int numOps, int len, int lDim, const int32_t* idx,
                  const int32_t* map, const double* matEls, const int32_t* fermi, const uint8_t* isDiag,
                  const double* mAll, const int32_t* choice, const int32_t* count, int Kmax, int32_t* sp,
                  double* matEl, void* stream);

/* Operator._get_O_loc: jVMC/operator/base.py:162-164. */
int jvmc_oloc_reduce(const double* matEl, const double* logPsiS, const double* logPsiSP, long long B, int K,
                     double* out, void* stream);

/* Operator.get_O_loc fast path (base.py:166-192) for (Cpx)RBM + branch-free strings flipping <= 2 sites,
 * lDim = 2, no fermionic strings.  errFlag (device int) is set to 1 if a string flips more sites. */
int jvmc_rbm_eloc_bfo(const int32_t* s, const double* tau, long long B, int N, int M, const double* tables,
                      int numOps, int len, int lDim, const int32_t* idx, const int32_t* map, const double* matEls,
                      const int32_t* fermi, const uint8_t* isDiag, int numDiag, const double* pref, double* out,
                      int* errFlag, void* stream);

/* NQS.gradients: jVMC/vqs.py:46-69,256-287.  layout 0: holomorphic [b, i b, W, i W]; 1: real params [b, W]. */
int jvmc_rbm_grad(const int32_t* s, const double* tau, long long B, int N, int M, int hasBias, int layout,
                  double* out, void* stream);

/* out[r,j] = sum_n wgt_n sigma_{n,r} (conjTau ? conj tau_nj : tau_nj): SampledObs mean / covar(grads,Eloc) /
 * MinSR -O^dagger x without forming O (jVMC/stats.py:50-58,204,245; jVMC/util/minsr.py:65).
 * workspace: jvmc_rbm_moments_chunks(B) * R * M complex128. */
int jvmc_rbm_moments_chunks(long long B);
int jvmc_rbm_moments(const int32_t* s, const double* tau, const double* wgt, long long B, int N, int M,
                     int hasBias, int conjTau, double* workspace, double* out, void* stream);

/* out[n] = sum_{r,j} sigma_{n,r} (conjTau ? conj tau_nj : tau_nj) x[r,j] = O_n . x without forming O: the per-sample
 * projection used by the centred MinSR kernel (jVMC/stats.py:332-336) and by matrix-free S.v products. */
int jvmc_rbm_krmatvec(const int32_t* s, const double* tau, const double* x, long long B, int N, int M,
                      int hasBias, int conjTau, double* out, void* stream);

/* sigT[R][ceil(B/32)] bit-packed transposed spins (bias pseudo-site row first). */
int jvmc_pack_sigma(const int32_t* s, long long B, int N, int hasBias, unsigned int* sigT, void* stream);

/* SampledObs.covar() of the gradients = S (jVMC/stats.py:235-245, jVMC/util/tdvp.py:142), Khatri-Rao form:
 * A[(r,j),(r',l)] = alpha sum_n sigma_nr sigma_nr' conj(Y_nj) Y_nl - kappa conj(mu_rj) mu_r'l, A row-major
 * [R*M, R*M] complex128, fully populated.  tile: 0 auto | 64 | 80. */
int jvmc_rbm_gram_S(const double* Y, long long B, int M, int R, const unsigned int* sigT, const double* mu,
                    double alpha, double kappa, double* A, int tile, void* stream);

/* Same contract as jvmc_rbm_gram_S on the tcgen05 INT8 tensor cores (Ozaki splitting, fp64-equivalent result):
 * jvmc_i8_layout -> scratch sizes; jvmc_i8_slice -> per-column scales (2 x column maximum) + 5 balanced base-255 int8 digits in
 * the UMMA canonical layout (digits must be zero-initialised); jvmc_rbm_gram_S_i8 -> A.  tiles: device int array, 8 ints
 * per tile (rowGroup, colGroup, NC, jlo, jhi, 0, 0, 0): 128 real rows from real column 8 rowGroup (complex row
 * 4 rowGroup), NC real columns (multiple of 16, <= cols of jvmc_i8_tile_shape) from real column 8 colGroup; the tile
 * writes the elements jlo <= j < jhi, l <= j (plus their Hermitian images); the list must cover every (j, l <= j)
 * exactly once (vmc_jax_b200/kernels.py:i8_tile_list). */
int jvmc_i8_layout(long long B, int M, long long* numChunks, int* numZGroups, long long* digitBytes);
int jvmc_i8_tile_shape(int* rows, int* cols);   /* tile of jvmc_rbm_gram_S_i8 in real columns */
int jvmc_i8_set_debug(int flags);   /* development ablations of the int8 Gram pipeline (timing only; results invalid) */
int jvmc_i8_set_trace(void* buf, int tile);   /* development: per-stage clock64 trace of one CTA (16 x 832 int64, device) */
int jvmc_i8_slice(const double* Y, long long B, int M, unsigned long long* colmax, double* scale,
                  signed char* digits, void* stream);
/* Heavy-tail diagnostic behind the choice between the int8 and the fp64 DMMA Gram (kernels.py:rbm_gram_S_auto):
 * ratios[z] (device, 2M) <- max_n|Z_nz| / rms_n(Z_nz) per real column of Y.  scratch: 4M doubles (device). */
int jvmc_i8_tail_ratios(const double* Y, long long B, int M, double* scratch, double* ratios, void* stream);
/* flag[n] (device bytes) <- 1 where some entry of sample n exceeds T x the rms of its column (scratch of the call above):
 * these few samples go through the exact fp64 kernel, the rest through the int8 kernel (rbm_gram_S_auto). */
int jvmc_i8_outlier_rows(const double* Y, long long B, int M, const double* scratch, double T, unsigned char* flag,
                         void* stream);
/* accumulate != 0: A += alpha G.  [pair0, pair0 + npairs): range of site pairs in the order (r0 ascending, r1 = r0..R-1),
 * npairs <= 0: all R(R+1)/2 -- block row r0 of the upper block triangle of A is final once its pairs are done. */
int jvmc_rbm_gram_S_i8(const signed char* digits, const double* scale, long long B, int M, int R,
                       const unsigned int* sigT, const int* tiles, int numTiles, const double* mu, double alpha,
                       double kappa, int accumulate, long long pair0, long long npairs, double* A, void* stream);

/* SampledObs.tangent_kernel (jVMC/stats.py:332-336; MinSR, jVMC/util/minsr.py:59-60), Khatri-Rao form:
 * T[n,m] = scale sqrt(p_n p_m) [ (sum_r sigma_nr sigma_mr)(sum_j tau_nj conj tau_mj) - v_n - conj(v_m) + c ],
 * v = O.conj(mu) (jvmc_rbm_krmatvec), c[0] = |mu|^2 (device scalar); sigR from jvmc_pack_sigma_rows
 * ([B][ceil(R/32)] words).  T row-major [B,B] complex128, fully populated Hermitian.  scale = 2: reference's doubled
 * holomorphic layout. */
int jvmc_pack_sigma_rows(const int32_t* s, long long B, int N, int hasBias, unsigned int* sigR, void* stream);
long long jvmc_rbm_gram_T_tiles(long long B);   /* tile pairs of the Hermitian half */
/* [tile0, tile0 + ntiles) of them (ntiles <= 0: all): disjoint ranges on different ranks + one SUM all-reduce form T once */
int jvmc_rbm_gram_T(const double* Y, long long B, int M, int R, const unsigned int* sigR, const double* p,
                    const double* v, const double* c, double scale, long long tile0, long long ntiles, double* T,
                    void* stream);

/* S = q(S0) (+ diagonal shift) in the reference's flat layout from A: jVMC/util/tdvp.py:140-146.
 * mode 0: Re -> double[P,P]; mode 1: i*Im -> complex128[P,P]; column-major. */
/* Multi-GPU reduction of the Hermitian A: only the suffix [r0 M, Pc) of every block row is communicated; this fills
 * A[i][k] = conj(A[k][i]) below the block diagonal afterwards (blocks of M x M, A row-major complex128 [Pc, Pc]). */
int jvmc_hermitian_mirror_blocks(double* A, int Pc, int M, void* stream);
long long jvmc_hermitian_packed_elems(int Pc, int M);   /* complex elements of the packed upper block triangle */
int jvmc_hermitian_pack_blocks(double* A, int Pc, int M, double* packed, int unpack, void* stream);
int jvmc_hermitian_pack_rows(double* A, int Pc, int M, int row0, int nrows, double* packed, int unpack, void* stream);

int jvmc_expand_S(const double* A, int M, int N, int hasBias, int mode, double shift, double* out, void* stream);

/* jnp.linalg.eigh (jVMC/util/tdvp.py:153-171): cuSOLVER Xsyevd, lower, vectors; column-major in/out. */
int jvmc_eigh_workspace(int n, int isComplex, long long* hostDeviceBytes, long long* hostHostBytes);
int jvmc_eigh(int n, int isComplex, double* A, double* w, void* work, long long deviceBytes, int* info,
              void* stream);

/* Regularised pseudo-inverse loop of TDVP.solve (jVMC/util/tdvp.py:193-209) on the device.
 * VtF, F complex128[n]; snr may be NULL (ExactSampler semantics); scal[0]=residual, scal[1]=cutoff. */
int jvmc_tdvp_regularize(int n, const double* ev, const double* VtF, const double* snr, const double* F,
                         double pinvTol, double pinvCutoff, double snrTol, double* pinvEv, double* scal,
                         void* stream);

/* Orbit-averaged (Cpx)RBM = SymNet(orbit, RBM, avgFun_Coefficients_Exp) (jVMC/nets/sym_wrapper.py:8-66, orbits from
 * jVMC/util/symmetries.py): log Psi(s) = log sum_g fac_g exp(logpsi_RBM(O_g s)), O_g signed permutations given as
 * x^g_i = esgn[g,i] * sigma[pmap[g,i]] (int32 [G,N] each; sigma = 2s-1), qmap[g,k] = the row i with pmap[g,i] = k.
 * jvmc_symrbm_logpsi: logpsi complex128[B]; weights (optional) complex128[B,G] = fac_g psi_g / sum   <- NQS.__call__
 * jvmc_symrbm_grad:   d log Psi / d theta, reference flat layout (layout 0: holomorphic [g, ig] per leaf; 1: plain)
 *                     from the weights of jvmc_symrbm_logpsi                                          <- NQS.gradients
 * jvmc_symrbm_mcmc:   Metropolis sampler with propose_spin_flip; same state / RNG / output conventions as
 *                     jvmc_rbm_mcmc; tanh(theta^g) of every group element is updated incrementally     <- MCSampler */
int jvmc_symrbm_logpsi(const int32_t* s, long long B, int N, int M, int G, const double* W, const double* bias,
                       const int32_t* pmap, const int32_t* esgn, const double* fac, double* logpsi, double* weights,
                       void* stream);
int jvmc_symrbm_grad(const int32_t* s, long long B, int N, int M, int G, const double* W, const double* bias,
                     const int32_t* pmap, const int32_t* esgn, const double* fac, const double* weights, int layout,
                     double* out, void* stream);
int jvmc_symrbm_mcmc(int32_t* states, long long C, int N, int M, int G, const double* W, const double* bias,
                     const int32_t* pmap, const int32_t* esgn, const int32_t* qmap, const double* fac,
                     const double* tables, unsigned long long seed, unsigned long long step0, long long chain0,
                     double mu, int sweepSteps, long long thermSteps, int numSamplesPerChain, int refreshEvery,
                     int32_t* out, unsigned long long* counters, void* stream);

/* Real-parameter CNN ansatz (jVMC/nets/cnn.py:18-81): wrap padding, strided cross-correlation, activation per layer,
 * sum / sqrt(size).  desc: HOST int array [nl, Lx, Ly, Fx, Fy, sx, sy, firstLayerBias, bias, ch_1..ch_nl, act_1..act_nl]
 * (act: 0 elu, 1 relu, 2 tanh, 3 poly5, 4 poly6, 5 square; Ly = Fy = sy = 1 for chains); theta: device float64[P], the
 * flat parameter vector in the reference's order (per layer: bias if present, kernel [Fx, Fy, Cin, Cout]).
 * jvmc_cnn_logpsi: complex128[B] (imaginary part 0)                                           <- NQS.__call__
 * jvmc_cnn_grad:   complex128[B, P], d log psi / d theta by per-sample back-propagation        <- NQS.gradients
 * jvmc_cnn_mcmc:   Metropolis sampler, full forward pass per proposal; state / RNG / output conventions and proposer
 *                  ids of jvmc_rbm_mcmc                                                        <- MCSampler */
int jvmc_cnn_num_parameters(const int* desc, int ndesc, int* P);
int jvmc_cnn_logpsi(const int* desc, int ndesc, const double* theta, const int32_t* s, long long B, double* logpsi,
                    void* stream);
int jvmc_cnn_grad(const int* desc, int ndesc, const double* theta, const int32_t* s, long long B, double* out,
                  void* stream);
int jvmc_cnn_mcmc(const int* desc, int ndesc, const double* theta, int32_t* states, long long C,
                  unsigned long long seed, unsigned long long step0, long long chain0, int proposer, double mu,
                  int sweepSteps, long long thermSteps, int numSamplesPerChain, int32_t* out,
                  unsigned long long* counters, void* stream);
/* Incremental fast paths for stride-1 nets (csrc/cnn_inc.cu): a move changing one or two sites only re-evaluates the
 * (l (F-1) + 1)^d affected positions of layer l from activations cached in shared memory.  Both answer
 * JVMC_ERR_UNSUPPORTED for nets outside their scope (stride > 1, > 4 layers, state beyond shared memory); jvmc_cnn_mcmc
 * tries jvmc_cnn_mcmc_inc first.  jvmc_cnn_eloc_bfo: fused local energy for nets.CNN <- Operator.get_O_loc
 * (base.py:166-192); operator tables as for jvmc_rbm_eloc_bfo.  jvmc_cnn_set_generic(1) disables both, (2) runs the incremental sampler with four instead
 * of two warps per chain (A/B knobs for the tests and tools/cnn_bench.py). */
int jvmc_cnn_mcmc_inc(const int* desc, int ndesc, const double* theta, int32_t* states, long long C,
                  unsigned long long seed, unsigned long long step0, long long chain0, int proposer, double mu,
                  int sweepSteps, long long thermSteps, int numSamplesPerChain, int32_t* out,
                  unsigned long long* counters, void* stream);
int jvmc_cnn_eloc_bfo(const int* desc, int ndesc, const double* theta, const int32_t* s, long long B, int numOps, int len,
                      int lDim, const int32_t* idx, const int32_t* map, const double* matEls, const int32_t* fermi,
                      const uint8_t* isDiag, int numDiag, const double* pref, double* out, int* errFlag, void* stream);
int jvmc_cnn_set_generic(int on);

/* ---- the whole regularised solve of one step as single entry points (so that a binding without Python between
 * kernels can call it): TDVP.solve, reference jVMC/util/tdvp.py:153-213, and MinSR.solve, jVMC/util/minsr.py:59-78.
 * mode 0 = makeReal 'real' (S real symmetric, double), mode 1 = 'imag' (S complex Hermitian).  S / T: column-major,
 * overwritten by the eigenvectors.  D: complex[B, n] centred data sqrt(w)(O - <O>) of the gradients, or NULL for the
 * ExactSampler semantics without SNR weighting (tdvp.py:203); e: complex[B] centred local energies; w: [B] weights.
 * comm: communicator of jvmc_comm_init or NULL (single rank).  The per-sample projections V^dagger x_n are a cuBLAS gemm. */
int jvmc_tdvp_solve_workspace(int n, int mode, long long B, long long* bytes);
int jvmc_tdvp_solve(int n, int mode, double* S, const double* F, long long B, const double* D, const double* e,
                    const double* w, double xre, double xim, double numSamplesGlobal, int useSnr, double snrTol, double pinvTol,
                    double pinvCutoff, void* comm, double* ev, double* VtF, double* rhoVar, double* snr, double* pinvEv,
                    double* update, double* scal, int* info, void* work, long long workBytes, void* stream);
int jvmc_minsr_solve_workspace(int n, int isComplex, long long* bytes);
int jvmc_minsr_solve(int n, int isComplex, double* T, const double* e, double rtol, double* x, double* ev, int* info,
                     void* work, long long workBytes, void* stream);

/* ---- building blocks of the Hermitian eigen-decomposition beyond cuSOLVER's dense-eigensolver limit (n <= 32768,
 * measured), used by kernels.py:eigh_large for the config-2 TDVP solve (P_c = 40 000): one level of Cuppen's divide
 * and conquer on the tridiagonal matrix (merge after LAPACK dlaed2/dlaed3: deflation on the host, O(n); secular equation
 * and Gu-Eisenstat vectors on the device).  hetrd / unmtr are cuSOLVER library calls (Zhetrd/Dsytrd, Zunmtr/Dormtr). */
int jvmc_hetrd_workspace(int n, int isComplex, long long* bytes);
int jvmc_hetrd(int n, int isComplex, double* A, double* d, double* e, double* tau, void* work, long long bytes, int* info,
               void* stream);
int jvmc_unmtr_workspace(int n, int ncols, int isComplex, long long* bytes);
int jvmc_unmtr(int n, int ncols, int isComplex, double* A, double* tau, double* C, void* work, long long bytes, int* info,
               void* stream);
int jvmc_unmqr_workspace(int m, int ncols, int k, int isComplex, int lda, int ldc, long long* bytes);
int jvmc_unmqr(int m, int ncols, int k, int isComplex, double* A, int lda, double* tau, double* C, int ldc, void* work,
               long long bytes, int* info, void* stream);
int jvmc_tridiag_dense(int m, const double* d, const double* e, double shiftFirst, double shiftLast, double* T, void* stream);
int jvmc_real_to_complex(long long count, const double* src, double* dst, void* stream);
int jvmc_secular_roots(int k, const double* d, const double* z2, double rho, double z2sum, int* orig, double* mu, void* stream);
int jvmc_secular_vectors(int k, const double* d, const double* z, double rho, const int* orig, const double* mu,
                         const int* rowIdx, const int* colIdx, double* U, long long ldu, double* zhat, void* stream);
int jvmc_apply_row_rotations(int ncols, long long ldu, double* U, int nrot, const int* ri, const int* rj, const double* c,
                             const double* s, void* stream);

/* ---- NCCL plumbing replacing jVMC/mpi_wrapper.py (global_sum/mean/variance/covariance :114-243, gather :278-292,
 * bcast_unknown_size :246-275) on device buffers, in-stream.  NCCL is resolved at run time (libnccl.so.2);
 * JVMC_ERR_UNSUPPORTED when it cannot be loaded.  The 128-byte id of rank 0 is shipped to the other ranks by the caller. */
int jvmc_comm_nccl_version(void);
int jvmc_comm_unique_id(void* hostId128);
int jvmc_comm_init(const void* hostId128, int rank, int world, void** comm);
int jvmc_comm_destroy(void* comm);
int jvmc_comm_allreduce_sum_f64(void* comm, double* buf, long long count, void* stream);
int jvmc_comm_allgather_bytes(void* comm, const void* send, void* recv, long long bytesPerRank, void* stream);
int jvmc_comm_bcast_bytes(void* comm, void* buf, long long bytes, int root, void* stream);
int jvmc_comm_reduce_scatter_sum_f64(void* comm, const double* send, double* recv, long long recvCount, void* stream);

#ifdef __cplusplus
}
#endif
#endif
