"""-m gpu: parity of every CUDA kernel (called through the C ABI) against the CPU oracle.
Tolerances: bit-exact for s'/matEl/counts; 1e-10 relative (fp64) for logpsi, tau, E_loc, F, S
(BASELINE.json north_star); chi-squared p > 1e-3 for sampled distributions."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bfo as obfo  # noqa: E402
import gpu_checks as G  # noqa: E402

with open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")) as f:
    REFG = json.load(f)


@pytest.mark.parametrize("N,M,B,bias,real", [(20, 40, 333, False, False), (20, 40, 8, True, False),
                                              (100, 400, 1001, False, False), (40, 80, 77, True, False),
                                              (6, 3, 1, True, False), (12, 7, 50, True, True),
                                              (400, 160, 64, False, False)])
def test_logpsi_tau(N, M, B, bias, real):
    G.check_logpsi(N, M, B, bias, seed=N + M, real=real)


def test_logpsi_empty_batch():
    import torch
    from vmc_jax_b200 import kernels as K
    W = G.dev(np.ones((4, 2), np.complex128))
    lp, tau = K.rbm_logpsi(torch.zeros((0, 4), dtype=torch.int32, device=G.DEV), W)
    assert lp.shape == (0,) and tau.shape == (0, 2)


@pytest.mark.parametrize("bias", [False, True])
def test_tables(bias):
    G.check_tables(bias=bias)


@pytest.mark.parametrize("name", ["tfim1d", "tfim_field", "sx_sysz", "splus", "heis", "single_diag", "same_site"])
def test_s_primes_bit_exact(name):
    N = 8
    G.check_s_primes(G.operator_zoo(N)[name], N)


def test_s_primes_2d_tfim_and_ragged():
    G.check_s_primes(obfo.tfim_strings((4, 4), 3.04), 16, B=33)
    G.check_s_primes(obfo.tfim_strings((10, 10), 3.04), 100, B=5)
    G.check_s_primes(obfo.tfim_strings((6,), 1.0), 6, B=1)


def test_s_primes_time_dependent_prefactor():
    strings = [(lambda t: 2.0 * t, [obfo.Sp(i)]) for i in range(3)]
    for t in [0.5, 2, 13.9]:
        G.check_s_primes(strings, 4, B=24, args=(t,))


def test_s_primes_fermionic_bit_exact():
    G.check_s_primes(G.hubbard_strings(4), 8, B=256)


def test_s_primes_all_zero_rows():
    # S+ on all-ones configurations: no nonzero entry at all -> Kmax = 0
    import torch
    from vmc_jax_b200 import kernels as K
    tab = obfo.Tables([(2., [obfo.Sp(i)]) for i in range(3)])
    s = np.ones((5, 4), np.int32)
    sp, m, cnt = K.bfo_s_primes(G.dev(s), G.op_tables_to_device(tab), G.dev(tab.eval_prefactors()))
    assert sp.shape == (0, 4) and m.shape == (5, 0) and int(cnt.sum()) == 0


@pytest.mark.parametrize("name", ["tfim1d", "tfim_field", "sx_sysz", "heis", "single_diag", "same_site"])
@pytest.mark.parametrize("bias", [False, True])
def test_eloc_fused_and_generic(name, bias):
    N = 10
    G.check_eloc(G.operator_zoo(N)[name], N=N, M=20, bias=bias)


@pytest.mark.parametrize("M,bias", [(33, False), (95, True), (512, False), (600, False), (1100, True), (2500, False),
                                    (4500, False)])
@pytest.mark.parametrize("name", ["tfim1d", "heis", "splus"])
def test_eloc_hidden_unit_splits(name, M, bias):
    """ragged unrolls (M not a multiple of 32), 2 / 4 / 8 warps per sample (M > 512) and the shared-memory path (M > 4096)."""
    G.check_eloc(G.operator_zoo(6)[name], N=6, M=M, B=37, bias=bias)


def test_eloc_2d_tfim_config2_shape():
    G.check_eloc(obfo.tfim_strings((10, 10), 3.04), N=100, M=400, B=64)


def test_grad_layouts():
    G.check_grad(bias=True)
    G.check_grad(bias=False)
    G.check_grad(N=20, M=40, B=9, bias=False)


@pytest.mark.parametrize("N,M,B,bias,uniform,tile", [
    (7, 24, 211, True, True, 0), (7, 24, 211, False, False, 0), (5, 80, 100, False, True, 0),
    (3, 40, 37, True, False, 0), (4, 100, 64, False, True, 0), (6, 160, 50, True, True, 0),
    (5, 80, 100, False, True, 64), (20, 40, 450, False, True, 0), (2, 70, 1, False, True, 0)])
def test_moments_and_gram(N, M, B, bias, uniform, tile):
    G.check_moments_gram(N, M, B, bias, uniform=uniform, tile=tile)


@pytest.mark.parametrize("N,M,B,bias,uniform", [
    (7, 24, 211, True, True), (7, 24, 211, False, False), (5, 80, 100, False, True), (3, 40, 37, True, False),
    (4, 100, 64, False, True), (6, 160, 50, True, True), (20, 40, 450, False, True), (2, 70, 1, False, True),
    (3, 33, 30000, True, False), (2, 130, 54000, False, True)])
def test_gram_int8_tensor_cores(N, M, B, bias, uniform):
    """tcgen05 INT8 Ozaki-split Gram: same 1e-10 bar as the fp64 kernel (ragged tiles, > 26624 samples = 2-3 launches)."""
    G.check_moments_gram(N, M, B, bias, uniform=uniform, backend="i8")


@pytest.mark.parametrize("n,cplx,kind", [(700, True, "gram"), (701, True, "random"), (512, False, "gram"),
                                         (600, True, "clustered"), (400, True, "blockdiag"), (333, False, "random")])
def test_eigh_large_cuppen_tearing(n, cplx, kind):
    """The eigen-decomposition path for n > 32768 (cuSOLVER's limit), exercised at sizes where the direct solver runs."""
    G.check_eigh_large(n, cplx, kind)


def test_eigh_inplace_tears_above_threshold():
    """eigh_inplace routes to the tearing path above kernels.EIGH_TEAR_ABOVE; the TDVP solve on top of it is unchanged."""
    from vmc_jax_b200 import kernels as K
    old = K.EIGH_TEAR_ABOVE
    try:
        K.EIGH_TEAR_ABOVE = 16
        G.check_tdvp_solve()
    finally:
        K.EIGH_TEAR_ABOVE = old


@pytest.mark.parametrize("B", [4096, 30000])
def test_gram_heavy_tailed_columns(B):
    """Columns of tau with max >> rms: entry-wise accuracy |dA_jl| <= 1e-10 sqrt(A_jj A_ll); the automatic choice sends the
    outlier samples of the heavy-tailed matrix through the fp64 DMMA Gram and everything else through the int8 tensor-core
    Gram (1 / 2 launches, accumulating into the fp64 part)."""
    G.check_gram_heavy_tail(B=B)


@pytest.mark.parametrize("N,M,B,bias,uniform", [(7, 24, 211, True, False), (7, 24, 64, False, True),
                                                 (5, 40, 129, False, False), (40, 7, 65, True, False),
                                                 (3, 16, 1, False, True), (12, 33, 300, True, True)])
def test_tangent_kernel_T(N, M, B, bias, uniform):
    G.check_gram_T(N, M, B, bias, uniform=uniform)


@pytest.mark.parametrize("bias", [False, True])
def test_minsr_solve(bias):
    G.check_minsr_solve(bias=bias)


@pytest.mark.parametrize("makeReal,x,shift,bias", [('real', 1.0, 2.0, False), ('real', 1.0, 0.0, True),
                                                  ('imag', 1.j, 0.0, False), ('imag', 1.j, 0.0, True)])
def test_tdvp_solve(makeReal, x, shift, bias):
    G.check_tdvp_solve(makeReal=makeReal, rhsPrefactor=x, shift=shift, bias=bias)


def test_sampler_reference_weights():
    """reference tests/sampler_test.py:32-75 (bare CpxRBM, fixed 16-float weight vector)."""
    G.check_sampler_chi2(N=4, M=2, weights=REFG["rbm_weights"], numSamples=1_000_000)


@pytest.mark.parametrize("N,M", [(4, 8), (8, 16), (12, 24)])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sampler_chi2_spin_flip(N, M, seed):
    G.check_sampler_chi2(N=N, M=M, numSamples=1_000_000, seed=seed)


def test_sampler_chi2_16_sites():
    G.check_sampler_chi2(N=16, M=32, numSamples=2_000_000, C=1184)


@pytest.mark.parametrize("bias", [False, True])
def test_sampler_chi2_z2(bias):
    G.check_sampler_chi2(N=8, M=16, proposer="spin_flip_Z2", bias=bias, numSamples=1_000_000)


@pytest.mark.parametrize("bias", [False, True])
def test_sampler_chi2_zero_mag(bias):
    G.check_sampler_chi2(N=8, M=16, proposer="spin_flip_zeroMag", bias=bias, sector=True, numSamples=1_000_000)


@pytest.mark.parametrize("M,proposer,bias", [(70, "spin_flip", False), (600, "spin_flip", False),
                                             (1100, "spin_flip_Z2", True), (2100, "spin_flip_Z2", False)])
def test_sampler_chi2_hidden_unit_splits(M, proposer, bias):
    """ragged unroll and 2 / 4 / 8 warps per chain (M > 512): same chi-squared bar; weights scaled so that
    theta stays O(1)."""
    G.check_sampler_chi2(N=6, M=M, proposer=proposer, bias=bias, numSamples=500_000, C=296, amp=(6 * M) ** -0.25)


def test_sampler_mu1_and_rare_refresh():
    G.check_sampler_chi2(N=8, M=16, mu=1.0, numSamples=1_000_000)
    G.check_sampler_chi2(N=8, M=16, numSamples=1_000_000, refreshEvery=64)


def test_sampler_stream_independent_of_chain_partition():
    """Philox is keyed by global chain id: two half-size launches reproduce one full launch (multi-GPU sharding)."""
    import torch
    from oracle import rbm as orbm
    from vmc_jax_b200 import kernels as K
    N, M, C = 8, 16, 64
    W, b = orbm.init_o1(N, M, False, 3)
    dW = G.dev(W)
    tables = K.rbm_tables(dW, None)

    def run(chain0, nchains):
        st = torch.zeros((nchains, N), dtype=torch.int32, device=G.DEV)
        cnt = torch.zeros(2, dtype=torch.int64, device=G.DEV)
        out = K.rbm_mcmc(st, dW, None, tables, 99, 0, chain0, "spin_flip", 2.0, N, 5 * N, 7, cnt)
        return G.host(out).reshape(7, nchains, N), G.host(cnt)
    full, cf = run(0, C)
    a, ca = run(0, C // 2)
    bb, cb = run(C // 2, C // 2)
    assert np.array_equal(full, np.concatenate([a, bb], axis=1))
    assert np.array_equal(cf, ca + cb)


# ---------------------------------------------------------------- orbit-averaged RBM (SymNet around CpxRBM)
@pytest.mark.parametrize("kind,L,M,args,bias,fac", [
    ("1d", 4, 2, ("translation", "reflection", "spinflip"), False, {}),
    ("1d", 6, 5, ("translation",), True, {"translation_factor": -1.0}),
    ("1d", 5, 3, ("reflection", "spinflip"), True, {"reflection_factor": -1.0, "spinflip_factor": -1.0}),
    ("1d", 7, 4, (), False, {}),
    ("2d", 3, 4, ("translation", "reflection", "rotation"), True, {}),
    ("2d", 2, 3, ("rotation", "spinflip"), False, {"rotation_factor": 1j}),
    ("1d", 20, 6, ("translation", "reflection", "spinflip"), False, {})])
def test_symnet_logpsi_and_gradients(kind, L, M, args, bias, fac):
    G.check_symrbm(kind, L, M, args, bias, **fac)


def test_symnet_sampler_reference_weights():
    """the reference's own sampler test case: SymNet(translation, reflection, spinflip) around CpxRBM(2), L = 4,
    fixed weight vector (tests/sampler_test.py:32-45)."""
    G.check_symrbm_sampler(L=4, M=2, weights=REFG["rbm_weights"])


@pytest.mark.parametrize("L,M,args,mu", [(6, 4, ("translation",), 2.0), (8, 3, ("translation", "reflection", "spinflip"), 2.0),
                                        (5, 2, ("reflection",), 1.0)])
def test_symnet_sampler_chi2(L, M, args, mu):
    G.check_symrbm_sampler(L=L, M=M, args=args, mu=mu)


# ---------------------------------------------------------------- real-parameter CNN (reference nets/cnn.py)
@pytest.mark.parametrize("shape,F,channels,strides,act,bias,flb", [
    ((6,), (3,), (3, 2), (1,), ("elu",), True, False),
    ((8,), (4,), (2,), (2,), ("tanh",), False, True),
    ((5,), (8,), (2, 2), (1,), ("poly5", "elu"), True, True),          # filter wider than the lattice (multiple wraps)
    ((4, 4), (2, 2), (3, 2), (1, 1), ("elu",), True, False),
    ((4, 6), (3, 2), (2, 3, 2), (2, 1), ("poly6", "relu", "square"), True, False),
    ((3, 3), (2, 2), (4,), (1, 1), ("elu",), False, False)])
def test_cnn_logpsi_and_gradients(shape, F, channels, strides, act, bias, flb):
    G.check_cnn(shape, F, channels, strides, act, bias, flb)


@pytest.mark.parametrize("shape,F,proposer,sector", [((6,), (3,), "spin_flip", False), ((2, 3), (2, 2), "spin_flip_Z2", False),
                                                   ((6,), (4,), "spin_flip_zeroMag", True)])
def test_cnn_sampler_chi2(shape, F, proposer, sector):
    G.check_cnn_sampler(shape=shape, F=F, proposer=proposer, sector=sector)


@pytest.mark.parametrize("shape,F,channels,act,proposer", [
    ((6, 6), (3, 3), (3, 2), ("elu",), "spin_flip_zeroMag"),
    ((6, 6), (3, 3), (3, 2), ("elu",), "spin_flip"),
    ((4, 6), (2, 3), (2, 3, 2), ("poly6", "tanh", "elu"), "spin_flip_Z2"),
    ((10,), (4,), (5,), ("poly5",), "spin_flip_zeroMag"),
    ((3, 4), (3, 3), (2, 2, 2), ("elu",), "spin_flip"),               # boxes wrap around the whole lattice
    ((12, 12), (3, 3), (6, 4), ("elu",), "spin_flip_zeroMag")])        # BASELINE configs[3] net
def test_cnn_incremental_sampler_equals_full_forward_sampler(shape, F, channels, act, proposer):
    G.check_cnn_inc_sampler(shape, F, channels, act, proposer)


def _heisenberg_strings(shape):
    Lx, Ly = shape if len(shape) == 2 else (shape[0], 1)
    out = []
    for x in range(Lx):
        for y in range(Ly):
            i = x * Ly + y
            nbrs = [x * Ly + (y + 1) % Ly] if Ly > 1 else []
            nbrs.append(((x + 1) % Lx) * Ly + y)
            for j in nbrs:
                if j == i:
                    continue
                out += [(-0.25, [G.obfo.Sx(i), G.obfo.Sx(j)]), (-0.25, [G.obfo.Sy(i), G.obfo.Sy(j)]),
                        (0.25, [G.obfo.Sz(i), G.obfo.Sz(j)])]
    return out


@pytest.mark.parametrize("shape,F,channels,act,kind", [
    ((4, 4), (3, 3), (3, 2), ("elu",), "heisenberg"),
    ((4, 4), (3, 3), (3, 2), ("elu",), "tfim"),
    ((4, 6), (2, 3), (2, 3, 2), ("poly6", "tanh", "elu"), "heisenberg"),
    ((10,), (4,), (5,), ("poly5",), "tfim"),
    ((10,), (3,), (4, 3), ("elu",), "heisenberg"),
    ((3, 4), (3, 3), (2, 2, 2), ("elu",), "heisenberg")])
def test_cnn_fused_eloc(shape, F, channels, act, kind):
    strings = _heisenberg_strings(shape) if kind == "heisenberg" else G.obfo.tfim_strings(shape, -0.8)
    G.check_cnn_eloc(strings, shape, F, channels, act)


@pytest.mark.parametrize("R,M", [(3, 5), (7, 24), (4, 64), (11, 33)])
def test_hermitian_mirror_blocks(R, M):
    """rebuilds the part of a Hermitian matrix below the block diagonal from the part above (multi-GPU reduction)."""
    import torch
    Pc = R * M
    g = torch.Generator(device="cuda").manual_seed(R * 100 + M)
    X = torch.randn((Pc, Pc), dtype=torch.float64, device="cuda", generator=g) + \
        1j * torch.randn((Pc, Pc), dtype=torch.float64, device="cuda", generator=g)
    A = (X + X.conj().T).contiguous()
    B = A.clone()
    blk = torch.arange(Pc, device="cuda") // M
    B[blk[:, None] > blk[None, :]] = float("nan")          # destroy everything below the block diagonal
    G.K.hermitian_mirror_blocks(B, M)
    assert torch.equal(B, A)
