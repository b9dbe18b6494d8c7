"""CPU tests of the host-side logic and of the C-ABI surface (no kernel launches)."""
import ctypes
import os

import numpy as np
import pytest
import torch

import vmc_jax_b200 as jVMC
import vmc_jax_b200.operator as op
from vmc_jax_b200 import _lib, mpi_wrapper as mpi
from vmc_jax_b200.util import stepper

from oracle import bfo as obfo, sampling as osamp


def test_library_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert len(names) >= 19 and "jvmc_rbm_mcmc" in names and "jvmc_rbm_gram_S" in names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib._SIGS.keys())
    l = _lib.load()
    assert l.jvmc_version() >= 100
    assert l.jvmc_error_string(-1) == b"invalid argument"
    assert l.jvmc_rbm_tables_elems(4, 3) == 4 * 3 + 4 + 3 + 1
    assert l.jvmc_rbm_moments_chunks(1) == 1 and l.jvmc_rbm_moments_chunks(10 ** 7) == 64


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        psi(torch.zeros((1, 3, 4), dtype=torch.int32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        jVMC.sampler.MCSampler(psi, (4,), 0, updateProposer=jVMC.sampler.propose_spin_flip)


def test_unsupported_components_raise():
    with pytest.raises(NotImplementedError):
        jVMC.vqs.NQS(object())
    with pytest.raises(NotImplementedError):
        jVMC.vqs.NQS((jVMC.nets.CpxRBM(), jVMC.nets.RBM()))        # two-network ansatz: pair of real RBMs only
    with pytest.raises(NotImplementedError):
        jVMC.nets.CNN(F=(3,), channels=(2,), periodicBoundary=False)
    with pytest.raises(NotImplementedError):
        jVMC.nets.CNN(F=(3,), channels=(2,), actFun=(abs,))
    assert jVMC.vqs.NQS((jVMC.nets.RBM(), jVMC.nets.RBM())).kind == "tworbm"


def test_opstr_algebra():
    """reference tests/operator_test.py:256-265."""
    op1, op2 = op.Sz(3), op.Sx(5)
    opstr1 = 13. * op1 * op2
    opstr2 = 1.j * opstr1 * op1
    assert np.allclose(opstr2[0](), 13.j)
    for o in opstr2[1:]:
        assert isinstance(o, (op.LocalOp, dict))
    s = op.scal_opstr(lambda t: 2.0 * t, (op.Sp(0),))
    assert s[0](3.0) == 6.0
    s2 = 0.5 * s
    assert s2[0](3.0) == 3.0
    with pytest.raises(RuntimeError):
        op.scal_opstr(2.0, op.Sz(0))


def _product_ham(L, g, h=None):
    H = op.BranchFreeOperator()
    S = []
    for l in range(L):
        H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz((l + 1) % L))))
        S.append((-1., [obfo.Sz(l), obfo.Sz((l + 1) % L)]))
        H.add(op.scal_opstr(g, (op.Sx(l),)))
        S.append((g, [obfo.Sx(l)]))
        if h is not None:
            H.add(op.scal_opstr(h, (op.Sy(l), op.Sz((l + 1) % L), op.Sp((l + 2) % L))))
            S.append((h, [obfo.Sy(l), obfo.Sz((l + 1) % L), obfo.Sp((l + 2) % L)]))
    return H, obfo.Tables(S)


def test_compile_tables_match_oracle():
    """BranchFreeOperator.compile (reference branch_free.py:348-441): padding with Id(0), reversed order."""
    H, T = _product_ham(5, 0.7, h=lambda t: 0.1 * t)
    tab = H.compile()
    assert np.array_equal(tab.idx, T.idx) and np.array_equal(tab.map, T.map)
    assert np.array_equal(tab.matEls, T.matEls) and np.array_equal(tab.fermionic, T.fermi)
    assert np.array_equal(tab.diag, T.diag)
    assert tab.maxOpStrLength == 3
    assert not tab.fused_ok(object())
    # the fused local-energy kernels are chosen by the kind of net and the shape of the strings (<= 2 changed sites)

    class _Psi:
        def __init__(self, kind, kr):
            self.kind, self.khatri_rao, self.logarithmic = kind, kr, True

        def flip_tables(self):
            return None
    assert tab.fused_ok(_Psi("rbm", True)) and not tab.cnn_fused_ok(_Psi("rbm", True))
    assert tab.cnn_fused_ok(_Psi("cnn", False)) and not tab.fused_ok(_Psi("cnn", False))
    H3 = op.BranchFreeOperator()
    H3.add(op.scal_opstr(1., (op.Sx(0), op.Sx(1), op.Sx(2))))       # three flipped sites: generic s' route
    t3 = H3.compile()
    assert not t3.cnn_fused_ok(_Psi("cnn", False)) and not t3.fused_ok(_Psi("rbm", True))
    # fermionic flags
    F = op.BranchFreeOperator()
    F.add(op.scal_opstr(1., (op.creation(1), op.annihilation(0))))
    F.add(op.scal_opstr(2., (op.number(0),)))
    ft = F.compile()
    assert ft.fermionic.tolist() == [[1, 1], [0, 0]] and ft.isDiag.tolist() == [0, 1]
    assert not ft.cnn_fused_ok(_Psi("cnn", False))                  # Jordan-Wigner strings: generic route


def test_td_prefactor_compiles():
    """reference tests/operator_test.py:164-172."""
    hamiltonian = op.BranchFreeOperator()
    hamiltonian.add((op.Sz(0),))
    hamiltonian.add((op.Sz(1),))
    hamiltonian.add(op.scal_opstr(0.1, (op.Sx(0), op.Sx(1))))
    hamiltonian.compile()


@pytest.mark.parametrize("n,commSize,chains", [(4096, 1, 500), (10, 3, 1), (65536, 8, 1184), (7, 2, 3)])
def test_distribute_sampling_matches_oracle(n, commSize, chains, monkeypatch):
    for rank in range(commSize):
        monkeypatch.setattr(mpi, "_refresh", lambda: None)
        monkeypatch.setattr(mpi, "rank", rank)
        monkeypatch.setattr(mpi, "commSize", commSize)
        assert mpi.distribute_sampling(n, localDevices=1, numChainsPerDevice=chains) == \
            osamp.distribute_sampling(n, commSize, rank, 1, chains)
        assert mpi.distribute_sampling(n) == osamp.distribute_sampling(n, commSize, rank)
    # first_sample_id partitions [0, n)
    firsts = []
    for rank in range(commSize):
        monkeypatch.setattr(mpi, "rank", rank)
        cnt, _ = mpi.distribute_sampling(n)
        firsts.append((mpi.first_sample_id(), cnt))
    pos = 0
    for f, c in firsts:
        assert f == pos
        pos += c
    assert pos == n


def test_steppers_on_linear_ode():
    """reference tests/stepper_test.py:14-39."""
    from scipy.linalg import expm
    rng = np.random.default_rng(0)
    A = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    y0 = rng.normal(size=4) + 0j
    f = lambda y, t, intStep=0: A @ y
    st = stepper.AdaptiveHeun(timeStep=1e-2, tol=1e-7)
    y, t = y0.copy(), 0.0
    while t < 0.3:
        y, dt = st.step(t, f, y)
        t += dt
    assert np.linalg.norm(y - expm(A * t) @ y0) < 1e-5
    At = torch.as_tensor(A)
    ft = lambda y, t, intStep=0: At @ y
    yt, dt = stepper.Heun(timeStep=1e-3).step(0., ft, torch.as_tensor(y0))
    assert np.allclose(yt.numpy(), expm(A * 1e-3) @ y0, atol=1e-8)
    ye, _ = stepper.Euler(timeStep=1e-3).step(0., f, y0)
    assert np.allclose(ye, y0 + 1e-3 * A @ y0)


def test_sampler_argument_validation():
    class FakeNet:
        is_generator = False
    with pytest.raises(RuntimeError):
        jVMC.sampler.MCSampler(FakeNet(), (4,), 0, updateProposer=None)
    with pytest.raises(NotImplementedError):
        jVMC.sampler.MCSampler(FakeNet(), (4,), 0, updateProposer=lambda k, s, i: s)


@pytest.mark.parametrize("rows,cols", [(64, 40), (64, 32), (64, 48)])
@pytest.mark.parametrize("M", [1, 3, 4, 7, 40, 63, 64, 65, 100, 130, 333, 400])
def test_i8_gram_tile_list_covers_lower_triangle_once(M, rows, cols):
    """The INT8 Gram kernel writes (j, l <= j) of every site pair once per tile that covers it and ADDS on later
    launches, so the host-built tile list must cover the lower triangle exactly once."""
    from vmc_jax_b200.kernels import i8_tile_list
    cover = np.zeros((M, M), dtype=np.int32)
    pad = max((2 * M + 127) // 128 * 128, (2 * M + 2 * cols - 1) // (2 * cols) * (2 * cols))   # jvmc_i8_layout
    for rg, cg, nc, lo, hi in i8_tile_list(M, rows, cols):
        assert 4 * rg <= lo and hi <= 4 * rg + rows                     # written rows lie inside the tile
        assert nc % 16 == 0 and 16 <= nc <= 2 * cols and (8 * cg) % 2 == 0 and 8 * cg + nc <= pad
        assert 8 * rg + 2 * rows <= pad
        j = np.arange(max(lo, 4 * rg), min(hi, M))
        l = np.arange(4 * cg, min(4 * cg + nc // 2, M))
        jj, ll = np.meshgrid(j, l, indexing="ij")
        cover[jj[ll <= jj], ll[ll <= jj]] += 1
    want = np.tril(np.ones((M, M), dtype=np.int32))
    assert np.array_equal(cover, want)


@pytest.mark.parametrize("kind,L,args,fac", [
    ("1d", 4, ("translation", "reflection", "spinflip"), {}),
    ("1d", 6, ("translation",), {"translation_factor": -1.0}),
    ("1d", 5, ("reflection", "spinflip"), {"reflection_factor": -1.0, "spinflip_factor": -1.0}),
    ("2d", 3, ("translation", "reflection", "rotation"), {"rotation_factor": 1j}),
    ("2d", 2, ("rotation", "spinflip"), {})])
def test_symmetry_orbits_match_oracle(kind, L, args, fac):
    """jVMC.util.symmetries mirror (index form used by the kernels) vs the dense restatement of the reference."""
    import warnings
    from oracle import symmetries as osym
    from vmc_jax_b200.util import symmetries as psym
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if kind == "1d":
            ls, (o, f) = psym.get_orbit_1D(L, *args, **fac), osym.orbit_1d(L, *args, **fac)
        else:
            ls, (o, f) = psym.get_orbit_2D_square(L, *args, **fac), osym.orbit_2d_square(L, *args, **fac)
    assert np.array_equal(ls.orbit, o) and np.allclose(ls.factor, f)
    n = o.shape[1]
    x = np.random.default_rng(0).normal(size=n)
    for g in range(o.shape[0]):
        assert np.allclose(o[g] @ x, ls.sign[g] * x[ls.perm[g]])
        assert np.array_equal(ls.perm[g][ls.inv[g]], np.arange(n))


def test_symmetry_orbit_sizes_reference_cases():
    """reference tests/symmetries_test.py:17-44: every on/off combination of the lattice symmetries (None entries in
    the argument list, as the reference test passes them), integer orbit matrices, expected group orders in 1-D."""
    import warnings
    from vmc_jax_b200.util import symmetries
    L = 3
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for rotation in (True, False):
            for reflection in (True, False):
                for translation in (True, False):
                    for spinflip in (True, False):
                        symms = ("rotation" if rotation else None, "reflection" if reflection else None,
                                 "translation" if translation else None, "spinflip" if spinflip else None)
                        orbit = symmetries.get_orbit_2D_square(L, *symms)
                        assert np.issubdtype(orbit.dtype, np.integer) and orbit.shape[1:] == (L * L, L * L)
        for translation in (True, False):
            for reflection in (True, False):
                for spinflip in (True, False):
                    symms = ("reflection" if reflection else None, "translation" if translation else None,
                             "spinflip" if spinflip else None)
                    orbit = symmetries.get_orbit_1D(L, *symms)
                    assert orbit.shape[0] == (2 if reflection else 1) * (L if translation else 1) * (2 if spinflip else 1)
                    assert np.issubdtype(orbit.dtype, np.integer)


def test_output_manager_layout(tmp_path):
    """reference tests/output_manager_test.py:17-60 against the npz-backed tree (h5py is not in this image): same group
    and dataset names, rows appended along axis 0, scalars stored as shape-(1,) rows."""
    from vmc_jax_b200.util.output_manager import OutputManager
    fn = str(tmp_path / "test.h5")
    outp = OutputManager(fn, append=False, backend="npz")
    outp.set_group("test")
    with np.load(outp.path) as z:
        assert list(z["__groups__"]) == ["/test"]
    x = np.array([13.2])
    outp.write_metadata(0.3, bla=x)
    assert outp.has("/test/metadata") and outp.has("/test/metadata/bla")
    assert np.allclose(outp.read_dataset("/test/metadata/bla")[0], x)
    y = 99.1
    outp.write_metadata(0.5, bla=y)
    assert np.allclose(outp.read_dataset("/test/metadata/bla"), np.array([x, [y]]))
    assert np.allclose(outp.read_dataset("/test/metadata/times"), [0.3, 0.5])
    x = np.random.uniform(1, 2, size=(13,))
    y = np.random.uniform(-1, 1, size=(3,))
    outp.write_observables(0.1, obs1={"mean": x}, obs2={"mean": torch.as_tensor(y)})
    assert outp.has("/test/observables/obs1")
    assert np.allclose(outp.read_dataset("/test/observables/obs1/mean")[0], x)
    assert np.allclose(outp.read_dataset("/test/observables/obs2/mean")[0], y)
    outp.write_observables(0.5, bla={"mean": 99.1})
    assert np.allclose(outp.read_dataset("/test/observables/bla/mean")[0], np.array([99.1]))
    # checkpoints: flat parameter vectors, retrieved by index or by nearest time; the archive on disk is complete
    w0, w1 = np.arange(6.0), np.arange(6.0) + 0.5
    outp.write_network_checkpoint(0.0, w0)
    outp.write_network_checkpoint(1.0, torch.as_tensor(w1))
    t, w = outp.get_network_checkpoint(time=0.9)
    assert t == 1.0 and np.allclose(w, w1)
    t, w = outp.get_network_checkpoint(idx=0)
    assert t == 0.0 and np.allclose(w, w0)
    outp.write_error_data("bad", np.array([1, 2, 3]))
    again = OutputManager(fn, group="test", append=True, backend="npz")
    assert np.allclose(again.read_dataset("/test/network_checkpoints/checkpoints"), np.stack([w0, w1]))
    assert np.array_equal(again.read_dataset("/error_data/bad"), [1, 2, 3])
    again.write_metadata(0.7, bla=1.0)
    assert again.read_dataset("/test/metadata/bla").shape == (3, 1)
    # timers
    outp.start_timing("a"); outp.stop_timing("a"); outp.add_timing("b", 2.0)
    assert outp.timings["a"]["count"] == 1 and outp.timings["b"]["total"] == 2.0
    outp.print_timings(indent=" ")
    assert outp.timings["b"]["last_total"] == 2.0
    assert OutputManager(None).timings == {}


def test_bench_reference_arm_and_argument_handling():
    """bench.py: the reference arm is recognised before numpy is imported (thread pinning) and the CPU step functions
    run at a reduced size; the driver's `--steps 20 --warmup 5` needs no pre-sized event pool any more."""
    import bench
    assert bench._is_reference_arm(["bench.py", "--impl", "reference"]) and bench._is_reference_arm(["x", "--impl=reference"])
    assert not bench._is_reference_arm(["bench.py", "--gpus", "1"])
    t, d = bench.cpu_reference_step((4, 4), 3.04, 2, False, 16, 8, gram_block=100)
    assert d["samples"] == 16 and d["gram_cols"] == 16 * 32 and t > 0
    t, d = bench.cpu_tdvp_step(L=6, alpha=1, nsamp=200, chains=50)
    assert d["samples"] == 200 and np.isfinite(d["update_norm"])
    # the MinSR workload (BASELINE configs[4]) has its own CPU step and is a registered workload of both arms
    t, d = bench.cpu_minsr_step((3, 3), 3.04, 2, 12, 4)
    assert d["samples"] == 12 and np.isfinite(d["update_max_abs"]) and d["update_max_abs"] > 0
    assert bench.MINSR_WORKLOAD in bench.WORKLOADS and bench.WORKLOADS[bench.MINSR_WORKLOAD][0] == (20, 20)
    src = open(bench.__file__).read()
    assert "args.steps + args.warmup + 8" not in src
