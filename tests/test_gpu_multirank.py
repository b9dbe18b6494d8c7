"""-m gpu: N-rank end-to-end equality -- 2 ranks x C chains give the SAME configurations and statistics as 1 rank x 2C
chains (the reference runs its whole suite under `mpirun -n 2`, .github/workflows/automatic_testing.yml:25,
tests/mpi_wrapper_test.py:24-40).  With two GPUs the ranks talk NCCL, on a one-GPU box they share cuda:0 over gloo."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, out, chains, samples, mu, wscale, gram):
    env = dict(os.environ)
    env["JVMC_GRAM_BACKEND"] = gram
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    worker = [os.path.join(HERE, "multirank_worker.py"), out, str(chains), str(samples), str(mu), str(wscale)]
    if world == 1:
        cmd = [sys.executable] + worker
    else:
        env["JVMC_DIST_BACKEND"] = "nccl" if torch.cuda.device_count() >= world else "gloo"
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + worker
    # own process group: a timeout kills the launcher AND the ranks
    proc = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, start_new_session=True)
    try:
        out_, _ = proc.communicate(timeout=300)
    except subprocess.TimeoutExpired:
        import signal
        os.killpg(proc.pid, signal.SIGKILL)
        out_, _ = proc.communicate()
        raise AssertionError("multi-rank worker timed out:\n" + out_[-3000:])
    assert proc.returncode == 0, out_[-3000:]
    return [np.load(out + ".rank%d.npz" % k) for k in range(world)]


@pytest.mark.parametrize("mu,wscale,gram", [(2.0, 0.08, "i8"), (1.0, 0.3, "dmma")])
def test_two_ranks_equal_one_rank(tmp_path, mu, wscale, gram):
    """wscale 0.08, Born weights, int8 tensor-core Gram: its column scales follow the rank-local maxima, so A of the 1- and
    2-rank runs agrees to the splitting accuracy; wscale 0.3, mu = 1 (non-uniform weights), fp64 Gram: theta reaches the
    poles of tanh and the 3840 samples of this 12-site chain repeat configurations many times (coherent digit errors),
    so the rank-equality is asserted on the exact kernel, to round-off."""
    chains, samples = 96, 96 * 40
    one = _run(1, str(tmp_path / "w1"), chains, samples, mu, wscale, gram)[0]
    two = _run(2, str(tmp_path / "w2"), chains, samples, mu, wscale, gram)
    assert int(one["nglob"][0]) == int(two[0]["nglob"][0]) == samples
    N = one["configs"].shape[-1]
    c1 = one["configs"].reshape(-1, chains, N)                         # time-major, chain-minor (sampler.py:323)
    c2 = np.concatenate([t["configs"].reshape(-1, chains // 2, N) for t in two], axis=1)
    assert np.array_equal(c1, c2)                                      # Philox keyed by the GLOBAL chain id: bit-equal
    l1 = one["logPsi"].reshape(-1, chains)
    l2 = np.concatenate([t["logPsi"].reshape(-1, chains // 2) for t in two], axis=1)
    assert np.allclose(l1, l2, rtol=1e-13, atol=1e-13)
    p1 = one["p"].reshape(-1, chains)
    p2 = np.concatenate([t["p"].reshape(-1, chains // 2) for t in two], axis=1)
    assert np.allclose(p1, p2, rtol=1e-12) and abs(p2.sum() - 1.0) < 1e-12
    for t in two:                                                      # global statistics, replicated on both ranks
        assert np.allclose(t["Emean"], one["Emean"], rtol=1e-12)
        assert np.allclose(t["Evar"], one["Evar"], rtol=1e-11)
        assert np.allclose(t["F"], one["F"], rtol=1e-10, atol=1e-12 * np.abs(one["F"]).max())
        # the int8 Gram's column scales follow the rank-local maxima: the 1- and 2-rank matrices agree to the splitting
        # accuracy (entry-wise, relative to sqrt(A_jj A_ll)), not to round-off
        nat = np.sqrt(np.outer(np.real(np.diag(one["A"])), np.real(np.diag(one["A"]))))
        if gram == "i8":
            assert np.max(np.abs(t["A"] - one["A"]) / nat) <= 2e-11
        else:   # A = second moment - mean correction: round-off is relative to the uncentred moment (= the largest |A|)
            assert np.abs(t["A"] - one["A"]).max() <= 1e-12 * np.abs(one["A"]).max()
        assert np.allclose(t["acc"], one["acc"], rtol=1e-12)
        assert float(t["herm_err"][0]) < 1e-12
        assert np.allclose(t["update"], one["update"], rtol=1e-6, atol=1e-8 * np.abs(one["update"]).max())
        assert np.allclose(t["update2"], one["update2"], rtol=1e-6, atol=1e-7 * np.abs(one["update2"]).max())
        assert float(t["comm_err"][0]) == 0.0
        assert bool(t["capi_comm"][0]) == (torch.cuda.device_count() >= 2)
        # MinSR: T rows / columns are ordered rank-major (gather), the one-rank run is time-major over all chains
        assert t["T"].shape == (samples, samples)
        i1 = np.arange(samples).reshape(-1, chains)
        perm = np.concatenate([i1[:, k * (chains // 2):(k + 1) * (chains // 2)].reshape(-1) for k in range(2)])
        T1 = one["T"][np.ix_(perm, perm)]
        assert np.abs(t["T"] - T1).max() <= 1e-12 * np.abs(T1).max()
        assert np.allclose(t["update_minsr"], one["update_minsr"], rtol=1e-6, atol=1e-8 * np.abs(one["update_minsr"]).max())
    assert np.array_equal(two[0]["A"], two[1]["A"])
    assert np.array_equal(two[0]["T"], two[1]["T"])
    # ... and the two-rank tangent kernel against the oracle (stats.py:332-336 on dense holomorphic gradients) on the
    # gathered (rank-major) samples
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import rbm as orbm, stats as ostats
    W, b = orbm.unflatten_params(two[0]["params"], N, 20, bias=True)
    s_all = np.concatenate([t["configs"].reshape(-1, N) for t in two])
    p_all = np.concatenate([t["p"].reshape(-1) for t in two])
    T_ref = ostats.SampledObs(orbm.gradients_holomorphic(s_all, W, b), p_all).tangent_kernel()
    assert np.abs(two[0]["T"] - T_ref).max() <= 1e-10 * np.abs(T_ref).max()
