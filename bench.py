#!/usr/bin/env python
"""bench.py -- VMC step throughput (sample + E_loc + S/F) on B200, next to the reference's CPU algorithm.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's algorithm (oracle port) on the host cores

Workload (BASELINE.json configs[1]): 2D TFIM 10x10 (pbc) at g = 3.04, CpxRBM alpha = 4 (N = 100, M = 400,
P_c = 40 000 complex parameters), 2^16 samples per GPU (weak scaling), Metropolis sampler with
sweepSteps = N, 25 thermalisation sweeps, propose_spin_flip.  One step = MCSampler.sample (MCMC kernel + logpsi/tau)
-> BranchFreeOperator.get_O_loc (fused kernel) -> <E>, Var E, mu = <O>, F = <O* E>_c -> S = <O* O>_c
(Hermitian P_c x P_c, DMMA Gram kernel) [-> all-reduce of mu, F, S over ranks].  Synthetic random-init weights
(Re, Im ~ U(-1/sqrt(N), 1/sqrt(N)), seed 4321).

Prints ONE JSON line (rank 0)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time


def _is_reference_arm(argv):
    for i, a in enumerate(argv):
        if a == "--impl=reference" or (a == "--impl" and i + 1 < len(argv) and argv[i + 1] == "reference"):
            return True
    return False


if _is_reference_arm(sys.argv):
    # the CPU arm uses every host core whatever the launcher exported (torch.distributed.run sets OMP_NUM_THREADS=1):
    # the BLAS pools read these when numpy is first imported
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (lattice shape, g, alpha, bias, samples per GPU, numChains per GPU)
    "tfim2d_10x10_cpxrbm_a4_2p16": ((10, 10), 3.04, 4, False, 2 ** 16, 2368),
    "tfim1d_L20_cpxrbm_a2_2p12": ((20,), -0.7, 2, False, 2 ** 12, 500),
    "tfim1d_L40_cpxrbm_a2_bias_2p16": ((40,), -0.7, 2, True, 2 ** 16, 2368),
    "tfim2d_6x6_cpxrbm_a4_2p14": ((6, 6), 3.04, 4, False, 2 ** 14, 1184),
    # BASELINE configs[4]: 2^17 samples per GPU (2^20 on 8), MinSR step (run_minsr_workload)
    "tfim2d_20x20_cpxrbm_a4_2p20": ((20, 20), 3.04, 4, False, 2 ** 17, 1184),
}
MINSR_WORKLOAD = "tfim2d_20x20_cpxrbm_a4_2p20"
MINSR_METRIC = "MinSR step samples/sec (sample + E_loc + tangent kernel T + pinv + update)"
DEFAULT_WORKLOAD = "tfim2d_10x10_cpxrbm_a4_2p16"
_REAL_STDOUT = 1
METRIC = "VMC step samples/sec (sample + E_loc + S/F)"
UNIT = "samples/s"
TDVP_METRIC = "TDVP step ms (sample + E_loc + O_k + S/F + eigh + SNR + regularised solve)"
TDVP_WORKLOAD = "tfim1d_L20_cpxrbm_a2_2p12"


def o1_weights(N, M, bias, seed=4321):
    rng = np.random.default_rng(seed)
    a = 1.0 / np.sqrt(N)
    W = rng.uniform(-a, a, (N, M)) + 1j * rng.uniform(-a, a, (N, M))
    b = (rng.uniform(-a, a, M) + 1j * rng.uniform(-a, a, M)) if bias else None
    return W, b


def flat_params(W, b):
    leaves = ([b] if b is not None else []) + [W]
    return np.concatenate([np.concatenate([x.ravel().real, x.ravel().imag]) for x in leaves])


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for k, nme in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [c for c, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------- CPU arm
REF_SAMPLES, REF_CHAINS = 64, 32          # bounded sample of the workload one CPU "step" runs


def host_threads():
    """Threads the BLAS pool really uses (threadpoolctl), else the core count."""
    try:
        from threadpoolctl import threadpool_info
        n = [int(x.get("num_threads", 0)) for x in threadpool_info() if x.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    return os.cpu_count() or 1


def cpu_reference_step(shape, g, alpha, bias, nsamp, chains, rng_seed=0, gram_block=4000):
    """The reference's algorithm on the host cores (oracle port): Metropolis with a full forward pass per proposal
    (jVMC/sampler.py:327-356), get_s_primes -> psi(s') local energy (jVMC/operator/base.py:166-192), materialised
    per-sample gradients and the zgemm Gram (jVMC/stats.py:52-58).  The whole P_c x P_c Gram is EXECUTED, in column
    blocks of ``gram_block`` (each block is reduced to a checksum and dropped: the 25.6 GB host matrix of config 2 does
    not fit every box).  The reference itself forms the doubled P x P = (2 P_c)^2 matrix, i.e. 4x this zgemm.
    Returns (seconds for nsamp samples, detail)."""
    from oracle import rbm as orbm, bfo as obfo, sampling as osamp
    N = int(np.prod(shape))
    M = alpha * N
    W, b = o1_weights(N, M, bias)
    ham = obfo.Tables(obfo.tfim_strings(shape, g, -1.0))
    f = lambda s: orbm.cpx_rbm_logpsi(s, W, b)
    spc = max(1, nsamp // chains)
    # thermalisation amortised as in the full workload: 25 sweeps per 28 emitted samples per chain
    therm = max(1, int(round(25.0 * spc / 28.0)))
    smp = osamp.MCSampler(lambda s: np.real(f(s)), N, numChains=chains, proposer="spin_flip",
                          thermalizationSweeps=therm, sweepSteps=N, seed=rng_seed)
    t0 = time.perf_counter()
    cfg, _ = smp.sample(spc * chains)
    lp = f(cfg)
    t1 = time.perf_counter()
    E = obfo.get_O_loc(ham, cfg, f, 0.0, logPsiS=lp)
    t2 = time.perf_counter()
    p = np.ones(cfg.shape[0]) / cfg.shape[0]
    th = np.tanh(orbm.theta(cfg, W, b))
    sig = 2.0 * cfg - 1.0
    O = (sig[:, :, None] * th[:, None, :]).reshape(cfg.shape[0], -1)        # Khatri-Rao order, [n, P_c]
    mu = p @ O
    D = np.sqrt(p)[:, None] * (O - mu[None, :])
    Fv = D.conj().T @ (np.sqrt(p) * (E - np.sum(p * E)))
    Dh = np.ascontiguousarray(D.conj().T)
    t3 = time.perf_counter()
    chk, Pc = 0.0, D.shape[1]
    for c0 in range(0, Pc, gram_block):
        Ablk = Dh @ D[:, c0:c0 + gram_block]
        chk += float(np.abs(Ablk[::97, ::89]).sum())
    t4 = time.perf_counter()
    total = t4 - t0
    detail = {"samples": int(cfg.shape[0]), "t_sample_s": t1 - t0, "t_eloc_s": t2 - t1, "t_grad_moments_s": t3 - t2,
              "t_gram_s": t4 - t3, "gram_cols": int(Pc), "gram_block": int(min(gram_block, Pc)),
              "checksum": chk + float(np.abs(Fv).sum())}
    return total, detail


def cpu_tdvp_step(L=20, alpha=2, nsamp=4096, chains=500, seed=0):
    """One complete SR/TDVP step of BASELINE configs[0] by the reference's algorithm on the host (oracle port):
    sample (full forward per proposal, 25 thermalisation sweeps) -> E_loc via s' -> dense gradients -> S, F ->
    eigh(P x P) -> SNR -> regularised solve (jVMC/util/tdvp.py:219-290).  Returns (seconds, energy per site)."""
    from oracle import rbm as orbm, bfo as obfo, sampling as osamp, solve as osolve
    M = alpha * L
    W, b = orbm.init_cpx_rbm(L, M, False)
    ham = obfo.Tables(obfo.tfim_strings((L,), -0.7, -1.0))
    f = lambda s: orbm.cpx_rbm_logpsi(s, W, b)
    smp = osamp.MCSampler(lambda s: np.real(f(s)), L, numChains=chains, proposer="spin_flip_Z2",
                          thermalizationSweeps=25, sweepSteps=L, seed=seed)
    tdvp = osolve.TDVP(snrTol=2, pinvTol=1e-8, makeReal='real', rhsPrefactor=1., diagonalShift=10)
    t0 = time.perf_counter()
    cfg, glob = smp.sample(nsamp)
    lp = f(cfg)
    E = obfo.get_O_loc(ham, cfg, f, 0.0, logPsiS=lp)
    p = np.ones(cfg.shape[0]) / cfg.shape[0]
    upd, res, cut = osolve.tdvp_rhs(W, b, cfg, p, E, tdvp, glob, orbm.gradients_holomorphic)
    t1 = time.perf_counter()
    return t1 - t0, {"samples": int(cfg.shape[0]), "energy_per_site": float(np.real(tdvp.ElocMean)) / L,
                     "residual": float(res), "update_norm": float(np.linalg.norm(upd))}


def cpu_minsr_step(shape, g, alpha, nsamp, chains, rng_seed=0):
    """One MinSR step of BASELINE configs[4] by the reference's algorithm on the host (oracle port), bounded sample:
    Metropolis with a full forward pass per proposal, s' -> psi(s') E_loc, dense holomorphic gradients [n, 2 P_c],
    T = Obar Obar^dagger, pinv(T) e, update = -Obar^dagger x (jVMC/util/minsr.py:53-80).  Returns (seconds, detail)."""
    from oracle import rbm as orbm, bfo as obfo, sampling as osamp, stats as ostats, solve as osolve
    N = int(np.prod(shape))
    M = alpha * N
    W, b = o1_weights(N, M, False)
    ham = obfo.Tables(obfo.tfim_strings(shape, g, -1.0))
    f = lambda s: orbm.cpx_rbm_logpsi(s, W, b)
    spc = max(1, nsamp // chains)
    therm = max(1, int(round(25.0 * spc / 110.0)))      # 25 sweeps per 110 emitted samples per chain, as in the workload
    smp = osamp.MCSampler(lambda s: np.real(f(s)), N, numChains=chains, proposer="spin_flip",
                          thermalizationSweeps=therm, sweepSteps=N, seed=rng_seed)
    t0 = time.perf_counter()
    cfg, _ = smp.sample(spc * chains)
    lp = f(cfg)
    t1 = time.perf_counter()
    E = obfo.get_O_loc(ham, cfg, f, 0.0, logPsiS=lp)
    t2 = time.perf_counter()
    p = np.ones(cfg.shape[0]) / cfg.shape[0]
    oE, oG = ostats.SampledObs(E, p), ostats.SampledObs(orbm.gradients_holomorphic(cfg, W, b), p)
    upd = osolve.minsr_solve(oE, oG, True, pinvTol=1e-8)
    t3 = time.perf_counter()
    return t3 - t0, {"samples": int(cfg.shape[0]), "t_sample_s": t1 - t0, "t_eloc_s": t2 - t1, "t_minsr_s": t3 - t2,
                     "update_max_abs": float(np.abs(upd).max())}


def run_reference_arm(args, wl):
    """`--impl reference`: the reference's own algorithm for the path, on the host cores (jax cannot be installed in this
    image, so this is the oracle port, kind = "port").  Rank 0 alone works; every step is a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    if args.metric == "tdvp":
        for _ in range(min(args.warmup, 1)):
            cpu_tdvp_step()
        times = []
        for k in range(args.steps):
            t, detail = cpu_tdvp_step(seed=k)
            times.append(t)
        ms = float(np.mean(times)) * 1e3
        sample = ("oracle port of the reference algorithm (jax not installable): one complete TDVP/SR step of configs[0] "
                  "(1D TFIM L=20, CpxRBM alpha=2, 500 chains, %d samples), full size" % detail["samples"])
        line = {"metric": TDVP_METRIC, "value": ms, "unit": "ms", "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": TDVP_WORKLOAD, "samples_per_step": detail["samples"], "bounded_sample": False},
                "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "detail": detail}
        print(json.dumps(line), flush=True)
        return
    shape, g, alpha, bias, nsamp_gpu, chains_gpu = WORKLOADS[wl]
    nsamp, chains = REF_SAMPLES, REF_CHAINS
    if wl == MINSR_WORKLOAD:
        for _ in range(min(args.warmup, 1)):
            cpu_minsr_step(shape, g, alpha, nsamp, chains, rng_seed=1000)
        times = []
        for k in range(args.steps):
            t, detail = cpu_minsr_step(shape, g, alpha, nsamp, chains, rng_seed=k)
            times.append(t)
        sec = float(np.mean(times))
        value = nsamp / sec
        sample = ("oracle port of the reference algorithm (jax not installable): %d samples/step from %d chains, full forward "
                  "pass per proposal, s'->psi(s') E_loc, dense holomorphic gradients, MinSR on all %d samples" % (nsamp, chains, nsamp))
        line = {"metric": MINSR_METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl, "samples_per_step": nsamp, "bounded_sample": True},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "detail": detail}
        print(json.dumps(line), flush=True)
        return
    for _ in range(args.warmup):
        cpu_reference_step(shape, g, alpha, bias, nsamp, chains, rng_seed=1000)
    times = []
    for k in range(args.steps):
        t, detail = cpu_reference_step(shape, g, alpha, bias, nsamp, chains, rng_seed=k)
        times.append(t)
    sec = float(np.mean(times))
    value = nsamp / sec
    sample = ("oracle port of the reference algorithm (jax not installable): %d samples/step from %d chains, full forward "
              "pass per proposal, s'->psi(s') E_loc, dense O and the whole %d x %d zgemm Gram executed in column blocks of %d"
              % (nsamp, chains, detail["gram_cols"], detail["gram_cols"], detail["gram_block"]))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "samples_per_step": nsamp, "bounded_sample": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "detail": detail}
    print(json.dumps(line), flush=True)


def reference_arm_subprocess(extra, timeout=600):
    """Runs `bench.py --impl reference ...` in a fresh interpreter (own BLAS thread pool, all host cores) and returns
    its JSON line: the cpu_baseline leg of the GPU arm."""
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference"] + extra, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    raise RuntimeError("reference arm printed no JSON: " + out.stderr[-500:])


# ----------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--metric", default="vmc", choices=["vmc", "tdvp"],
                    help="vmc: samples/s of sample + E_loc + S/F at configs[1] (headline); tdvp: ms of one complete "
                         "TDVP/SR step at configs[0], the size the reference's CPU path runs in full")
    ap.add_argument("--samples", type=int, default=0, help="override samples per GPU")
    ap.add_argument("--minsr-samples", type=int, default=16384,
                    help="config-5 workload: global number N_T of samples the MinSR kernel T is formed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tdvp", action="store_true")
    ap.add_argument("--gram", default="i8", choices=["i8", "dmma"], help="Gram backend: tcgen05 INT8 (default) or fp64 DMMA")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = args.workload
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    if args.metric == "tdvp":
        run_tdvp_metric(args)
        return
    if wl == MINSR_WORKLOAD:
        run_minsr_workload(args)
        return
    os.environ["JVMC_GRAM_BACKEND"] = args.gram
    # stdout carries exactly ONE JSON line: everything libraries print to fd 1 (e.g. "NCCL version ..." at communicator
    # creation) is diverted to stderr, the JSON line is written to the saved descriptor at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import vmc_jax_b200 as jVMC
    from vmc_jax_b200 import kernels as K, mpi_wrapper as mpi
    from vmc_jax_b200.stats import SampledObs, RBMGradientObs
    import vmc_jax_b200.operator as op

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    mpi.init_distributed()
    rank, world = mpi.rank, mpi.commSize
    dev = jVMC.global_defs.myDevice
    shape, g, alpha, bias, nsamp, chains = WORKLOADS[wl]
    if args.samples:
        nsamp = args.samples
    N = int(np.prod(shape))
    M = alpha * N
    Pc = (N + (1 if bias else 0)) * M

    # ---- problem set-up (untimed): net, Hamiltonian, sampler
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=bias), seed=1234)
    psi(torch.zeros((1, 1) + shape, dtype=torch.int32, device=dev))
    W, b = o1_weights(N, M, bias)
    P_host = torch.as_tensor(flat_params(W, b)).pin_memory()
    psi.set_parameters(P_host.to(dev))
    H = op.BranchFreeOperator()
    if len(shape) == 1:
        L = shape[0]
        for l in range(L):
            H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz((l + 1) % L))))
            H.add(op.scal_opstr(g, (op.Sx(l),)))
    else:
        Lx, Ly = shape
        for x in range(Lx):
            for y in range(Ly):
                l = x * Ly + y
                H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(x * Ly + (y + 1) % Ly))))
                H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(((x + 1) % Lx) * Ly + y))))
                H.add(op.scal_opstr(g, (op.Sx(l),)))
    smp = jVMC.sampler.MCSampler(psi, shape, 4321, updateProposer=jVMC.sampler.propose_spin_flip, numChains=chains,
                                 sweepSteps=N, thermalizationSweeps=25, numSamples=nsamp * world)
    smp.refreshEvery = 8

    gram_events = []          # one (start, stop) CUDA-event pair per Gram call, appended as the steps run

    def vmc_step():
        s, logPsi, p = smp.sample()
        Eloc = H.get_O_loc(s, psi, logPsi, 0.0)
        E = SampledObs(Eloc, p)
        G = RBMGradientObs(psi, s, p)
        Emean, Evar = E.mean()[0], E.var()[0]
        F = G.covar(E)
        G.kr_mean()
        G._sigT = K.pack_sigma(G._s, G.hasBias)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        A = G.gram_A()
        g1.record()
        gram_events.append((g0, g1))
        return s, Emean, Evar, F, A

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(args.warmup):
        out = vmc_step()
    barrier()
    nLocal = out[0].shape[1]
    nGlobal = smp.get_last_number_of_samples()
    del out

    # ---- timed region: K steps, device-resident inputs
    clocks = ClockSampler(dev.index if dev.index is not None else 0)
    clocks.start()
    time.sleep(0.3)
    from vmc_jax_b200 import _lib as _jl
    launches0 = _jl.LAUNCHES
    i0 = len(gram_events)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = vmc_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _jl.LAUNCHES - launches0
    i1 = len(gram_events)
    gram_ms = float(np.mean([a.elapsed_time(b_) for a, b_ in gram_events[i0:i1]]))
    energy = complex(out[1].item())
    del out
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    ms_per_step = ms / args.steps
    value = nGlobal / (ms_per_step * 1e-3)

    # ---- e2e: same step through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    res_host = torch.empty(2 + 2 * Pc, dtype=torch.float64).pin_memory()
    e2e_steps = max(1, args.steps)
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        psi.set_parameters(P_host.to(dev, non_blocking=True))             # H2D: flat parameter vector
        s, Emean, Evar, F, A = vmc_step()
        res = torch.cat([torch.view_as_real(Emean.reshape(1)).reshape(-1)[:1], Evar.reshape(1).to(torch.float64),
                         torch.view_as_real(F.reshape(-1)[:Pc].contiguous()).reshape(-1)])
        res_host.copy_(res, non_blocking=True)                            # D2H: <E>, Var E, F (S stays on device)
        torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    clk = clocks.stop()
    tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_value = nGlobal / (float(tms.item()) / e2e_steps * 1e-3)
    h2d = int(P_host.numel() * 8)
    d2h = int(res_host.numel() * 8)
    del s, Emean, Evar, F, A

    # ---- per-phase breakdown (diagnostic, untimed w.r.t. the headline)
    phases = {}
    if rank == 0 or world > 1:
        def timed(fn):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = fn(); b_.record(); torch.cuda.synchronize()
            return r, a.elapsed_time(b_)
        (s, logPsi, p), phases["sampling_ms"] = timed(lambda: smp.sample())
        Eloc, phases["eloc_ms"] = timed(lambda: H.get_O_loc(s, psi, logPsi, 0.0))
        def mom():
            E = SampledObs(Eloc, p); G = RBMGradientObs(psi, s, p)
            return E, G, E.mean(), E.var(), G.covar(E), G.kr_mean()
        (E, G, *_), phases["moments_F_ms"] = timed(mom)
        _, phases["gram_S_ms"] = timed(lambda: G.gram_A())
        phases["acceptance"] = float(smp.acceptance_ratio())
        # logpsi alone (it is part of sampling_ms: the sampler re-evaluates logpsi of the emitted configurations)
        flat_ = s.reshape(s.shape[0] * s.shape[1], -1).contiguous()
        Wd_, bd_ = psi._cW()
        _, phases["logpsi_ms"] = timed(lambda: K.rbm_logpsi(flat_, Wd_, bd_))
        del E, G, Eloc, s, logPsi, p, flat_
    torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel (Gram): fp64 tensor pipe
    R = N + (1 if bias else 0)
    algo_flop = 4.0 * nLocal * float(Pc) ** 2                 # 4 N_s P_c^2 (complex Hermitian half)
    TS = 80 if (M % 80 == 0 or M == 40) else 64
    nT = (M + TS - 1) // TS
    exec_flop = 8.0 * nLocal * (R * (R + 1) / 2) * (nT * (nT + 1) / 2) * TS * TS   # DMMA flops issued (incl. padding)
    peak_tf, peak_src = None, None
    try:
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        c = torch.empty_like(a)
        best = 1e9
        for _ in range(4):
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record(); torch.matmul(a, a, out=c); x1.record(); torch.cuda.synchronize()
            best = min(best, x0.elapsed_time(x1))
        peak_tf = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        peak_src = "cuBLAS DGEMM 8192^3 measured live (best of 4); MEASURED_PEAKS.json holds no fp64 figure"
        del a, c
    except Exception as ex:  # pragma: no cover
        peak_tf, peak_src = 37.0, "fallback: DMMA issue-rate probe tools/fp64_probe.cu (37.05 TF/s); live DGEMM failed: %r" % (ex,)
    achieved = algo_flop / (gram_ms * 1e-3) / 1e12
    if args.gram == "dmma":
        roofline = {"bound": "tensor", "kernel": "gram_s_kernel (fp64 DMMA m8n8k4)", "achieved": achieved, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": None, "peak_source": peak_src,
                    "launch_ms": gram_ms,
                    "algorithmic_flop_per_launch": algo_flop,
                    "executed_tensor_flop_per_launch": exec_flop,
                    "executed_tflops": exec_flop / (gram_ms * 1e-3) / 1e12,
                    "note": "algorithmic = SURVEY 8d figure 4*N_s*P_c^2 (Hermitian half); the kernel additionally uses the "
                            "site-pair symmetry of the Khatri-Rao Gram and issues about half of that on the tensor pipe"}
    else:
        # INT8 tensor-core path: S is computed to fp64-equivalent accuracy by error-free splitting into 5 int8 digits,
        # 15 digit-pair products (levels k+k' <= 6) on tcgen05 kind::i8 with exact int32 accumulation.
        tile_cols = int(K._i8_tiles(M, dev)[:, 2].sum())      # real columns over all tiles (narrow tiles at the diagonal)
        stages = (nLocal + 31) // 32
        launches_i8 = (stages + 799) // 800
        exec_ops = 2.0 * (R * (R + 1) / 2) * stages * (15 * tile_cols) * 128 * 32
        # minimal work of the method: symmetric half of the 2M x 2M real Gram per site pair, 15 digit-pair products
        algo_ops = 2.0 * (R * (R + 1) / 2) * (2 * M * (2 * M + 1) / 2) * nLocal * 15
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as fh:
                pk = json.load(fh)
            # the Gram runs for seconds inside the step under the board power cap: sustained figure
            bf16 = float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"]))
            i8_src = ("2 x bf16_tflops_sustained of MEASURED_PEAKS.json (measured; B200 dense INT8 rate = 2 x dense bf16 "
                      "rate; burst figure bf16_tflops = %.1f)" % float(pk["bf16_tflops"]))
        else:
            bf16, i8_src = 1400.0, "2 x 1.4 PFLOP/s sustained bf16 fallback of B200_PROFILING.md"
        i8_peak = 2.0 * bf16
        ach = algo_ops / (gram_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "gram_s_i8_kernel (tcgen05.mma kind::i8, UTCIMMA) + i8 slicing",
                    "achieved": ach, "peak": i8_peak, "unit": "TOP/s", "frac": ach / i8_peak, "traffic": None,
                    "peak_source": i8_src, "launch_ms": gram_ms, "launches_per_step": launches_i8,
                    "algorithmic_int8_op_per_step": algo_ops,
                    "executed_int8_op_per_step": exec_ops,
                    "executed_tops": exec_ops / (gram_ms * 1e-3) / 1e12,
                    "executed_frac": exec_ops / (gram_ms * 1e-3) / 1e12 / i8_peak,
                    "algorithmic_flop_per_step": algo_flop,
                    "fp64_equivalent": {"achieved_tflops": achieved, "dgemm_peak_tflops": peak_tf,
                                        "ratio": achieved / peak_tf, "dgemm_peak_source": peak_src},
                    "note": "fp64-equivalent Gram (4*N_s*P_c^2 algorithmic flop, SURVEY 8d) executed as 15 int8 digit-pair "
                            "products per sample pair; achieved counts the symmetric half of each site pair's 2M x 2M "
                            "real Gram, executed also the padding of the 128 x 80 tiles that straddle the diagonal; "
                            "the tensor pipe waits on shared-memory bandwidth (TMA writes + sign-pass reads + UTCIMMA "
                            "B reads = 91 KB per 32-sample stage) under the board power cap, see DESIGN.md 4.2"}
    roofline["gram_backend_last_step"] = dict(K.LAST_GRAM) if hasattr(K, "LAST_GRAM") else None
    tp = os.path.join(ROOT, "profiles", "gram_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as fh:
                tj = json.load(fh)
            if tj.get("workload") == wl and tj.get("backend", "dmma") == args.gram:
                roofline["traffic"] = tj.get("dram_bytes_per_launch")
                roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # ---- TDVP step (sample + E_loc + O_k + S/F + SNR + solve) on the CPU-runnable config 1, for the "TDVP step ms" half
    tdvp_info = None
    other_configs = {}
    if not args.no_tdvp and world == 1:
        try:
            tdvp_info = tdvp_step_ms(jVMC, op, torch)
        except Exception as ex:  # pragma: no cover
            tdvp_info = {"error": repr(ex)}
        try:
            tdvp_info["config3"] = tdvp_step_ms_config3(jVMC, op, torch)
        except Exception as ex:  # pragma: no cover
            tdvp_info["config3"] = {"error": repr(ex)}
        torch.cuda.empty_cache()
        # the remaining BASELINE configs, bounded (auxiliary: not part of the headline metric)
        for name, fn in (("config4_cnn_heisenberg", lambda: aux_config4_cnn(jVMC, op, torch)),
                         ("config5_minsr_20x20", lambda: aux_config5_minsr(jVMC, op, torch, K))):
            try:
                other_configs[name] = fn()
            except Exception as ex:  # pragma: no cover
                other_configs[name] = {"error": repr(ex)}
            torch.cuda.empty_cache()

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    # ---- the fp64-pipe kernels (report achieved fp64 flop/s AND achieved HBM bytes/s, from the algorithmic
    # figures, against the measured peaks; the binding resource of all of them is the fp64 pipe, not HBM)
    kernels = None
    if phases:
        # fp64 denominator: cuBLAS DGEMM 8192^3 measured live above (MEASURED_PEAKS.json holds no fp64 figure)
        FP64_PEAK = float(peak_tf)
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            hbm_peak, hbm_src = 6550.0, "fallback of B200_PROFILING.md"
        acc = phases["acceptance"]
        spc = nLocal // chains
        props = chains * N * (25 + spc)
        Kmax = N + 1

        def entry(name, ms, flop, hbm_bytes, note):
            tf, gbs = flop / (ms * 1e-3) / 1e12, hbm_bytes / (ms * 1e-3) / 1e9
            return {"kernel": name, "ms": ms, "bound": "fp64", "fp64_tflops": tf, "fp64_frac": tf / FP64_PEAK,
                    "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak, "algorithmic": note}
        t_mc = max(phases["sampling_ms"] - phases["logpsi_ms"], 1e-6)
        kernels = {
            "fp64_peak_tflops": FP64_PEAK, "fp64_peak_source": peak_src,
            "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
            "list": [
                entry("rbm_mcmc_flip_kernel (sampler sweep incl. %d thermalisation sweeps)" % 25, t_mc,
                      props * M * (14.0 + 30.0 * acc), nLocal * (4.0 * N + 16.0),
                      "K M (14 + 30 acc) flop and 4N+16 HBM bytes per emitted sample; %.3g proposals, weight rows "
                      "stream from L2 (%.1f TB/s)" % (props, props * M * 16.0 / (t_mc * 1e-3) / 1e12)),
                entry("rbm_logpsi_kernel", phases["logpsi_ms"], nLocal * (4.0 * N * M + 100.0 * M),
                      nLocal * (4.0 * N + 16.0 * M + 16.0), "4NM + ~100M flop, 4N + 16M + 16 bytes per sample"),
                entry("rbm_eloc_kernel (fused E_loc)", phases["eloc_ms"], nLocal * 14.0 * N * M,
                      nLocal * (4.0 * N + 16.0 * M + 16.0), "14 N M flop, 4N + 16M + 16 bytes per sample"),
                entry("rbm_moments_kernel x2 + reductions (<O>, F)", phases["moments_F_ms"], nLocal * 16.0 * N * M,
                      2.0 * nLocal * (4.0 * N + 16.0 * M + 24.0), "8 N M flop and 4N + 16M + 24 bytes per sample and pass"),
            ]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = reference_arm_subprocess(["--workload", wl, "--steps", "8", "--warmup", "1"])
            cpu = dict(ref["cpu_baseline"], steps=ref["steps"], ms_per_step=ref["ms_per_step"], detail=ref.get("detail"))
        except Exception as ex:  # pragma: no cover
            cpu = {"error": repr(ex)}
        if tdvp_info is not None and "error" not in tdvp_info:
            try:
                ref = reference_arm_subprocess(["--metric", "tdvp", "--steps", "3", "--warmup", "1"])
                tdvp_info["reference_cpu"] = {"ms_per_step": ref["value"], "cores": ref["cpu_baseline"]["cores"],
                                              "kind": "port", "detail": ref.get("detail")}
                tdvp_info["speedup_vs_reference_cpu"] = ref["value"] / tdvp_info["ms_per_step"]
            except Exception as ex:  # pragma: no cover
                tdvp_info["reference_cpu"] = {"error": repr(ex)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl, "lattice": list(shape), "g": g, "numHidden": M, "P_complex": Pc,
                           "samples_per_gpu": nLocal, "samples_global": nGlobal, "numChains_per_gpu": chains,
                           "sweepSteps": N, "thermalizationSweeps": 25, "proposer": "spin_flip",
                           "l2": "inputs_exceed_L2 (tau %.0f MB, S %.1f GB per GPU)" % (nLocal * M * 16 / 1e6, Pc * Pc * 16 / 1e9),
                           "parallelism": "chains sharded, %d rank(s), all-reduce of mu/F/S" % world},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "note": "parameters H2D from pinned memory; <E>, VarE, F D2H; S stays on device for the solver"},
                "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
                "phases_ms": phases, "kernels": kernels, "energy_mean": [energy.real, energy.imag], "tdvp_step": tdvp_info, "other_configs": other_configs}
        sys.stdout.flush()
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def tdvp_step_ms(jVMC, op, torch, L=20, alpha=2, nsamp=4096, chains=500, steps=5, warm=2):
    """Full TDVP/SR step (TDVP.__call__: sample + E_loc + gradients + S,F + eigh + SNR + regularised solve) on
    BASELINE configs[0] (1D TFIM L=20, CpxRBM alpha=2, 500 chains, 2^12 -> 4500 samples, ex0-style SR)."""
    dev = jVMC.global_defs.myDevice
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=alpha * L, bias=False), seed=1234)
    psi(torch.zeros((1, 1, L), dtype=torch.int32, device=dev))
    H = op.BranchFreeOperator()
    for l in range(L):
        H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz((l + 1) % L))))
        H.add(op.scal_opstr(-0.7, (op.Sx(l),)))
    smp = jVMC.sampler.MCSampler(psi, (L,), 4321, updateProposer=jVMC.sampler.propose_spin_flip_Z2, numChains=chains,
                                 sweepSteps=L, numSamples=nsamp, thermalizationSweeps=25)
    tdvp = jVMC.util.TDVP(smp, rhsPrefactor=1., pinvTol=1e-8, diagonalShift=10, makeReal='real')
    stepper = jVMC.util.stepper.Euler(timeStep=1e-2)
    times = []
    for k in range(steps + warm):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        dp, _ = stepper.step(0, tdvp, psi.get_parameters(), hamiltonian=H, psi=psi, numSamples=None)
        psi.set_parameters(dp)
        b_.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b_))
    return {"config": "1D TFIM L=%d, CpxRBM alpha=%d (P=%d), %d chains, %d samples, SR (makeReal=real, diagonalShift=10)"
                      % (L, alpha, 2 * L * alpha * L, chains, smp.get_last_number_of_samples()),
            "ms_per_step": float(np.mean(times[warm:])), "ms_median": float(np.median(times[warm:])),
            "steps": steps, "timer": "CUDA events around Euler.step(TDVP.__call__) + set_parameters",
            "energy_per_site": float(tdvp.ElocMean0.real) / L}


def run_tdvp_metric(args):
    """`--metric tdvp`: the "TDVP step ms" half of BASELINE.json's metric, on configs[0] -- the configuration the
    reference's CPU path runs at full size, so `--impl reference --metric tdvp` is its exact counterpart."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    import torch
    import vmc_jax_b200 as jVMC
    from vmc_jax_b200 import mpi_wrapper as mpi, _lib as _jl
    import vmc_jax_b200.operator as op
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    mpi.init_distributed()
    clocks = ClockSampler(jVMC.global_defs.myDevice.index or 0)
    clocks.start()
    l0 = _jl.LAUNCHES
    info = tdvp_step_ms(jVMC, op, torch, steps=args.steps, warm=max(args.warmup, 3))
    launches = (_jl.LAUNCHES - l0) * args.steps // (args.steps + max(args.warmup, 3))
    clk = clocks.stop()
    if mpi.rank == 0:
        ms = info["ms_per_step"]
        line = {"metric": TDVP_METRIC, "value": ms, "unit": "ms", "n_gpus": mpi.commSize, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": False, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": TDVP_WORKLOAD, "detail": info["config"],
                           "l2": "latency-bound configuration (P = 1600): inputs fit L2, nothing to flush"},
                "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16,
                        "note": "TDVP.__call__ through the public API; the step reads <E> back (ElocMean0)"},
                "gpu_launches": launches, "clocks": clk, "tdvp_step": info}
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    if mpi.commSize > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_minsr_workload(args):
    """BASELINE configs[4] / north_star target lattice: 2D TFIM 20x20, CpxRBM alpha=4 (P_c = 640 000 complex parameters),
    2^17 samples per GPU (2^20 on 8 GPUs), one MinSR step (reference jVMC/util/minsr.py:82-165):
      sample + E_loc over ALL samples (<E>, Var E global), then the params > samples solve on N_T = --minsr-samples of
      them (every (N_s/N_T)-th sample of each rank, uniform weights): T = 2 Obar Obar^dagger formed ONCE over the ranks
      (all-gathered Khatri-Rao factors, rank k computes the k-th range of tile pairs, SUM all-reduce), pinv(T) e
      (replicated eigen-decomposition), update = -Obar^dagger x (local moments, all-reduced).
    The dense kernel on all 2^20 samples is 17.6 TB -- infeasible on any number of B200s, as it is for the reference."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import vmc_jax_b200 as jVMC
    from vmc_jax_b200 import kernels as K, mpi_wrapper as mpi, _lib as _jl
    from vmc_jax_b200.stats import SampledObs, RBMGradientObs
    from vmc_jax_b200.util.minsr import pinv_hermitian_apply
    import vmc_jax_b200.operator as op
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    mpi.init_distributed()
    rank, world = mpi.rank, mpi.commSize
    dev = jVMC.global_defs.myDevice
    wl = args.workload
    shape, g, alpha, bias, nsamp, chains = WORKLOADS[wl]
    if args.samples:
        nsamp = args.samples
    Lx, Ly = shape
    N = Lx * Ly
    M = alpha * N
    Pc = N * M
    nT_loc = max(64, args.minsr_samples // world)
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=bias), seed=1234)
    psi(torch.zeros((1, 1) + shape, dtype=torch.int32, device=dev))
    W, b = o1_weights(N, M, bias)
    P_host = torch.as_tensor(flat_params(W, b)).pin_memory()
    psi.set_parameters(P_host.to(dev))
    H = op.BranchFreeOperator()
    for x in range(Lx):
        for y in range(Ly):
            l = x * Ly + y
            H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(x * Ly + (y + 1) % Ly))))
            H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(((x + 1) % Lx) * Ly + y))))
            H.add(op.scal_opstr(g, (op.Sx(l),)))
    smp = jVMC.sampler.MCSampler(psi, shape, 4321, updateProposer=jVMC.sampler.propose_spin_flip, numChains=chains,
                                 sweepSteps=N, thermalizationSweeps=25, numSamples=nsamp * world)
    smp.refreshEvery = 8
    minsr = jVMC.util.MinSR(smp, pinvTol=1e-8)

    def subset(s, Eloc):
        nl = s.shape[1]
        idx = torch.arange(nT_loc, device=dev) * (nl // nT_loc) if nl >= nT_loc else torch.arange(nl, device=dev)
        ps = torch.full((1, idx.numel()), 1.0 / (idx.numel() * world), dtype=torch.float64, device=dev)
        return s[:, idx].contiguous(), Eloc[:, idx].contiguous(), ps

    def step():
        s, logPsi, p = smp.sample()
        Eloc = H.get_O_loc(s, psi, logPsi, 0.0)
        E = SampledObs(Eloc, p)
        Emean, Evar = E.mean()[0], E.var()[0]
        ss, Es, ps = subset(s, Eloc)
        upd = minsr.solve(SampledObs(Es, ps), RBMGradientObs(psi, ss, ps), holomorphic=True)
        return s, Emean, Evar, upd

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        out = step()
    barrier()
    nLocal, nGlobal = out[0].shape[1], smp.get_last_number_of_samples()
    del out
    clocks = ClockSampler(dev.index if dev.index is not None else 0)
    clocks.start()
    time.sleep(0.3)
    launches0 = _jl.LAUNCHES
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    launches = _jl.LAUNCHES - launches0
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    energy = complex(out[1].item())
    upd_norm = float(out[3].abs().max())
    del out
    # e2e: parameters from pinned host memory, <E>, Var E and the update read back every step
    res_host = torch.empty(2 + 4 * Pc, dtype=torch.float64).pin_memory()
    barrier()
    e0.record()
    for _ in range(args.steps):
        psi.set_parameters(P_host.to(dev, non_blocking=True))
        s, Emean, Evar, upd = step()
        res = torch.cat([Emean.real.reshape(1).to(torch.float64), Evar.reshape(1).to(torch.float64),
                         torch.view_as_real(upd.to(torch.complex128).contiguous()).reshape(-1)])
        res_host.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    clk = clocks.stop()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    del s, Emean, Evar, upd, res

    # per-phase breakdown (diagnostic): the pieces of MinSR.solve called one by one, max over ranks
    def timed(fn):
        barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b_.record(); torch.cuda.synchronize()
        return r, max_over_ranks(a.elapsed_time(b_))
    ph = {}
    (s, logPsi, p), ph["sampling_ms"] = timed(lambda: smp.sample())
    Eloc, ph["eloc_ms"] = timed(lambda: H.get_O_loc(s, psi, logPsi, 0.0))
    ss, Es, ps = subset(s, Eloc)
    G, ph["tau_of_subset_ms"] = timed(lambda: RBMGradientObs(psi, ss, ps))
    Eo = SampledObs(Es, ps)
    mu, _ = timed(lambda: G.kr_mean())
    if world > 1:
        (s_all, tau_all, p_all, v_all), ph["gather_factors_ms"] = timed(
            lambda: (mpi.gather(G._s[None]), mpi.gather(G._tau[None]), mpi.gather(G._p[None]),
                     mpi.gather(K.rbm_krmatvec(G._s, G._tau, mu.conj().contiguous(), False)[None])))
        T, ph["tangent_kernel_tiles_ms"] = timed(
            lambda: K.rbm_gram_T(s_all, tau_all, p_all, mu, False, 2.0, part=(rank, world), v=v_all))
        T, ph["allreduce_T_ms"] = timed(lambda: mpi._all_reduce_sum(T))
        del s_all, tau_all, p_all, v_all
    else:
        T, ph["tangent_kernel_tiles_ms"] = timed(lambda: K.rbm_gram_T(G._s, G._tau, G._p, mu, False, 2.0))
    e_all = mpi.gather(Eo._data).reshape(-1).to(T.dtype)
    x, ph["pinv_eigh_replicated_ms"] = timed(lambda: pinv_hermitian_apply(T, 1e-8, e_all))
    del T
    _, ph["update_contraction_ms"] = timed(lambda: G.minsr_contract(x))
    ph["acceptance"] = float(smp.acceptance_ratio())
    NT = nT_loc * world
    t_k = ph["tangent_kernel_tiles_ms"]
    ph["tangent_kernel_tflops_per_gpu"] = 4.0 * NT * NT * M / world / (t_k * 1e-3) / 1e12
    # roofline of the dominant OWN kernel (the sampler sweep; the replicated cuSOLVER eigen-decomposition is library code)
    try:
        n = 8192
        a_ = torch.randn(n, n, dtype=torch.float64, device=dev)
        c_ = torch.empty_like(a_)
        best = 1e9
        for _ in range(4):
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record(); torch.matmul(a_, a_, out=c_); x1.record(); torch.cuda.synchronize()
            best = min(best, x0.elapsed_time(x1))
        fp64_peak, peak_src = 2.0 * n ** 3 / (best * 1e-3) / 1e12, "cuBLAS DGEMM 8192^3 measured live (best of 4)"
        del a_, c_
    except Exception as ex:  # pragma: no cover
        fp64_peak, peak_src = 37.0, "fallback 37 TF/s (tools/fp64_probe.cu); live DGEMM failed: %r" % (ex,)
    acc = ph["acceptance"]
    props = chains * N * (25 + nLocal // chains)
    flop = props * M * (14.0 + 30.0 * acc)
    tf = flop / (ph["sampling_ms"] * 1e-3) / 1e12
    roofline = {"bound": "fp64", "kernel": "rbm_mcmc_flip_kernel (sampler sweep, 4 warps per chain) + logpsi of the emitted samples",
                "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak, "traffic": None,
                "peak_source": peak_src, "launch_ms": ph["sampling_ms"],
                "note": "K M (14 + 30 acc) flop per proposal; the weight rows (10.2 MB) are L2-resident, HBM traffic is "
                        "4N + 16 bytes per emitted sample -- the fp64 pipe, not HBM or the tensor pipe, bounds this workload's "
                        "own kernels; the step is dominated by the replicated cuSOLVER eigen-decomposition of T"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = reference_arm_subprocess(["--workload", wl, "--steps", "2", "--warmup", "0"])
            cpu = dict(ref["cpu_baseline"], steps=ref["steps"], ms_per_step=ref["ms_per_step"], detail=ref.get("detail"))
        except Exception as ex:  # pragma: no cover
            cpu = {"error": repr(ex)}
    if rank == 0:
        line = {"metric": MINSR_METRIC, "value": nGlobal / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl, "lattice": list(shape), "g": g, "numHidden": M, "P_complex": Pc,
                           "samples_per_gpu": nLocal, "samples_global": nGlobal, "numChains_per_gpu": chains,
                           "minsr_samples_global": NT, "minsr_samples_per_gpu": nT_loc, "pinvTol": 1e-8,
                           "l2": "inputs_exceed_L2 (tau of the MinSR subset %.0f MB, T %.1f GB)" % (NT * M * 16 / 1e6, NT * NT * 16 / 1e9),
                           "parallelism": "chains sharded over %d rank(s); T tile pairs sharded + SUM all-reduce; eigh replicated" % world},
                "e2e": {"value": nGlobal / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(P_host.numel() * 8),
                        "d2h_bytes_per_step": int(res_host.numel() * 8), "steps": args.steps,
                        "note": "parameters H2D from pinned memory; <E>, Var E and the MinSR update D2H"},
                "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
                "phases_ms": ph, "energy_mean": [energy.real, energy.imag], "update_max_abs": upd_norm}
        sys.stdout.flush()
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def tdvp_step_ms_config3(jVMC, op, torch, L=40, alpha=2, nsamp=2 ** 16, chains=2368, steps=2):
    """One real-time TDVP right-hand side (TDVP.__call__ with rhsPrefactor=1j, makeReal='imag', SNR-regularised solve)
    on BASELINE configs[2]: 1D TFIM L=40 quench, CpxRBM alpha=2 with bias (P = 6560), 2^16 samples.  An AdaptiveHeun
    step costs 5 of these per attempt (stepper.py:146-158)."""
    dev = jVMC.global_defs.myDevice
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=alpha * L, bias=True), seed=1234)
    psi(torch.zeros((1, 1, L), dtype=torch.int32, device=dev))
    W, b = o1_weights(L, alpha * L, True)
    psi.set_parameters(torch.as_tensor(flat_params(0.3 * W, 0.3 * b)))
    H = op.BranchFreeOperator()
    for l in range(L):
        H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz((l + 1) % L))))
        H.add(op.scal_opstr(-0.3, (op.Sx(l),)))
    smp = jVMC.sampler.MCSampler(psi, (L,), 4321, updateProposer=jVMC.sampler.propose_spin_flip, numChains=chains,
                                 sweepSteps=L, numSamples=nsamp, thermalizationSweeps=25)
    tdvp = jVMC.util.TDVP(smp, snrTol=2, pinvTol=1e-8, rhsPrefactor=1.j, makeReal='imag')
    times = []
    for k in range(steps + 1):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        tdvp(psi.get_parameters(), 0.0, hamiltonian=H, psi=psi, numSamples=None)
        b_.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b_))
    P = 2 * (alpha * L + L * alpha * L)
    return {"config": "1D TFIM L=%d quench, CpxRBM alpha=%d with bias (P=%d), %d chains, %d samples, real-time TDVP "
                      "(rhsPrefactor=1j, makeReal=imag, snrTol=2)" % (L, alpha, P, chains, smp.get_last_number_of_samples()),
            "ms_per_rhs": float(np.median(times[1:])), "rhs_per_adaptive_heun_attempt": 5,
            "solver_residual": _maybe_float(lambda: tdvp.get_residuals()[1])}


def aux_config5_minsr(jVMC, op, torch, K, L=20, alpha=4, nsamp=2 ** 14, chains=1184):
    """BASELINE configs[4] lattice on ONE GPU, bounded: 2D TFIM 20x20, CpxRBM alpha=4 (P_c = 640 000), sample + E_loc +
    the MinSR tangent kernel T on N_T samples + (for N_T <= 4096) the full MinSR.__call__ incl. the pseudo-inverse."""
    dev = jVMC.global_defs.myDevice
    N, M = L * L, alpha * L * L
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=False), seed=1234)
    psi(torch.zeros((1, 1, L, L), dtype=torch.int32, device=dev))
    W, _ = o1_weights(N, M, False)
    psi.set_parameters(torch.as_tensor(flat_params(W, None)))
    H = op.BranchFreeOperator()
    for x in range(L):
        for y in range(L):
            l = x * L + y
            H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(x * L + (y + 1) % L))))
            H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(((x + 1) % L) * L + y))))
            H.add(op.scal_opstr(3.04, (op.Sx(l),)))
    smp = jVMC.sampler.MCSampler(psi, (L, L), 4321, updateProposer=jVMC.sampler.propose_spin_flip, numChains=chains,
                                 sweepSteps=N, numSamples=nsamp, thermalizationSweeps=25)
    smp.refreshEvery = 8

    def timed(fn):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b_.record(); torch.cuda.synchronize()
        return r, a.elapsed_time(b_)
    smp.sample()
    (s, logPsi, p), t_s = timed(lambda: smp.sample())
    H.get_O_loc(s, psi, logPsi)
    Eloc, t_e = timed(lambda: H.get_O_loc(s, psi, logPsi))
    from vmc_jax_b200.stats import RBMGradientObs
    G = RBMGradientObs(psi, s, p)
    G.tangent_kernel()
    T, t_t = timed(lambda: G.tangent_kernel())
    B = s.shape[1]
    out = {"config": "2D TFIM %dx%d, CpxRBM alpha=%d (P_c=%d), %d chains, %d samples on one GPU" % (L, L, alpha, N * M, chains, B),
           "sampling_ms": t_s, "eloc_ms": t_e, "tangent_kernel_ms": t_t,
           "tangent_kernel_tflops_dmma": 4.0 * B * B * M / (t_t * 1e-3) / 1e12,
           "samples_per_s_sample_eloc_T": B / ((t_s + t_e + t_t) * 1e-3),
           "note": "2^20 samples make the dense N_s x N_s tangent kernel (17.6 TB) infeasible on any number of B200s; "
                   "T is formed on N_T = %d samples per step" % B}
    del T, G
    minsr = jVMC.util.MinSR(smp, pinvTol=1e-8)
    nsmall = 4096
    minsr(psi.get_parameters(), 0.0, psi=psi, hamiltonian=H, numSamples=nsmall)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    minsr(psi.get_parameters(), 0.0, psi=psi, hamiltonian=H, numSamples=nsmall)
    torch.cuda.synchronize()
    out["minsr_step_ms_4096_samples"] = (time.perf_counter() - t0) * 1e3
    return out


def aux_config4_cnn(jVMC, op, torch, L=12, nsamp=2 ** 13, chains=1184):
    """BASELINE configs[3] on ONE GPU, bounded: 2D Heisenberg J1 12x12 (Marshall-rotated), real CNN, exchange proposer:
    sample (incremental sampler, csrc/cnn_inc.cu) + E_loc (fused, incremental psi(s')/psi(s)) + dense gradients + S, F
    (DESIGN 4.5).  Every phase runs once untimed first (module load, shared-memory attributes, cuBLAS handles)."""
    dev = jVMC.global_defs.myDevice
    H = op.BranchFreeOperator(ElocBatchSize=2048)
    for x in range(L):
        for y in range(L):
            i = x * L + y
            for j in (x * L + (y + 1) % L, ((x + 1) % L) * L + y):
                H.add(op.scal_opstr(-0.25, (op.Sx(i), op.Sx(j))))
                H.add(op.scal_opstr(-0.25, (op.Sy(i), op.Sy(j))))
                H.add(op.scal_opstr(0.25, (op.Sz(i), op.Sz(j))))
    psi = jVMC.vqs.NQS(jVMC.nets.CNN(F=(3, 3), channels=(6, 4), strides=(1, 1), bias=True, firstLayerBias=False), seed=7)
    neel = np.indices((L, L)).sum(0) % 2
    smp = jVMC.sampler.MCSampler(psi, (L, L), 99, updateProposer=jVMC.sampler.propose_spin_flip_zeroMag, numChains=chains,
                                 numSamples=nsamp, thermalizationSweeps=10, sweepSteps=L * L, initState=neel)

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        return r, (time.perf_counter() - t0) * 1e3
    from vmc_jax_b200.stats import SampledObs
    s, logPsi, p = smp.sample()
    Eloc = H.get_O_loc(s, psi, logPsi)

    def stats():
        E = SampledObs(Eloc, p)
        G = SampledObs(psi.gradients(s), p)
        return E.mean(), G.covar(E), G.covar()
    stats()
    (s, logPsi, p), t_s = timed(lambda: smp.sample())
    Eloc, t_e = timed(lambda: H.get_O_loc(s, psi, logPsi))
    (Em, F, S), t_g = timed(stats)
    B = s.shape[1]
    return {"config": "2D Heisenberg %dx%d, CNN F=(3,3) channels=(6,4) (P=%d real), exchange proposer, %d chains, %d "
                      "samples on one GPU" % (L, L, psi.numParameters, chains, B),
            "sampling_ms": t_s, "eloc_ms": t_e, "gradients_S_F_ms": t_g,
            "samples_per_s": B / ((t_s + t_e + t_g) * 1e-3), "energy_per_site": float(Em.real.reshape(-1)[0]) / (L * L),
            "acceptance": float(smp.acceptance_ratio())}


def _maybe_float(f):
    try:
        return float(f())
    except Exception:
        return None


if __name__ == "__main__":
    main()
