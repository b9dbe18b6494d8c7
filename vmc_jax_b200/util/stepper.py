"""ODE integrators driving the TDVP right-hand side (mirror of jVMC/util/stepper.py; host side).
The right-hand side is called as f(y, t, **rhsArgs, intStep=k)."""
import numpy as np
import torch


def _default_norm(x):
    if isinstance(x, torch.Tensor):
        return float(torch.linalg.norm(x))
    return float(np.linalg.norm(x))


class Euler:
    """First-order explicit step y + dt f(y, t) (reference :6-40)."""

    def __init__(self, timeStep=1e-3):
        self.dt = timeStep

    def step(self, t, f, yInitial, **rhsArgs):
        dy = f(yInitial, t, **rhsArgs, intStep=0)
        return yInitial + self.dt * dy, self.dt


class Heun:
    """Second-order consistent step (reference :45-93); rhsArgs["dt"] overrides the stored step."""

    def __init__(self, timeStep=1e-3):
        self.dt = timeStep

    def set_dt(self, timeStep):
        self.dt = timeStep

    def step(self, t, f, yInitial, **rhsArgs):
        dt = rhsArgs.get("dt", self.dt)
        k0 = f(yInitial, t, **rhsArgs, intStep=0)
        k1 = f(yInitial + dt * k0, t + dt, **rhsArgs, intStep=1)
        return yInitial + 0.5 * dt * (k0 + k1), dt


class AdaptiveHeun:
    """Heun with step-size control: one full step is compared with two half steps (5 right-hand-side
    evaluations per attempt) and dt is rescaled by clip(0.9 (tol/err)^(1/3), 0.2, 2) until the error
    estimate is below ``tol`` (reference :98-182)."""

    def __init__(self, timeStep=1e-3, tol=1e-8, maxStep=1):
        self.dt = timeStep
        self.tolerance = tol
        self.maxStep = maxStep

    def step(self, t, f, y, normFunction=_default_norm, **rhsArgs):
        y0 = y.clone() if isinstance(y, torch.Tensor) else np.array(y, copy=True)
        dt = self.dt
        fe = 0.5
        while fe < 1.:
            k0 = f(y0, t, **rhsArgs, intStep=0)
            k1 = f(y0 + dt * k0, t + dt, **rhsArgs, intStep=1)
            dy0 = 0.5 * dt * (k0 + k1)
            k10 = f(y0 + 0.5 * dt * k0, t + 0.5 * dt, **rhsArgs, intStep=2)
            dy1 = 0.25 * dt * (k0 + k10)
            ymid = y0 + dy1
            k01 = f(ymid, t + 0.5 * dt, **rhsArgs, intStep=3)
            k11 = f(ymid + 0.5 * dt * k01, t + dt, **rhsArgs, intStep=4)
            dy1 = dy1 + 0.25 * dt * (k01 + k11)
            err = float(normFunction(dy1 - dy0))
            fe = float('inf') if err == 0. else self.tolerance / err    # identical estimates: accept, dt doubles
            realDt = dt
            dt = min(dt * min(max(0.9 * fe ** 0.33333, 0.2), 2.), self.maxStep)
        self.dt = dt
        return y0 + dy1, realDt
